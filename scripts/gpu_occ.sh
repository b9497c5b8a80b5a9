P=${1:-524288}
for b in 1 2 3 4 5; do echo "blocks/SM=$b"; ORX_BLOCKS_PER_SM=$b timeout 600 python scripts/perf_quick.py $P 1 2>&1 | tail -1; done
