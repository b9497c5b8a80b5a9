set -x
nproc
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 1 --warmup 1 --playouts 65536 --no-cpu 2>&1 | tail -3
timeout 900 python bench.py --steps 2 --warmup 1 2>&1 | tail -3
