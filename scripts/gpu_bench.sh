mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_${1:-x}.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_${1:-x}.json | cut -c1-250
