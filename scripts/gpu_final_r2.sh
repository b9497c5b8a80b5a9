# Final round-2 check on one B200: smoke, the whole -m gpu suite, then both bench arms exactly as the driver runs them.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2final_smoke.log 2>&1; tail -n 2 gpurun_out/r2final_smoke.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2final_tests.log 2>&1; tail -n 3 gpurun_out/r2final_tests.log
timeout 400 python bench.py --impl reference --gpus 1 > gpurun_out/r2final_bench_reference_n1.json 2> gpurun_out/r2final_bench_reference_n1.err
timeout 600 python bench.py --gpus 1 > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.err
tail -n 1 gpurun_out/r2final_bench_reference_n1.json | cut -c1-300
tail -n 1 gpurun_out/r2final_bench_n1.json | cut -c1-1500
