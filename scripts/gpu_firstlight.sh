set -x

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
make -s -C oracle
timeout 600 python -m pytest tests/test_gpu_first.py -x -q -m gpu 2>&1 | tail -30
