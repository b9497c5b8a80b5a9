# usage: bash scripts/gpu_scale.sh N   -> bench.py at N GPUs (full size), JSON line saved under gpurun_out/
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/scale_n1.json | cut -c1-330
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/scale_n$N.json | cut -c1-330
fi
