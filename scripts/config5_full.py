"""BASELINE.json configs[4] at full size: 4,096 leaf states x 16,384 playouts per state, states partitioned over the GPUs, no collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/config5_full.py [playouts_per_state]

Every rank plays its slice of the states in ONE orx_rollout_batch_device call (game ids disjoint per state); time = device work bracketed by
synchronize + barrier, max over ranks; rank 0 prints one JSON line.  (bench.py's extra_configs block runs a 1/16 slice of this on every bench.)
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oponn_b200.engine import RolloutEngine  # noqa: E402
from oponn_b200.maps import classic42  # noqa: E402
from oponn_b200.state import LEAF_RESULT_DTYPE, Rules  # noqa: E402


def main():
    ppl = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = RolloutEngine(classic42(6), Rules(), device=local)
    l64 = np.fromfile(os.path.join(ROOT, "tests", "golden", "classic42_mid_64.leaves"), dtype=np.uint8).reshape(64, -1)
    n_states = 4096
    lo, hi = n_states * rank // world, n_states * (rank + 1) // world
    mine = np.ascontiguousarray(np.tile(l64, (n_states // 64, 1))[lo:hi])
    ld = torch.from_numpy(mine).to(dev)
    words = LEAF_RESULT_DTYPE.itemsize // 8
    agg = torch.zeros((hi - lo) * words, dtype=torch.int64, device=dev)
    eng.rollout_batch_device(ld.data_ptr(), 1, 4096, 1, 0, agg.data_ptr()); eng.sync(); agg.zero_()   # warm-up launch
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    eng.rollout_batch_device(ld.data_ptr(), hi - lo, ppl, 20261019, lo * ppl, agg.data_ptr())
    eng.sync()
    dt = time.perf_counter() - t0
    kms = eng.last_kernel_ms
    t = torch.tensor([dt, kms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    a = agg.cpu().numpy().view(LEAF_RESULT_DTYPE)
    ok = torch.tensor([int((a["n"] == ppl).all()), int(a["flagged"].sum())], dtype=torch.int64, device=dev)
    wins = torch.from_numpy(a["wins"].sum(axis=0).astype(np.int64)).to(dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.SUM); dist.all_reduce(wins, op=dist.ReduceOp.SUM)   # (reporting only: after the timed region)
    if rank == 0:
        print(json.dumps({"config": "BASELINE.json configs[4]: 4,096 leaf states x %d playouts per state, classic42, 6 players, all-SimRandom; states "
                                    "partitioned over the GPUs, one launch per GPU, no data-path collective" % ppl,
                          "n_gpus": world, "states": n_states, "playouts_per_state": ppl, "total_playouts": n_states * ppl,
                          "seconds_max_over_ranks": float(t[0]), "kernel_ms_max_over_ranks": float(t[1]),
                          "playouts_per_s": n_states * ppl / float(t[0]), "ranks_with_every_state_complete": int(ok[0]), "flagged_playouts": int(ok[1]),
                          "wins_per_player": [int(x) for x in wins[:6]], "kernel": "orx::rollout_kernel<2,false,false>"}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
