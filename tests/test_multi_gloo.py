"""The N > 1 path on CPU: two gloo ranks shard a move's playouts and merge with one all-reduce; the result must
equal the single-rank batch exactly.  The per-rank rollout provider is the oracle (no GPU on this box)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, RUN_SEED


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oponn_b200.maps import classic42
    from oponn_b200.rootparallel import ShardedRollout
    from oponn_b200.state import Rules
    from oracle.pyoracle import Oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o = Oracle(classic42(6), Rules())
    leaves = np.fromfile(os.path.join(GOLDEN, "classic42_mid_64.leaves"), dtype=np.uint8).reshape(64, -1)[:3]
    merged = ShardedRollout(o, rank, world).rollout(leaves, 7, RUN_SEED, first_game=1000)
    np.save(os.path.join(out_dir, f"agg{rank}.npy"), merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path, oracle):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a0 = np.load(tmp_path / "agg0.npy"); a1 = np.load(tmp_path / "agg1.npy")
    assert a0.tobytes() == a1.tobytes()
    leaves = np.fromfile(os.path.join(GOLDEN, "classic42_mid_64.leaves"), dtype=np.uint8).reshape(64, -1)[:3]
    want, _ = oracle.rollout_batch(leaves, 7, RUN_SEED, first_game=1000, per_game=False)
    assert a0.tobytes() == want.tobytes()


def test_shard_covers_everything():
    from oponn_b200.rootparallel import shard
    for total in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            parts = [shard(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (lo, c), (lo2, _) in zip(parts, parts[1:]):
                assert lo + c == lo2
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
