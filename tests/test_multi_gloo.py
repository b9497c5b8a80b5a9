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


def _merge_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oponn_b200.tree import ROOT_CHILD_DTYPE, merge_root_children
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    ch = np.zeros(5, dtype=ROOT_CHILD_DTYPE)
    ch["move"]["type"] = 4; ch["move"]["a"] = np.arange(5)
    for f in ("wins", "losses", "win_len_sum", "loss_len_sum"):
        ch[f] = rng.integers(0, 1000, size=ch[f].shape)
    ch["visits"] = rng.integers(0, 5000, size=5)
    np.save(os.path.join(out_dir, f"local{rank}.npy"), ch)
    np.save(os.path.join(out_dir, f"merged{rank}.npy"), merge_root_children(ch))
    dist.barrier()
    dist.destroy_process_group()


def test_root_parallel_merge(tmp_path):
    """independent trees merge by summing the integer root statistics with one all-reduce"""
    port = 31500 + os.getpid() % 2000
    mp.spawn(_merge_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    l0, l1 = np.load(tmp_path / "local0.npy"), np.load(tmp_path / "local1.npy")
    m0, m1 = np.load(tmp_path / "merged0.npy"), np.load(tmp_path / "merged1.npy")
    assert m0.tobytes() == m1.tobytes()
    for f in ("visits", "wins", "losses", "win_len_sum", "loss_len_sum"):
        assert np.array_equal(m0[f], l0[f] + l1[f])
    assert m0["move"].tobytes() == l0["move"].tobytes()
