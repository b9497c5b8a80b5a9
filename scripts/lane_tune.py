"""Times the rollout kernels on the bench leaf under different scheduler settings (ORX_LANE_TUNE) in one process.

    python scripts/lane_tune.py [playouts] [tune ...]      tune = "thr_new,thr_turn,thr_sel,thr_fin,age_new,age_turn,age_sel,age_fin,coop_min" | warp

Prints playouts/s from the kernel's own CUDA-event time (orx_last_kernel_ms).  A tuning aid, not the bench.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oponn_b200.engine import RolloutEngine  # noqa: E402
from oponn_b200.maps import classic42  # noqa: E402
from oponn_b200.state import Rules  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    tunes = sys.argv[2:] or ["warp", "4,8,8,8,64,32,8,8,1000000000"]
    leaf = np.fromfile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "classic42_mid_s1.leaf"), dtype=np.uint8)
    ref = None
    for t in tunes:
        if t == "warp":
            os.environ["ORX_KERNEL"] = "warp"
        else:
            os.environ["ORX_KERNEL"] = "lanes"
            parts = t.split("/")
            os.environ["ORX_LANE_TUNE"] = parts[0]
            if len(parts) > 1:
                os.environ["ORX_LANE_BLOCKS_PER_SM"] = parts[1]
            else:
                os.environ.pop("ORX_LANE_BLOCKS_PER_SM", None)
        with RolloutEngine(classic42(6), Rules(), 0) as e:
            agg, _ = e.rollout_batch(leaf[None, :], n, 20261019, per_game=False)
            ms = e.last_kernel_ms
            wins = [int(x) for x in agg["wins"][0][:6]]
            if ref is None:
                ref = wins
            print(f"{t:40s} kernel={e.last_kernel_kind:5s} {ms:10.1f} ms  {n / ms * 1e3:12.0f} playouts/s  same_result={wins == ref} blocks_per_sm={e.kernel_info(n)['blocks_per_sm']} wins={wins}", flush=True)


if __name__ == "__main__":
    main()
