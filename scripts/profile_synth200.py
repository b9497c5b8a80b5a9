"""one launch of the 200-territory kernel (BASELINE.json config 4) for ncu: python scripts/profile_synth200.py [playouts]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oponn_b200.engine import RolloutEngine
from oponn_b200.maps import synth200
from oponn_b200.state import Rules
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
leaf = np.fromfile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "synth200_mid_s1.leaf"), dtype=np.uint8)
with RolloutEngine(synth200(), Rules(max_player_turns=4096), 0) as e:
    agg, _ = e.rollout_batch(leaf[None, :], n, 20261019, per_game=False)
    print(f"synth200 {n} playouts {e.last_kernel_ms:.1f} ms {n / e.last_kernel_ms * 1e3:.0f} playouts/s", e.kernel_info(n))
