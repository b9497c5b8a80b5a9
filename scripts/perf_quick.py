"""Quick device-resident timing of the rollout kernel: python scripts/perf_quick.py [playouts] [reps] [agents]

``agents``: comma-separated ORX_AGENT_* ids per player (opponent modelling, turn limit 2), e.g. 0,1,2,1,2,0"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oponn_b200.engine import RolloutEngine
from oponn_b200.maps import classic42
from oponn_b200.state import Rules, LEAF_RESULT_DTYPE
P = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = RolloutEngine(classic42(6), Rules(), 0)
if len(sys.argv) > 3:
    eng.set_sim_agents([int(x) for x in sys.argv[3].split(",")], 2)
leaf = np.fromfile("tests/golden/classic42_mid_s1.leaf", dtype=np.uint8)
ld = torch.from_numpy(leaf.copy()).cuda(); agg = torch.zeros(20, dtype=torch.int64, device="cuda")
for r in range(reps):
    eng.rollout_batch_device(ld.data_ptr(), 1, P, 20261019, 0, agg.data_ptr()); eng.sync()
    ms = eng.last_kernel_ms
    print(f"{os.environ.get('ORX_LIB','default')}: {P} playouts {ms:.1f} ms  {P/ms*1e3:.0f} playouts/s  wins {agg[2:8].tolist()}")
