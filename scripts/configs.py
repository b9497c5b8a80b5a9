"""Throughput of the other BASELINE.json configs on one GPU (reported in DESIGN.md; not bench lines).
config 3 kernel-only form: 64 root children x equal playouts;  config 4: synth200;  config 5 slice: many leaves."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oponn_b200.engine import RolloutEngine
from oponn_b200.maps import classic42, synth200
from oponn_b200.state import Rules, LEAF_RESULT_DTYPE

def run(name, m, r, leaves, ppl, reps=2):
    eng = RolloutEngine(m, r, 0)
    ld = torch.from_numpy(leaves.copy()).cuda()
    n = leaves.shape[0]
    agg = torch.zeros(n * 20, dtype=torch.int64, device="cuda")
    best = None
    for k in range(reps):
        eng.rollout_batch_device(ld.data_ptr(), n, ppl, 20261019, k * n * ppl, agg.data_ptr()); eng.sync()
        ms = eng.last_kernel_ms
        best = ms if best is None else min(best, ms)
    a = agg.cpu().numpy().view(LEAF_RESULT_DTYPE)
    out = dict(config=name, leaves=n, playouts_per_leaf=ppl, kernel_ms=best, playouts_per_s=n * ppl / best * 1e3,
               mean_rounds=float(a["len_sum"].sum() / a["n"].sum()), flagged=int(a["flagged"].sum()))
    print(json.dumps(out)); eng.close()

g = "tests/golden/"
l64 = np.fromfile(g + "classic42_mid_64.leaves", dtype=np.uint8).reshape(64, -1)
run("cfg3_kernel_only_64_children", classic42(6), Rules(), l64, 16384)
ls = np.fromfile(g + "synth200_mid_s1.leaf", dtype=np.uint8)[None, :]
run("cfg4_synth200_N200_P8_cap4096", synth200(), Rules(max_player_turns=4096), ls, 131072)
run("cfg5_slice_4096_states_x_256", classic42(6), Rules(), np.tile(l64, (64, 1)), 256)
