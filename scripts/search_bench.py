"""Whole-move search throughput: batched UCT (host, C++) + GPU rollouts.  python scripts/search_bench.py [iters] [ppl] [lif]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oponn_b200.engine import RolloutEngine
from oponn_b200.maps import classic42
from oponn_b200.state import ATTACK_START, PHASE_ATTACK, BoardState, Rules
from oponn_b200.tree import SearchTree, TreeOptions, move_str
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ppl = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lifs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1024, 4096, 16384]
m = classic42(6)
eng = RolloutEngine(m, Rules(), 0)
leaf = np.fromfile("tests/golden/classic42_mid_s1.leaf", dtype=np.uint8)
bs = BoardState.unpack(m, leaf); bs.currentPhase, bs.attackState = PHASE_ATTACK, ATTACK_START
for t in range(42):
    if bs.owner[t] == 0: bs.armies[t] = 6
for lif in lifs:
    tree = SearchTree(m, Rules(), TreeOptions(agent_id=0, playouts_per_leaf=ppl, leaves_in_flight=lif, seed=7), eng)
    t0 = time.perf_counter(); move, ch, st = tree.runMCTS(bs, iters); dt = time.perf_counter() - t0
    print(json.dumps(dict(iterations=int(st["iterations"]), playouts_per_leaf=ppl, leaves_in_flight=lif, seconds=dt,
                          playouts_per_s=int(st["playouts"]) / dt, leaves_per_s=int(st["iterations"]) / dt, nodes=int(st["nodes"]),
                          max_depth=int(st["max_depth"]), batches=int(st["batches"]), select_ms=float(st["select_ms"]),
                          rollout_ms=float(st["rollout_ms"]), backprop_ms=float(st["backprop_ms"]), best=move_str(move),
                          agent_win_rate=float(ch["wins"][:, 0].sum() / ch["visits"].sum()))))
    tree.close()
