"""Headless match: the GPU MCTS agent (batched UCT + GPU rollouts, player 0) against five uniformly-random movers on
the classic map, from new games to the end.  End-to-end exercise of rows f1 + f2 + f3 on top of the rollout engine.

    python scripts/match.py [games] [iterations_per_decision] [max_decisions]
"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oponn_b200.engine import RolloutEngine
from oponn_b200.maps import classic42
from oponn_b200.state import BoardState, Rules
from oponn_b200.tree import MOVE_ATTACK_OUTCOME, MOVE_NEXT_PHASE, SearchTree, TreeOptions, new_game

games = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 6000
m = classic42(6)
eng = RolloutEngine(m, Rules(), 0)
tree = SearchTree(m, Rules(), TreeOptions(agent_id=0, playouts_per_leaf=8, leaves_in_flight=1024, seed=5), eng)
rng = np.random.default_rng(7)
results = []
t0 = time.perf_counter()
for g in range(games):
    leaf, extra, on = new_game(m, 500 + g), None, True
    decisions = searches = playouts = 0
    while on and decisions < cap:
        mv = tree.getPossibleMoves(leaf, extra)
        cp = int(leaf[14])
        if int(mv[0]["type"]) == MOVE_ATTACK_OUTCOME:
            p = np.asarray(mv["probability"], dtype=np.float64); k = int(rng.choice(len(mv), p=p / p.sum()))
        elif len(mv) == 1:
            k = 0
        elif cp == 0:
            ch, st = tree.search(leaf, iters, extra)
            k = tree.choose(ch, tie_break=decisions); searches += 1; playouts += int(st["playouts"])
        else:
            k = int(rng.integers(len(mv)))
        leaf, extra, on = tree.updateState(leaf, mv[k], extra)
        decisions += 1
    s = BoardState.unpack(m, leaf)
    owners = sorted(set(s.owner))
    results.append(dict(game=g, finished=not on, decisions=decisions, agent_searches=searches, playouts=playouts,
                        winner=owners[0] if len(owners) == 1 else None, agent_territories=s.owner.count(0), round=s.turnCount))
    print(json.dumps(results[-1]), flush=True)
print(json.dumps(dict(games=games, agent_wins=sum(1 for r in results if r["winner"] == 0), seconds=time.perf_counter() - t0,
                      iterations_per_decision=iters)))
