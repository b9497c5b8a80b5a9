"""State-parallel self-play (the shape of BASELINE.json config 5): G independent games advance in lockstep; at every
decision all legal moves of all games are evaluated by ONE rollout batch (flat Monte-Carlo, playouts per child move)
and each game plays its best move; chance nodes are sampled from the battle tables.  Headless: new games from
orx_new_game, moves from the native leaf producer, playouts on the GPU.

    python scripts/selfplay.py [games] [playouts_per_move] [decisions]
"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oponn_b200.engine import RolloutEngine
from oponn_b200.maps import classic42
from oponn_b200.state import Rules
from oponn_b200.tree import MOVE_ATTACK_OUTCOME, SearchTree, TreeOptions, new_game

G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ppl = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
m = classic42(6)
eng = RolloutEngine(m, Rules(), int(os.environ.get("LOCAL_RANK", 0)))
tree = SearchTree(m, Rules(), TreeOptions(), None)
rng = np.random.default_rng(1)
states = [(new_game(m, 1000 + g), None, True) for g in range(G)]
playouts = decisions = forced = 0
t0 = time.perf_counter()
for step in range(steps):
    cand, owner_of = [], []            # child leaves of every undecided game, and (game, move index, extra)
    nxt = list(states)
    for g, (leaf, extra, alive) in enumerate(states):
        if not alive:
            continue
        mv = tree.getPossibleMoves(leaf, extra)
        if int(mv[0]["type"]) == MOVE_ATTACK_OUTCOME:      # chance node: nature moves
            p = np.asarray(mv["probability"], dtype=np.float64)
            l2, e2, on = tree.updateState(leaf, mv[int(rng.choice(len(mv), p=p / p.sum()))], extra)
            nxt[g] = (l2, e2, on); continue
        if len(mv) == 1:
            l2, e2, on = tree.updateState(leaf, mv[0], extra); nxt[g] = (l2, e2, on); forced += 1; continue
        for k in range(len(mv)):
            l2, e2, on = tree.updateState(leaf, mv[k], extra)
            cand.append(l2); owner_of.append((g, k, e2, on, int(leaf[14])))
    if cand:
        agg, _ = eng.rollout_batch(np.stack(cand), ppl, 777, first_game=playouts)
        playouts += len(cand) * ppl
        best = {}
        for i, (g, k, e2, on, cp) in enumerate(owner_of):
            score = 1.0 if not on else agg["wins"][i][cp] / max(int(agg["n"][i]), 1)
            if g not in best or score > best[g][0]:
                best[g] = (score, cand[i], e2, on)
        for g, (score, l2, e2, on) in best.items():
            nxt[g] = (l2, e2, on); decisions += 1
    states = nxt
dt = time.perf_counter() - t0
print(json.dumps(dict(games=G, playouts_per_move=ppl, lockstep_steps=steps, decisions=decisions, forced_moves=forced,
                      playouts=playouts, seconds=dt, playouts_per_s=playouts / dt, decisions_per_s=decisions / dt,
                      games_still_running=sum(1 for s in states if s[2]))))
