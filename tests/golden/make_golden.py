"""Regenerates the committed fixtures under tests/golden/ with the CPU oracle (oracle/oponn_oracle.c).

    python tests/golden/make_golden.py

The Java reference cannot run in this image (no JDK; Lux SDK not vendored), so game-level goldens are
outputs of the oracle restatement, frozen here so that (a) the oracle itself is regression-pinned and
(b) the CUDA engine is compared against bytes that do not depend on the oracle being rebuilt.

Files written:
  classic42_mid_s1.leaf    one packed leaf record (SURVEY.md 8d recipe, state seed 1)
  classic42_mid_64.leaves  64 packed records, state seeds 1..64 (distinct mid-game positions)
  synth200_mid_s1.leaf     the large-map analogue (BASELINE.json config 4)
  golden_games.npz         per-game result records for those leaves under run seed 20261019
                           (+ "classic42_mid_s1_modelled": the same leaf with the line-up MODELLED_AGENTS behind
                           ModellingBoard, turn limit 2 -- written by --modelled without touching the other keys)
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oponn_b200.maps import classic42, synth200  # noqa: E402
from oponn_b200.state import PHASE_INITIAL_PLACEMENT, BoardState, Rules  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

RUN_SEED = 20261019
MODELLED_AGENTS = [0, 1, 2, 1, 2, 0]     # SimRandom, SimAngry, SimStinky, ... (oponn_b200.state.AGENT_*)


def new_game_leaf(m, state_seed: int) -> np.ndarray:
    """SimulationBoard.setupNewGame analogue (SimulationBoard.java:141-176): shuffled territories dealt
    round-robin, one army each, INITIAL_PLACEMENT about to start with player 0."""
    rng = np.random.default_rng(state_seed)
    perm = rng.permutation(m.n_territories)
    owner = [0] * m.n_territories
    for i, cc in enumerate(perm):
        owner[int(cc)] = i % m.n_players
    bs = BoardState(owner=owner, armies=[1] * m.n_territories, currentPlayer=0, currentPhase=PHASE_INITIAL_PLACEMENT,
                    turnCount=0, playerNumberOfCards=[0] * m.n_players, numberOfPlaceableArmies=4, cardProgressionPos=1)
    return bs.pack(m)


def mid_game_leaf(o: Oracle, m, state_seed: int, rounds: int = 3):
    """Initial placement + `rounds` full rounds under SimRandom, snapshot at the start of a CARDS phase with
    every player alive; the seed is bumped by 1000 until that holds (returns the seed used)."""
    seed = state_seed
    while True:
        alive, leaf = o.advance_to_leaf(new_game_leaf(m, seed), seed, 0, 1 + rounds)
        if alive == m.n_players:
            return seed, leaf
        seed += 1000


def main():
    out = {}
    m, r = classic42(6), Rules()
    o = Oracle(m, r)
    used, leaf = mid_game_leaf(o, m, 1)
    leaf.tofile(os.path.join(HERE, "classic42_mid_s1.leaf"))
    print("classic42_mid_s1: state seed used", used)
    _, games = o.rollout_batch(leaf[None, :], 512, RUN_SEED, per_game=True, threads=8)
    out["classic42_mid_s1"] = games
    leaves = []
    seeds = []
    for s in range(1, 65):
        u, lf = mid_game_leaf(o, m, s)
        leaves.append(lf); seeds.append(u)
    leaves = np.stack(leaves)
    leaves.tofile(os.path.join(HERE, "classic42_mid_64.leaves"))
    _, games = o.rollout_batch(leaves, 4, RUN_SEED, per_game=True, threads=8)
    out["classic42_mid_64"] = games
    out["classic42_mid_64_seeds"] = np.array(seeds)
    o.close()

    m2, r2 = synth200(), Rules(max_player_turns=4096)
    o2 = Oracle(m2, r2)
    used, leaf2 = mid_game_leaf(o2, m2, 1)
    leaf2.tofile(os.path.join(HERE, "synth200_mid_s1.leaf"))
    print("synth200_mid_s1: state seed used", used)
    _, games = o2.rollout_batch(leaf2[None, :], 64, RUN_SEED, per_game=True, threads=8)
    out["synth200_mid_s1"] = games
    o2.close()
    np.savez_compressed(os.path.join(HERE, "golden_games.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype.names if v.dtype.names else "")


def workload_large(n: int = 2048):
    """the same counters for the 200-territory map (BASELINE.json config 4) -> workload_synth200_mid_s1.json"""
    import json
    m, r = synth200(), Rules(max_player_turns=4096)
    o = Oracle(m, r)
    leaf = np.fromfile(os.path.join(HERE, "synth200_mid_s1.leaf"), dtype=np.uint8)
    agg, games, ct = o.rollout_batch(leaf[None, :], n, RUN_SEED, per_game=True, threads=8, counters=True)
    per = {k: v / n for k, v in ct.items()}
    A = per["tv"] + per["ev"] + per["cdf_steps"] + per["writes"] + 16 * per["r_words"]
    out = dict(leaf="synth200_mid_s1", run_seed=RUN_SEED, playouts=n, per_playout=per, algorithmic_int_ops_per_playout=A,
               formula="A = TV + EV + L(cdf_steps) + W + 16*R(words)  (SURVEY.md 8d)", step_cap_player_turns=4096,
               mean_sim_turn_count=float(games["sim_turn_count"].mean()), flagged=int(agg["flagged"][0]),
               generated_by="python tests/golden/make_golden.py --workload-large  (oracle counters, 8 threads)")
    json.dump(out, open(os.path.join(HERE, "workload_synth200_mid_s1.json"), "w"), indent=1)
    print("A =", A, "flagged", int(agg["flagged"][0]))


def workload(n: int = 16384):
    """Algorithmic work per playout of the bench leaf (A = TV + EV + L + W + 16 R), from the
    oracle's counters -> workload_classic42_mid_s1.json (read by bench.py for roofline.achieved)."""
    import json
    m, r = classic42(6), Rules()
    o = Oracle(m, r)
    leaf = np.fromfile(os.path.join(HERE, "classic42_mid_s1.leaf"), dtype=np.uint8)
    agg, games, ct = o.rollout_batch(leaf[None, :], n, RUN_SEED, per_game=True, threads=8, counters=True)
    per = {k: v / n for k, v in ct.items()}
    A = per["tv"] + per["ev"] + per["cdf_steps"] + per["writes"] + 16 * per["r_words"]
    out = dict(leaf="classic42_mid_s1", run_seed=RUN_SEED, playouts=n, per_playout=per,
               algorithmic_int_ops_per_playout=A, formula="A = TV + EV + L(cdf_steps) + W + 16*R(words)  (SURVEY.md 8d)",
               mean_sim_turn_count=float(games["sim_turn_count"].mean()), wins=[int(x) for x in agg["wins"][0]],
               generated_by="python tests/golden/make_golden.py --workload  (oracle counters, 8 threads)")
    json.dump(out, open(os.path.join(HERE, "workload_classic42_mid_s1.json"), "w"), indent=1)
    print("A =", A)


def modelled():
    """per-game records with modelled opponents, added to golden_games.npz."""
    path = os.path.join(HERE, "golden_games.npz")
    out = dict(np.load(path))
    leaf = np.fromfile(os.path.join(HERE, "classic42_mid_s1.leaf"), dtype=np.uint8)
    o = Oracle(classic42(6), Rules(), sim_agents=MODELLED_AGENTS, modelling_turn_limit=2)
    _, games = o.rollout_batch(leaf[None, :], 256, RUN_SEED, per_game=True, threads=8)
    o.close()
    out["classic42_mid_s1_modelled"] = games
    np.savez_compressed(path, **out)
    print("classic42_mid_s1_modelled", games.shape)


def results_only():
    """Round 2: the stream layout and the deal generator changed (include/orx_state.h), the LEAVES did not.  Keeps the
    committed leaf files (the bench workload stays the same mid-game position), stamps each record's deal_seed with
    deck_hash(deck) -- the stand-in for simDeck.hashCode() -- and regenerates every per-game golden from them."""
    from oponn_b200.state import BoardState, leaf_layout
    out = {}
    old = dict(np.load(os.path.join(HERE, "golden_games.npz")))
    m, r = classic42(6), Rules()

    def stamp(path, mp):
        L = leaf_layout(mp.n_territories, mp.n_players)
        a = np.fromfile(path, dtype=np.uint8).reshape(-1, L.stride)
        for i in range(len(a)):
            bs = BoardState.unpack(mp, a[i]); bs.dealSeed = None; bs.dealFromStream = False
            a[i] = bs.pack(mp)
        a.tofile(path)
        return a

    leaf = stamp(os.path.join(HERE, "classic42_mid_s1.leaf"), m)[0]
    leaves = stamp(os.path.join(HERE, "classic42_mid_64.leaves"), m)
    o = Oracle(m, r)
    out["classic42_mid_s1"] = o.rollout_batch(leaf[None, :], 512, RUN_SEED, per_game=True, threads=8)[1]
    out["classic42_mid_64"] = o.rollout_batch(leaves, 4, RUN_SEED, per_game=True, threads=8)[1]
    out["classic42_mid_64_seeds"] = old["classic42_mid_64_seeds"]
    o.close()
    m2, r2 = synth200(), Rules(max_player_turns=4096)
    leaf2 = stamp(os.path.join(HERE, "synth200_mid_s1.leaf"), m2)[0]
    o2 = Oracle(m2, r2)
    out["synth200_mid_s1"] = o2.rollout_batch(leaf2[None, :], 64, RUN_SEED, per_game=True, threads=8)[1]
    o2.close()
    np.savez_compressed(os.path.join(HERE, "golden_games.npz"), **out)
    modelled()
    workload()


if __name__ == "__main__":
    if "--workload-large" in sys.argv:
        workload_large()
    elif "--results" in sys.argv:
        results_only()
    elif "--workload" in sys.argv:
        workload()
    elif "--modelled" in sys.argv:
        modelled()
    else:
        main()
