"""Game-level checks of the oracle (CPU): hand-derived scenarios, frozen goldens, rule invariants."""
import os

import numpy as np
import pytest

import scenarios
from conftest import GOLDEN, RUN_SEED
from oponn_b200.maps import classic42, synth200
from oponn_b200.state import (ATTACK_EXECUTE, ATTACK_INVADE, FLAG_ERROR, FLAG_STEP_LIMIT, FLAG_STREAM_END, PHASE_ATTACK, PHASE_CARDS,
                              PHASE_COUNTRY_SELECTION, PHASE_FORTIFICATION, PHASE_INITIAL_PLACEMENT, PHASE_PLACEMENT,
                              BoardState, Rules)
from oracle.pyoracle import Oracle


@pytest.mark.parametrize("make", scenarios.ALL, ids=lambda f: f.__name__)
def test_hand_derived_scenarios(make):
    m, r, leaf, words, exp = make()
    o = Oracle(m, r)
    games = o.rollout_replay(leaf.pack(m)[None, :], words[None, :])
    o.close()
    scenarios.check(games[0], exp, m.n_players)


@pytest.mark.parametrize("make", scenarios.MODELLED, ids=lambda f: f.__name__)
def test_hand_derived_modelled_opponents(make):
    """SimAngry / SimStinky behind ModellingBoard, worked out by hand from the reference's source."""
    m, r, leaf, words, exp, agents, limit = make()
    o = Oracle(m, r, sim_agents=agents, modelling_turn_limit=limit)
    games = o.rollout_replay(leaf.pack(m)[None, :], words[None, :])
    o.close()
    scenarios.check(games[0], exp, m.n_players)


def test_golden_regression(oracle, mid_leaf, golden_games):
    """The committed per-game records are what today's oracle produces (guards the checker itself)."""
    gold = golden_games["classic42_mid_s1"]
    _, games = oracle.rollout_batch(mid_leaf[None, :], len(gold), RUN_SEED, per_game=True, threads=8)
    assert games.tobytes() == gold.tobytes()
    leaves = np.fromfile(os.path.join(GOLDEN, "classic42_mid_64.leaves"), dtype=np.uint8).reshape(64, -1)
    gold = golden_games["classic42_mid_64"]
    _, games = oracle.rollout_batch(leaves, 4, RUN_SEED, per_game=True, threads=8)
    assert games.tobytes() == gold.tobytes()


def test_golden_regression_modelled_opponents(classic, rules, mid_leaf, golden_games):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    gold = golden_games["classic42_mid_s1_modelled"]
    o = Oracle(classic, rules, sim_agents=mg.MODELLED_AGENTS, modelling_turn_limit=2)
    _, games = o.rollout_batch(mid_leaf[None, :], len(gold), RUN_SEED, per_game=True, threads=8)
    assert games.tobytes() == gold.tobytes()
    # an all-SimRandom line-up behind ModellingBoard is the plain board (ModellingBoard.java:29-47)
    o.set_sim_agents([0] * 6, 2)
    _, games = o.rollout_batch(mid_leaf[None, :], 64, RUN_SEED, per_game=True, threads=8)
    o.close()
    assert games.tobytes() == golden_games["classic42_mid_s1"][:64].tobytes()


def test_modelled_opponents_play_every_phase_to_the_end(classic, rules, mid_leaf):
    """SimAngry / SimStinky line-ups from every phase; SimStinky cannot draft (its pickCountry returns -1 and
    SimulationBoard.selectCountry throws, SimulationBoard.java:485-494) -> those playouts carry FLAG_ERROR."""
    from leafzoo import phase_zoo
    from oponn_b200.state import AGENT_ANGRY, AGENT_STINKY
    for agents in ([AGENT_ANGRY] * 6, [AGENT_STINKY] * 6, [0, 1, 2, 1, 2, 0]):
        for limit in (2, 1 << 20):
            o = Oracle(classic, rules, sim_agents=agents, modelling_turn_limit=limit)
            for base in (None, mid_leaf):
                leaves, names = phase_zoo(classic, base)
                _, games = o.rollout_batch(leaves, 4, 11, per_game=True, threads=8)
                games = games.reshape(len(names), 4)
                for i, name in enumerate(names):
                    if name.startswith("country_selection") and AGENT_STINKY in agents:
                        continue
                    assert np.all(games[i]["flags"] == 0), (agents, limit, name)
                    assert np.all((games[i]["player_out_on_turn"][:, :6] == -1).sum(axis=1) == 1), (agents, limit, name)
            o.close()


def test_golden_regression_large_map(golden_games):
    m, r = synth200(), Rules(max_player_turns=4096)
    o = Oracle(m, r)
    leaf = np.fromfile(os.path.join(GOLDEN, "synth200_mid_s1.leaf"), dtype=np.uint8)
    gold = golden_games["synth200_mid_s1"]
    _, games = o.rollout_batch(leaf[None, :], len(gold), RUN_SEED, per_game=True, threads=8)
    o.close()
    assert games.tobytes() == gold.tobytes()


def test_result_invariants(oracle, mid_leaf):
    agg, games = oracle.rollout_batch(mid_leaf[None, :], 256, 99, per_game=True, threads=8)
    pot = games["player_out_on_turn"][:, :6]
    assert np.all(games["flags"] == 0)
    assert np.all((pot == -1).sum(axis=1) == 1)                       # exactly one winner (SimulationBoard.java:862-866)
    assert np.all(pot[pot != -1] >= 0)
    assert np.all(pot.max(axis=1) <= games["sim_turn_count"])         # nobody dies after the last round
    assert np.all(games["player_out_on_turn"][:, 6:] == 0)
    # the aggregate is the exact sum of the records (what MCTSTree.backpropagation folds, MCTSTree.java:399-431)
    assert int(agg["n"][0]) == 256 and int(agg["len_sum"][0]) == int(games["sim_turn_count"].sum())
    for p in range(6):
        won = pot[:, p] == -1
        assert int(agg["wins"][0][p]) == int(won.sum())
        assert int(agg["win_len_sum"][0][p]) == int(games["sim_turn_count"][won].sum())


def test_sharding_invariance(oracle, mid_leaf):
    """game ids, not batch boundaries, select the stream: any split of a batch gives the same records"""
    _, whole = oracle.rollout_batch(mid_leaf[None, :], 64, 7, first_game=100, per_game=True, threads=4)
    _, a = oracle.rollout_batch(mid_leaf[None, :], 24, 7, first_game=100, per_game=True)
    _, b = oracle.rollout_batch(mid_leaf[None, :], 40, 7, first_game=124, per_game=True, threads=3)
    assert whole.tobytes() == a.tobytes() + b.tobytes()


def test_invalid_leaves_are_flagged(oracle, classic, mid_leaf):
    from leafzoo import invalid_leaves
    bad = invalid_leaves(classic, mid_leaf)
    _, games = oracle.rollout_batch(bad, 1, 5, per_game=True)
    assert np.all(games["flags"] == FLAG_ERROR)
    for f in ("sim_turn_count", "stream_pos", "board_hash"):
        assert np.all(games[f] == 0)
    assert np.all(games["player_out_on_turn"] == 0)


def test_step_limit_and_stream_end(classic, mid_leaf):
    o = Oracle(classic, Rules(max_player_turns=5))
    _, games = o.rollout_batch(mid_leaf[None, :], 16, 5, per_game=True)
    o.close()
    assert np.all(games["flags"] == FLAG_STEP_LIMIT)
    assert np.all(games["sim_turn_count"] <= 1)
    o = Oracle(classic, Rules())
    words = np.random.default_rng(1).integers(0, 1 << 32, size=(4, 50), dtype=np.uint64).astype(np.uint32)
    games = o.rollout_replay(np.repeat(mid_leaf[None, :], 4, axis=0), words)
    o.close()
    assert np.all(games["flags"] == FLAG_STREAM_END) and np.all(games["board_hash"] == 0)


@pytest.mark.parametrize("from_mid_game", [False, True])
def test_every_phase_plays_to_the_end(oracle, classic, mid_leaf, from_mid_game):
    from leafzoo import phase_zoo
    leaves, names = phase_zoo(classic, mid_leaf if from_mid_game else None)
    _, games = oracle.rollout_batch(leaves, 4, 11, per_game=True, threads=8)
    games = games.reshape(len(names), 4)
    for i, name in enumerate(names):
        assert np.all(games[i]["flags"] == 0), name
        assert np.all((games[i]["player_out_on_turn"][:, :6] == -1).sum(axis=1) == 1), name
