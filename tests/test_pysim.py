"""The C oracle against a second, independent restatement of the rollout path written from the Java in pure Python
(oracle/pysim.py): per-game records must be identical.  This is the substitute for running the Java reference (no JVM
in this image): a misreading of SimulationBoard / SimRandom / CardManager would have to be made twice, in two
languages and two program structures, to go unnoticed.  > 500 playouts: every phase and attack sub-state, rule
variants, 2..8 players, random maps with big stacks (single rolls above 100, the gaussian branch), both deal modes,
and a few games played to the end."""
import os
import sys
from functools import lru_cache

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)

from leafzoo import phase_zoo, random_leaves, random_map  # noqa: E402
from oponn_b200.maps import classic42  # noqa: E402
from oponn_b200.state import BoardState, Rules  # noqa: E402
from oracle import pyoracle, pysim  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")
pysim.STRICT_LOG = pyoracle.fdlibm_log


@lru_cache(maxsize=None)
def outcomes(a: int, d: int):
    pr, al, dl, inv = pyoracle.get_attack_outcomes(a, d)      # BattleSim's tables are pinned on their own (test_oracle_battlesim.py)
    return [(float(pr[i]), int(al[i]), int(dl[i])) for i in range(len(pr))]


def compare(m, r, leaves, ppl, seed, agents=None, limit=2):
    """pysim v oracle on every (leaf, playout); returns the number of games compared"""
    o = Oracle(m, r) if agents is None else Oracle(m, r, sim_agents=agents, modelling_turn_limit=limit)
    try:
        _, want = o.rollout_batch(leaves, ppl, seed, per_game=True, threads=8)
    finally:
        o.close()
    k = 0
    for i in range(len(leaves)):
        bs = BoardState.unpack(m, leaves[i])
        for j in range(ppl):
            got = pysim.play(m, r, bs, outcomes, seed=seed, game=i * ppl + j, sim_agents=agents, modelling_turn_limit=limit)
            w = want[i * ppl + j]
            exp = dict(player_out_on_turn=[int(x) for x in w["player_out_on_turn"]], sim_turn_count=int(w["sim_turn_count"]),
                       flags=int(w["flags"]), stream_pos=int(w["stream_pos"]), board_hash=int(w["board_hash"]))
            assert got == exp, f"leaf {i} playout {j}: pysim {got} != oracle {exp}"
            k += 1
    return k


def mid_leaf():
    return np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)


def test_every_phase_and_attack_state():
    m = classic42(6)
    n = 0
    for base in (None, mid_leaf()):
        leaves, _ = phase_zoo(m, base)
        n += compare(m, Rules(max_player_turns=30), leaves, 3, 4242)
    assert n >= 120


@pytest.mark.parametrize("rules", [
    Rules(use_cards=False, max_player_turns=40), Rules(transfer_cards=False, max_player_turns=200), Rules(continent_increase=5, max_player_turns=40),
    Rules(card_progression="4, 4, 6, 6, 6, 8, 8, 8, 8, 10...", max_player_turns=60), Rules(card_progression="5, 10, 15... ^10%", max_player_turns=60),
    Rules(card_progression="0", max_player_turns=40), Rules(card_progression="nonsense", max_player_turns=40),
], ids=lambda r: f"cards{int(r.use_cards)}_tr{int(r.transfer_cards)}_inc{r.continent_increase}_{r.card_progression[:10]}")
def test_rule_variants(rules):
    assert compare(classic42(6), rules, mid_leaf()[None, :], 12, 606) == 12


@pytest.mark.parametrize("players", [2, 3, 8])
def test_other_player_counts(players):
    m = classic42(players)
    leaves, _ = phase_zoo(m, None, seed=players)
    assert compare(m, Rules(max_player_turns=25), leaves, 2, 99) >= 40


def test_random_maps_with_big_stacks():
    """random_leaves draws armies up to 2000 per territory: single rolls above 100 and the > 500 v > 500 gaussian branch
    happen in the first turns"""
    rng = np.random.default_rng(20261023)
    n = 0
    for k in range(5):
        nt = int(rng.choice([5, 12, 33, 42, 64])); p = int(rng.integers(2, 9))
        m = random_map(rng, nt, p, int(rng.integers(1, 9)))
        r = Rules(use_cards=bool(rng.random() < 0.8), transfer_cards=bool(rng.integers(2)), continent_increase=int(rng.choice([0, 3])),
                  max_player_turns=int(rng.choice([6, 12])))
        n += compare(m, r, random_leaves(m, rng, 30), 1, 1234 + k)
    assert n == 150


def test_deals_from_the_game_stream_option():
    m = classic42(6)
    bs = BoardState.unpack(m, mid_leaf()); bs.dealFromStream = True
    assert compare(m, Rules(max_player_turns=60), bs.pack(m)[None, :], 8, 5) == 8


def test_all_playouts_of_a_leaf_see_the_same_hidden_hands():
    """CardManager.simReset: `random = new Random(simDeck.hashCode())` on every setupState (CardManager.java:240-248) --
    in the default deal mode the hidden hands after setupState are the same for every playout of a leaf and depend on
    the seed the host injects; with deals from the game stream they differ from playout to playout."""
    m = classic42(6)
    bs = BoardState.unpack(m, mid_leaf())

    def hands(seed_val, game, from_stream=False):
        b = pysim.SimulationBoard(m, Rules(), outcomes)
        bs.dealSeed, bs.dealFromStream = seed_val, from_stream
        b.setup_state(bs, pysim.Stream(77, game))
        return [tuple(h) for h in b.hands]

    assert hands(123, 0) == hands(123, 1) == hands(123, 999)
    assert hands(123, 0) != hands(124, 0)
    assert hands(123, 0, True) != hands(123, 1, True)
    assert sum(len(h) for h in hands(123, 0)) == len(bs.hand) + sum(n for p, n in enumerate(bs.playerNumberOfCards) if p != bs.currentPlayer)


def test_games_played_to_the_end():
    """full-length playouts (hundreds of rounds, hundreds of thousands of dice): a handful, they are slow in pure Python"""
    assert compare(classic42(6), Rules(), mid_leaf()[None, :], 16, 20261019) == 16


def test_hand_derived_scenarios_through_pysim():
    """the hand-derived expectations of tests/scenarios.py also hold for the independent restatement"""
    import scenarios
    for make in scenarios.ALL:
        m, r, leaf, words, exp = make()
        got = pysim.play(m, r, BoardState.unpack(m, leaf.pack(m)), outcomes, words=list(words))
        assert got["player_out_on_turn"][:m.n_players] == exp["pot"], make.__name__
        assert (got["sim_turn_count"], got["flags"], got["stream_pos"], got["board_hash"]) == (exp["turns"], exp["flags"], exp["stream_pos"], exp["hash"]), make.__name__


# ---- opponent modelling (row f4): ModellingBoard + SimAngry + SimStinky, restated a second time from the Java -----------------

LINEUPS = [([1] * 6, 2), ([2] * 6, 2), ([0, 1, 2, 1, 2, 0], 2), ([1, 2, 0, 0, 2, 1], 1), ([2, 1, 1, 2, 1, 2], 1 << 20)]


@pytest.mark.parametrize("agents,limit", LINEUPS, ids=lambda x: "".join(map(str, x)) if isinstance(x, list) else str(x))
def test_modelled_opponents_every_phase(agents, limit):
    """SimAngry / SimStinky line-ups behind ModellingBoard from every phase and attack sub-state (SimStinky cannot draft: its
    playouts from COUNTRY_SELECTION leaves carry FLAG_ERROR in both restatements)"""
    m = classic42(6)
    n = 0
    for base in (None, mid_leaf()):
        leaves, _ = phase_zoo(m, base)
        n += compare(m, Rules(max_player_turns=40), leaves, 2, 777, agents, limit)
    assert n >= 80


def test_modelled_opponents_random_maps_and_rules():
    rng = np.random.default_rng(20261024)
    n = 0
    for k in range(6):
        nt = int(rng.choice([5, 12, 33, 42, 64])); p = int(rng.integers(2, 9))
        m = random_map(rng, nt, p, int(rng.integers(1, 9)))
        r = Rules(use_cards=bool(rng.random() < 0.8), transfer_cards=bool(rng.integers(2)), continent_increase=int(rng.choice([0, 3])),
                  max_player_turns=int(rng.choice([8, 16])))
        agents = [int(x) for x in rng.integers(0, 3, size=p)]
        n += compare(m, r, random_leaves(m, rng, 20), 1, 4321 + k, agents, int(rng.choice([1, 2, 3])))
    assert n == 120


def test_modelled_games_played_to_the_end():
    assert compare(classic42(6), Rules(), mid_leaf()[None, :], 6, 20261019, [0, 1, 2, 1, 2, 1], 2) == 6


def test_hand_derived_modelled_scenarios_through_pysim():
    import scenarios
    for make in scenarios.MODELLED:
        m, r, leaf, words, exp, agents, limit = make()
        got = pysim.play(m, r, BoardState.unpack(m, leaf.pack(m)), outcomes, words=list(words), sim_agents=agents, modelling_turn_limit=limit)
        assert got["player_out_on_turn"][:m.n_players] == exp["pot"], make.__name__
        assert (got["sim_turn_count"], got["flags"], got["stream_pos"], got["board_hash"]) == (exp["turns"], exp["flags"], exp["stream_pos"], exp["hash"]), make.__name__
