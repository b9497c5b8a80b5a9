"""GPU parity of the lane-per-game kernel (oponn_b200/csrc/orx_lanes.cuh), forced with ORX_KERNEL=lanes so that the
small batches the oracle can follow run on it too.  Through the C ABI; bit-exact per-game records."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, RUN_SEED
from leafzoo import invalid_leaves, phase_zoo, random_leaves, random_map
from oponn_b200.maps import classic42
from oponn_b200.state import BoardState, Rules

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def same(a, b):
    if a.tobytes() == b.tobytes():
        return
    bad = [i for i in range(len(a)) if a[i].tobytes() != b[i].tobytes()]
    raise AssertionError(f"{len(bad)} of {len(a)} records differ; first {bad[0]}: gpu={a[bad[0]]} oracle={b[bad[0]]}")


# scheduler settings (ORX_LANE_TUNE): the defaults; every dice battle that can be rolled by the whole warp is; no
# cooperative dice and no waiting -- results must not depend on any of it
TUNES = {"default": None, "coop_all": "4,8,8,8,64,32,8,8,1", "no_coop_no_wait": "1,1,1,1,0,0,0,0,1000000000"}


@pytest.fixture(autouse=True, params=list(TUNES))
def force_lanes(monkeypatch, request):
    monkeypatch.setenv("ORX_KERNEL", "lanes")
    if TUNES[request.param] is None:
        monkeypatch.delenv("ORX_LANE_TUNE", raising=False)
    else:
        monkeypatch.setenv("ORX_LANE_TUNE", TUNES[request.param])


def engines(m, r):
    from oponn_b200.engine import RolloutEngine
    from oracle.pyoracle import Oracle
    return RolloutEngine(m, r, 0), Oracle(m, r)


def run(m, r, leaves, ppl, seed, first_game=0):
    e, o = engines(m, r)
    try:
        agg, games = e.rollout_batch(leaves, ppl, seed, first_game=first_game, per_game=True)
        assert e.last_kernel_kind == "lanes"
        oagg, ogames = o.rollout_batch(leaves, ppl, seed, first_game=first_game, per_game=True, threads=THREADS)
        same(games, ogames)
        assert agg.tobytes() == oagg.tobytes()
    finally:
        e.close(); o.close()


def test_goldens_on_the_lane_kernel(golden_games):
    from oponn_b200.engine import RolloutEngine
    m = classic42(6)
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    leaves = np.fromfile(os.path.join(GOLDEN, "classic42_mid_64.leaves"), dtype=np.uint8).reshape(64, -1)
    with RolloutEngine(m, Rules(), 0) as e:
        _, games = e.rollout_batch(leaf[None, :], 512, RUN_SEED, per_game=True)
        assert e.last_kernel_kind == "lanes"
        same(games, golden_games["classic42_mid_s1"])
        _, games = e.rollout_batch(leaves, 4, RUN_SEED, per_game=True)
        same(games, golden_games["classic42_mid_64"])


def test_mid_game_vs_oracle():
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    run(classic42(6), Rules(), leaf[None, :], 4096, 4242, first_game=10**12)


def test_every_phase_and_attack_state():
    m = classic42(6)
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    for base in (None, leaf):
        leaves, _ = phase_zoo(m, base)
        run(m, Rules(), leaves, 8, 11)


def test_invalid_leaves():
    m = classic42(6)
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    run(m, Rules(), invalid_leaves(m, leaf), 3, 5)


@pytest.mark.parametrize("rules", [
    Rules(use_cards=False), Rules(transfer_cards=False), Rules(continent_increase=5),
    Rules(card_progression="4, 4, 6, 6, 6, 8, 8, 8, 8, 10..."), Rules(card_progression="5, 10, 15... ^10%"), Rules(max_player_turns=40),
], ids=lambda r: f"cards{int(r.use_cards)}_tr{int(r.transfer_cards)}_inc{r.continent_increase}_{r.card_progression[:12]}_{r.max_player_turns}")
def test_rule_variants(rules):
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    run(classic42(6), rules, leaf[None, :], 96, 606)


@pytest.mark.parametrize("players", [2, 3, 8])
def test_other_player_counts(players):
    m = classic42(players)
    leaves, _ = phase_zoo(m, None, seed=players)
    run(m, Rules(), leaves, 4, 99)


def test_random_maps_and_leaves_fuzz():
    rng = np.random.default_rng(20261022)
    progs = ["4, 6, 8, 10, 15, 20...", "5, 5, 5...", "4, 5, 6...", "3, 6, 9...", "4, 6, 8, 10, 15, 20, 25, 10, 10, 10..."]
    for k in range(8):
        n = int(rng.choice([5, 12, 31, 32, 33, 42, 64]))
        p = int(rng.integers(2, 9))
        m = random_map(rng, n, p, int(rng.integers(1, 9)))
        r = Rules(use_cards=bool(rng.random() < 0.8), transfer_cards=bool(rng.integers(2)), continent_increase=int(rng.choice([0, 0, 3, 10])),
                  card_progression=str(rng.choice(progs)), max_player_turns=int(rng.choice([300, 2000])))
        run(m, r, random_leaves(m, rng, 160), 2, 1234 + k)


def test_game_stream_deal_mode():
    m = classic42(6)
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    bs = BoardState.unpack(m, leaf); bs.dealFromStream = True
    run(m, Rules(), bs.pack(m)[None, :], 256, 5)


def test_all_playouts_of_a_leaf_get_the_same_hidden_hands():
    """CardManager.simReset re-seeds from simDeck.hashCode() on every setupState (CardManager.java:245): with cards
    that cannot be cashed (use_cards off changes nothing about dealing) the first deal after the leaf is the same card
    in every playout; seen through the result: two batches with different run seeds but one leaf, stopped after one
    player-turn, end on boards that differ only by dice -- while with deal_mode = GAME_STREAM they also differ by cards.
    The direct check is on the oracle side (tests/test_oracle_game.py); here: both kernels agree with it in both modes."""
    m = classic42(6)
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    for stream in (False, True):
        bs = BoardState.unpack(m, leaf); bs.dealFromStream = stream; bs.dealSeed = 12345
        run(m, Rules(max_player_turns=3), bs.pack(m)[None, :], 64, 17)


def test_auto_choice_and_agreement_of_the_two_kernels(monkeypatch):
    """with ORX_LANE_MIN_TOTAL set, a large batch goes to the lane kernel and a small one does not; without it every
    batch runs on the warp-per-game kernel; the two kernels' records are identical"""
    from oponn_b200.engine import RolloutEngine
    m = classic42(6)
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    n = 1 << 17
    monkeypatch.delenv("ORX_KERNEL")
    monkeypatch.setenv("ORX_LANE_MIN_TOTAL", str(1 << 16))
    with RolloutEngine(m, Rules(), 0) as e:
        agg, games = e.rollout_batch(leaf[None, :], n, 99, per_game=True)
        assert e.last_kernel_kind == "lanes"
        small, _ = e.rollout_batch(leaf[None, :], 1024, 99, per_game=False)
        assert e.last_kernel_kind == "warp"
    monkeypatch.delenv("ORX_LANE_MIN_TOTAL")
    with RolloutEngine(m, Rules(), 0) as e:
        agg2, games2 = e.rollout_batch(leaf[None, :], n, 99, per_game=True)
        assert e.last_kernel_kind == "warp"
    same(games, games2)
    assert agg.tobytes() == agg2.tobytes()
