"""The lane-per-game kernel's game logic, compiled for the CPU (tests/tools/lane_emulator), against the oracle.

Bit-exact per-game records on the goldens, every phase / attack sub-state, rule variants, other player counts,
random maps and invalid records -- the same cases the GPU parity suite runs, here without a GPU.  This checks the
kernel's RULES (orx_lanes.cuh is compiled unchanged); the CUDA build itself is checked by tests/test_gpu_parity.py.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)

import lane_emulator as le  # noqa: E402
from leafzoo import invalid_leaves, phase_zoo, random_leaves, random_map  # noqa: E402
from oponn_b200.maps import classic42  # noqa: E402
from oponn_b200.state import Rules  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

RUN_SEED = 20261019
GOLDEN = os.path.join(HERE, "golden")


def check(m, r, leaves, ppl, seed, tune=le.DEFAULT_TUNE, n_warps=1):
    o = Oracle(m, r)
    try:
        _, want = o.rollout_batch(leaves, ppl, seed, per_game=True, threads=8)
    finally:
        o.close()
    got = le.emulate(m, r, leaves, ppl, seed, tune=tune, n_warps=n_warps)
    bad = [i for i in range(len(want)) if got[i].tobytes() != want[i].tobytes()]
    assert not bad, f"{len(bad)} of {len(want)} games differ, first {bad[0]}: {got[bad[0]]} != {want[bad[0]]}"


def test_goldens():
    m, r = classic42(6), Rules()
    gold = np.load(os.path.join(GOLDEN, "golden_games.npz"))
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    got = le.emulate(m, r, leaf[None, :], 128, RUN_SEED)
    assert got.tobytes() == gold["classic42_mid_s1"][:128].tobytes()
    leaves = np.fromfile(os.path.join(GOLDEN, "classic42_mid_64.leaves"), dtype=np.uint8).reshape(64, -1)
    got = le.emulate(m, r, leaves, 4, RUN_SEED, n_warps=3)
    assert got.tobytes() == gold["classic42_mid_64"].tobytes()


@pytest.mark.parametrize("tune", [(1, 1, 1, 1, 0, 0, 0, 0, 1 << 30), (32, 32, 32, 32, 5, 5, 5, 5, 1 << 30), (4, 8, 8, 8, 64, 32, 8, 8, 48),
                                  (4, 8, 8, 8, 64, 32, 8, 8, 1)])
def test_results_do_not_depend_on_the_schedule(tune):
    m, r = classic42(6), Rules()
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    check(m, r, leaf[None, :], 48, 77, tune=tune, n_warps=2)


def test_every_phase_and_attack_state():
    m, r = classic42(6), Rules()
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    for base in (leaf, None):
        leaves, names = phase_zoo(m, base)
        check(m, r, leaves, 6, 4242)


def test_invalid_leaves():
    m, r = classic42(6), Rules()
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    check(m, r, invalid_leaves(m, leaf), 2, 9)


@pytest.mark.parametrize("rules", [
    Rules(use_cards=False), Rules(transfer_cards=False), Rules(continent_increase=5), Rules(continent_increase=-10),
    Rules(card_progression="5, 5, 5..."), Rules(card_progression="4, 4, 6, 6, 6, 8, 8, 8, 8, 10..."),
    Rules(card_progression="5, 10, 15... ^10%"), Rules(card_progression="0"), Rules(max_player_turns=40),
], ids=lambda r: f"cards{int(r.use_cards)}_tr{int(r.transfer_cards)}_inc{r.continent_increase}_{r.card_progression[:12]}_{r.max_player_turns}")
def test_rule_variants(rules):
    m = classic42(6)
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    check(m, rules, leaf[None, :], 40, 606)


@pytest.mark.parametrize("players", [2, 3, 8])
def test_other_player_counts(players):
    m = classic42(players)
    leaves, names = phase_zoo(m, None, seed=players)
    check(m, Rules(), leaves, 3, 99)


def test_random_maps_and_leaves_fuzz():
    rng = np.random.default_rng(20261021)
    progs = ["4, 6, 8, 10, 15, 20...", "5, 5, 5...", "4, 5, 6...", "3, 6, 9...", "4, 6, 8, 10, 15, 20, 25, 10, 10, 10..."]
    for k in range(8):
        n = int(rng.choice([5, 12, 31, 32, 33, 42, 64]))
        p = int(rng.integers(2, 9))
        m = random_map(rng, n, p, int(rng.integers(1, 9)))
        r = Rules(use_cards=bool(rng.random() < 0.8), transfer_cards=bool(rng.integers(2)), continent_increase=int(rng.choice([0, 0, 3, 10])),
                  card_progression=str(rng.choice(progs)), max_player_turns=int(rng.choice([300, 2000])))
        leaves = random_leaves(m, rng, 80)
        check(m, r, leaves, 2, 1234 + k, n_warps=2)


def test_game_stream_deal_mode():
    """ORX_DEAL_GAME_STREAM (the option): deals draw from the playout's own stream"""
    from oponn_b200.state import BoardState
    m, r = classic42(6), Rules()
    leaf = np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)
    bs = BoardState.unpack(m, leaf); bs.dealFromStream = True
    check(m, r, bs.pack(m)[None, :], 40, 5)


def test_cooperative_dice_algorithm_equals_sequential_rolls(tmp_path):
    """tests/tools/lane_emulator/coop_model.cpp: the warp-cooperative dice block of lane_coop_dice (pair layout for every
    stream alignment, the REDUX shortcut, the prefix-sum scan for the stopping roll), transliterated with explicit
    lanes, against rolling the same doubles one after the other -- ~185 000 random (attackers, defenders, position)."""
    import subprocess
    exe = str(tmp_path / "coop_model")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", exe, os.path.join(HERE, "tools", "lane_emulator", "coop_model.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert out.strip().endswith(", 0 bad"), out
