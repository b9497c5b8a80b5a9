"""The parity tests proper: the CUDA engine, called through the C ABI (ctypes -> liborx.so), against the CPU oracle
on the same seeded inputs, against the committed goldens, against hand-derived scenarios, and -- at the full
BASELINE size -- through size-independent properties.  Integer work: the bar is bit-exact."""
import os

import numpy as np
import pytest

import scenarios
from battlesim_ref import CONSTANTS, junit_pairs
from conftest import GOLDEN, RUN_SEED
from javarandom import lcg_words
from leafzoo import invalid_leaves, phase_zoo
from oponn_b200.maps import classic42, synth200
from oponn_b200.state import FLAG_ERROR, FLAG_STEP_LIMIT, FLAG_STREAM_END, LEAF_RESULT_DTYPE, GAME_RESULT_DTYPE, Rules

pytestmark = pytest.mark.gpu

THREADS = os.cpu_count() or 1


def same(a: np.ndarray, b: np.ndarray):
    if a.tobytes() == b.tobytes():
        return
    bad = [i for i in range(len(a)) if a[i].tobytes() != b[i].tobytes()]
    raise AssertionError(f"{len(bad)} of {len(a)} records differ; first {bad[0]}: gpu={a[bad[0]]} oracle={b[bad[0]]}")


def pair(m, r):
    from oponn_b200.engine import RolloutEngine
    from oracle.pyoracle import Oracle
    return RolloutEngine(m, r, 0), Oracle(m, r)


# ---- stream --------------------------------------------------------------------------------------------------
def test_stream_counter_mode(engine):
    from oracle import pyoracle
    ops = [-2] * 300 + [0] * 10 + [7, 10, 1, 2, 3, 1 << 20, (1 << 30) + 1, 100, 1023, 1024, 1025, 44, 6] * 8 + [-1] * 5 + [0, 5, -1, -1]
    for seed, game in ((123456789123, 987654321), (0, 0), (2**64 - 1, 2**64 - 1), (RUN_SEED, 5)):
        g, gp, gf = engine.stream_draws(ops, seed=seed, game=game)
        o, op, of = pyoracle.stream_draws(ops, seed=seed, game=game)
        assert (gp, gf) == (op, of)
        assert np.array_equal(g.view(np.uint64), o.view(np.uint64))


def test_stream_replay_mode_and_java_known_answers(engine):
    from oracle import pyoracle
    w = lcg_words(0, 64)
    v, _, _ = engine.stream_draws([-1, -1, -1, -1], words=w)
    assert list(v) == [0.8025330637390305, -0.9015460884175122, 2.080920790428163, 0.7637707684364894]
    v, _, _ = engine.stream_draws([0], words=w)
    assert v[0] == 0.730967787376657
    rng = np.random.default_rng(5)
    words = rng.integers(0, 1 << 32, size=400, dtype=np.uint64).astype(np.uint32)
    ops = [int(x) for x in rng.choice([0, -1, -2, 3, 6, 7, 42, 1000, 4096, (1 << 31) - 1], size=150)]
    g, gp, gf = engine.stream_draws(ops, words=words)
    o, op, of = pyoracle.stream_draws(ops, words=words)
    assert (gp, gf) == (op, of) and np.array_equal(g.view(np.uint64), o.view(np.uint64))
    # running out of words, and an illegal bound
    _, _, gf = engine.stream_draws([-2, -2, -2], words=words[:2])
    assert gf & FLAG_STREAM_END


# ---- BattleSim -----------------------------------------------------------------------------------------------
def test_every_outcome_table_is_bit_exact(engine):
    """all 100 x 100 tables built on the device == the oracle's float32 sweep (probabilities, losses, order)"""
    from oracle import pyoracle
    for a in range(1, 101):
        for d in range(1, 101):
            gp, ga, gd, gi = engine.get_attack_outcomes(a, d)
            op, oa, od, oi = pyoracle.get_attack_outcomes(a, d)
            assert len(gp) == len(op), (a, d)
            assert gp.tobytes() == op.tobytes(), (a, d)
            assert np.array_equal(ga, oa) and np.array_equal(gd, od) and np.array_equal(gi, oi), (a, d)


def test_battles_match_oracle(engine):
    from oracle import pyoracle
    cases = [(1, 1), (3, 2), (10, 10), (50, 80), (100, 100), (101, 5), (5, 101), (150, 150), (400, 90), (90, 400), (501, 501),
             (600, 1000), (5000, 3000), (700, 520), (101, 1), (1, 101), (102, 2), (2, 102), (103, 3), (163, 64), (164, 65),
             (165, 66), (64, 163), (228, 228), (229, 229), (100000, 3), (3, 100000), (20000, 499), (499, 20000)]
    for k, (a, d) in enumerate(cases):
        ga, gd = engine.simulate_battles(a, d, 1500, seed=7 + k, first_game=11)
        oa, od = pyoracle.simulate_battles(a, d, 1500, seed=7 + k, first_game=11)
        assert np.array_equal(ga, oa) and np.array_equal(gd, od), (a, d)
    rng = np.random.default_rng(9)
    for k in range(40):
        a, d = int(rng.integers(1, 700)), int(rng.integers(1, 700))
        ga, gd = engine.simulate_battles(a, d, 300, seed=1000 + k)
        oa, od = pyoracle.simulate_battles(a, d, 300, seed=1000 + k)
        assert np.array_equal(ga, oa) and np.array_equal(gd, od), (a, d)


def test_junit_individual_battle_port(engine):
    """individualBattleTest (BattleSimTests.java:77-118): 5,000,000 draws per case, tolerance 0.001"""
    for (pa, pd), consts in CONSTANTS.items():
        al = engine.simulate_individual(pa, pd, 5_000_000, seed=17)
        freq = np.bincount(al, minlength=3) / len(al)
        want = list(consts) + [0.0] * (3 - len(consts))
        assert np.all(np.abs(freq - np.array(want)) < 0.001), (pa, pd, freq)


def test_junit_simulate_battle_port(engine):
    """simulateBattleTest (BattleSimTests.java:15-53): 500,000 draws per (a, d), tolerance 0.01 (see the CPU port for
    the note on AttackOutcome.equals)"""
    from oracle import pyoracle
    for k, (a, d) in enumerate(junit_pairs(150, 10)):
        pr, al, dl, inv = pyoracle.get_attack_outcomes(a, d)
        n = 500_000
        xa, xd = engine.simulate_battles(a, d, n, seed=100 + k)
        for p, x, y in zip(pr, al, dl):
            obs = np.count_nonzero((xa == x) & (xd == y)) / n
            assert abs(obs - float(p)) < 0.01, (a, d, x, y, obs, p)


def test_card_cash_values():
    from oponn_b200.engine import RolloutEngine
    from oracle import pyoracle
    progs = ["0", "5, 5, 5...", "4, 4, 6, 6, 6, 8, 8, 8, 8, 10...", "4, 5, 6...", "4, 6, 8...", "3, 6, 9...",
             "4, 6, 8, 10, 15, 20...", "5, 10, 15...", "4, 6, 8, 10, 15, 20, 25, 10, 10, 10...", "5, 10, 15... ^10%",
             "4, 6, 8... ^5%", "unknown progression"]
    for prog in progs:
        with RolloutEngine(classic42(6), Rules(card_progression=prog), 0) as e:
            for pos in list(range(0, 40)) + [100, 333, 1023]:
                assert e.card_cash_value(pos) == pyoracle.card_cash_value(prog, pos), (prog, pos)


# ---- the ABI from plain C --------------------------------------------------------------------------------------
def test_c_program_gets_the_same_numbers(tmp_path):
    """examples/orx_demo.c (gcc, no Python in the process) against the ctypes binding on the same map and leaf"""
    import subprocess
    from test_abi_and_host import build_c_demo
    from oponn_b200.engine import RolloutEngine
    from oponn_b200.state import PHASE_CARDS, BoardState, RiskMap
    exe = build_c_demo(tmp_path)
    r = subprocess.run([exe, "-", "4096", "77"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    words = r.stdout.split()
    wins_c = [int(x) for x in words[words.index("wins") + 1:words.index("wins") + 7]]
    n = 42
    m = RiskMap(name="ring42", n_players=6, adjacency=[[(t - 1) % n, (t + 1) % n] for t in range(n)],
                continent=[t // 7 for t in range(n)], continent_bonus=[5, 2, 5, 3, 7, 2])
    bs = BoardState(owner=[t % 6 for t in range(n)], armies=[3] * n, currentPlayer=0, currentPhase=PHASE_CARDS, turnCount=1,
                    playerNumberOfCards=[0] * 6, hand=[], deck=list(range(n)) + [0xFF, 0xFF], cardProgressionPos=1,
                    attackUntilDead=True)
    with RolloutEngine(m, Rules(), 0) as e:
        agg, _ = e.rollout_batch(bs.pack(m)[None, :], 4096, 77, per_game=False)
    assert [int(x) for x in agg["wins"][0][:6]] == wins_c and int(agg["n"][0]) == 4096
    assert f"playouts 4096 flagged {int(agg['flagged'][0])}" in r.stdout


# ---- whole games ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("make", scenarios.ALL, ids=lambda f: f.__name__)
def test_hand_derived_scenarios(make):
    from oponn_b200.engine import RolloutEngine
    m, r, leaf, words, exp = make()
    with RolloutEngine(m, r, 0) as e:
        games = e.rollout_replay(leaf.pack(m)[None, :], words[None, :])
    scenarios.check(games[0], exp, m.n_players)


@pytest.mark.parametrize("make", scenarios.MODELLED, ids=lambda f: f.__name__)
def test_hand_derived_modelled_opponents(make):
    """SimAngry / SimStinky behind ModellingBoard, expectations worked out by hand."""
    from oponn_b200.engine import RolloutEngine
    m, r, leaf, words, exp, agents, limit = make()
    with RolloutEngine(m, r, 0, sim_agents=agents, modelling_turn_limit=limit) as e:
        games = e.rollout_replay(leaf.pack(m)[None, :], words[None, :])
    scenarios.check(games[0], exp, m.n_players)


def test_modelled_opponents_match_oracle(classic, rules, mid_leaf, golden_games):
    """Every per-game record with modelled line-ups equals the oracle's: mid-game leaf, the phase zoo (pre-game and
    mid-game), turn limit 2 (the reference's) and "never switch back"; then modelling off again == the plain goldens."""
    from leafzoo import phase_zoo
    from oponn_b200.engine import RolloutEngine
    from oracle.pyoracle import Oracle
    e, o = RolloutEngine(classic, rules, 0), Oracle(classic, rules)
    gold = golden_games["classic42_mid_s1_modelled"]
    e.set_sim_agents([0, 1, 2, 1, 2, 0], 2)
    _, games = e.rollout_batch(mid_leaf[None, :], len(gold), RUN_SEED, per_game=True)
    same(games, gold)
    for agents in ([1] * 6, [2] * 6, [0, 1, 2, 1, 2, 0], [2, 0, 0, 1, 0, 0]):
        for limit in (2, 1 << 20):
            e.set_sim_agents(agents, limit); o.set_sim_agents(agents, limit)
            _, g = e.rollout_batch(mid_leaf[None, :], 192, 77, per_game=True)
            _, w = o.rollout_batch(mid_leaf[None, :], 192, 77, per_game=True, threads=8)
            same(g, w)
            for base in (None, mid_leaf):
                leaves, _ = phase_zoo(classic, base)
                _, g = e.rollout_batch(leaves, 3, 13, per_game=True)
                _, w = o.rollout_batch(leaves, 3, 13, per_game=True, threads=8)
                same(g, w)
    e.set_sim_agents(None)
    _, games = e.rollout_batch(mid_leaf[None, :], 64, RUN_SEED, per_game=True)
    same(games, golden_games["classic42_mid_s1"][:64])
    e.close(); o.close()


def test_modelled_opponents_large_map(golden_games):
    """the 8-word board scans (territories > 64) of the modelled policies"""
    from leafzoo import phase_zoo
    m, r = synth200(), Rules(max_player_turns=4096)
    e, o = pair(m, r)
    leaf = np.fromfile(os.path.join(GOLDEN, "synth200_mid_s1.leaf"), dtype=np.uint8)
    agents = [1, 2, 0, 1, 2, 1, 2, 0]
    e.set_sim_agents(agents, 3); o.set_sim_agents(agents, 3)
    _, g = e.rollout_batch(leaf[None, :], 24, 5, per_game=True)
    _, w = o.rollout_batch(leaf[None, :], 24, 5, per_game=True, threads=8)
    same(g, w)
    leaves, _ = phase_zoo(m, None)
    _, g = e.rollout_batch(leaves, 2, 5, per_game=True)
    _, w = o.rollout_batch(leaves, 2, 5, per_game=True, threads=8)
    same(g, w)
    e.close(); o.close()


def test_goldens(engine, mid_leaf, golden_games):
    gold = golden_games["classic42_mid_s1"]
    agg, games = engine.rollout_batch(mid_leaf[None, :], len(gold), RUN_SEED, per_game=True)
    same(games, gold)
    leaves = np.fromfile(os.path.join(GOLDEN, "classic42_mid_64.leaves"), dtype=np.uint8).reshape(64, -1)
    _, games = engine.rollout_batch(leaves, 4, RUN_SEED, per_game=True)
    same(games, golden_games["classic42_mid_64"])


def test_golden_large_map(golden_games):
    from oponn_b200.engine import RolloutEngine
    leaf = np.fromfile(os.path.join(GOLDEN, "synth200_mid_s1.leaf"), dtype=np.uint8)
    with RolloutEngine(synth200(), Rules(max_player_turns=4096), 0) as e:
        _, games = e.rollout_batch(leaf[None, :], 64, RUN_SEED, per_game=True)
    same(games, golden_games["synth200_mid_s1"])


def test_classic_mid_game_vs_oracle(engine, oracle, mid_leaf):
    n = 2048
    agg, games = engine.rollout_batch(mid_leaf[None, :], n, 4242, first_game=10**12, per_game=True)
    oagg, ogames = oracle.rollout_batch(mid_leaf[None, :], n, 4242, first_game=10**12, per_game=True, threads=THREADS)
    same(games, ogames)
    assert agg.tobytes() == oagg.tobytes()


def test_many_leaves_vs_oracle(engine, oracle):
    leaves = np.fromfile(os.path.join(GOLDEN, "classic42_mid_64.leaves"), dtype=np.uint8).reshape(64, -1)
    agg, games = engine.rollout_batch(leaves, 8, 31337, per_game=True)
    oagg, ogames = oracle.rollout_batch(leaves, 8, 31337, per_game=True, threads=THREADS)
    same(games, ogames)
    assert agg.tobytes() == oagg.tobytes()


def test_every_phase_and_attack_state(engine, oracle, classic, mid_leaf):
    for base in (None, mid_leaf):
        leaves, names = phase_zoo(classic, base)
        _, games = engine.rollout_batch(leaves, 8, 11, per_game=True)
        _, ogames = oracle.rollout_batch(leaves, 8, 11, per_game=True, threads=THREADS)
        for i, name in enumerate(names):
            a, b = games[8 * i:8 * i + 8], ogames[8 * i:8 * i + 8]
            assert a.tobytes() == b.tobytes(), (name, a, b)


def test_invalid_leaves(engine, oracle, classic, mid_leaf):
    bad = invalid_leaves(classic, mid_leaf)
    agg, games = engine.rollout_batch(bad, 2, 5, per_game=True)
    _, ogames = oracle.rollout_batch(bad, 2, 5, per_game=True)
    same(games, ogames)
    assert np.all(games["flags"] == FLAG_ERROR) and np.all(agg["flagged"] == 2)


def test_replay_mode_vs_oracle(engine, oracle, classic, mid_leaf):
    rng = np.random.default_rng(77)
    leaves, names = phase_zoo(classic, mid_leaf)
    # short streams: every game runs out of words; long streams: most games finish
    for width in (40, 3000, 400_000):
        n = len(names) if width < 100_000 else 6
        words = rng.integers(0, 1 << 32, size=(n, width), dtype=np.uint64).astype(np.uint32)
        g = engine.rollout_replay(leaves[:n], words)
        o = oracle.rollout_replay(leaves[:n], words)
        same(g, o)
    assert np.any(g["flags"] == 0)


@pytest.mark.parametrize("rules", [
    Rules(use_cards=False), Rules(transfer_cards=False), Rules(continent_increase=5), Rules(continent_increase=-10),
    Rules(card_progression="5, 5, 5..."), Rules(card_progression="4, 4, 6, 6, 6, 8, 8, 8, 8, 10..."),
    Rules(card_progression="5, 10, 15... ^10%"), Rules(card_progression="0"), Rules(max_player_turns=40),
], ids=lambda r: f"cards{int(r.use_cards)}_tr{int(r.transfer_cards)}_inc{r.continent_increase}_{r.card_progression[:12]}_{r.max_player_turns}")
def test_rule_variants(classic, mid_leaf, rules):
    e, o = pair(classic, rules)
    try:
        n = 96
        agg, games = e.rollout_batch(mid_leaf[None, :], n, 606, per_game=True)
        oagg, ogames = o.rollout_batch(mid_leaf[None, :], n, 606, per_game=True, threads=THREADS)
        same(games, ogames)
        assert agg.tobytes() == oagg.tobytes()
    finally:
        e.close(); o.close()


@pytest.mark.parametrize("players", [2, 3, 8])
def test_other_player_counts(players):
    m = classic42(players)
    e, o = pair(m, Rules())
    try:
        leaves, names = phase_zoo(m, None, seed=players)
        _, games = e.rollout_batch(leaves, 4, 99, per_game=True)
        _, ogames = o.rollout_batch(leaves, 4, 99, per_game=True, threads=THREADS)
        same(games, ogames)
    finally:
        e.close(); o.close()


def test_large_map_zoo():
    m = synth200()
    e, o = pair(m, Rules(max_player_turns=600))
    try:
        leaves, names = phase_zoo(m, None, seed=8)
        _, games = e.rollout_batch(leaves, 2, 5, per_game=True)
        _, ogames = o.rollout_batch(leaves, 2, 5, per_game=True, threads=THREADS)
        same(games, ogames)
    finally:
        e.close(); o.close()


# ---- full size: properties that do not need the oracle ---------------------------------------------------------
def test_full_size_properties(engine, mid_leaf):
    """BASELINE.json configs[1]: 1,048,576 playouts of one leaf on one GPU"""
    n = 1 << 20
    agg, games = engine.rollout_batch(mid_leaf[None, :], n, RUN_SEED, per_game=True)
    pot = games["player_out_on_turn"][:, :6]
    assert np.all(games["flags"] == 0)
    assert np.all((pot == -1).sum(axis=1) == 1)
    assert np.all(pot[pot != -1] >= 0) and np.all(pot.max(axis=1) <= games["sim_turn_count"])
    assert int(agg["n"][0]) == n and int(agg["flagged"][0]) == 0
    assert int(agg["len_sum"][0]) == int(games["sim_turn_count"].astype(np.int64).sum())
    for p in range(6):
        won = pot[:, p] == -1
        assert int(agg["wins"][0][p]) == int(won.sum())
        assert int(agg["win_len_sum"][0][p]) == int(games["sim_turn_count"][won].astype(np.int64).sum())
    # the committed goldens are the first 512 games of this very batch
    gold = np.load(os.path.join(GOLDEN, "golden_games.npz"))["classic42_mid_s1"]
    same(games[:512], gold)
    # sharding invariance: two half batches (as two GPUs would play them) are the whole batch
    h = 1 << 16
    _, a = engine.rollout_batch(mid_leaf[None, :], h, RUN_SEED, first_game=0, per_game=True)
    _, b = engine.rollout_batch(mid_leaf[None, :], h, RUN_SEED, first_game=h, per_game=True)
    assert a.tobytes() + b.tobytes() == games[:2 * h].tobytes()


def test_device_buffers_equal_host_buffers(engine, mid_leaf):
    import torch
    n = 4096
    agg, games = engine.rollout_batch(mid_leaf[None, :], n, 3, per_game=True)
    ld = torch.from_numpy(mid_leaf.copy()).cuda()
    ad = torch.zeros(LEAF_RESULT_DTYPE.itemsize // 8, dtype=torch.int64, device="cuda")
    gd = torch.zeros(n * GAME_RESULT_DTYPE.itemsize // 4, dtype=torch.int32, device="cuda")
    engine.rollout_batch_device(ld.data_ptr(), 1, n, 3, 0, ad.data_ptr(), gd.data_ptr())
    engine.sync()
    assert ad.cpu().numpy().tobytes() == agg.tobytes()
    assert gd.cpu().numpy().tobytes() == games.tobytes()
    assert engine.last_kernel_ms > 0


def test_reference_facing_board_api(engine, classic, mid_leaf):
    from oponn_b200.engine import BattleSim, ModellingBoard
    from oponn_b200.state import BoardState
    board = ModellingBoard(engine, seed=RUN_SEED)
    board.setupState(BoardState.unpack(classic, mid_leaf))
    gold = np.load(os.path.join(GOLDEN, "golden_games.npz"))["classic42_mid_s1"]
    for g in range(3):
        pot = board.simulate()
        assert pot == [int(x) for x in gold[g]["player_out_on_turn"][:6]]
        assert board.getSimTurnCount() == int(gold[g]["sim_turn_count"])
    bs = BattleSim(engine, seed=1)
    assert bs.simulateIndividualBattle(3, 2) in (0, 1, 2)
    with pytest.raises(ValueError):
        bs.simulateIndividualBattle(4, 2)
    rows = bs.getAttackOutcomes(3, 2)
    assert abs(sum(r["probability"] for r in rows) - 1.0) < 1e-3


def test_two_slots_in_flight_equal_the_blocking_call(engine, mid_leaf):
    """orx_rollout_begin / _end on both slots at once == two orx_rollout_batch calls"""
    leaves = np.fromfile(os.path.join(GOLDEN, "classic42_mid_64.leaves"), dtype=np.uint8).reshape(64, -1)
    t0 = engine.rollout_begin(0, leaves[:40], 6, 9, first_game=1000, per_game=True)
    t1 = engine.rollout_begin(1, mid_leaf[None, :], 300, 9, first_game=5000, per_game=True)
    a1, g1 = engine.rollout_end(1, t1)
    a0, g0 = engine.rollout_end(0, t0)
    b0, h0 = engine.rollout_batch(leaves[:40], 6, 9, first_game=1000, per_game=True)
    b1, h1 = engine.rollout_batch(mid_leaf[None, :], 300, 9, first_game=5000, per_game=True)
    assert a0.tobytes() == b0.tobytes() and g0.tobytes() == h0.tobytes()
    assert a1.tobytes() == b1.tobytes() and g1.tobytes() == h1.tobytes()
    from oponn_b200.engine import OrxError
    t0 = engine.rollout_begin(0, mid_leaf[None, :], 2, 1)
    with pytest.raises(OrxError):
        engine.rollout_begin(0, mid_leaf[None, :], 2, 1)        # slot still in flight
    engine.rollout_end(0, t0)


def test_65536_games_vs_oracle(engine, oracle, mid_leaf):
    """BASELINE config 2: every per-game record of the first 65,536 games of the bench batch equals the oracle's
    (about 30 s of oracle time on 16 host threads)"""
    n = 1 << 16
    agg, games = engine.rollout_batch(mid_leaf[None, :], n, RUN_SEED, per_game=True)
    oagg, ogames = oracle.rollout_batch(mid_leaf[None, :], n, RUN_SEED, per_game=True, threads=THREADS)
    same(games, ogames)
    assert agg.tobytes() == oagg.tobytes()


def test_random_maps_and_leaves_fuzz():
    """random graphs (both kernel widths), 2..8 players, random rules and random legal records: every game record equal"""
    from leafzoo import random_leaves, random_map
    rng = np.random.default_rng(20261020)
    progs = ["4, 6, 8, 10, 15, 20...", "5, 5, 5...", "4, 5, 6...", "3, 6, 9...", "4, 6, 8, 10, 15, 20, 25, 10, 10, 10..."]
    for k in range(12):
        n = int(rng.choice([5, 12, 31, 32, 33, 42, 64, 65, 90, 150]))
        p = int(rng.integers(2, 9))
        m = random_map(rng, n, p, int(rng.integers(1, 9)))
        r = Rules(use_cards=bool(rng.random() < 0.8), transfer_cards=bool(rng.integers(2)), continent_increase=int(rng.choice([0, 0, 3, 10])),
                  card_progression=str(rng.choice(progs)), max_player_turns=int(rng.choice([300, 2000])))
        leaves = random_leaves(m, rng, 160 if n <= 64 else 60)
        e, o = pair(m, r)
        try:
            agg, games = e.rollout_batch(leaves, 2, 1234 + k, per_game=True)
            oagg, ogames = o.rollout_batch(leaves, 2, 1234 + k, per_game=True, threads=THREADS)
            same(games, ogames)
            assert agg.tobytes() == oagg.tobytes()
        finally:
            e.close(); o.close()


def test_random_maps_and_leaves_fuzz_modelled_opponents():
    """the same fuzz with random SimRandom / SimAngry / SimStinky line-ups and modelling turn limits (f4)"""
    from leafzoo import random_leaves, random_map
    rng = np.random.default_rng(20261021)
    for k in range(10):
        n = int(rng.choice([5, 12, 32, 33, 42, 64, 65, 120]))
        p = int(rng.integers(2, 9))
        m = random_map(rng, n, p, int(rng.integers(1, 9)))
        r = Rules(use_cards=bool(rng.random() < 0.8), transfer_cards=bool(rng.integers(2)), continent_increase=int(rng.choice([0, 0, 5])),
                  max_player_turns=int(rng.choice([200, 1000])))
        leaves = random_leaves(m, rng, 120 if n <= 64 else 40)
        agents = [int(x) for x in rng.integers(0, 3, size=p)]
        limit = int(rng.choice([1, 2, 3, 1 << 20]))
        e, o = pair(m, r)
        try:
            e.set_sim_agents(agents, limit); o.set_sim_agents(agents, limit)
            agg, games = e.rollout_batch(leaves, 2, 4321 + k, per_game=True)
            oagg, ogames = o.rollout_batch(leaves, 2, 4321 + k, per_game=True, threads=THREADS)
            same(games, ogames)
            assert agg.tobytes() == oagg.tobytes()
        finally:
            e.close(); o.close()


def test_empty_batches_and_the_largest_map():
    """edge sizes: empty batches return without work; the largest map the ABI accepts (252 territories, 8 players,
    32 continents, one hub with the maximum 32 neighbours, directed and duplicate-free) against the oracle, with and
    without modelled opponents; armies at the record's upper bound"""
    from leafzoo import random_leaves
    from oponn_b200.state import MAX_CONTINENTS, MAX_DEGREE, MAX_PLAYERS, MAX_TERRITORIES, RiskMap
    n = MAX_TERRITORIES
    rng = np.random.default_rng(99)
    adjacency = [[(t - 1) % n, (t + 1) % n] for t in range(n)]               # a ring ...
    hub = 17
    spokes = [int(x) for x in rng.choice([t for t in range(n) if abs(t - hub) > 1], size=MAX_DEGREE - 2, replace=False)]
    adjacency[hub] += spokes                                                  # ... with one hub of degree 32
    for t in spokes:
        adjacency[t].append(hub)
    m = RiskMap(name="max252", n_players=MAX_PLAYERS, adjacency=adjacency, continent=[t % MAX_CONTINENTS for t in range(n)],
                continent_bonus=[(c % 5) + 1 for c in range(MAX_CONTINENTS)])
    r = Rules(max_player_turns=600)
    e, o = pair(m, r)
    try:
        leaves = random_leaves(m, rng, 24)
        agg, games = e.rollout_batch(leaves[:0], 4, 1, per_game=True)
        assert len(agg) == 0 and len(games) == 0
        agg, games = e.rollout_batch(leaves, 0, 1, per_game=True)
        assert len(games) == 0 and int(agg["n"].sum()) == 0
        agg, games = e.rollout_batch(leaves, 2, 31, per_game=True)
        oagg, ogames = o.rollout_batch(leaves, 2, 31, per_game=True, threads=THREADS)
        same(games, ogames)
        big = leaves.copy()
        L = e.layout
        arm = big[:, L.off_armies:L.off_armies + 4 * n].view("<i4")
        arm[:, ::7] = 1 << 20                                                 # ORX_MAX_ARMIES on every seventh territory
        agg, games = e.rollout_batch(big, 1, 32, per_game=True)
        oagg, ogames = o.rollout_batch(big, 1, 32, per_game=True, threads=THREADS)
        same(games, ogames)
        agents = [1, 2, 0, 1, 2, 0, 1, 2]
        e.set_sim_agents(agents, 2); o.set_sim_agents(agents, 2)
        agg, games = e.rollout_batch(leaves, 2, 33, per_game=True)
        oagg, ogames = o.rollout_batch(leaves, 2, 33, per_game=True, threads=THREADS)
        same(games, ogames)
    finally:
        e.close(); o.close()


def test_win_rates_under_normal_seeding(engine, oracle, mid_leaf):
    """North-star correctness part 2: with ordinary (different) seeds the engine's root win-rate estimates agree with
    the CPU simulator's within a stated confidence interval: per player, |p_gpu - p_cpu| < 4.5 sigma of the difference
    of two binomial estimates (a 1-in-10^5 event per player if the two samplers are the same distribution)."""
    n_gpu, n_cpu = 1 << 17, 1 << 13
    agg, _ = engine.rollout_batch(mid_leaf[None, :], n_gpu, 555)
    oagg, _ = oracle.rollout_batch(mid_leaf[None, :], n_cpu, 99991, per_game=False, threads=THREADS)
    for p in range(6):
        a, b = agg["wins"][0][p] / n_gpu, oagg["wins"][0][p] / n_cpu
        sigma = np.sqrt(a * (1 - a) / n_gpu + b * (1 - b) / n_cpu)
        assert abs(a - b) < 4.5 * sigma, (p, a, b, sigma)
    la, lb = agg["len_sum"][0] / n_gpu, oagg["len_sum"][0] / n_cpu      # mean game length in rounds (std ~ 330)
    assert abs(la - lb) < 4.5 * 330 * np.sqrt(1 / n_gpu + 1 / n_cpu), (la, lb)
