"""Pins the oracle's BattleSim against the reference's known answers (O/BattleSim.java:15-44) and ports of its
three JUnit tests (O/tests/BattleSimTests.java:15-118).  CPU only."""
import numpy as np
import pytest

from battlesim_ref import CONSTANTS, exact_single_roll, junit_pairs, numpy_outcome_table
from oracle import pyoracle


def test_constants_are_exact_odds_to_4dp():
    """The 14 hard-coded floats equal the exact dice probabilities rounded to four decimals."""
    for (pa, pd), consts in CONSTANTS.items():
        exact = exact_single_roll(pa, pd)
        assert len(exact) == len(consts)
        for e, c in zip(exact, consts):
            assert round(float(e), 4) == c


def test_individual_battle_is_a_threshold_on_k():
    """simulateIndividualBattle (O/BattleSim.java:152-162): r <= c0 -> 0, r <= c1 -> 1, else 2, with the CDF rows of
    :36-44 (float sums widened to double).  Checked on replayed words at the exact boundaries."""
    f = np.float32
    for (pa, pd), consts in CONSTANTS.items():
        c0 = float(f(consts[0]))
        c1 = float(f(f(consts[0]) + f(consts[1]))) if len(consts) == 3 else 1.0
        for thr, below, above in ((c0, 0, 1), (c1, 1, 2)):
            if thr >= 1.0:
                continue
            k = int(thr * (1 << 53))  # largest k with k * 2^-53 <= thr
            for kk, want in ((k, below), (k + 1, above)):
                words = np.array([(kk >> 27) << 6, (kk & ((1 << 27) - 1)) << 5], dtype=np.uint32)
                # a till-dead battle with > 100 armies on one side starts with exactly this single roll
                got, _, _ = pyoracle.stream_draws([0], words=words)
                assert got[0] == kk / float(1 << 53)
        al = pyoracle.simulate_individual(pa, pd, 200000, seed=3)
        freq = np.bincount(al, minlength=3) / len(al)
        want = list(consts) + [0.0] * (3 - len(consts))
        assert np.allclose(freq, want, atol=0.004)   # individualBattleTest uses 0.001 at 5M draws (GPU test does that)


@pytest.mark.parametrize("a,d", [(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (3, 2), (3, 3), (4, 2), (5, 5), (7, 4), (2, 9), (10, 10), (12, 6)])
def test_tables_match_independent_float32_sweep(a, d):
    rows = numpy_outcome_table(a, d)
    pr, al, dl, inv = pyoracle.get_attack_outcomes(a, d)
    assert len(rows) == len(pr)
    for i, (p, x, y, v) in enumerate(rows):
        assert np.float32(p).tobytes() == pr[i].tobytes()
        assert (x, y, bool(v)) == (int(al[i]), int(dl[i]), bool(inv[i]))


def test_table_shape_and_mass():
    for a, d in [(1, 1), (5, 3), (20, 30), (49, 49), (100, 100)]:
        pr, al, dl, inv = pyoracle.get_attack_outcomes(a, d)
        assert len(pr) <= a + d
        assert np.all(np.diff(pr) <= 0)                      # descending probability
        assert abs(float(pr.astype(np.float64).sum()) - 1.0) < 1e-2   # the 3v2 row sums to 1.0001: mass drifts up (1.0043 at 100v100), not renormalised
        # every row is a finished battle: one side wiped out
        assert np.all((al == a) ^ (dl == d))
        assert np.array_equal(inv.astype(bool), dl == d)


def test_outcome_duplicates_port():
    """outcomeDuplicatesTest (BattleSimTests.java:55-75): 100 (a, d) pairs from new Random(0).nextInt(200)."""
    for a, d in junit_pairs(200, 100):
        pr, al, dl, inv = pyoracle.get_attack_outcomes(a, d)
        keys = set(zip(al.tolist(), dl.tolist(), inv.tolist()))
        assert len(keys) == len(pr), (a, d)


def test_simulate_battle_port():
    """simulateBattleTest (BattleSimTests.java:15-53): observed frequencies vs table probabilities within 0.01, for
    the 10 (a, d) pairs new Random(0).nextInt(150) yields.  The reference compares with AttackOutcome.equals, which
    also compares `invading`, although simulateBattle always returns invading=false (BattleSim.java:144): taken
    literally no conquest could ever match.  This port compares (attackerLosses, defenderLosses), the evident intent.
    100k draws per pair here (CPU); the GPU port uses the reference's 500k."""
    for k, (a, d) in enumerate(junit_pairs(150, 10)):
        pr, al, dl, inv = pyoracle.get_attack_outcomes(a, d)
        n = 100000
        xa, xd = pyoracle.simulate_battles(a, d, n, seed=100 + k)
        for p, x, y in zip(pr, al, dl):
            obs = np.count_nonzero((xa == x) & (xd == y)) / n
            assert abs(obs - float(p)) < 0.01, (a, d, x, y, obs, p)


def test_large_battle_paths():
    """> 100 armies: single rolls until both sides are <= 100 (BattleSim.java:99-116); > 500 both: gaussian (:71-93)."""
    xa, xd = pyoracle.simulate_battles(300, 150, 2000, seed=5)
    assert np.all((xa == 300) ^ (xd == 150))                  # fought to the death
    xa, xd = pyoracle.simulate_battles(2000, 1000, 2000, seed=6)
    assert np.all(xd == 1000) and np.all(xa <= 1999)          # matchedAttackers = 922 <= 2000: defenders die
    assert abs(xa.mean() - 922) < 5 and 20 < xa.std() < 45    # variance ~ sqrt(1000)
    xa, xd = pyoracle.simulate_battles(600, 1000, 2000, seed=7)
    assert np.all(xa == 600) and np.all(xd <= 999)            # matchedAttackers = 922 > 600: attackers die
    assert abs(xd.mean() - 600 / 0.922) < 5


def test_card_values_and_starting_armies():
    prog = "4, 6, 8, 10, 15, 20..."
    assert [pyoracle.card_cash_value(prog, p) for p in range(1, 8)] == [4, 6, 8, 10, 15, 20, 25]
    assert [pyoracle.card_cash_value("4, 4, 6, 6, 6, 8, 8, 8, 8, 10...", p) for p in range(1, 11)] == [4, 4, 6, 6, 6, 8, 8, 8, 8, 10]
    assert [pyoracle.card_cash_value("4, 6, 8, 10, 15, 20, 25, 10, 10, 10...", p) for p in range(1, 11)] == [4, 6, 8, 10, 15, 20, 25, 10, 10, 10]
    assert [pyoracle.card_cash_value("5, 5, 5...", p) for p in range(1, 4)] == [5, 5, 5]
    assert pyoracle.card_cash_value("no such progression", 3) == 2          # the catch block (CardManager.java:333-339)
    assert pyoracle.card_cash_value("5, 10, 15... ^10%", 3) == 15           # base <= 25 stays linear
    assert pyoracle.card_cash_value("5, 10, 15... ^10%", 8) == int(40 * 1.1 ** 2)
    # calculateStartingArmies (SimulationBoard.java:518-533), P=6, N=42
    assert pyoracle.starting_armies(42, 6) == [20, 20, 22, 23, 24, 25]


def test_table_key_and_dense_row_offsets():
    """BattleSim keys its HashMap of tables by szudzikMapping(a, d) (O/BattleSim.java:288-291).  The mapping is Szudzik's
    pairing, injective on the naturals, so a table is identified by (a, d) alone -- which is what lets the engine
    replace the map by a dense array.  The engine lays table (a, d) out with capacity a + d at the closed-form row
    offset used by table_battle (csrc/orx_device.cuh); check the closed form against the running sum the table
    builder uses (csrc/orx_api.cu), that blocks do not overlap, and that a + d bounds the row count."""
    szudzik = lambda a, b: a * a + a + b if a >= b else a + b * b
    keys = {szudzik(a, d) for a in range(0, 201) for d in range(0, 201)}
    assert len(keys) == 201 * 201
    off, seen_end = 0, 0
    for a in range(1, 101):
        for d in range(1, 101):
            closed = 50 * a * (a - 1) + 5050 * (a - 1) + a * (d - 1) + (d * (d - 1)) // 2
            assert closed == off and closed >= seen_end
            seen_end = closed + a + d
            off += a + d
    assert off == 2 * 100 * (100 * 101 // 2)            # kTableRows
    for a, d in ((1, 1), (2, 1), (3, 2), (7, 5), (20, 20), (3, 40)):
        prob, al, dl, inv = pyoracle.get_attack_outcomes(a, d)
        assert len(prob) <= a + d                        # outcomes: (0, 1..d) and (1..a, 0)
