"""First-light GPU checks (through the C ABI) against the oracle."""
import numpy as np
import pytest

from conftest import RUN_SEED

pytestmark = pytest.mark.gpu


def test_stream_counter(engine):
    from oracle import pyoracle
    ops = [-2] * 300 + [0] * 10 + [7, 10, 1, 2, 3, 1 << 20, (1 << 30) + 1, 100] * 8 + [-1] * 5 + [0, 5, -1, -1]
    g, gp, gf = engine.stream_draws(ops, seed=123456789123, game=987654321)
    o, op, of = pyoracle.stream_draws(ops, seed=123456789123, game=987654321)
    assert (gp, gf) == (op, of)
    assert np.array_equal(g.view(np.uint64), o.view(np.uint64))


def test_tables_match_oracle(engine):
    from oracle import pyoracle
    for (a, d) in [(1, 1), (1, 2), (2, 1), (3, 2), (5, 7), (10, 10), (49, 50), (100, 100), (100, 1), (1, 100), (37, 93)]:
        gp, ga, gd, gi = engine.get_attack_outcomes(a, d)
        op, oa, od, oi = pyoracle.get_attack_outcomes(a, d)
        assert len(gp) == len(op), (a, d)
        assert np.array_equal(gp.view(np.uint32), op.view(np.uint32)), (a, d)
        assert np.array_equal(ga, oa) and np.array_equal(gd, od) and np.array_equal(gi, oi), (a, d)


def test_battles_match_oracle(engine):
    from oracle import pyoracle
    for (a, d) in [(1, 1), (3, 2), (10, 10), (50, 80), (100, 100), (101, 5), (5, 101), (150, 150), (400, 90), (501, 501),
                   (600, 1000), (5000, 3000), (700, 520)]:
        ga, gd = engine.simulate_battles(a, d, 2000, seed=7, first_game=11)
        oa, od = pyoracle.simulate_battles(a, d, 2000, seed=7, first_game=11)
        assert np.array_equal(ga, oa) and np.array_equal(gd, od), (a, d)


def test_rollout_matches_golden(engine, mid_leaf, golden_games):
    gold = golden_games["classic42_mid_s1"]
    agg, games = engine.rollout_batch(mid_leaf[None, :], len(gold), RUN_SEED, per_game=True)
    bad = [i for i in range(len(gold)) if games[i].tobytes() != gold[i].tobytes()]
    assert not bad, (bad[:5], games[bad[0]], gold[bad[0]])
    assert int(agg["n"][0]) == len(gold)
