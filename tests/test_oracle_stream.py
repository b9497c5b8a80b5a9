"""The oracle's random plumbing against published known answers (CPU only)."""
import numpy as np

from javarandom import JavaRandom, lcg_words
from oracle import pyoracle


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert pyoracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert pyoracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert pyoracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_counter_stream_layout():
    """word i = Philox(ctr = {i/4, 0, game_lo, game_hi}, key = seed)[i % 4] (orx_state.h)"""
    seed, game = 0x1122334455667788, 0x99AABBCCDDEEFF00
    vals, pos, flags = pyoracle.stream_draws([-2] * 300, seed=seed, game=game)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    for i in (0, 1, 31, 32, 95, 127, 128, 200, 299):
        ctr = [i // 4, 0, game & 0xFFFFFFFF, game >> 32]
        assert int(vals[i]) == pyoracle.philox4x32_10(ctr, key)[i % 4]
    assert pos == 300 and flags == 0


def test_java_known_answers_through_the_stream():
    """Feeding the stream the LCG's next(32) outputs must reproduce java.util.Random exactly."""
    w = lcg_words(0, 64)
    v, _, _ = pyoracle.stream_draws([0], words=w)
    assert v[0] == 0.730967787376657          # new Random(0).nextDouble()
    v, _, _ = pyoracle.stream_draws([-1, -1, -1, -1], words=w)
    # new Random(0).nextGaussian() x4 (the 2nd and 4th are the cached partners of the polar method)
    assert list(v) == [0.8025330637390305, -0.9015460884175122, 2.080920790428163, 0.7637707684364894]
    v, _, _ = pyoracle.stream_draws([-2], words=w)
    assert int(v[0]) == (-1155484576) & 0xFFFFFFFF   # new Random(0).nextInt()


def test_derivations_match_lcg_model():
    for seed in (0, 1, 42, 20261019):
        w = lcg_words(seed, 4096)
        jr = JavaRandom(seed)
        ops, want = [], []
        rng = np.random.default_rng(seed)
        for _ in range(600):
            if rng.random() < 0.3:
                ops.append(0); want.append(jr.nextDouble())
            else:
                b = int(rng.choice([1, 2, 3, 5, 6, 7, 10, 42, 44, 64, 100, 1000, 65536, (1 << 30) + 1, (1 << 31) - 1]))
                ops.append(b); want.append(float(jr.nextInt(b)))
        got, pos, flags = pyoracle.stream_draws(ops, words=w)
        assert flags == 0
        assert np.array_equal(got, np.array(want))


def test_stream_end_and_bad_bound():
    _, pos, flags = pyoracle.stream_draws([-2, -2, -2], words=np.array([1, 2], dtype=np.uint32))
    assert flags & 4
    _, _, flags = pyoracle.stream_draws([0], words=np.zeros(1, dtype=np.uint32))
    assert flags & 4


def test_fdlibm_log_matches_libm_on_exact_cases():
    import math
    assert pyoracle.fdlibm_log(1.0) == 0.0
    for x in (0.5, 0.25, 0.1, 0.3, 0.7, 0.999, 1e-3, 0.6180339887):
        assert abs(pyoracle.fdlibm_log(x) - math.log(x)) <= 2.3e-16 * abs(math.log(x)) + 1e-300
