"""Independent statements of BattleSim facts used to pin the oracle (and, on the GPU, the engine)."""
import itertools
from fractions import Fraction

import numpy as np

from javarandom import JavaRandom

# O/BattleSim.java:15-33, as written in the reference source
CONSTANTS = {
    (1, 1): (0.4167, 0.5833), (1, 2): (0.2546, 0.7454), (2, 1): (0.5787, 0.4213),
    (2, 2): (0.2276, 0.3241, 0.4483), (3, 1): (0.6597, 0.3403), (3, 2): (0.3717, 0.3358, 0.2926),
}


def exact_single_roll(pa: int, pd: int):
    """Exact Risk dice odds by enumerating all 6^(pa+pd) rolls: P(attacker loses k), k = 0..min(pa,pd)."""
    n = min(pa, pd)
    counts = [0] * (n + 1)
    for roll in itertools.product(range(1, 7), repeat=pa + pd):
        a = sorted(roll[:pa], reverse=True); d = sorted(roll[pa:], reverse=True)
        loss = sum(1 for i in range(n) if a[i] <= d[i])
        counts[loss] += 1
    tot = 6 ** (pa + pd)
    return [Fraction(c, tot) for c in counts]


def numpy_outcome_table(a: int, d: int):
    """generateProbabilities + generateAttackOutcomes (O/BattleSim.java:190-286) restated a second time, in
    numpy float32 scalars with the reference's scatter order; returns rows (p, al, dl, inv) sorted like Java."""
    f = np.float32
    C = {k: [f(x) for x in v] for k, v in CONSTANTS.items()}
    rl = d + 1
    last = np.zeros((a + 1) * rl, dtype=np.float32)
    last[a * rl + d] = f(1.0)
    absorbed = False
    while not absorbed:
        new = np.zeros_like(last)
        absorbed = True
        for ra in range(a + 1):
            for rd in range(d + 1):
                v = last[ra * rl + rd]
                if v == 0:
                    continue
                if ra == 0 or rd == 0:
                    new[ra * rl + rd] = f(new[ra * rl + rd] + v)
                    continue
                pa, pd = min(ra, 3), min(rd, 2)
                n = min(pa, pd)
                for x, c in enumerate(C[(pa, pd)]):       # attacker loses x, defender n - x
                    t = (ra - x) * rl + (rd - (n - x))
                    new[t] = f(new[t] + f(v * c))
                absorbed = False
        last = new
    rows = []
    for ra in range(a + 1):
        for rd in range(d + 1):
            p = last[ra * rl + rd]
            if p != 0:
                rows.append((p, a - ra, d - rd, rd == 0))
    # Collections.sort(reverseOrder()): stable, descending probability
    rows.sort(key=lambda r: -float(r[0]))
    return rows


def junit_pairs(bound: int, cases: int):
    """The (a, d) pairs BattleSimTests draws: new Random(0), nextInt(bound) twice per case."""
    r = JavaRandom(0)
    return [(r.nextInt(bound), r.nextInt(bound)) for _ in range(cases)]
