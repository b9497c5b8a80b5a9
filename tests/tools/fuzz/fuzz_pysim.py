"""One-off differential fuzz: the C oracle against the independent pure-Python restatement (oracle/pysim.py) on random maps, rules, leaves
and -- for half of the maps -- random SimRandom / SimAngry / SimStinky line-ups behind ModellingBoard.  Every per-game record must be identical.

    python tests/tools/fuzz/fuzz_pysim.py [rng seed] [seconds]

Not collected by pytest (tests/test_pysim.py holds the fixed cases); DESIGN.md section 2 records the last run.
"""
import sys, os, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..", ".."))); sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import numpy as np
import test_pysim as T
from leafzoo import random_leaves, random_map, phase_zoo
from oponn_b200.state import Rules
from oponn_b200.maps import classic42
t0 = time.time(); n = 0; nm = 0
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 600
k = 0
while time.time() - t0 < budget:
    nt = int(rng.choice([4, 5, 7, 12, 20, 33, 42, 64])); p = int(rng.integers(2, 9))
    m = random_map(rng, nt, p, int(rng.integers(1, 9)))
    r = Rules(use_cards=bool(rng.random() < 0.8), transfer_cards=bool(rng.integers(2)), continent_increase=int(rng.choice([0, 0, 3, 10])),
              card_progression=str(rng.choice(["4, 6, 8, 10, 15, 20...", "5, 5, 5...", "4, 4, 6, 6, 6, 8, 8, 8, 8, 10...", "5, 10, 15... ^10%", "3, 6, 9..."])),
              max_player_turns=int(rng.choice([6, 12, 30])))
    leaves = random_leaves(m, rng, 16)
    if rng.random() < 0.5:
        n += T.compare(m, r, leaves, 2, 9000 + k)
    else:
        agents = [int(x) for x in rng.integers(0, 3, size=p)]
        nm += T.compare(m, r, leaves, 2, 9000 + k, agents, int(rng.choice([1, 2, 3, 1 << 20])))
    k += 1
print("maps", k, "SimRandom games", n, "modelled games", nm, "all identical; seconds", round(time.time() - t0))
