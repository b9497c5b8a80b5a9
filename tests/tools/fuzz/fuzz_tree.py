"""One-off differential fuzz of the native leaf producer (orx_tree_possible_moves / orx_tree_apply, host C++ in liborx.so) against the
pure-Python restatement oracle/pyexplore.py: random maps, rules, pruning switches and start positions, random walks; every move list
(order, fields, float32 probabilities) and every successor state must be identical.  CPU only.

    python tests/tools/fuzz/fuzz_tree.py [rng seed] [seconds]

Not collected by pytest (tests/test_tree_vs_pyexplore.py holds the fixed cases); DESIGN.md section 6b records the last run.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..", ".."))); sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import test_tree_vs_pyexplore as T  # noqa: E402
from leafzoo import random_leaves, random_map  # noqa: E402
from oponn_b200.maps import classic42  # noqa: E402
from oponn_b200.state import BoardState, Rules  # noqa: E402
from oponn_b200.tree import TreeOptions, new_game  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 600
rng = np.random.default_rng(seed)
t0 = time.time(); walks = steps = maps = 0
while time.time() - t0 < budget:
    n, p = int(rng.choice([4, 6, 9, 13, 20, 30, 42, 70])), int(rng.integers(2, 9))
    m = classic42(p) if (n == 42 and p <= 6 and rng.random() < 0.5) else random_map(rng, n, p, int(rng.integers(1, 7)))
    rules = Rules(use_cards=bool(rng.random() < 0.8), transfer_cards=bool(rng.integers(2)), continent_increase=int(rng.choice([0, 0, 5])),
                  card_progression=str(rng.choice(["4, 6, 8, 10, 15, 20...", "5, 5, 5...", "3, 6, 9...", "4, 4, 6, 6, 6, 8, 8, 8, 8, 10..."])))
    b = lambda: bool(rng.integers(2))
    opts = TreeOptions(placement_partial_order=b(), attack_aggressive=b(), attack_repeat_moves=b(), attack_partial_order=b(),
                       attack_until_dead=b(), invasion=b(), fortification_outwards=b())
    maps += 1
    leaves = list(random_leaves(m, rng, 4))
    if rng.random() < 0.3:
        leaves.append(new_game(m, int(rng.integers(1 << 30))))
    for leaf in leaves:
        bs = BoardState.unpack(m, leaf); bs.armies = [min(a, 40) for a in bs.armies]   # keep the outcome tables cheap
        steps += T.walk(m, rules, opts, bs.pack(m), rng, 80); walks += 1
print("maps", maps, "walks", walks, "steps", steps, "all identical; seconds", round(time.time() - t0))
