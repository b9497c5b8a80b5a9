"""End-to-end guard of the PTX uniformity pass (oponn_b200/ptx_uniform.py).

The shipped library promises ``bra.uni`` on the branches a home-made data-flow proves warp-uniform; an unsound
promise would be undefined behaviour that a parity suite might only catch on other inputs.  So the same sources are
also built WITHOUT the pass (``liborx_nouniform.so``, built by ``__graft_entry__.build()``), and every per-game record
of the phase zoo, the fuzz maps, the goldens and the modelled line-ups must be identical in the two libraries."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import sys, os, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from leafzoo import phase_zoo, random_leaves, random_map
from oponn_b200.engine import RolloutEngine
from oponn_b200.maps import classic42, synth200
from oponn_b200.state import Rules
out = {}
m = classic42(6)
leaf = np.fromfile(os.path.join(sys.argv[1], "tests", "golden", "classic42_mid_s1.leaf"), dtype=np.uint8)
with RolloutEngine(m, Rules(), 0) as e:
    out["mid"] = e.rollout_batch(leaf[None, :], 2048, 20261019, per_game=True)[1]
    for k, base in enumerate((None, leaf)):
        out[f"zoo{k}"] = e.rollout_batch(phase_zoo(m, base)[0], 8, 11, per_game=True)[1]
    e.set_sim_agents([0, 1, 2, 1, 2, 0], 2)
    out["modelled"] = e.rollout_batch(phase_zoo(m, leaf)[0], 4, 13, per_game=True)[1]
rng = np.random.default_rng(77)
for k in range(6):
    n = int(rng.choice([5, 31, 42, 64, 65, 150])); p = int(rng.integers(2, 9))
    mm = random_map(rng, n, p, int(rng.integers(1, 9)))
    with RolloutEngine(mm, Rules(max_player_turns=600), 0) as e:
        out[f"fuzz{k}"] = e.rollout_batch(random_leaves(mm, rng, 100 if n <= 64 else 40), 2, 99 + k, per_game=True)[1]
leaf2 = np.fromfile(os.path.join(sys.argv[1], "tests", "golden", "synth200_mid_s1.leaf"), dtype=np.uint8)
with RolloutEngine(synth200(), Rules(max_player_turns=1024), 0) as e:
    out["large"] = e.rollout_batch(leaf2[None, :], 64, 5, per_game=True)[1]
np.savez(sys.argv[2], **out)
'''


def run_with(lib: str, path: str):
    env = dict(os.environ, ORX_LIB=lib, ORX_KERNEL="warp")
    subprocess.check_call([sys.executable, "-c", WORKER, ROOT, path], env=env)
    return dict(np.load(path))


def test_records_are_identical_with_and_without_the_ptx_pass(tmp_path):
    from oponn_b200 import build as b
    if not os.path.exists(b.LIB_NOUNIFORM):
        pytest.skip("liborx_nouniform.so has not been built (__graft_entry__.build())")
    a = run_with(b.LIB, str(tmp_path / "uniform.npz"))
    c = run_with(b.LIB_NOUNIFORM, str(tmp_path / "plain.npz"))
    assert sorted(a) == sorted(c)
    for k in a:
        assert a[k].tobytes() == c[k].tobytes(), f"records of '{k}' differ between the bra.uni build and the plain build"
