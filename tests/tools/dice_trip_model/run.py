"""Trip-level model of the dice loop ("where the instructions go").

The CPU oracle (test infrastructure) is text-patched HERE, into a scratch copy, to hand every dice battle's roll sequence (start
armies, stream position, attacker loss of each roll) to policy.h, which replays it through the rollout kernel's block / scan / scalar
stepping policy and through alternatives (sum criterion, 64-wide scan, register-resident Philox, speculative block) and adds up trips
and per-trip instruction costs taken from the ncu SASS page of the shipped kernel.

    [ENUM=8] [SPEC_MIN=1] python tests/tools/dice_trip_model/run.py [playouts=2048]

Analysis tooling that drives the oracle, hence kept under tests/; not collected by pytest, not used by the product or the bench.
"""
import sys, os, subprocess, numpy as np, ctypes as C
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
OUT = os.path.join(ROOT, "gpurun_out", "dice_trip_model")
os.makedirs(OUT, exist_ok=True)
sys.path.insert(0, ROOT)
s = open(ROOT + "/oracle/oponn_oracle.c").read()
edits = [
 ('#include "../include/orx_state.h"', '#include "%s/include/orx_state.h"\n#include "%s/policy.h"' % (ROOT, HERE)),
 ("        int remainingAttackers = attackers, remainingDefenders = defenders;\n", "        int remainingAttackers = attackers, remainingDefenders = defenders; XN = 0; A0 = attackers; D0 = defenders; POS0 = s->pos;\n"),
 ("                if (ct) ct->single_rolls++;", "                if (ct) ct->single_rolls++; XS[XN++] = (unsigned char)attackerLoss;"),
 ("        if (remainingAttackers != 0 && remainingDefenders != 0 && !*s->flags) {\n            if (remainingAttackers < 0", "        analyze();\n        if (remainingAttackers != 0 && remainingDefenders != 0 && !*s->flags) {\n            if (remainingAttackers < 0"),
]
for a, b in edits:
    assert a in s, a
    s = s.replace(a, b, 1)
open(os.path.join(OUT, "o.c"), "w").write(s)
so = os.path.join(OUT, "libo.so")
subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11", "-ffp-contract=off", "-shared", "-DENUM=%s" % os.environ.get("ENUM", "8"), "-DSPEC_MIN=%s" % os.environ.get("SPEC_MIN", "1"), "-o", so, os.path.join(OUT, "o.c"), "-lm", "-lpthread"])
import oracle.pyoracle as po
po._LIB_PATH = so; po.build = lambda force=False: so
from oponn_b200.maps import classic42
from oponn_b200.state import Rules
L = po.lib()
leaf = np.fromfile(ROOT + "/tests/golden/classic42_mid_s1.leaf", dtype=np.uint8)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
po.Oracle(classic42(6), Rules()).rollout_batch(leaf, n, 20261019, threads=1)
L.orc_policy_report(n)
