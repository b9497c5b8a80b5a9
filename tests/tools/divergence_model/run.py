"""Trace-driven model of "how many warp instructions would another work decomposition need?".

Question answered: the rollout kernel gives one WARP to one game and runs the game's scalar
decisions redundantly in 32 lanes.  Would one LANE per game (32 games per warp), or a sub-warp group per game, need
fewer warp instructions once lanes/groups that sit in different parts of the game have to take turns?

Method: the CPU oracle (test infrastructure) is text-patched HERE, into a scratch copy, to log one byte per state
visit of every playout (N new game, P turn preamble, S attack selection, R single dice roll, T table battle,
G gaussian battle); two small C programs replay those per-game state sequences through a warp of 32 lanes
(lane_sim.c) or G groups (group_sim.c) under several scheduling policies and add up per-state instruction costs.

    python tests/tools/divergence_model/run.py [playouts=512]

Analysis tooling that drives the oracle, hence kept under tests/ (the oracle is test infrastructure); nothing here is
used by the product, collected by pytest or run by the bench.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
OUT = os.path.join(ROOT, "gpurun_out", "divergence_model")
sys.path.insert(0, ROOT)


def patched_oracle() -> str:
    s = open(os.path.join(ROOT, "oracle", "oponn_oracle.c")).read()
    hook = ('#include "%s"\n#include <stdio.h>\nstatic __thread FILE *g_tr; static void TR(int c) { if (g_tr) fputc(c, g_tr); }\n'
            'void orc_trace_open(const char *p) { g_tr = fopen(p, "wb"); } void orc_trace_close(void) { fclose(g_tr); g_tr = 0; }'
            % os.path.join(ROOT, "include", "orx_state.h"))
    edits = [
        ('#include "../include/orx_state.h"', hook),
        ("if (ct) ct->gauss_battles++;", "if (ct) ct->gauss_battles++; TR('G');"),
        ("if (ct) ct->single_rolls++;", "if (ct) ct->single_rolls++; TR('R');"),
        ("double r = jr_next_double(s);\n                double total = 0;", "TR('T'); double r = jr_next_double(s);\n                double total = 0;"),
        ("        int attackerIndex = jr_next_int(&b->rng, nAtt);", "        TR('S'); int attackerIndex = jr_next_int(&b->rng, nAtt);"),
        ("    int possibleDefenders[MAXN], nDef;", "    int possibleDefenders[MAXN], nDef; TR('P');"),
        ("    setupState(b, leaf);\n", "    TR('N'); setupState(b, leaf);\n"),
    ]
    for a, b in edits:
        assert a in s, a
        s = s.replace(a, b)
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(OUT, "oracle_trace.c")
    open(src, "w").write(s)
    so = os.path.join(OUT, "liborc_trace.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11", "-ffp-contract=off", "-shared", "-o", so, src, "-lm", "-lpthread"])
    return so


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    import oracle.pyoracle as po
    so = patched_oracle()
    po._LIB_PATH = so
    po.build = lambda force=False: so
    from oponn_b200.maps import classic42
    from oponn_b200.state import Rules
    L = po.lib()
    L.orc_trace_open.argtypes = [C.c_char_p]
    leaf = np.fromfile(os.path.join(ROOT, "tests", "golden", "classic42_mid_s1.leaf"), dtype=np.uint8)
    trace = os.path.join(OUT, "trace.bin")
    L.orc_trace_open(trace.encode())
    po.Oracle(classic42(6), Rules()).rollout_batch(leaf, n, 20261019, threads=1)
    L.orc_trace_close()
    for name in ("lane_sim", "group_sim"):
        subprocess.check_call(["gcc", "-O2", "-w", "-o", os.path.join(OUT, name), os.path.join(HERE, name + ".c")])
    print("== one lane per game: rolls per ROLL step, policy 0 = run every state present, 3 = most lanes + ageing")
    for rolls in (2, 8):
        for pol in (0, 1, 3):
            print(subprocess.run([os.path.join(OUT, "lane_sim"), str(rolls), str(pol), str(n), trace], capture_output=True, text=True).stdout.splitlines()[0])
    print("== G games per warp (32/G lanes each), costs calibrated on the shipped kernel")
    for g in (1, 2, 4, 8, 32):
        for pol in (0, 3):
            print(subprocess.run([os.path.join(OUT, "group_sim"), str(g), str(pol), str(n), trace], capture_output=True, text=True).stdout.splitlines()[0])


if __name__ == "__main__":
    main()
