"""Builds and drives the CPU emulation of the lane-per-game rollout kernel (tests/tools/lane_emulator).

TEST INFRASTRUCTURE ONLY: compiles the per-lane logic of ``oponn_b200/csrc/orx_lanes.cuh`` with g++ and runs the
kernel's section loop lane after lane, so the kernel's rules can be checked against the oracle without a GPU and the
scheduler's section statistics can be read.  Nothing under ``oponn_b200/`` imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))
SRC = os.path.join(HERE, "tools", "lane_emulator", "lanes_host.cpp")
OUT_DIR = os.path.join(HERE, "tools", "lane_emulator", "_build")
LIB = os.path.join(OUT_DIR, "liblanes_host.so")
DEPS = [SRC] + [os.path.join(ROOT, "oponn_b200", "csrc", f) for f in ("orx_lanes.cuh", "orx_devmap.h", "orx_mapimage.h")] + \
    [os.path.join(ROOT, "include", "orx_state.h")]
DEFAULT_TUNE = (4, 8, 8, 8, 64, 32, 8, 8, 1 << 30)
SECTIONS = ("NEW", "TURN", "SEL", "DICE", "FIN")
_lib = None
_tables = None


def build() -> str:
    from oracle import pyoracle
    pyoracle.build()
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    ora = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w", "-o", LIB, SRC,
                           "-L" + ora, "-loponn_oracle", "-Wl,-rpath," + ora, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.lanes_emulate.restype = C.c_int
    return _lib


def battle_tables():
    """(meta u32[100*100], rows u64[]) as build_tables_kernel lays them out, from the oracle's float tables:
    row = min(floor(running double sum of float probabilities * 2^53), 2^53 - 1) << 11 | invading << 7 | survivors."""
    global _tables
    if _tables is None:
        from oracle import pyoracle
        meta = np.zeros(100 * 100, dtype=np.uint32)
        rows = []
        off = 0
        for a in range(1, 101):
            for d in range(1, 101):
                pr, al, dl, inv = pyoracle.get_attack_outcomes(a, d)
                total = np.cumsum(pr.astype(np.float64))          # sequential double additions, in row order
                thr = np.minimum(np.floor(total * 9007199254740992.0), 9007199254740991.0).astype(np.uint64)
                code = np.where(inv != 0, 0x80 | (a - al), d - dl).astype(np.uint64)
                rows.append((thr << np.uint64(11)) | code)
                meta[(a - 1) * 100 + (d - 1)] = (off << 8) | len(pr)
                pad = a + d - len(pr)
                if pad:
                    rows.append(np.zeros(pad, dtype=np.uint64))
                off += a + d
        _tables = (meta, np.concatenate(rows))
    return _tables


def emulate(risk_map, rules, leaves: np.ndarray, playouts_per_leaf: int, seed: int, first_game: int = 0,
            tune=DEFAULT_TUNE, n_warps: int = 1, want_stats: bool = False):
    from oponn_b200.state import GAME_RESULT_DTYPE, leaf_layout
    L = leaf_layout(risk_map.n_territories, risk_map.n_players)
    leaves = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, L.stride)
    n = leaves.shape[0]
    games = np.zeros(n * playouts_per_leaf, dtype=GAME_RESULT_DTYPE)
    meta, rows = battle_tables()
    mc, rc = risk_map.to_c(), rules.to_c()
    t = np.array(tune, dtype=np.int32)
    stats = np.zeros(5 * 34 + 2, dtype=np.uint64)
    r = lib().lanes_emulate(C.byref(mc), C.byref(rc), leaves.ctypes.data_as(C.c_void_p), n, playouts_per_leaf,
                            C.c_uint64(seed), C.c_uint64(first_game), meta.ctypes.data_as(C.c_void_p),
                            rows.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), n_warps,
                            games.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p))
    if r != 0:
        raise RuntimeError(f"lanes_emulate failed: {r}")
    if want_stats:
        return games, stats
    return games
