"""ctypes binding of the CPU oracle (oracle/oponn_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this module; nothing under ``oponn_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboponn_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oponn_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "orx_state.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "liboponn_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u32p, i32p, u8p, f32p, f64p, u64p = (C.POINTER(t) for t in
                                             (C.c_uint32, C.c_int32, C.c_uint8, C.c_float, C.c_double, C.c_uint64))
        L.orc_init.argtypes = [C.c_int]
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_fdlibm_log.argtypes = [C.c_double]; L.orc_fdlibm_log.restype = C.c_double
        L.orc_stream_draws.argtypes = [C.c_uint64, C.c_uint64, u32p, C.c_uint32, i32p, C.c_int, f64p, u32p, u32p]
        L.orc_get_attack_outcomes.argtypes = [C.c_int, C.c_int, f32p, i32p, i32p, i32p, C.c_int]
        L.orc_get_attack_outcomes.restype = C.c_int
        L.orc_simulate_battles.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, i32p, i32p]
        L.orc_simulate_individual.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, i32p]
        L.orc_card_cash_value.argtypes = [C.c_char_p, C.c_int]; L.orc_card_cash_value.restype = C.c_int
        L.orc_starting_armies.argtypes = [C.c_int, C.c_int, i32p]; L.orc_starting_armies.restype = C.c_int
        L.orc_create.argtypes = [C.c_void_p, C.c_void_p]; L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_sim_agents.argtypes = [C.c_void_p, i32p, C.c_int]; L.orc_set_sim_agents.restype = C.c_int
        L.orc_rollout_batch.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p,
                                        C.c_void_p, u64p, C.c_int]
        L.orc_rollout_batch.restype = C.c_int
        L.orc_rollout_replay.argtypes = [C.c_void_p, u8p, C.c_int, u32p, C.c_uint32, C.c_void_p]
        L.orc_rollout_replay.restype = C.c_int
        L.orc_advance_to_leaf.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_int, u8p]
        L.orc_advance_to_leaf.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def philox4x32_10(ctr: Sequence[int], key: Sequence[int]) -> Tuple[int, ...]:
    c = np.array(ctr, dtype=np.uint32); k = np.array(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return tuple(int(x) for x in o)


def fdlibm_log(x: float) -> float:
    return lib().orc_fdlibm_log(x)


def stream_draws(ops: Sequence[int], seed: int = 0, game: int = 0, words: Optional[np.ndarray] = None):
    """ops: >0 nextInt(bound); 0 nextDouble; -1 nextGaussian; -2 raw word.  Returns (values, pos, flags)."""
    o = np.array(ops, dtype=np.int32); out = np.zeros(len(o), dtype=np.float64)
    pos = C.c_uint32(); flags = C.c_uint32()
    if words is not None:
        w = np.ascontiguousarray(words, dtype=np.uint32)
        lib().orc_stream_draws(seed, game, _p(w, C.c_uint32), len(w), _p(o, C.c_int32), len(o), _p(out, C.c_double),
                               C.byref(pos), C.byref(flags))
    else:
        lib().orc_stream_draws(seed, game, None, 0, _p(o, C.c_int32), len(o), _p(out, C.c_double), C.byref(pos),
                               C.byref(flags))
    return out, pos.value, flags.value


def get_attack_outcomes(a: int, d: int):
    """BattleSim.getAttackOutcomes(a, d) -> (probability f32[], attackerLosses, defenderLosses, invading)."""
    lib().orc_init(8)
    cap = max(a + d + 2, 4)
    pr = np.zeros(cap, dtype=np.float32); al = np.zeros(cap, dtype=np.int32); dl = np.zeros(cap, dtype=np.int32)
    inv = np.zeros(cap, dtype=np.int32)
    n = lib().orc_get_attack_outcomes(a, d, _p(pr, C.c_float), _p(al, C.c_int32), _p(dl, C.c_int32),
                                      _p(inv, C.c_int32), cap)
    return pr[:n], al[:n], dl[:n], inv[:n]


def simulate_battles(a: int, d: int, n: int, seed: int = 1, first_game: int = 0):
    lib().orc_init(8)
    al = np.zeros(n, dtype=np.int32); dl = np.zeros(n, dtype=np.int32)
    lib().orc_simulate_battles(seed, first_game, a, d, n, _p(al, C.c_int32), _p(dl, C.c_int32))
    return al, dl


def simulate_individual(a: int, d: int, n: int, seed: int = 1, first_game: int = 0):
    al = np.zeros(n, dtype=np.int32)
    lib().orc_simulate_individual(seed, first_game, a, d, n, _p(al, C.c_int32))
    return al


def card_cash_value(progression: str, pos: int) -> int:
    return lib().orc_card_cash_value(progression.encode(), pos)


def starting_armies(n_territories: int, n_players: int):
    out = np.zeros(8, dtype=np.int32)
    lib().orc_starting_armies(n_territories, n_players, _p(out, C.c_int32))
    return [int(x) for x in out[:n_players]]


COUNTER_NAMES = ("tv", "ev", "r_words", "battles", "cdf_steps", "writes", "player_turns", "attacks", "conquests",
                 "single_rolls", "gauss_battles", "table_battles",
                 "b_big_v_1", "b_big_v_2", "b_big_v_3to100", "b_big_v_big", "b_1_v_big", "b_2_v_big", "b_3to100_v_big", "b_small_v_small")


class Oracle:
    """setupState + simulate + getters, batched, on the CPU restatement."""

    def __init__(self, risk_map, rules, sim_agents=None, modelling_turn_limit=0):
        from oponn_b200.state import GAME_RESULT_DTYPE, LEAF_RESULT_DTYPE, leaf_layout
        self._gdt, self._ldt = GAME_RESULT_DTYPE, LEAF_RESULT_DTYPE
        self.map, self.rules = risk_map, rules
        self._mc, self._rc = risk_map.to_c(), rules.to_c()
        self._h = lib().orc_create(C.byref(self._mc), C.byref(self._rc))
        if not self._h:
            raise ValueError("orc_create rejected the map")
        self.layout = leaf_layout(risk_map.n_territories, risk_map.n_players)
        if sim_agents is not None:
            self.set_sim_agents(sim_agents, modelling_turn_limit)

    def set_sim_agents(self, sim_agents, modelling_turn_limit=2):
        """new ModellingBoard(bs, sas, cm, cid, modellingTurnLimit): agent ids per player (oponn_b200.state.AGENT_*)"""
        ids = np.asarray(list(sim_agents), dtype=np.int32)
        if len(ids) != self.map.n_players:
            raise ValueError("one agent id per player")
        if lib().orc_set_sim_agents(self._h, _p(ids, C.c_int32), int(modelling_turn_limit)) != 0:
            raise ValueError("orc_set_sim_agents rejected the agent ids")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h); self._h = None

    def __del__(self):
        self.close()

    def rollout_batch(self, leaves: np.ndarray, playouts_per_leaf: int, seed: int, first_game: int = 0,
                      per_game: bool = True, threads: int = 1, counters: bool = False):
        leaves = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, self.layout.stride)
        n = leaves.shape[0]
        agg = np.zeros(n, dtype=self._ldt)
        games = np.zeros(n * playouts_per_leaf, dtype=self._gdt) if per_game else None
        ct = np.zeros(20, dtype=np.uint64)
        rc = lib().orc_rollout_batch(self._h, _p(leaves, C.c_uint8), n, playouts_per_leaf, seed, first_game,
                                     agg.ctypes.data, games.ctypes.data if per_game else None,
                                     _p(ct, C.c_uint64), threads)
        if rc != 0:
            raise RuntimeError(f"orc_rollout_batch failed: {rc}")
        if counters:
            return agg, games, dict(zip(COUNTER_NAMES, (int(x) for x in ct)))
        return agg, games

    def rollout_replay(self, leaves: np.ndarray, words: np.ndarray):
        leaves = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, self.layout.stride)
        words = np.ascontiguousarray(words, dtype=np.uint32)
        n = leaves.shape[0]
        assert words.shape[0] == n
        games = np.zeros(n, dtype=self._gdt)
        rc = lib().orc_rollout_replay(self._h, _p(leaves, C.c_uint8), n, _p(words, C.c_uint32), words.shape[1],
                                      games.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"orc_rollout_replay failed: {rc}")
        return games

    def advance_to_leaf(self, leaf: np.ndarray, seed: int, game: int, stop_at_sim_turn: int):
        leaf = np.ascontiguousarray(leaf, dtype=np.uint8)
        out = np.zeros(self.layout.stride, dtype=np.uint8)
        alive = lib().orc_advance_to_leaf(self._h, _p(leaf, C.c_uint8), seed, game, stop_at_sim_turn,
                                          _p(out, C.c_uint8))
        return alive, out
