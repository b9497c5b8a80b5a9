"""Host mirror of the reference's rollout interface over ``liborx.so`` (include/orx.h).

The reference's seam is three Java calls made from ``MCTSTree.simulation`` / ``backpropagation``
(src/com/mld46/oponn/MCTSTree.java:372-373,409-410)::

    simulationBoards[id].setupState(boardState); simulationBoards[id].simulate();
    ... board.getPlayerOutOnTurn(); board.getSimTurnCount();

:class:`ModellingBoard` keeps those names and meanings for one playout; :class:`RolloutEngine` is the
batched form the GPU wants (many leaves x many playouts per call).  :class:`BattleSim` mirrors the
static methods of src/com/mld46/oponn/BattleSim.java.  Everything runs in CUDA kernels inside
``liborx.so``; nothing here computes game logic, and nothing falls back to a CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

from . import build as _build
from .state import (GAME_RESULT_DTYPE, LEAF_RESULT_DTYPE, BoardState, OrxMapC, OrxRulesC, RiskMap, Rules,
                    leaf_layout, pack_leaves)

ORX_OK, ORX_E_INVALID, ORX_E_CUDA, ORX_E_NOMEM = 0, -1, -2, -4

_lib = None


class OrxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"liborx error {code}: {message}")
        self.code = code


def lib_path() -> str:
    return os.environ.get("ORX_LIB", _build.LIB)


def load_library():
    """dlopen ``liborx.so``.  Raises if the CUDA library has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m oponn_b200.build` "
                          "(__graft_entry__.build()); oponn_b200 has no CPU rollout path")
    L = C.CDLL(path)
    vp, i32, u32, u64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64
    p = C.POINTER
    L.orx_create.argtypes = [p(OrxMapC), p(OrxRulesC), C.c_int, p(vp)]; L.orx_create.restype = C.c_int
    L.orx_destroy.argtypes = [vp]; L.orx_destroy.restype = None
    L.orx_last_error.argtypes = []; L.orx_last_error.restype = C.c_char_p
    L.orx_leaf_stride.argtypes = [vp]; L.orx_leaf_stride.restype = C.c_int
    L.orx_rollout_batch.argtypes = [vp, vp, C.c_int, C.c_int, u64, u64, vp, vp]; L.orx_rollout_batch.restype = C.c_int
    L.orx_rollout_batch_device.argtypes = [vp, vp, C.c_int, C.c_int, u64, u64, vp, vp]
    L.orx_rollout_batch_device.restype = C.c_int
    L.orx_rollout_begin.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, u64, u64, C.c_int]; L.orx_rollout_begin.restype = C.c_int
    L.orx_rollout_end.argtypes = [vp, C.c_int, vp, vp]; L.orx_rollout_end.restype = C.c_int
    L.orx_rollout_replay.argtypes = [vp, vp, C.c_int, vp, u32, vp]; L.orx_rollout_replay.restype = C.c_int
    L.orx_sync.argtypes = [vp]; L.orx_sync.restype = C.c_int
    L.orx_stream.argtypes = [vp]; L.orx_stream.restype = vp
    L.orx_last_kernel_ms.argtypes = [vp]; L.orx_last_kernel_ms.restype = C.c_float
    L.orx_kernel_launches.argtypes = [vp]; L.orx_kernel_launches.restype = u64
    L.orx_last_kernel_kind.argtypes = [vp]; L.orx_last_kernel_kind.restype = C.c_int
    if hasattr(L, "orx_kernel_info"):   # (absent from libraries built before round 2: variant timing only)
        L.orx_kernel_info.argtypes = [vp, u64, p(i32), p(i32), p(i32), p(i32)]; L.orx_kernel_info.restype = C.c_int
    L.orx_get_attack_outcomes.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int]
    L.orx_get_attack_outcomes.restype = C.c_int
    L.orx_simulate_battles.argtypes = [vp, C.c_int, C.c_int, C.c_int, u64, u64, vp, vp]
    L.orx_simulate_battles.restype = C.c_int
    L.orx_simulate_individual.argtypes = [vp, C.c_int, C.c_int, C.c_int, u64, u64, vp]
    L.orx_simulate_individual.restype = C.c_int
    L.orx_stream_draws.argtypes = [vp, u64, u64, vp, u32, vp, C.c_int, vp, p(u32), p(u32)]
    L.orx_stream_draws.restype = C.c_int
    L.orx_card_cash_value.argtypes = [vp, C.c_int, p(i32)]; L.orx_card_cash_value.restype = C.c_int
    L.orx_set_sim_agents.argtypes = [vp, p(i32), C.c_int]; L.orx_set_sim_agents.restype = C.c_int
    _lib = L
    return L


#: every symbol include/orx.h declares (checked by tests/test_abi.py)
EXPORTS = ("orx_create", "orx_destroy", "orx_last_error", "orx_leaf_stride", "orx_rollout_batch",
           "orx_rollout_batch_device", "orx_rollout_begin", "orx_rollout_end", "orx_rollout_replay", "orx_sync", "orx_stream", "orx_last_kernel_ms",
           "orx_kernel_launches", "orx_last_kernel_kind", "orx_kernel_info", "orx_get_attack_outcomes", "orx_simulate_battles", "orx_simulate_individual",
           "orx_stream_draws", "orx_card_cash_value", "orx_set_sim_agents")


class RolloutEngine:
    """One ``OrxCtx``: a map + rules staged on one GPU, with the battle tables built there.

    Replaces ``new ModellingBoard(...) x CORES`` (src/com/mld46/oponn/Oponn.java:106-110)."""

    def __init__(self, risk_map: RiskMap, rules: Optional[Rules] = None, device: int = 0,
                 sim_agents: Optional[Sequence[int]] = None, modelling_turn_limit: int = 2):
        self._h = None
        self._L = load_library()
        self.map, self.rules, self.device = risk_map, rules or Rules(), device
        self._mc, self._rc = risk_map.to_c(), self.rules.to_c()
        h = C.c_void_p()
        rc = self._L.orx_create(C.byref(self._mc), C.byref(self._rc), device, C.byref(h))
        if rc != ORX_OK:
            raise OrxError(rc, self._L.orx_last_error().decode())
        self._h = h
        self.layout = leaf_layout(risk_map.n_territories, risk_map.n_players)
        assert self._L.orx_leaf_stride(self._h) == self.layout.stride
        if sim_agents is not None:
            self.set_sim_agents(sim_agents, modelling_turn_limit)

    def set_sim_agents(self, sim_agents: Optional[Sequence[int]], modelling_turn_limit: int = 2):
        """The ``SimAgent[] sas`` and ``modellingTurnLimit`` of ``new ModellingBoard(bs, sas, cm, cid, limit)``
        (ModellingBoard.java:13-19; Oponn.createSimAgents, Oponn.java:146-162): one ``state.AGENT_*`` id per player
        (``state.agent_from_name`` maps Lux agent names); ``None`` switches opponent modelling off."""
        if sim_agents is None:
            self._ck(self._L.orx_set_sim_agents(self._h, None, 0))
            return
        ids = np.asarray(list(sim_agents), dtype=np.int32)
        if len(ids) != self.map.n_players:
            raise ValueError("one agent id per player")
        self._ck(self._L.orx_set_sim_agents(self._h, ids.ctypes.data_as(C.POINTER(C.c_int32)), int(modelling_turn_limit)))

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if self._h:
            self._L.orx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc: int):
        if rc < 0:
            raise OrxError(rc, self._L.orx_last_error().decode())
        return rc

    # -- rollouts ------------------------------------------------------------------------
    def _leaves(self, leaves) -> np.ndarray:
        if isinstance(leaves, BoardState):
            leaves = [leaves]
        if not isinstance(leaves, np.ndarray):
            leaves = pack_leaves(self.map, list(leaves))
        return np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, self.layout.stride)

    def rollout_batch(self, leaves, playouts_per_leaf: int, seed: int, first_game: int = 0, per_game: bool = False):
        """HOST buffers in, HOST results out (the reference-facing call).  Returns (agg, games|None)."""
        lv = self._leaves(leaves)
        n = lv.shape[0]
        agg = np.zeros(n, dtype=LEAF_RESULT_DTYPE)
        games = np.zeros(n * playouts_per_leaf, dtype=GAME_RESULT_DTYPE) if per_game else None
        self._ck(self._L.orx_rollout_batch(self._h, lv.ctypes.data, n, playouts_per_leaf, seed, first_game,
                                           agg.ctypes.data, games.ctypes.data if per_game else None))
        return agg, games

    def rollout_batch_device(self, leaves_ptr: int, n_leaves: int, playouts_per_leaf: int, seed: int,
                             first_game: int, agg_ptr: int, games_ptr: Optional[int] = None):
        """DEVICE pointers (e.g. ``torch.Tensor.data_ptr()``), asynchronous on the engine's stream."""
        self._ck(self._L.orx_rollout_batch_device(self._h, leaves_ptr, n_leaves, playouts_per_leaf, seed, first_game,
                                                  agg_ptr, games_ptr))

    def rollout_begin(self, slot: int, leaves, playouts_per_leaf: int, seed: int, first_game: int = 0, per_game: bool = False):
        """launch on one of the two in-flight slots; pair with :meth:`rollout_end`"""
        lv = self._leaves(leaves)
        self._ck(self._L.orx_rollout_begin(self._h, slot, lv.ctypes.data, lv.shape[0], playouts_per_leaf, seed, first_game,
                                           int(per_game)))
        return lv.shape[0], playouts_per_leaf, per_game

    def rollout_end(self, slot: int, ticket):
        n, ppl, per_game = ticket
        agg = np.zeros(n, dtype=LEAF_RESULT_DTYPE)
        games = np.zeros(n * ppl, dtype=GAME_RESULT_DTYPE) if per_game else None
        self._ck(self._L.orx_rollout_end(self._h, slot, agg.ctypes.data, games.ctypes.data if per_game else None))
        return agg, games

    def rollout_replay(self, leaves, words: np.ndarray) -> np.ndarray:
        """Game g plays leaf g on the injected word stream ``words[g]`` (replayed dice/move stream)."""
        lv = self._leaves(leaves)
        w = np.ascontiguousarray(words, dtype=np.uint32)
        assert w.ndim == 2 and w.shape[0] == lv.shape[0]
        games = np.zeros(lv.shape[0], dtype=GAME_RESULT_DTYPE)
        self._ck(self._L.orx_rollout_replay(self._h, lv.ctypes.data, lv.shape[0], w.ctypes.data, w.shape[1],
                                            games.ctypes.data))
        return games

    def sync(self):
        self._ck(self._L.orx_sync(self._h))

    @property
    def stream(self) -> int:
        return int(self._L.orx_stream(self._h) or 0)

    @property
    def last_kernel_ms(self) -> float:
        return float(self._L.orx_last_kernel_ms(self._h))

    @property
    def last_kernel_kind(self) -> str:
        """which rollout kernel the last launch used: "warp" (one warp per game) or "lanes" (one lane per game)"""
        return {1: "warp", 2: "lanes"}.get(int(self._L.orx_last_kernel_kind(self._h)), "none")

    def kernel_info(self, total_playouts: int) -> dict:
        """launch shape of the rollout kernel a batch of that many playouts would use (orx_kernel_info)"""
        k, b, t, sm = (C.c_int32() for _ in range(4))
        self._ck(self._L.orx_kernel_info(self._h, int(total_playouts), C.byref(k), C.byref(b), C.byref(t), C.byref(sm)))
        return {"kernel": {1: "warp", 2: "lanes"}[k.value], "blocks_per_sm": b.value, "threads_per_block": t.value,
                "smem_bytes_per_block": sm.value, "warps_per_sm": b.value * t.value // 32}

    @property
    def kernel_launches(self) -> int:
        return int(self._L.orx_kernel_launches(self._h))

    # -- BattleSim / stream / cards ------------------------------------------------------
    def get_attack_outcomes(self, a: int, d: int):
        cap = a + d + 2
        pr = np.zeros(cap, dtype=np.float32); al = np.zeros(cap, dtype=np.int32); dl = np.zeros(cap, dtype=np.int32)
        inv = np.zeros(cap, dtype=np.int32)
        n = self._ck(self._L.orx_get_attack_outcomes(self._h, a, d, pr.ctypes.data, al.ctypes.data, dl.ctypes.data,
                                                     inv.ctypes.data, cap))
        return pr[:n], al[:n], dl[:n], inv[:n]

    def simulate_battles(self, a: int, d: int, n: int, seed: int = 1, first_game: int = 0):
        al = np.zeros(n, dtype=np.int32); dl = np.zeros(n, dtype=np.int32)
        self._ck(self._L.orx_simulate_battles(self._h, a, d, n, seed, first_game, al.ctypes.data, dl.ctypes.data))
        return al, dl

    def simulate_individual(self, a: int, d: int, n: int, seed: int = 1, first_game: int = 0):
        al = np.zeros(n, dtype=np.int32)
        self._ck(self._L.orx_simulate_individual(self._h, a, d, n, seed, first_game, al.ctypes.data))
        return al

    def stream_draws(self, ops: Sequence[int], seed: int = 0, game: int = 0, words: Optional[np.ndarray] = None):
        o = np.array(ops, dtype=np.int32); out = np.zeros(len(o), dtype=np.float64)
        pos, flags = C.c_uint32(), C.c_uint32()
        if words is not None:
            w = np.ascontiguousarray(words, dtype=np.uint32)
            # an empty replay stream is still a replay stream: pass a non-null pointer
            wp = w.ctypes.data if len(w) else np.zeros(1, dtype=np.uint32).ctypes.data
            self._ck(self._L.orx_stream_draws(self._h, seed, game, wp, len(w), o.ctypes.data, len(o), out.ctypes.data,
                                              C.byref(pos), C.byref(flags)))
        else:
            self._ck(self._L.orx_stream_draws(self._h, seed, game, None, 0, o.ctypes.data, len(o), out.ctypes.data,
                                              C.byref(pos), C.byref(flags)))
        return out, pos.value, flags.value

    def card_cash_value(self, pos: int) -> int:
        v = C.c_int32()
        self._ck(self._L.orx_card_cash_value(self._h, pos, C.byref(v)))
        return v.value


class ModellingBoard:
    """The reference's rollout board (src/com/mld46/oponn/sim/boards/ModellingBoard.java, base class
    SimulationBoard.java) as seen from ``MCTSTree``: same method names and meanings, one playout per
    ``simulate()``.  ``simulateBatch`` is the form to use for throughput."""

    def __init__(self, engine: RolloutEngine, seed: int = 0):
        self.engine, self.seed = engine, seed
        self._leaf: Optional[np.ndarray] = None
        self._game = 0
        self._result = None

    def setupState(self, boardState: BoardState) -> None:      # SimulationBoard.java:178-190
        self._leaf = boardState.pack(self.engine.map)

    def simulate(self):                                          # SimulationBoard.java:365-481
        if self._leaf is None:
            raise ValueError("setupState has not been called")
        _, games = self.engine.rollout_batch(self._leaf[None, :], 1, self.seed, self._game, per_game=True)
        self._game += 1
        self._result = games[0]
        if self._result["flags"] & 1:
            raise ValueError("illegal board state (the reference would have thrown)")
        return self.getPlayerOutOnTurn()

    def simulateBatch(self, playouts: int):
        if self._leaf is None:
            raise ValueError("setupState has not been called")
        agg, _ = self.engine.rollout_batch(self._leaf[None, :], playouts, self.seed, self._game)
        self._game += playouts
        return agg[0]

    def getPlayerOutOnTurn(self):                                # SimulationBoard.java:1041-1044
        return [int(x) for x in self._result["player_out_on_turn"][:self.engine.map.n_players]]

    def getSimTurnCount(self) -> int:                            # SimulationBoard.java:1031-1034
        return int(self._result["sim_turn_count"])


class BattleSim:
    """src/com/mld46/oponn/BattleSim.java's public static methods, bound to an engine."""

    def __init__(self, engine: RolloutEngine, seed: int = 0):
        self.engine, self.seed, self._n = engine, seed, 0

    def simulateBattle(self, attackers: int, defenders: int) -> Tuple[int, int]:       # BattleSim.java:66-145
        al, dl = self.engine.simulate_battles(attackers, defenders, 1, self.seed, self._n)
        self._n += 1
        return int(al[0]), int(dl[0])

    def simulateIndividualBattle(self, attackers: int, defenders: int) -> int:         # BattleSim.java:152-162
        if attackers > 3 or defenders > 2 or attackers < 1 or defenders < 1:
            raise ValueError(f"Error: {attackers}:{defenders}")
        al = self.engine.simulate_individual(attackers, defenders, 1, self.seed, self._n)
        self._n += 1
        return int(al[0])

    def getAttackOutcomes(self, attackers: int, defenders: int):                       # BattleSim.java:174-184
        pr, al, dl, inv = self.engine.get_attack_outcomes(attackers, defenders)
        return [dict(probability=float(p), attackerLosses=int(a), defenderLosses=int(d), invading=bool(i))
                for p, a, d, i in zip(pr, al, dl, inv)]

    @staticmethod
    def favourableForAttacker(freeAttackers: int, defenders: int) -> bool:              # BattleSim.java:164-167
        return (freeAttackers >= 3) or (freeAttackers == 3 and defenders == 1)

    @staticmethod
    def unfavourableForAttacker(freeAttackers: int, defenders: int) -> bool:            # BattleSim.java:169-172
        return freeAttackers == 1 and defenders >= 2
