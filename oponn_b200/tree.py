"""Host mirror of the search side (include/orx_tree.h): the leaf producer (``ExplorationBoard``: legal moves, move
application) and the batched UCT driver (``MCTSTree.runMCTS`` with thousands of leaves in flight under virtual loss).

Reference: src/com/mld46/oponn/sim/boards/ExplorationBoard.java, src/com/mld46/oponn/MCTSTree.java,
src/com/mld46/oponn/moves/*.java.  All logic lives in ``liborx.so`` (C++ host code + the CUDA rollout engine).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .engine import ORX_OK, OrxError, RolloutEngine, load_library
from .state import MAX_PLAYERS, BoardState, OrxMapC, OrxRulesC, RiskMap, Rules, leaf_layout

(MOVE_COUNTRY_SELECTION, MOVE_INITIAL_PLACEMENT, MOVE_CARD_CASH, MOVE_PLACEMENT, MOVE_ATTACK, MOVE_ATTACK_OUTCOME,
 MOVE_INVASION, MOVE_FORTIFICATION, MOVE_NEXT_PHASE) = range(9)
MOVE_NAMES = ("CountrySelection", "InitialPlacement", "CardCash", "Placement", "Attack", "AttackOutcome", "Invasion",
              "Fortification", "NextPhase")

MOVE_DTYPE = np.dtype([("type", "<i4"), ("phase", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4"), ("d", "<i4"),
                       ("probability", "<f4"), ("flag", "<i4")])
ROOT_CHILD_DTYPE = np.dtype([("move", MOVE_DTYPE), ("visits", "<i8"), ("wins", "<i8", (MAX_PLAYERS,)),
                             ("losses", "<i8", (MAX_PLAYERS,)), ("win_len_sum", "<i8", (MAX_PLAYERS,)),
                             ("loss_len_sum", "<i8", (MAX_PLAYERS,))])
STATS_DTYPE = np.dtype([("iterations", "<i8"), ("playouts", "<i8"), ("batches", "<i8"), ("nodes", "<i8"),
                        ("terminal_hits", "<i8"), ("max_depth", "<i8"), ("select_ms", "<f8"), ("rollout_ms", "<f8"),
                        ("backprop_ms", "<f8")])
assert MOVE_DTYPE.itemsize == 32 and ROOT_CHILD_DTYPE.itemsize == 32 + 8 + 4 * 64


class OrxTreeOptionsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("placement_partial_order", "attack_aggressive", "attack_repeat_moves",
                                         "attack_partial_order", "attack_until_dead", "invasion",
                                         "fortification_outwards", "agent_id", "playouts_per_leaf", "leaves_in_flight",
                                         "max_child_policy", "reserved")] + [("seed", C.c_uint64)]


@dataclass
class TreeOptions:
    """ExplorationBoard's pruning switches (defaults: Oponn.java:46-52) + search settings."""
    placement_partial_order: bool = True
    attack_aggressive: bool = True
    attack_repeat_moves: bool = True
    attack_partial_order: bool = False
    attack_until_dead: bool = True
    invasion: bool = True
    fortification_outwards: bool = True
    agent_id: int = 0
    playouts_per_leaf: int = 8          # "cores"
    leaves_in_flight: int = 1024        # 1 = the reference's one-leaf-at-a-time loop
    max_child_policy: bool = False
    seed: int = 1

    def to_c(self) -> OrxTreeOptionsC:
        return OrxTreeOptionsC(*[int(getattr(self, n)) for n in
                                 ("placement_partial_order", "attack_aggressive", "attack_repeat_moves",
                                  "attack_partial_order", "attack_until_dead", "invasion", "fortification_outwards",
                                  "agent_id", "playouts_per_leaf", "leaves_in_flight", "max_child_policy")], 0, self.seed)


def _bind(L):
    if getattr(L, "_tree_bound", False):
        return L
    vp = C.c_void_p
    L.orx_tree_default_options.argtypes = [C.POINTER(OrxTreeOptionsC)]; L.orx_tree_default_options.restype = None
    L.orx_tree_create.argtypes = [C.POINTER(OrxMapC), C.POINTER(OrxRulesC), C.POINTER(OrxTreeOptionsC), vp, C.POINTER(vp)]
    L.orx_tree_create.restype = C.c_int
    L.orx_tree_destroy.argtypes = [vp]; L.orx_tree_destroy.restype = None
    L.orx_tree_possible_moves.argtypes = [vp, vp, vp, vp, C.c_int]; L.orx_tree_possible_moves.restype = C.c_int
    L.orx_tree_apply.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int32)]; L.orx_tree_apply.restype = C.c_int
    L.orx_tree_search.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int32), vp]; L.orx_tree_search.restype = C.c_int
    L.orx_tree_choose.argtypes = [vp, vp, C.c_int, C.c_uint64]; L.orx_tree_choose.restype = C.c_int
    L.orx_new_game.argtypes = [C.POINTER(OrxMapC), C.c_uint64, vp]; L.orx_new_game.restype = C.c_int
    L._tree_bound = True
    return L


TREE_EXPORTS = ("orx_new_game", "orx_tree_default_options", "orx_tree_create", "orx_tree_destroy", "orx_tree_possible_moves",
                "orx_tree_apply", "orx_tree_search", "orx_tree_choose")


def new_game(risk_map: RiskMap, seed: int) -> np.ndarray:
    """SimulationBoard.setupNewGame (SimulationBoard.java:141-176) as a packed leaf record."""
    L = _bind(load_library())
    mc = risk_map.to_c()
    out = np.zeros(leaf_layout(risk_map.n_territories, risk_map.n_players).stride, dtype=np.uint8)
    rc = L.orx_new_game(C.byref(mc), seed, out.ctypes.data)
    if rc != ORX_OK:
        raise OrxError(rc, "orx_new_game rejected the map")
    return out


def move_str(m) -> str:
    return f"{MOVE_NAMES[int(m['type'])]}(a={int(m['a'])}, b={int(m['b'])}, c={int(m['c'])}, d={int(m['d'])}, flag={int(m['flag'])})"


class SearchTree:
    """ExplorationBoard + MCTSTree over one map / rule set.  ``engine`` may be None for move generation only."""

    def __init__(self, risk_map: RiskMap, rules: Optional[Rules] = None, options: Optional[TreeOptions] = None,
                 engine: Optional[RolloutEngine] = None):
        self._h = None
        self._L = _bind(load_library())
        self.map, self.rules, self.options, self.engine = risk_map, rules or Rules(), options or TreeOptions(), engine
        self._mc, self._rc, self._oc = risk_map.to_c(), self.rules.to_c(), self.options.to_c()
        h = C.c_void_p()
        rc = self._L.orx_tree_create(C.byref(self._mc), C.byref(self._rc), C.byref(self._oc),
                                     engine._h if engine is not None else None, C.byref(h))
        if rc != ORX_OK:
            raise OrxError(rc, "orx_tree_create rejected the map")
        self._h = h
        self.layout = leaf_layout(risk_map.n_territories, risk_map.n_players)
        self.n = risk_map.n_territories

    def close(self):
        if self._h:
            self._L.orx_tree_destroy(self._h); self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers
    def _leaf(self, leaf) -> np.ndarray:
        if isinstance(leaf, BoardState):
            leaf = leaf.pack(self.map)
        return np.ascontiguousarray(leaf, dtype=np.uint8)

    def make_extra(self, moveable: Sequence[int], fortified: Sequence[int]) -> np.ndarray:
        e = np.zeros(5 * self.n, dtype=np.uint8)
        e[:4 * self.n] = np.asarray(moveable, dtype="<i4").view(np.uint8)
        e[4 * self.n:] = np.asarray(fortified, dtype=np.uint8)
        return e

    def split_extra(self, extra: np.ndarray):
        return extra[:4 * self.n].view("<i4").copy(), extra[4 * self.n:].copy()

    # -- ExplorationBoard
    def getPossibleMoves(self, leaf, extra: Optional[np.ndarray] = None) -> np.ndarray:
        lf = self._leaf(leaf)
        cap = 4096
        out = np.zeros(cap, dtype=MOVE_DTYPE)
        n = self._L.orx_tree_possible_moves(self._h, lf.ctypes.data, extra.ctypes.data if extra is not None else None,
                                            out.ctypes.data, cap)
        if n < 0:
            raise OrxError(n, "illegal board state for move generation")
        return out[:min(n, cap)].copy()

    def updateState(self, leaf, move, extra: Optional[np.ndarray] = None):
        """returns (successor leaf record, successor extra, simOngoing)"""
        lf = self._leaf(leaf)
        mv = np.ascontiguousarray(np.asarray(move, dtype=MOVE_DTYPE).reshape(1))
        out = np.zeros(self.layout.stride, dtype=np.uint8)
        eout = np.zeros(5 * self.n, dtype=np.uint8)
        on = C.c_int32()
        rc = self._L.orx_tree_apply(self._h, lf.ctypes.data, extra.ctypes.data if extra is not None else None,
                                    mv.ctypes.data, out.ctypes.data, eout.ctypes.data, C.byref(on))
        if rc != ORX_OK:
            raise OrxError(rc, "illegal move for this board state")
        return out, eout, bool(on.value)

    # -- MCTSTree
    def search(self, leaf, iterations: int, extra: Optional[np.ndarray] = None):
        """runMCTS up to (not including) the final choice: (root children in getPossibleMoves order, stats)"""
        if self.engine is None:
            raise OrxError(-2, "search needs a RolloutEngine: the playouts run on the GPU, there is no CPU path")
        lf = self._leaf(leaf)
        cap = 4096
        ch = np.zeros(cap, dtype=ROOT_CHILD_DTYPE)
        st = np.zeros(1, dtype=STATS_DTYPE)
        n = C.c_int32()
        rc = self._L.orx_tree_search(self._h, lf.ctypes.data, extra.ctypes.data if extra is not None else None, iterations,
                                     ch.ctypes.data, cap, C.byref(n), st.ctypes.data)
        if rc != ORX_OK:
            raise OrxError(rc, self._L.orx_last_error().decode() or "search failed")
        return ch[:n.value].copy(), st[0]

    def choose(self, children: np.ndarray, tie_break: int = 0) -> int:
        ch = np.ascontiguousarray(children)
        return self._L.orx_tree_choose(self._h, ch.ctypes.data, len(ch), tie_break)

    def runMCTS(self, leaf, iterations: int, extra: Optional[np.ndarray] = None):
        """MCTSTree.runMCTS (MCTSTree.java:159-210): the move to play."""
        ch, st = self.search(leaf, iterations, extra)
        return ch[self.choose(ch)]["move"], ch, st


def merge_root_children(children: np.ndarray, device=None) -> np.ndarray:
    """Root-parallel mode: sum the integer statistics of independent trees over all ranks (one all-reduce per move)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return children
    out = children.copy()
    fields = ("visits", "wins", "losses", "win_len_sum", "loss_len_sum")
    flat = np.concatenate([np.asarray(out[f], dtype=np.int64).reshape(len(out), -1) for f in fields], axis=1)
    t = torch.from_numpy(np.ascontiguousarray(flat))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    flat = t.cpu().numpy()
    k = 0
    for f in fields:
        w = 1 if f == "visits" else MAX_PLAYERS
        out[f] = flat[:, k:k + w].reshape(out[f].shape)
        k += w
    return out
