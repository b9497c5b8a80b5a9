"""Pure-Python restatement of MCTSTree.runMCTS (one leaf at a time, as the reference runs it).  TEST INFRASTRUCTURE ONLY.

Checker for the native batched search (orx_tree_search): only tests/ may import it.  Follows
/root/reference/src/com/mld46/oponn/MCTSTree.java: runMCTS :159-210, selection :214-331 (deterministic UCT with
win/loss-length shaping, probability matching at Attack nodes), expansion :335-355, backpropagation :399-431 (running
means in double, exactly as written), chooseRobustMove :541-568, SearchNode :621-674.  The board is
oracle/pyexplore.ExplorationBoard; `rollout(leaf_state) -> [(playerOutOnTurn[], simTurnCount)] * cores` is injected.

Deliberate difference shared with the native search: a selection that ends in a finished game scores that game (the
reference re-reads the previous iteration's boards there).
PARITY PIN STATUS: the reference has no test for MCTSTree => "parity unpinned".
"""
from __future__ import annotations

import math
import random
from typing import Callable, List

from . import pyexplore as px

C, A, EPSILON = 1.0, 0.05, 1e-10          # MCTSTree.java:23-27


class SearchNode:
    def __init__(self, current_player, n_players, parent=None, move=None):
        if parent is not None:
            current_player = move[4] if move[0] == px.M_NEXT_PHASE else parent.currentPlayer
        self.currentPlayer, self.move = current_player, move
        self.visits = 0
        self.wins = [0] * n_players; self.losses = [0] * n_players
        self.averageWinLength = [0.0] * n_players; self.averageLossLength = [0.0] * n_players
        self.children: List[SearchNode] = []; self.unvisitedChildren: list = []


def run_mcts(board: px.ExplorationBoard, setup: Callable[[], None], rollout: Callable, agent_id: int, n_players: int,
             loops: int, rng: random.Random):
    root = SearchNode(agent_id, n_players)
    setup()
    root.unvisitedChildren = list(board.possible_moves())
    root_moves = list(root.unvisitedChildren)
    i = 0
    while i < loops and not (len(root.children) == 1 and not root.unvisitedChildren):
        i += 1
        path = [root]
        setup()
        node, ongoing = root, True
        while ongoing:
            if node.unvisitedChildren or not node.children:
                break
            if node.move is not None and node.move[0] == px.M_ATTACK:      # stochasticSelection
                best, pick = float(-(1 << 31)), None
                for ch in node.children:
                    sc = float(ch.move[6])
                    if node.visits != 0:
                        sc -= ch.visits / float(node.visits)
                    if sc > best:
                        best, pick = sc, ch
            else:                                                           # deterministicSelection
                logv = math.log(node.visits + 1)
                aw = [ch.averageWinLength[ch.currentPlayer] for ch in node.children]
                al = [ch.averageLossLength[ch.currentPlayer] for ch in node.children]
                dW, dL = max(aw) - min(aw), max(al) - min(al)
                a = A * min(max(len(node.children) - 2, 0), 100) / 100.0
                best, pick = -math.inf, None
                for ch in node.children:
                    cp = ch.currentPlayer
                    wp = ch.wins[cp] / float(ch.visits)
                    ws, wls, lls = (1 - a) * wp, 0.0, 0.0
                    if wp != 0.0:
                        wls = a * wp
                        if dW != 0.0:
                            wls *= (max(aw) - ch.averageWinLength[cp]) / dW
                    if wp != 1.0:
                        lls = a * (1 - wp)
                        if dL != 0.0:
                            lls *= (ch.averageLossLength[cp] - min(al)) / dL
                    sc = C * math.sqrt(logv / (ch.visits + 1)) + ws + wls + lls + rng.random() * EPSILON
                    if sc > best:
                        best, pick = sc, ch
            path.append(pick)
            ongoing = board.update(pick.move)
            node = pick
        results = None
        if ongoing:                                                         # expansion
            mv = node.unvisitedChildren.pop(rng.randrange(len(node.unvisitedChildren)))
            child = SearchNode(None, n_players, node, mv)
            node.children.append(child)
            ongoing = board.update(mv)
            if ongoing:
                child.unvisitedChildren = list(board.possible_moves())
            path.append(child)
        if ongoing:
            results = rollout(board)                                        # simulation: `cores` playouts of this leaf
        else:
            results = [(list(board.out_on_turn), 0)] * rollout.cores
        for pot, final_turn in results:                                     # backpropagation
            for nd in path:
                for p in range(n_players):
                    if pot[p] == -1:
                        nd.averageWinLength[p] = (final_turn + nd.wins[p] * nd.averageWinLength[p]) / (nd.wins[p] + 1)
                        nd.wins[p] += 1
                    else:
                        nd.averageLossLength[p] = (final_turn + nd.losses[p] * nd.averageLossLength[p]) / (nd.losses[p] + 1)
                        nd.losses[p] += 1
                nd.visits += 1
    return root, root_moves


def choose_robust(root: SearchNode, agent_id: int, rng: random.Random):
    best_v, best_l, best = -(1 << 31), float((1 << 31) - 1), []
    for ch in root.children:
        if ch.visits > best_v or (ch.visits == best_v and ch.averageWinLength[agent_id] < best_l):
            best_v, best_l, best = ch.visits, ch.averageWinLength[agent_id], [ch.move]
        elif ch.visits == best_v and ch.averageWinLength[agent_id] == best_l:
            best.append(ch.move)
    return best[int(rng.random() * len(best))]
