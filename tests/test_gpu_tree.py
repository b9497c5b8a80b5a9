"""The batched UCT driver (orx_tree_search) on the GPU: bookkeeping identities, agreement of the batched search with
the one-leaf-at-a-time loop of the reference, forced and winning positions."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oponn_b200.maps import classic42
from oponn_b200.state import ATTACK_START, PHASE_ATTACK, BoardState, RiskMap, Rules
from oponn_b200.tree import MOVE_ATTACK, MOVE_NEXT_PHASE, SearchTree, TreeOptions, move_str

pytestmark = pytest.mark.gpu


def attack_leaf(classic, mid_leaf):
    """the bench leaf moved to player 0's attack phase with armies to spend: a position with real choices"""
    bs = BoardState.unpack(classic, mid_leaf)
    bs.currentPhase, bs.attackState = PHASE_ATTACK, ATTACK_START
    for t in range(42):
        if bs.owner[t] == 0:
            bs.armies[t] = 6
    return bs


def test_search_bookkeeping(engine, classic, mid_leaf):
    bs = attack_leaf(classic, mid_leaf)
    tree = SearchTree(classic, Rules(), TreeOptions(agent_id=0, playouts_per_leaf=4, leaves_in_flight=256, seed=3), engine)
    ch, st = tree.search(bs, 2000)
    moves = tree.getPossibleMoves(bs)
    assert len(ch) == len(moves) and all(ch[i]["move"].tobytes() == moves[i].tobytes() for i in range(len(ch)))
    assert int(st["iterations"]) == 2000 and int(st["playouts"]) == 8000
    assert int(ch["visits"].sum()) == 8000                     # every playout passes through exactly one root child
    assert np.all(ch["wins"][:, :6].sum(axis=1) == ch["visits"])   # every playout has exactly one winner
    assert np.all(ch["wins"][:, :6] + ch["losses"][:, :6] == ch["visits"][:, None])
    assert int(st["batches"]) <= 2000 // 128 + 3 and int(st["nodes"]) > 1000   # two slots of leaves_in_flight / 2
    best = tree.choose(ch)
    assert 0 <= best < len(ch) and ch[best]["visits"] == ch["visits"].max()   # ROBUST_CHILD
    tree.close()


def test_batched_agrees_with_one_leaf_at_a_time(engine, classic, mid_leaf):
    """leaves_in_flight = 1 is MCTSTree.runMCTS as written; the batched search must estimate the same root values.
    Stated confidence interval: the agent's win rate over all root playouts, 4 sigma of the binomial."""
    bs = attack_leaf(classic, mid_leaf)
    est = {}
    for lif in (1, 512):
        tree = SearchTree(classic, Rules(), TreeOptions(agent_id=0, playouts_per_leaf=8, leaves_in_flight=lif, seed=11), engine)
        ch, st = tree.search(bs, 400)
        tree.close()
        n = int(ch["visits"].sum())
        est[lif] = (ch["wins"][:, 0].sum() / n, n, ch)
    (p1, n1, c1), (p2, n2, c2) = est[1], est[512]
    sigma = np.sqrt(p1 * (1 - p1) / n1 + p2 * (1 - p2) / n2)
    assert abs(p1 - p2) < 4 * sigma + 0.01, (p1, p2, sigma)
    # both concentrate their visits on attacking rather than passing with 6-army stacks
    for c in (c1, c2):
        top = c[np.argmax(c["visits"])]["move"]
        assert int(top["type"]) in (MOVE_ATTACK, MOVE_NEXT_PHASE)


def test_search_with_modelled_opponents(classic, mid_leaf):
    """Opponent modelling is a property of the rollout context (orx_set_sim_agents): the same search, with every
    opponent of player 0 modelled (Oponn.createSimAgents: simAgents[ID] itself stays SimRandom), keeps the bookkeeping
    identities and evaluates the root differently from the all-SimRandom line-up."""
    from oponn_b200.engine import RolloutEngine
    from oponn_b200.state import AGENT_ANGRY, AGENT_RANDOM, AGENT_STINKY
    bs = attack_leaf(classic, mid_leaf)
    rate = {}
    for name, agents in (("random", None), ("modelled", [AGENT_RANDOM, AGENT_ANGRY, AGENT_STINKY, AGENT_ANGRY, AGENT_STINKY, AGENT_ANGRY])):
        with RolloutEngine(classic, Rules(), 0, sim_agents=agents, modelling_turn_limit=2) as e:
            tree = SearchTree(classic, Rules(), TreeOptions(agent_id=0, playouts_per_leaf=8, leaves_in_flight=256, seed=3), e)
            ch, st = tree.search(bs, 1000)
            tree.close()
        assert int(ch["visits"].sum()) == 8000 and np.all(ch["wins"][:, :6].sum(axis=1) == ch["visits"])
        rate[name] = ch["wins"][:, :6].sum(axis=0) / 8000.0
    # two methodical rounds of five opponents change who tends to win from this position (8000 playouts each: the
    # binomial sigma of a rate is below 0.006)
    assert np.abs(rate["random"] - rate["modelled"]).max() > 0.03, rate


def test_forced_move_stops_early(engine, classic, mid_leaf):
    """one legal move: the loop ends after the first expansion (MCTSTree.java:179)"""
    tree = SearchTree(classic, Rules(), TreeOptions(playouts_per_leaf=2, leaves_in_flight=64), engine)
    ch, st = tree.search(mid_leaf, 500)       # CARDS phase, no set in hand: only NextPhase
    assert len(ch) == 1 and int(st["iterations"]) <= 64
    tree.close()


def test_finds_the_winning_attack(engine):
    m = RiskMap(name="line4", n_players=2, adjacency=[[1], [0, 2], [1, 3], [2]], continent=[0, 0, 1, 1], continent_bonus=[2, 3])
    from oponn_b200.engine import RolloutEngine
    with RolloutEngine(m, Rules(use_cards=False), 0) as e:
        tree = SearchTree(m, Rules(use_cards=False), TreeOptions(agent_id=0, playouts_per_leaf=4, leaves_in_flight=64, seed=5), e)
        bs = BoardState(owner=[0, 0, 0, 1], armies=[1, 1, 30, 1], currentPlayer=0, currentPhase=PHASE_ATTACK,
                        attackState=ATTACK_START, playerNumberOfCards=[0, 0])
        move, ch, st = tree.runMCTS(bs, 600)
        assert int(move["type"]) == MOVE_ATTACK and (int(move["a"]), int(move["b"])) == (2, 3), move_str(move)
        assert int(st["terminal_hits"]) > 0       # the search walked into finished games inside the tree
        won = ch[np.argmax(ch["visits"])]
        assert won["wins"][0] / won["visits"] > 0.9
        tree.close()


def test_search_without_engine_refuses(classic, mid_leaf):
    from oponn_b200.engine import OrxError
    tree = SearchTree(classic, Rules())
    with pytest.raises(OrxError):
        tree.search(mid_leaf, 10)
    tree.close()


def test_native_search_vs_python_restatement(engine, oracle, classic, mid_leaf):
    """orx_tree_search at leaves_in_flight = 1 against oracle/pymcts.py (MCTSTree.runMCTS restated in Python over
    oracle/pyexplore.py, playouts by the CPU simulator).  Different random streams, so the comparison is statistical,
    with the interval stated: the agent's root win rate within 4.5 sigma, and the same most-visited move family."""
    import functools
    import random

    from oracle import pyexplore, pymcts, pyoracle

    @functools.lru_cache(maxsize=None)
    def outcomes(a, d):
        pr, al, dl, inv = pyoracle.get_attack_outcomes(a, d)
        return [(np.float32(p), int(x), int(y), bool(v)) for p, x, y, v in zip(pr, al, dl, inv)]

    bs = attack_leaf(classic, mid_leaf)
    iters, cores = 300, 8
    tree = SearchTree(classic, Rules(), TreeOptions(agent_id=0, playouts_per_leaf=cores, leaves_in_flight=1, seed=21), engine)
    ch, st = tree.search(bs, iters)

    board = pyexplore.ExplorationBoard(classic, Rules(), outcomes)
    game = [0]

    class Rollout:
        def __init__(self):
            self.cores = cores

        def __call__(self, b):
            s = b.state()
            leaf = BoardState(owner=s["owner"], armies=s["armies"], currentPlayer=s["currentPlayer"], currentPhase=s["currentPhase"],
                              turnCount=s["turnCount"], playerNumberOfCards=s["playerNumberOfCards"],
                              numberOfPlaceableArmies=s["numberOfPlaceableArmies"], hasInvadedACountry=s["hasInvadedACountry"],
                              attackState=s["attackState"], attackingCC=s["attackingCC"], defendingCC=s["defendingCC"],
                              defendingPlayer=s["defendingPlayer"], attackUntilDead=True, cardProgressionPos=b.real_pos,
                              hand=b.real_hand, deck=b.real_deck).pack(classic)
            _, games = oracle.rollout_batch(leaf[None, :], cores, 4711, first_game=game[0], per_game=True, threads=8)
            game[0] += cores
            return [([int(x) for x in g["player_out_on_turn"][:6]], int(g["sim_turn_count"])) for g in games]

    rng = random.Random(5)
    root, root_moves = pymcts.run_mcts(board, lambda: board.setup(bs), Rollout(), 0, 6, iters, rng)
    tree.close()
    # same root move list, in the same order
    assert [(int(m["type"]), int(m["a"]), int(m["b"]), int(m["flag"])) for m in ch["move"]] == [(m[0], m[2], m[3], m[7]) for m in root_moves]
    n1, n2 = int(ch["visits"].sum()), root.visits
    assert n1 == n2 == iters * cores
    p1, p2 = ch["wins"][:, 0].sum() / n1, root.wins[0] / n2
    sigma = np.sqrt(p1 * (1 - p1) / n1 + p2 * (1 - p2) / n2)
    assert abs(p1 - p2) < 4.5 * sigma + 0.02, (p1, p2, sigma)
    # both searches expand every root move first and then spread their visits the same way: at 300 iterations over 17
    # root children the UCT exploration term dominates, so the visit shares stay close to each other
    v1 = ch["visits"] / n1
    v2 = np.zeros(len(ch))
    for c in root.children:
        v2[root_moves.index(c.move)] = c.visits / n2
    assert np.all(v1 > 0) and np.all(v2 > 0)
    assert np.max(np.abs(v1 - v2)) < 0.1, (v1, v2)
