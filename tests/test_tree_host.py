"""The leaf producer (ExplorationBoard restatement in liborx.so's host code) on hand-checked positions.  CPU only:
move generation and application need no GPU.  Expectations are derived by hand from
src/com/mld46/oponn/sim/boards/ExplorationBoard.java (cited per test)."""
import numpy as np
import pytest

from battlesim_ref import numpy_outcome_table
from oponn_b200.maps import classic42
from oponn_b200.state import (ATTACK_EXECUTE, ATTACK_INVADE, ATTACK_START, CARD_WILD, PHASE_ATTACK, PHASE_CARDS,
                              PHASE_COUNTRY_SELECTION, PHASE_FORTIFICATION, PHASE_INITIAL_PLACEMENT, PHASE_NEXT_PHASE,
                              PHASE_PLACEMENT, BoardState, RiskMap, Rules)
from oponn_b200.tree import (MOVE_ATTACK, MOVE_ATTACK_OUTCOME, MOVE_CARD_CASH, MOVE_COUNTRY_SELECTION, MOVE_FORTIFICATION,
                             MOVE_INITIAL_PLACEMENT, MOVE_INVASION, MOVE_NEXT_PHASE, MOVE_PLACEMENT, SearchTree, TreeOptions)


def line5(p=2):
    # 0 - 1 - 2 - 3 - 4
    return RiskMap(name="line5", n_players=p, adjacency=[[1], [0, 2], [1, 3], [2, 4], [3]], continent=[0, 0, 1, 1, 1],
                   continent_bonus=[2, 3])


def tup(m):
    return (int(m["type"]), int(m["a"]), int(m["b"]), int(m["c"]), int(m["d"]), int(m["flag"]))


def moves(tree, bs, extra=None):
    return [tup(m) for m in tree.getPossibleMoves(bs, extra)]


@pytest.fixture
def tree():
    t = SearchTree(line5(), Rules())
    yield t
    t.close()


def test_card_moves(tree):
    # getCardMoves (ExplorationBoard.java:333-356): best set (if any) + NextPhase while fewer than 5 cards
    bs = BoardState(owner=[0, 0, 1, 1, 1], armies=[1, 3, 2, 2, 2], currentPlayer=0, currentPhase=PHASE_CARDS,
                    playerNumberOfCards=[0, 0], hand=[0, 3, CARD_WILD])
    assert moves(tree, bs) == [(MOVE_CARD_CASH, 0, 3, CARD_WILD, 0, 0), (MOVE_NEXT_PHASE, PHASE_PLACEMENT, PHASE_CARDS, 0, 0, 0)]
    bs.hand = [0, 1, 3]                      # symbols 0, 1, 0: not a set
    assert moves(tree, bs) == [(MOVE_NEXT_PHASE, PHASE_PLACEMENT, PHASE_CARDS, 0, 0, 0)]
    bs.hand = [0, 1, 2, 3, 4]                # five cards: must cash (no NextPhase); declared getBestSet: most owned territories
    assert moves(tree, bs) == [(MOVE_CARD_CASH, 0, 1, 2, 0, 0)]


def test_cash_then_placement_income(tree):
    bs = BoardState(owner=[0, 0, 1, 1, 1], armies=[1, 3, 2, 2, 2], currentPlayer=0, currentPhase=PHASE_CARDS,
                    playerNumberOfCards=[0, 0], hand=[0, 3, CARD_WILD], cardProgressionPos=2)
    mv = tree.getPossibleMoves(bs)
    leaf, extra, on = tree.updateState(bs, mv[0])           # cashCards (SimulationBoard.java:577-618)
    s = BoardState.unpack(tree.map, leaf)
    assert on and s.armies == [3, 3, 2, 2, 2] and s.numberOfPlaceableArmies == 6 and s.currentPhase == PHASE_CARDS
    nxt = tree.getPossibleMoves(leaf, extra)
    leaf, extra, on = tree.updateState(leaf, nxt[-1], extra)  # NextPhase -> PLACEMENT: += max(2/3, 3) + bonus 2 (continent 0)
    s = BoardState.unpack(tree.map, leaf)
    assert s.currentPhase == PHASE_PLACEMENT and s.numberOfPlaceableArmies == 6 + 3 + 2


def test_placement_partial_order_and_sparse_indices():
    # getPlacementMoves (:429-481): every candidate but the last gets sparseIndices(n); the last gets "all n"
    m = RiskMap(name="tri", n_players=2, adjacency=[[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]], continent=[0, 0, 0, 0],
                continent_bonus=[1])
    t = SearchTree(m, Rules())
    bs = BoardState(owner=[0, 0, 0, 1], armies=[1, 1, 1, 5], currentPlayer=0, currentPhase=PHASE_PLACEMENT,
                    numberOfPlaceableArmies=3, playerNumberOfCards=[0, 0])
    assert moves(t, bs) == [(MOVE_PLACEMENT, 0, a, 0, 0, 0) for a in (1, 2, 3)] + [(MOVE_PLACEMENT, 1, a, 0, 0, 0) for a in (1, 2, 3)] + \
        [(MOVE_PLACEMENT, 2, 3, 0, 0, 0)]
    bs.numberOfPlaceableArmies = 40        # sparseIndices(40): inc = (39 / 9) = 4 (integer division) -> 1, 5, ..., 37
    got = moves(t, bs)
    assert [g[2] for g in got[:10]] == [1 + 4 * i for i in range(10)] and got[-1] == (MOVE_PLACEMENT, 2, 40, 0, 0, 0)
    # minPlacementCC is board-internal (not part of BoardState): a fresh setupState starts the order again (:85-91),
    # exactly as each runMCTS call of the reference does
    leaf, extra, _ = t.updateState(bs, t.getPossibleMoves(bs)[10])
    assert {g[1] for g in moves(t, leaf, extra)} == {0, 1, 2}
    assert BoardState.unpack(m, leaf).numberOfPlaceableArmies == 39 and BoardState.unpack(m, leaf).armies[1] == 2
    t2 = SearchTree(m, Rules(), TreeOptions(placement_partial_order=False))
    bs.numberOfPlaceableArmies = 2
    assert moves(t2, bs) == [(MOVE_PLACEMENT, c, a, 0, 0, 0) for c in (0, 1, 2) for a in (1, 2)]
    t.close(); t2.close()


def test_attack_start_moves(tree):
    # getAttackStartMoves (:497-557): single-roll and (attackers > 2) till-dead attack per legal pair; an attack with one
    # free attacker against 2+ defenders is "unfavourable" and dropped; two players left + aggressive: NextPhase only
    # when no attack has 3+ free attackers
    bs = BoardState(owner=[0, 0, 1, 1, 1], armies=[1, 5, 2, 2, 2], currentPlayer=0, currentPhase=PHASE_ATTACK,
                    attackState=ATTACK_START, playerNumberOfCards=[0, 0])
    assert moves(tree, bs) == [(MOVE_ATTACK, 1, 2, 5, 2, 0), (MOVE_ATTACK, 1, 2, 5, 2, 1)]
    bs.armies = [1, 3, 2, 2, 2]              # 2 free attackers: not favourable -> may also end the phase
    assert moves(tree, bs) == [(MOVE_ATTACK, 1, 2, 3, 2, 0), (MOVE_ATTACK, 1, 2, 3, 2, 1),
                               (MOVE_NEXT_PHASE, PHASE_FORTIFICATION, PHASE_ATTACK, 0, 0, 0)]
    bs.armies = [1, 2, 2, 2, 2]              # 1 free attacker v 2: unfavourable
    assert moves(tree, bs) == [(MOVE_NEXT_PHASE, PHASE_FORTIFICATION, PHASE_ATTACK, 0, 0, 0)]
    bs.armies = [1, 2, 1, 2, 2]              # 1 v 1 is allowed, single roll only (attackers == 2)
    assert moves(tree, bs)[0] == (MOVE_ATTACK, 1, 2, 2, 1, 0)


def test_attack_outcomes_and_invasion(tree):
    bs = BoardState(owner=[0, 0, 1, 1, 1], armies=[1, 5, 2, 2, 2], currentPlayer=0, currentPhase=PHASE_ATTACK,
                    attackState=ATTACK_START, playerNumberOfCards=[0, 0])
    mv = tree.getPossibleMoves(bs)
    # till-dead attack -> EXECUTE: the chance outcomes are BattleSim.getAttackOutcomes(4, 2) in table order (:563-569)
    leaf, extra, _ = tree.updateState(bs, mv[1])
    s = BoardState.unpack(tree.map, leaf)
    assert (s.attackState, s.attackingCC, s.defendingCC, s.defendingPlayer) == (ATTACK_EXECUTE, 1, 2, 1)
    out = tree.getPossibleMoves(leaf, extra)
    want = numpy_outcome_table(4, 2)
    assert [(int(o["a"]), int(o["b"]), bool(o["flag"])) for o in out] == [(x, y, bool(v)) for _, x, y, v in want]
    assert [np.float32(o["probability"]).tobytes() for o in out] == [np.float32(p).tobytes() for p, *_ in want]
    # single-roll attack: 3 v 2 dice rows (:571-607)
    leaf1, extra1, _ = tree.updateState(bs, mv[0])
    out1 = tree.getPossibleMoves(leaf1, extra1)
    assert [(int(o["a"]), int(o["b"]), int(o["flag"])) for o in out1] == [(0, 2, 1), (1, 1, 0), (2, 0, 0)]
    assert [round(float(o["probability"]), 4) for o in out1] == [0.3717, 0.3358, 0.2926]
    # conquering outcome (0, 2): finishAttack + setupInvasion (:167-177): owner moves now, armies wait for the Invasion move
    win = [o for o in out if int(o["a"]) == 0 and int(o["b"]) == 2][0]
    leaf2, extra2, on = tree.updateState(leaf, win, extra)
    s2 = BoardState.unpack(tree.map, leaf2)
    assert on and s2.owner == [0, 0, 0, 1, 1] and s2.attackState == ATTACK_INVADE
    inv = moves(tree, leaf2, extra2)
    # attacker's hostile neighbours: none (0 and 2 are ours now) -> contained in the defender's -> only the largest move (:615-652)
    assert inv == [(MOVE_INVASION, 1, 2, 4, 0, 0)]
    leaf3, extra3, on = tree.updateState(leaf2, tree.getPossibleMoves(leaf2, extra2)[0], extra2)
    s3 = BoardState.unpack(tree.map, leaf3)
    assert s3.armies[:3] == [1, 1, 4] and s3.attackState == ATTACK_START and s3.hasInvadedACountry
    # the attackRepeatMoves bookkeeping and the next attack options from the conquered territory
    assert (MOVE_ATTACK, 2, 3, 4, 2, 1) in moves(tree, leaf3, extra3)


def test_invasion_sparse_moves_when_attacker_keeps_a_front():
    # 1 has two hostile neighbours (0 and 2); after taking 2, territory 0 is not adjacent to 2 -> all invasion sizes offered
    m = line5()
    t = SearchTree(m, Rules())
    bs = BoardState(owner=[1, 0, 0, 1, 1], armies=[2, 9, 1, 2, 2], currentPlayer=0, currentPhase=PHASE_ATTACK,
                    attackState=ATTACK_INVADE, attackingCC=1, defendingCC=2, defendingPlayer=1, playerNumberOfCards=[0, 0])
    got = moves(t, bs)
    assert got == [(MOVE_INVASION, 1, 2, k, 0, 0) for k in range(3, 9)]      # min(8, 3) .. 8
    t.close()


def test_elimination_ends_the_search_path(tree):
    bs = BoardState(owner=[0, 0, 0, 0, 0], armies=[1, 1, 1, 6, 1], currentPlayer=0, currentPhase=PHASE_ATTACK,
                    attackState=ATTACK_INVADE, attackingCC=3, defendingCC=4, defendingPlayer=1, playerNumberOfCards=[0, 2])
    mv = tree.getPossibleMoves(bs)
    leaf, extra, on = tree.updateState(bs, mv[0])
    assert not on                                   # updateState returns remainingPlayers != 1 (:282)


def test_fortification_moves(tree):
    # getFortificationMoves (:660-706) with the outwards optimisation: weights = BFS distance from the border (:868-902)
    bs = BoardState(owner=[0, 0, 0, 1, 1], armies=[4, 1, 2, 2, 2], currentPlayer=0, currentPhase=PHASE_FORTIFICATION,
                    playerNumberOfCards=[0, 0])
    # first fortifiable country is 0 (4 armies, friendly neighbour 1); weights: 2 -> 0, 1 -> 1, 0 -> 2: moving 0 -> 1 goes outwards
    assert moves(tree, bs) == [(MOVE_FORTIFICATION, a, 0, 1, 0, 0) for a in (1, 2, 3)] + [(MOVE_FORTIFICATION, -1, 0, -1, 0, 0)]
    mv = tree.getPossibleMoves(bs)
    leaf, extra, _ = tree.updateState(bs, mv[2])    # 3 armies 0 -> 1
    s = BoardState.unpack(tree.map, leaf)
    assert s.armies == [1, 4, 2, 2, 2]
    mvb, ft = tree.split_extra(extra)
    assert list(mvb) == [1, 1, 2, 0, 0]              # moveable of the source dropped (fortifyArmies :908-954)
    # territory 1 now holds 4 armies but only 1 may move (the 3 that arrived cannot): 1 -> 2 outwards, or 1 -> 0 is inwards (dropped)
    assert moves(tree, leaf, extra) == [(MOVE_FORTIFICATION, 1, 1, 2, 0, 0), (MOVE_FORTIFICATION, -1, 1, -1, 0, 0)]
    leaf, extra, _ = tree.updateState(leaf, tree.getPossibleMoves(leaf, extra)[-1], extra)   # "done with 1"
    # country 2 sits on the border (weight 0); its only friendly neighbour 1 is further in (weight 1): inwards, dropped
    assert moves(tree, leaf, extra) == [(MOVE_FORTIFICATION, -1, 2, -1, 0, 0)]
    # nothing left to fortify -> NextPhase(CARDS) for the next living player
    leaf, extra, _ = tree.updateState(leaf, tree.getPossibleMoves(leaf, extra)[-1], extra)
    assert moves(tree, leaf, extra) == [(MOVE_NEXT_PHASE, PHASE_CARDS, PHASE_FORTIFICATION, 1, 0, 0)]
    leaf, extra, _ = tree.updateState(leaf, tree.getPossibleMoves(leaf, extra)[0], extra)
    s = BoardState.unpack(tree.map, leaf)
    assert (s.currentPlayer, s.currentPhase, s.turnCount) == (1, PHASE_CARDS, 0)


def test_pre_game_phases():
    m = line5()
    t = SearchTree(m, Rules())
    bs = BoardState(owner=[0, -1, 1, -1, 0], armies=[1] * 5, currentPlayer=1, currentPhase=PHASE_COUNTRY_SELECTION,
                    playerNumberOfCards=[0, 0])
    assert moves(t, bs) == [(MOVE_COUNTRY_SELECTION, 1, 0, 0, 0, 0), (MOVE_COUNTRY_SELECTION, 3, 0, 0, 0, 0)]
    leaf, extra, _ = t.updateState(bs, t.getPossibleMoves(bs)[0])
    s = BoardState.unpack(m, leaf)
    assert s.owner == [0, 1, 1, -1, 0] and s.currentPlayer == 0
    t.close()


def test_random_walk_produces_legal_leaves(classic, mid_leaf, oracle):
    """every state the leaf producer emits must be a leaf the rollout accepts (oracle: flags == 0) and play to an end"""
    t = SearchTree(classic, Rules())
    rng = np.random.default_rng(5)
    leaf, extra = mid_leaf.copy(), None
    visited = []
    for step in range(400):
        mv = t.getPossibleMoves(leaf, extra)
        assert len(mv) > 0
        # chance nodes by probability, decisions uniformly
        if int(mv[0]["type"]) == MOVE_ATTACK_OUTCOME:
            p = np.asarray(mv["probability"], dtype=np.float64); k = int(rng.choice(len(mv), p=p / p.sum()))
        else:
            k = int(rng.integers(len(mv)))
        leaf, extra, on = t.updateState(leaf, mv[k], extra)
        visited.append(leaf.copy())
        if not on:
            break
    t.close()
    _, games = oracle.rollout_batch(np.stack(visited), 1, 77, per_game=True, threads=8)
    assert np.all(games["flags"] == 0)
    assert np.all((games["player_out_on_turn"][:, :6] == -1).sum(axis=1) == 1)


def test_new_game_setup(classic, oracle):
    """setupNewGame (SimulationBoard.java:141-176): a shuffled round-robin deal, one army each, player 0 about to place"""
    from javarandom import JavaRandom
    from oponn_b200.tree import new_game
    leaf = new_game(classic, 12345)
    s = BoardState.unpack(classic, leaf)
    # Collections.shuffle(list, rnd): for i = size .. 2 swap(i - 1, rnd.nextInt(i))
    r = JavaRandom(12345); ccs = list(range(42))
    for i in range(42, 1, -1):
        j = r.nextInt(i); ccs[i - 1], ccs[j] = ccs[j], ccs[i - 1]
    want = [0] * 42
    for k, cc in enumerate(ccs):
        want[cc] = k % 6
    assert s.owner == want and s.armies == [1] * 42
    assert (s.currentPlayer, s.currentPhase, s.turnCount, s.numberOfPlaceableArmies) == (0, PHASE_INITIAL_PLACEMENT, 0, 4)
    assert len(s.hand) == 0 and len(s.deck) == 44 and s.deck.count(CARD_WILD) == 2
    # the record is a legal rollout start and a legal tree root
    _, games = oracle.rollout_batch(leaf[None, :], 4, 3, per_game=True)
    assert np.all(games["flags"] == 0) and np.all((games["player_out_on_turn"][:, :6] == -1).sum(axis=1) == 1)
    t = SearchTree(classic, Rules())
    assert all(int(m["type"]) == MOVE_INITIAL_PLACEMENT for m in t.getPossibleMoves(leaf))
    t.close()
