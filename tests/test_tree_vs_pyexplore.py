"""The native leaf producer (C++ in liborx.so) against the pure-Python restatement oracle/pyexplore.py: identical legal
move lists (order, fields, float32 probabilities) and identical successor states along random walks, over the classic
map, random maps, every option set and the pre-game phases.  CPU only."""
import functools

import numpy as np
import pytest

from leafzoo import random_leaves, random_map
from oponn_b200.maps import classic42
from oponn_b200.state import BoardState, Rules
from oponn_b200.tree import MOVE_ATTACK_OUTCOME, SearchTree, TreeOptions, new_game
from oracle import pyexplore, pyoracle


@functools.lru_cache(maxsize=None)
def outcomes(a, d):
    pr, al, dl, inv = pyoracle.get_attack_outcomes(a, d)
    return [(np.float32(p), int(x), int(y), bool(v)) for p, x, y, v in zip(pr, al, dl, inv)]


def tup(m):
    return (int(m["type"]), int(m["phase"]), int(m["a"]), int(m["b"]), int(m["c"]), int(m["d"]),
            np.float32(m["probability"]).tobytes(), int(m["flag"]))


def pytup(m):
    return m[:6] + (np.float32(m[6]).tobytes(), m[7])


def walk(m, rules, opts: TreeOptions, leaf, rng, steps):
    tree = SearchTree(m, rules, opts)
    py = pyexplore.ExplorationBoard(m, rules, outcomes, placement_partial_order=opts.placement_partial_order,
                                    attack_aggressive=opts.attack_aggressive, attack_repeat_moves=opts.attack_repeat_moves,
                                    attack_partial_order=opts.attack_partial_order, attack_until_dead=opts.attack_until_dead,
                                    invasion=opts.invasion, fortification_outwards=opts.fortification_outwards)
    extra = None
    n = m.n_territories
    for step in range(steps):
        bs = BoardState.unpack(m, leaf)
        mvb, ft = (None, None) if extra is None else tree.split_extra(extra)
        try:
            py.setup(bs, None if mvb is None else [int(x) for x in mvb], None if ft is None else [int(x) for x in ft],
                     explore_until_dead=bool(leaf[22]))
        except IndexError:
            # more hidden cards than the deck holds (the real hand is counted twice once the turn passes on a tiny
            # map): the reference's simDeal throws / returns null here; both restatements must refuse the record
            from oponn_b200.engine import OrxError
            with pytest.raises(OrxError):
                tree.getPossibleMoves(leaf, extra)
            break
        try:
            want = py.possible_moves()
        except IndexError:
            # the reference throws here (e.g. possiblePlacements.get(-1) when the mover has no border country,
            # ExplorationBoard.java:427-466): the native producer must refuse the position too
            from oponn_b200.engine import OrxError
            with pytest.raises(OrxError):
                tree.getPossibleMoves(leaf, extra)
            break
        got = tree.getPossibleMoves(leaf, extra)
        assert [tup(g) for g in got] == [pytup(w) for w in want], (step, bs)
        if len(got) == 0:          # no legal move in this position (both restatements agree on that): nowhere to walk on
            break
        if int(got[0]["type"]) == MOVE_ATTACK_OUTCOME:
            p = np.asarray(got["probability"], dtype=np.float64); k = int(rng.choice(len(got), p=p / p.sum()))
        else:
            k = int(rng.integers(len(got)))
        leaf2, extra2, on = tree.updateState(leaf, got[k], extra)
        on_py = py.update(want[k])
        st = py.state()
        s2 = BoardState.unpack(m, leaf2)
        mv2, ft2 = tree.split_extra(extra2)
        assert on == on_py
        assert (s2.owner, s2.armies) == (st["owner"], st["armies"])
        assert (s2.currentPlayer, s2.currentPhase, s2.turnCount) == (st["currentPlayer"], st["currentPhase"], st["turnCount"])
        assert list(s2.playerNumberOfCards) == st["playerNumberOfCards"]
        assert (s2.numberOfPlaceableArmies, bool(s2.hasInvadedACountry), s2.attackState) == \
            (st["numberOfPlaceableArmies"], st["hasInvadedACountry"], st["attackState"])
        assert (s2.attackingCC, s2.defendingCC, s2.defendingPlayer) == (st["attackingCC"], st["defendingCC"], st["defendingPlayer"])
        assert [int(x) for x in mv2] == st["moveable"] and [int(x) for x in ft2] == st["fortified"]
        assert bool(leaf2[22]) == st["explore_until_dead"]
        # the real hand / deck / progression position ride along unchanged (CardManager owns them)
        assert (s2.hand, s2.deck, s2.cardProgressionPos) == (bs.hand, bs.deck, bs.cardProgressionPos)
        leaf, extra = leaf2, extra2
        if not on:
            break
    tree.close()
    return step + 1


def test_classic_mid_game_walks(classic, mid_leaf):
    rng = np.random.default_rng(1)
    total = 0
    for _ in range(6):
        total += walk(classic, Rules(), TreeOptions(), mid_leaf.copy(), rng, 300)
    assert total > 600


def test_new_game_walk_through_initial_placement(classic):
    rng = np.random.default_rng(2)
    assert walk(classic, Rules(), TreeOptions(), new_game(classic, 99), rng, 400) > 100


@pytest.mark.parametrize("opts", [
    TreeOptions(placement_partial_order=False), TreeOptions(attack_aggressive=False), TreeOptions(attack_repeat_moves=False),
    TreeOptions(attack_partial_order=True), TreeOptions(attack_until_dead=False), TreeOptions(invasion=False),
    TreeOptions(fortification_outwards=False),
], ids=["no_ppo", "no_aggr", "no_repeat", "attack_po", "no_until_dead", "no_invasion_opt", "no_outwards"])
def test_every_pruning_switch(classic, mid_leaf, opts):
    rng = np.random.default_rng(3)
    walk(classic, Rules(), opts, mid_leaf.copy(), rng, 200)


def test_random_maps_and_positions():
    rng = np.random.default_rng(4)
    for k in range(8):
        n, p = int(rng.choice([6, 13, 30, 42, 70])), int(rng.integers(2, 7))
        m = random_map(rng, n, p, int(rng.integers(1, 6)))
        rules = Rules(use_cards=bool(rng.integers(2)), transfer_cards=bool(rng.integers(2)),
                      card_progression=str(rng.choice(["4, 6, 8, 10, 15, 20...", "5, 5, 5...", "3, 6, 9..."])))
        for leaf in random_leaves(m, rng, 6):
            # keep armies small enough for the outcome tables to stay cheap
            bs = BoardState.unpack(m, leaf); bs.armies = [min(a, 40) for a in bs.armies]
            walk(m, rules, TreeOptions(), bs.pack(m), rng, 60)
