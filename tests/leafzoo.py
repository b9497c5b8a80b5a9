"""Leaf records exercising every phase / attack sub-state / rule edge of setupState + simulate."""
import numpy as np

from oponn_b200.state import (ATTACK_CASH, ATTACK_EXECUTE, ATTACK_INVADE, ATTACK_PLACE, ATTACK_START, CARD_WILD, OWNER_NONE,
                              PHASE_ATTACK, PHASE_CARDS, PHASE_COUNTRY_SELECTION, PHASE_FORTIFICATION,
                              PHASE_INITIAL_PLACEMENT, PHASE_PLACEMENT, BoardState, leaf_layout)


def _enemy_pair(m, bs, player):
    """(attacker territory with > 1 army owned by player, adjacent enemy territory)"""
    for t in range(m.n_territories):
        if bs.owner[t] == player and bs.armies[t] > 1:
            for x in m.adjacency[t]:
                if bs.owner[x] != player:
                    return t, x
    # give the first border territory some armies
    for t in range(m.n_territories):
        if bs.owner[t] == player:
            for x in m.adjacency[t]:
                if bs.owner[x] != player:
                    bs.armies[t] = 7
                    return t, x
    raise ValueError("player has no border")


def phase_zoo(m, base: np.ndarray = None, seed: int = 3):
    """One leaf per (phase, attack state) the reference can hand to setupState, derived from a mid-game record
    (or a random deal when base is None)."""
    rng = np.random.default_rng(seed)
    n, p = m.n_territories, m.n_players
    if base is None:
        owner = [int(x) for x in rng.permutation(n) % p]
        armies = [int(x) for x in rng.integers(1, 6, size=n)]
        bs0 = BoardState(owner=owner, armies=armies, currentPlayer=1, currentPhase=PHASE_CARDS, turnCount=3,
                         playerNumberOfCards=[2] * p, cardProgressionPos=2, hand=[0, 4, CARD_WILD])
    else:
        bs0 = BoardState.unpack(m, base)
    out, names = [], []

    def add(name, **kw):
        d = dict(owner=list(bs0.owner), armies=list(bs0.armies), currentPlayer=bs0.currentPlayer, turnCount=bs0.turnCount,
                 playerNumberOfCards=list(bs0.playerNumberOfCards), cardProgressionPos=bs0.cardProgressionPos,
                 hand=list(bs0.hand), deck=None if bs0.deck is None else list(bs0.deck))
        d.update(kw)
        if "hand" in kw and "deck" not in kw:
            d["deck"] = None                       # the rest of the cards, in code order
        b = BoardState(**d)
        out.append(b); names.append(name)
        return b

    cp = bs0.currentPlayer
    add("cards", currentPhase=PHASE_CARDS, numberOfPlaceableArmies=0)
    add("cards_carry", currentPhase=PHASE_CARDS, numberOfPlaceableArmies=7)
    add("placement", currentPhase=PHASE_PLACEMENT, numberOfPlaceableArmies=9)
    add("placement_zero", currentPhase=PHASE_PLACEMENT, numberOfPlaceableArmies=0)
    add("attack_start", currentPhase=PHASE_ATTACK, attackState=ATTACK_START)
    add("attack_start_invaded", currentPhase=PHASE_ATTACK, attackState=ATTACK_START, hasInvadedACountry=True)
    b = add("attack_execute", currentPhase=PHASE_ATTACK, attackState=ATTACK_EXECUTE)
    a, d = _enemy_pair(m, b, cp)
    b.attackingCC, b.defendingCC, b.defendingPlayer = a, d, b.owner[d]
    b = add("attack_execute_single_roll", currentPhase=PHASE_ATTACK, attackState=ATTACK_EXECUTE, attackUntilDead=False)
    a, d = _enemy_pair(m, b, cp)
    b.attackingCC, b.defendingCC, b.defendingPlayer = a, d, b.owner[d]
    b = add("attack_execute_weak_defender", currentPhase=PHASE_ATTACK, attackState=ATTACK_EXECUTE)
    a, d = _enemy_pair(m, b, cp)
    b.armies[a], b.armies[d] = 40, 1
    b.attackingCC, b.defendingCC, b.defendingPlayer = a, d, b.owner[d]
    # INVADE: the tree has already moved ownership and zeroed the defender (ExplorationBoard.java:174-177)
    b = add("attack_invade", currentPhase=PHASE_ATTACK, attackState=ATTACK_INVADE, hasInvadedACountry=False)
    a, d = _enemy_pair(m, b, cp)
    b.armies[a] = max(b.armies[a], 6)
    b.attackingCC, b.defendingCC, b.defendingPlayer = a, d, b.owner[d]
    b.owner[d] = cp; b.armies[d] = 1   # armies are overwritten by executeInvasion
    # the same as ExplorationBoard.getBoardState() really leaves it: the conquered country at 0 armies
    b = add("attack_invade_zero_armies", currentPhase=PHASE_ATTACK, attackState=ATTACK_INVADE, hasInvadedACountry=False)
    a, d = _enemy_pair(m, b, cp)
    b.armies[a] = max(b.armies[a], 9)
    b.attackingCC, b.defendingCC, b.defendingPlayer = a, d, b.owner[d]
    b.owner[d] = cp; b.armies[d] = 0
    add("attack_cash", currentPhase=PHASE_ATTACK, attackState=ATTACK_CASH, numberOfPlaceableArmies=0,
        hand=[0, 1, 2, 3, 4, 5])
    add("attack_place", currentPhase=PHASE_ATTACK, attackState=ATTACK_PLACE, numberOfPlaceableArmies=11)
    add("fortification", currentPhase=PHASE_FORTIFICATION)
    add("fortification_last_player", currentPhase=PHASE_FORTIFICATION, currentPlayer=p - 1, hand=[])
    # pre-game phases (SimulationBoard.java:485-568)
    owner = [int(x) for x in rng.permutation(n) % p]
    add("initial_placement", owner=owner, armies=[1] * n, currentPlayer=0, currentPhase=PHASE_INITIAL_PLACEMENT, turnCount=0,
        numberOfPlaceableArmies=4, playerNumberOfCards=[0] * p, hand=[], deck=None, cardProgressionPos=1)
    half = [o if i % 2 == 0 else -1 for i, o in enumerate(owner)]
    add("country_selection", owner=half, armies=[1] * n, currentPlayer=2 % p, currentPhase=PHASE_COUNTRY_SELECTION, turnCount=0,
        playerNumberOfCards=[0] * p, hand=[], deck=None, cardProgressionPos=1)
    add("country_selection_empty", owner=[-1] * n, armies=[1] * n, currentPlayer=0, currentPhase=PHASE_COUNTRY_SELECTION,
        turnCount=0, playerNumberOfCards=[0] * p, hand=[], deck=None, cardProgressionPos=1)
    # big stacks: > 100 and > 500 battle paths, overflow cut-off (SimulationBoard.java:984-1015)
    big = [int(x) for x in rng.integers(90, 700, size=n)]
    add("big_stacks", armies=big, currentPhase=PHASE_ATTACK, attackState=ATTACK_START)
    huge = [int(x) for x in rng.integers(2000, 4000, size=n)]
    add("overflow", armies=huge, currentPhase=PHASE_FORTIFICATION)
    # a big hand (transfer after an elimination leaves > 5 cards: forced cashing, SimulationBoard.java:620-638)
    add("big_hand", currentPhase=PHASE_CARDS, hand=list(range(9)) + [CARD_WILD], playerNumberOfCards=[1] * p)
    leaves = np.stack([b.pack(m) for b in out])
    return leaves, names


def invalid_leaves(m, base: np.ndarray) -> np.ndarray:
    """Records that violate the "Leaf validity" contract of include/orx_state.h: all must come back ORX_FLAG_ERROR."""
    L = leaf_layout(m.n_territories, m.n_players)
    n, p = m.n_territories, m.n_players
    out = []

    def mutate(fn):
        b = base.copy(); fn(b); out.append(b)

    mutate(lambda b: b.__setitem__(14, p))                         # current_player out of range
    mutate(lambda b: b.__setitem__(14, 0xFF))                      # current_player -1
    mutate(lambda b: b.__setitem__(15, 6))                         # phase NEXT_PHASE
    mutate(lambda b: b.__setitem__(16, 5))                         # attack_state out of range
    mutate(lambda b: b.__setitem__(L.off_owner + 3, p))            # owner out of range
    mutate(lambda b: b.__setitem__(L.off_owner + 5, OWNER_NONE))   # unowned territory outside the draft
    mutate(lambda b: b.__setitem__(slice(L.off_armies + 8, L.off_armies + 12), 0))   # zero armies
    mutate(lambda b: b.__setitem__(slice(L.off_armies, L.off_armies + 4), np.array([-5], dtype="<i4").view(np.uint8)))
    mutate(lambda b: b.__setitem__(slice(L.off_armies, L.off_armies + 4), np.array([(1 << 20) + 1], dtype="<i4").view(np.uint8)))
    mutate(lambda b: b.__setitem__(L.off_cards + 1, n))            # card code out of range
    mutate(lambda b: b.__setitem__(20, 40))                        # n_hand + n_deck > N + 2
    mutate(lambda b: b.__setitem__(slice(8, 10), np.array([n], dtype="<i2").view(np.uint8)))     # attacking_cc == N
    mutate(lambda b: b.__setitem__(slice(10, 12), np.array([-2], dtype="<i2").view(np.uint8)))   # defending_cc < -1
    mutate(lambda b: b.__setitem__(slice(12, 14), np.array([-1], dtype="<i2").view(np.uint8)))   # card_pos < 0
    mutate(lambda b: b.__setitem__(18, p))                         # defending_player out of range
    mutate(lambda b: b.__setitem__(slice(4, 8), np.array([(1 << 20) + 1], dtype="<i4").view(np.uint8)))  # placeable too big
    # dynamic errors the reference would throw on
    def execute_without_countries(b):                                # ATTACK/EXECUTE with attackingCC == -1
        b[15] = PHASE_ATTACK; b[16] = ATTACK_EXECUTE
    mutate(execute_without_countries)
    def invade_without_defender(b):                                  # ATTACK/INVADE, defendingPlayer == -1
        b[15] = PHASE_ATTACK; b[16] = ATTACK_INVADE
    mutate(invade_without_defender)
    def too_many_hidden_cards(b):                                    # more hidden cards than the deck holds
        b[L.off_ncards:L.off_ncards + p] = 40
    mutate(too_many_hidden_cards)
    return np.stack(out)


def random_map(rng, n_territories: int, n_players: int, n_continents: int):
    """a connected random graph with bounded degree, random continents (possibly empty ones) and bonuses"""
    from oponn_b200.state import RiskMap
    n = n_territories
    nb = [set() for _ in range(n)]
    order = rng.permutation(n)
    for i in range(1, n):                       # random spanning tree
        a = int(order[i]); b = int(order[rng.integers(i)])
        if len(nb[a]) < 12 and len(nb[b]) < 12:
            nb[a].add(b); nb[b].add(a)
        else:
            c = next(int(x) for x in order[:i] if len(nb[int(x)]) < 12)
            nb[a].add(c); nb[c].add(a)
    for _ in range(int(n * rng.uniform(0.3, 1.5))):  # extra borders
        a, b = int(rng.integers(n)), int(rng.integers(n))
        if a != b and len(nb[a]) < 12 and len(nb[b]) < 12:
            nb[a].add(b); nb[b].add(a)
    adjacency = []
    for t in range(n):
        lst = sorted(nb[t]); rng.shuffle(lst); adjacency.append([int(x) for x in lst])
    cont = [int(x) for x in rng.integers(n_continents, size=n)]
    bonus = [int(x) for x in rng.integers(0, 8, size=n_continents)]
    sym = [int(x) for x in rng.integers(3, size=n)] if rng.random() < 0.5 else None
    return RiskMap(name=f"rand{n}", n_players=n_players, adjacency=adjacency, continent=cont, continent_bonus=bonus, card_symbol=sym)


def random_leaves(m, rng, count: int) -> np.ndarray:
    """random legal mid-game records in the four turn phases (every player may or may not be alive)"""
    n, p = m.n_territories, m.n_players
    out = []
    for _ in range(count):
        alive = [int(x) for x in rng.permutation(p)[:int(rng.integers(2, p + 1))]]
        owner = [alive[int(rng.integers(len(alive)))] for _ in range(n)]
        cp = owner[int(rng.integers(n))]
        scale = int(rng.choice([3, 10, 60, 400, 2000]))
        armies = [int(x) for x in rng.integers(1, scale + 1, size=n)]
        cards = list(range(n)) + [CARD_WILD, CARD_WILD]
        rng.shuffle(cards)
        nh = int(rng.integers(0, 7))
        hand, rest = [int(c) for c in cards[:nh]], [int(c) for c in cards[nh:]]
        ncards = [int(rng.integers(0, 5)) if q in alive and q != cp else 0 for q in range(p)]
        while sum(ncards) > len(rest):
            ncards[int(np.argmax(ncards))] -= 1
        phase = int(rng.choice([PHASE_CARDS, PHASE_PLACEMENT, PHASE_ATTACK, PHASE_FORTIFICATION]))
        b = BoardState(owner=owner, armies=armies, currentPlayer=cp, currentPhase=phase, turnCount=int(rng.integers(0, 30)),
                       playerNumberOfCards=ncards, numberOfPlaceableArmies=int(rng.integers(0, 40)) if phase != PHASE_ATTACK else 0,
                       hasInvadedACountry=bool(rng.integers(2)), attackState=ATTACK_START, cardProgressionPos=int(rng.integers(0, 30)),
                       hand=hand, deck=rest)
        out.append(b.pack(m))
    return np.stack(out)
