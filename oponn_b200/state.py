"""Host-side mirror of ``include/orx_state.h``: maps, rules and the packed leaf record.

The leaf record is the flat replacement for the Java object graph that crosses the
rollout seam in the reference: ``BoardState`` (src/com/mld46/oponn/BoardState.java:16-55),
the ``Country[]`` it borrows (:118-139) and ``CardManager``'s real hand / unseen deck
(src/com/mld46/oponn/CardManager.java:20-31,240-248).  Byte layout == ``orx_leaf_layout``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

MAX_TERRITORIES = 252
MAX_PLAYERS = 8
MAX_CONTINENTS = 32
MAX_DEGREE = 32
CARD_WILD = 0xFF
OWNER_NONE = 0xFF

# SimulationBoard.Phase / AttackState ordinals (SimulationBoard.java:19-35)
PHASE_COUNTRY_SELECTION, PHASE_INITIAL_PLACEMENT, PHASE_CARDS, PHASE_PLACEMENT, PHASE_ATTACK, \
    PHASE_FORTIFICATION, PHASE_NEXT_PHASE = range(7)
ATTACK_START, ATTACK_EXECUTE, ATTACK_INVADE, ATTACK_CASH, ATTACK_PLACE = range(5)

FLAG_ERROR, FLAG_STEP_LIMIT, FLAG_STREAM_END = 1, 2, 4

# playout policies (orx_state.h ORX_AGENT_*): the SimAgent subclasses this engine models (SimAgent.java:248-307)
AGENT_RANDOM, AGENT_ANGRY, AGENT_STINKY = range(3)
MODELLING_TURN_LIMIT = 2            # Oponn.java:55


def agent_from_name(name: str) -> int:
    """SimAgent.getSimAgent(name): case-insensitive; agents not modelled here play as SimRandom."""
    return {"angry": AGENT_ANGRY, "stinky": AGENT_STINKY}.get((name or "").lower(), AGENT_RANDOM)

HEADER_BYTES = 32

# OrxLeafHeader.deal_mode (orx_state.h): where CardManager.simDeal's nextInt comes from
DEAL_LEAF_LCG, DEAL_GAME_STREAM = 0, 1


def deck_hash(deck: Sequence[int]) -> int:
    """``orx_deck_hash``: the stand-in for ``simDeck.hashCode()`` (CardManager.java:245) used by hosts without a JVM:
    FNV-1a over the deck's card codes folded to a signed 32-bit value."""
    h = 1469598103934665603
    for c in deck:
        h = ((h ^ (int(c) & 0xFF)) * 1099511628211) & ((1 << 64) - 1)
    v = (h ^ (h >> 32)) & 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v



class OrxMapC(C.Structure):
    _fields_ = [("n_territories", C.c_int32), ("n_continents", C.c_int32), ("n_players", C.c_int32),
                ("reserved", C.c_int32),
                ("adj_offsets", C.POINTER(C.c_int32)), ("adj", C.POINTER(C.c_int32)),
                ("continent", C.POINTER(C.c_int32)), ("continent_bonus", C.POINTER(C.c_int32)),
                ("card_symbol", C.POINTER(C.c_int32))]


class OrxRulesC(C.Structure):
    _fields_ = [("use_cards", C.c_int32), ("transfer_cards", C.c_int32), ("immediate_cash", C.c_int32),
                ("continent_increase", C.c_int32), ("card_progression", C.c_char_p),
                ("max_player_turns", C.c_int32), ("reserved", C.c_int32)]


GAME_RESULT_DTYPE = np.dtype([("player_out_on_turn", "<i4", (MAX_PLAYERS,)), ("sim_turn_count", "<i4"),
                              ("flags", "<u4"), ("stream_pos", "<u4"), ("board_hash", "<u4")])
LEAF_RESULT_DTYPE = np.dtype([("n", "<i8"), ("len_sum", "<i8"), ("wins", "<i8", (MAX_PLAYERS,)),
                              ("win_len_sum", "<i8", (MAX_PLAYERS,)), ("flagged", "<i8"), ("reserved", "<i8")])
assert GAME_RESULT_DTYPE.itemsize == 48 and LEAF_RESULT_DTYPE.itemsize == 160


@dataclass
class RiskMap:
    """The Country graph + board constants (SimulationBoard.java:102-139)."""
    name: str
    n_players: int
    adjacency: List[List[int]]          # Country.getAdjoiningList() order, per territory code
    continent: List[int]                # Country.getContinent()
    continent_bonus: List[int]          # initialContinentBonuses
    territory_names: Optional[List[str]] = None
    card_symbol: Optional[List[int]] = None

    def __post_init__(self):
        n = len(self.adjacency)
        if not (1 <= n <= MAX_TERRITORIES):
            raise ValueError(f"map has {n} territories, supported 1..{MAX_TERRITORIES}")
        if not (2 <= self.n_players <= MAX_PLAYERS):
            raise ValueError("n_players out of range")
        if len(self.continent) != n:
            raise ValueError("continent[] length mismatch")
        if not (1 <= len(self.continent_bonus) <= MAX_CONTINENTS):
            raise ValueError("n_continents out of range")
        for t, nb in enumerate(self.adjacency):
            if len(nb) > MAX_DEGREE:
                raise ValueError(f"territory {t} has degree {len(nb)} > {MAX_DEGREE}")
            for x in nb:
                if not (0 <= x < n) or x == t:
                    raise ValueError(f"bad neighbour {x} of {t}")
        self._off = np.zeros(n + 1, dtype=np.int32)
        self._off[1:] = np.cumsum([len(a) for a in self.adjacency])
        self._adj = np.array([x for a in self.adjacency for x in a], dtype=np.int32)
        self._cont = np.array(self.continent, dtype=np.int32)
        self._bonus = np.array(self.continent_bonus, dtype=np.int32)
        self._sym = None if self.card_symbol is None else np.array(self.card_symbol, dtype=np.int32)

    @property
    def n_territories(self) -> int:
        return len(self.adjacency)

    @property
    def n_continents(self) -> int:
        return len(self.continent_bonus)

    def symbol(self, code: int) -> int:
        return code % 3 if self.card_symbol is None else self.card_symbol[code]

    def to_c(self) -> OrxMapC:
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        return OrxMapC(self.n_territories, self.n_continents, self.n_players, 0, p(self._off), p(self._adj),
                       p(self._cont), p(self._bonus),
                       p(self._sym) if self._sym is not None else C.POINTER(C.c_int32)())


@dataclass
class Rules:
    """Game settings (BoardState.java:22-26) + the engine's step cap."""
    use_cards: bool = True
    transfer_cards: bool = True
    immediate_cash: bool = False
    continent_increase: int = 0
    card_progression: str = "4, 6, 8, 10, 15, 20..."
    max_player_turns: int = 0

    def to_c(self) -> OrxRulesC:
        self._cp = self.card_progression.encode()
        return OrxRulesC(int(self.use_cards), int(self.transfer_cards), int(self.immediate_cash),
                         int(self.continent_increase), self._cp, int(self.max_player_turns), 0)


@dataclass
class LeafLayout:
    off_owner: int
    off_armies: int
    off_ncards: int
    off_cards: int
    stride: int


def leaf_layout(n_territories: int, n_players: int) -> LeafLayout:
    off = HEADER_BYTES
    off_owner = off; off += (n_territories + 3) & ~3
    off_armies = off; off += 4 * n_territories
    off_ncards = off; off += (n_players + 3) & ~3
    off_cards = off; off += (n_territories + 2 + 3) & ~3
    return LeafLayout(off_owner, off_armies, off_ncards, off_cards, (off + 15) & ~15)


@dataclass
class BoardState:
    """The mutable half of the reference's ``BoardState`` (BoardState.java:28-55) plus what
    ``CardManager`` contributes to ``setupState`` (real hand, unseen deck, progression position).
    Field names follow the Java fields."""
    owner: Sequence[int]
    armies: Sequence[int]
    currentPlayer: int = 0
    currentPhase: int = PHASE_CARDS
    turnCount: int = 0
    playerNumberOfCards: Sequence[int] = ()
    numberOfPlaceableArmies: int = 0
    hasInvadedACountry: bool = False
    attackState: int = ATTACK_START
    attackingCC: int = -1
    defendingCC: int = -1
    defendingPlayer: int = -1
    attackUntilDead: bool = True        # the board's sticky bit, SURVEY 8.1 quirk 2
    cardProgressionPos: int = 1
    hand: Sequence[int] = ()            # CardManager.inHand (codes, CARD_WILD for the wildcard)
    deck: Optional[Sequence[int]] = None  # CardManager.notInHand in list order; None = all other cards
    dealSeed: Optional[int] = None      # what simDeck.hashCode() returned (CardManager.java:245); None = deck_hash(deck)
    dealFromStream: bool = False        # True = ORX_DEAL_GAME_STREAM: deals draw from the playout's word stream (orx_state.h)

    def deal_seed(self, deck: Sequence[int]) -> int:
        return deck_hash(deck) if self.dealSeed is None else int(self.dealSeed)

    def pack(self, m: RiskMap) -> np.ndarray:
        n, p = m.n_territories, m.n_players
        L = leaf_layout(n, p)
        if len(self.owner) != n or len(self.armies) != n:
            raise ValueError("owner/armies length mismatch")
        hand = list(self.hand)
        if self.deck is None:
            full = list(range(n)) + [CARD_WILD, CARD_WILD]
            for c in hand:
                full.remove(c)
            deck = full
        else:
            deck = list(self.deck)
        if len(hand) + len(deck) > n + 2:
            raise ValueError("hand + deck exceed the number of cards")
        buf = np.zeros(L.stride, dtype=np.uint8)
        hdr = np.zeros(1, dtype=np.dtype([("turn_count", "<i4"), ("placeable", "<i4"), ("attacking_cc", "<i2"),
                                          ("defending_cc", "<i2"), ("card_pos", "<i2"), ("current_player", "i1"),
                                          ("phase", "i1"), ("attack_state", "i1"), ("has_invaded", "i1"),
                                          ("defending_player", "i1"), ("attack_until_dead", "i1"),
                                          ("n_hand", "u1"), ("n_deck", "u1"), ("tree_attack_until_dead", "u1"),
                                          ("deal_mode", "u1"), ("deal_seed", "<i4"), ("reserved", "u1", (4,))]))
        assert hdr.dtype.itemsize == HEADER_BYTES
        hdr["turn_count"] = self.turnCount
        hdr["placeable"] = self.numberOfPlaceableArmies
        hdr["attacking_cc"] = self.attackingCC
        hdr["defending_cc"] = self.defendingCC
        hdr["card_pos"] = self.cardProgressionPos
        hdr["current_player"] = self.currentPlayer
        hdr["phase"] = self.currentPhase
        hdr["attack_state"] = self.attackState
        hdr["has_invaded"] = int(self.hasInvadedACountry)
        hdr["defending_player"] = self.defendingPlayer
        hdr["attack_until_dead"] = int(self.attackUntilDead)
        hdr["n_hand"] = len(hand)
        hdr["n_deck"] = len(deck)
        hdr["deal_mode"] = DEAL_GAME_STREAM if self.dealFromStream else DEAL_LEAF_LCG
        hdr["deal_seed"] = self.deal_seed(deck)
        buf[:HEADER_BYTES] = hdr.view(np.uint8)
        own = np.array([OWNER_NONE if o < 0 else o for o in self.owner], dtype=np.uint8)
        buf[L.off_owner:L.off_owner + n] = own
        buf[L.off_armies:L.off_armies + 4 * n] = np.array(self.armies, dtype="<i4").view(np.uint8)
        nc = np.zeros(p, dtype=np.uint8)
        pc = list(self.playerNumberOfCards) or [0] * p
        nc[:] = pc
        buf[L.off_ncards:L.off_ncards + p] = nc
        cards = np.array(hand + deck, dtype=np.uint8)
        buf[L.off_cards:L.off_cards + len(cards)] = cards
        return buf

    @staticmethod
    def unpack(m: RiskMap, buf: np.ndarray) -> "BoardState":
        n, p = m.n_territories, m.n_players
        L = leaf_layout(n, p)
        b = np.asarray(buf, dtype=np.uint8)
        i4 = lambda o: int(b[o:o + 4].view("<i4")[0])
        i2 = lambda o: int(b[o:o + 2].view("<i2")[0])
        i1 = lambda o: int(b[o:o + 1].view("i1")[0])
        n_hand, n_deck = int(b[20]), int(b[21])
        owner = [(-1 if x == OWNER_NONE else int(x)) for x in b[L.off_owner:L.off_owner + n]]
        armies = [int(x) for x in b[L.off_armies:L.off_armies + 4 * n].view("<i4")]
        cards = [int(x) for x in b[L.off_cards:L.off_cards + n_hand + n_deck]]
        return BoardState(owner=owner, armies=armies, currentPlayer=i1(14), currentPhase=i1(15), turnCount=i4(0),
                          playerNumberOfCards=[int(x) for x in b[L.off_ncards:L.off_ncards + p]],
                          numberOfPlaceableArmies=i4(4), hasInvadedACountry=bool(i1(17)), attackState=i1(16),
                          attackingCC=i2(8), defendingCC=i2(10), defendingPlayer=i1(18), attackUntilDead=bool(i1(19)),
                          cardProgressionPos=i2(12), hand=cards[:n_hand], deck=cards[n_hand:],
                          dealSeed=i4(24), dealFromStream=int(b[23]) == DEAL_GAME_STREAM)


def pack_leaves(m: RiskMap, states: Sequence[BoardState]) -> np.ndarray:
    """A batch of leaf records, ``stride`` bytes apart (what ``orx_rollout_batch`` takes)."""
    L = leaf_layout(m.n_territories, m.n_players)
    out = np.zeros((len(states), L.stride), dtype=np.uint8)
    for i, s in enumerate(states):
        out[i] = s.pack(m)
    return out
