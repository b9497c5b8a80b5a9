"""pysim -- a second, independent restatement of the rollout path in pure Python.  TEST INFRASTRUCTURE ONLY.

Written from the Java sources, NOT from oracle/oponn_oracle.c: its purpose is to catch a misreading that the C oracle
and the CUDA kernels (same author) would share.  tests/test_pysim.py compares per-game records of the two on several
hundred playouts over the phase zoo, rule variants and random maps.  Java semantics are kept literally (java.util
Lists with remove(Object) / remove(int), int arithmetic, (int) casts, Math.round, the order of every draw).

Restated (O/ = /root/reference/src/com/mld46/oponn/):
  O/sim/boards/SimulationBoard.java:178-359 setupState, :365-481 simulate, :485-1015 the rules, :1076-1111 income / next player
  O/sim/boards/ModellingBoard.java, O/sim/agents/SimAngry.java, SimStinky.java, MapHelper.java:132-191 (class ModellingBoard below)
  O/sim/agents/SimRandom.java (whole), O/sim/agents/SimAgent.java:122-131
  O/BattleSim.java:66-162 (the samplers; outcome tables are taken from the caller -- they have their own independent
                           restatement in tests/battlesim_ref.py)
  O/CardManager.java:240-346 (the simulation half), java.util.Random (JDK documentation)
Engine-level conventions that are NOT the reference's (include/orx_state.h): the leaf-record validity rules, the step
cap, the "first fault only" flags, the stream position of the terminal move and the board hash.

Declared assumptions (Lux SDK not vendored), same as everywhere in this repository: Card.isASet / containsASet /
getRandomSet (wildcard, three equal or three distinct symbols; uniform over valid triples in lexicographic order).
"""
from __future__ import annotations

import math
import struct
from typing import Callable, List, Optional, Sequence, Tuple

M32 = 0xFFFFFFFF
MASK48 = (1 << 48) - 1
WILD = 0xFF
(COUNTRY_SELECTION, INITIAL_PLACEMENT, CARDS, PLACEMENT, ATTACK, FORTIFICATION, NEXT_PHASE) = range(7)
(A_START, A_EXECUTE, A_INVADE, A_CASH, A_PLACE) = range(5)
FLAG_ERROR, FLAG_STEP_LIMIT, FLAG_STREAM_END = 1, 2, 4
ATTACKERS_DESTROYED, DEFENDERS_DESTROYED, ATTACK_INCONCLUSIVE = 13, 7, 0
MAX_ARMIES = 1 << 20


#: StrictMath.log.  fdlibm's __ieee754_log is third-party arithmetic (pinned against real Java outputs in
#: tests/test_oracle_stream.py); tests/test_pysim.py plugs the pinned restatement in here.  The platform log differs from
#: it in the last bit on rare inputs, which only matters to (int)(gaussian * sqrt(n)).
STRICT_LOG = math.log


class JavaError(Exception):
    """the Java code would have thrown (or looped forever): the playout is flagged ORX_FLAG_ERROR"""


def i32(x: int) -> int:
    x &= M32
    return x - (1 << 32) if x & 0x80000000 else x


def java_int_cast(d: float) -> int:
    """(int) d : truncate toward zero, saturate, NaN -> 0"""
    if d != d:
        return 0
    if d >= 2147483647.0:
        return 2147483647
    if d <= -2147483648.0:
        return -2147483648
    return int(d)


# ---- Philox4x32-10 (Salmon et al., SC'11) and the word stream of include/orx_state.h ------------------------------------
def philox4x32_10(ctr: Sequence[int], key: Sequence[int]) -> Tuple[int, int, int, int]:
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


class Stream:
    """word i = Philox(ctr = {i / 4, 0, game_lo, game_hi}, key = {seed_lo, seed_hi})[i % 4], or explicit words"""

    def __init__(self, seed: int = 0, game: int = 0, words: Optional[Sequence[int]] = None):
        self.key = (seed & M32, (seed >> 32) & M32)
        self.game = (game & M32, (game >> 32) & M32)
        self.words = None if words is None else [int(w) for w in words]
        self.pos = 0
        self.cur = -1
        self.out = (0, 0, 0, 0)
        self.ended = False

    def word(self) -> int:
        i = self.pos
        self.pos += 1
        if self.words is not None:
            if i >= len(self.words):
                self.ended = True
                raise StreamEnd()
            return self.words[i]
        if i >> 2 != self.cur:
            self.cur = i >> 2
            self.out = philox4x32_10((self.cur & M32, 0, self.game[0], self.game[1]), self.key)
        return self.out[i & 3]


class StreamEnd(Exception):
    pass


class InjectedRandom:
    """java.util.Random's derivations over next(bits) = word >>> (32 - bits) (JDK documentation of nextInt(bound),
    nextDouble, nextGaussian)"""

    def __init__(self, stream: Stream):
        self.s = stream
        self.have_gauss = False
        self.next_gauss = 0.0

    def next(self, bits: int) -> int:
        return self.s.word() >> (32 - bits)

    def nextInt(self, bound: int) -> int:
        if bound <= 0:
            raise JavaError("nextInt: bound must be positive")
        r = self.next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if i32(u - r + m) >= 0:
                return r
            u = self.next(31)

    def nextDouble(self) -> float:
        return float((self.next(26) << 27) + self.next(27)) * (1.0 / (1 << 53))

    def nextGaussian(self) -> float:
        if self.have_gauss:
            self.have_gauss = False
            return self.next_gauss
        while True:
            v1 = 2 * self.nextDouble() - 1
            v2 = 2 * self.nextDouble() - 1
            s = v1 * v1 + v2 * v2
            if s < 1 and s != 0:
                break
        multiplier = math.sqrt(-2 * STRICT_LOG(s) / s)   # StrictMath.sqrt is correctly rounded; StrictMath.log is fdlibm's
        self.next_gauss = v2 * multiplier
        self.have_gauss = True
        return v1 * multiplier


class LcgRandom:
    """java.util.Random itself: CardManager.random after `new Random(simDeck.hashCode())` (O/CardManager.java:245)"""

    def __init__(self, seed: int):
        self.seed = (seed ^ 0x5DEECE66D) & MASK48

    def next(self, bits: int) -> int:
        self.seed = (self.seed * 0x5DEECE66D + 0xB) & MASK48
        return self.seed >> (48 - bits)

    def nextInt(self, bound: int) -> int:
        r = self.next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if i32(u - r + m) >= 0:
                return r
            u = self.next(31)


# ---- CardManager.getCardCashValue (O/CardManager.java:283-341) ---------------------------------------------------------------
def card_cash_value(cp: str, pos: int) -> int:
    try:
        if "%" in cp:
            if cp[0] == "5":
                base, offset = 5 * pos, 6
            else:
                base, offset = (2 * (pos + 1) if pos <= 4 else 5 * (pos - 2)), 8
            exp = 1 + int(cp[cp.index("^") + 1:len(cp) - 1]) / 100.0
            return java_int_cast(math.floor(base if base <= 25 else base * math.pow(exp, pos - offset)))
        if cp == "0":
            return 0
        if cp == "5, 5, 5...":
            return 5
        if cp == "4, 4, 6, 6, 6, 8, 8, 8, 8, 10...":
            return java_int_cast(math.ceil((math.sqrt(8 * pos + 9) - 1) / 2) * 2)
        if cp == "4, 5, 6...":
            return pos + 3
        if cp == "4, 6, 8...":
            return (pos + 1) * 2
        if cp == "3, 6, 9...":
            return pos * 3
        if cp == "4, 6, 8, 10, 15, 20...":
            return (pos + 1) * 2 if pos <= 4 else (pos - 2) * 5
        if cp == "5, 10, 15...":
            return pos * 5
        if cp == "4, 6, 8, 10, 15, 20, 25, 10, 10, 10...":
            return (pos + 1) * 2 if pos <= 4 else ((pos - 2) * 5 if pos <= 7 else 10)
        raise ValueError
    except Exception:
        return 2


def orx_mix(t: int, owner: int, armies: int) -> int:
    x = ((armies * 2654435761) & M32) ^ (((owner + 1) * 0x85EBCA77) & M32) ^ (((t + 1) * 0xC2B2AE3D) & M32)
    x ^= x >> 15
    x = (x * 0x2C1B3C6D) & M32
    x ^= x >> 12
    return x


# BattleSim.outcomes (O/BattleSim.java:36-44): float sums widened to double
def _f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


_SINGLE = [(_f32(0.4167), 1.0), (_f32(0.2546), 1.0), (_f32(0.5787), 1.0), (_f32(0.2276), _f32(_f32(0.2276) + _f32(0.3241))),
           (_f32(0.6597), 1.0), (_f32(0.3717), _f32(_f32(0.3717) + _f32(0.3358)))]


class SimulationBoard:
    """One board: setupState(leaf) then simulate().  `outcomes(a, d)` returns BattleSim.getAttackOutcomes(a, d) as a list
    of (probability as a Python float holding the float32 value, attackerLosses, defenderLosses)."""

    def __init__(self, risk_map, rules, outcomes: Callable[[int, int], List[Tuple[float, int, int]]]):
        self.m, self.r, self.outcomes = risk_map, rules, outcomes
        self.N, self.P, self.C = risk_map.n_territories, risk_map.n_players, risk_map.n_continents
        self.adj = risk_map.adjacency
        self.continent = risk_map.continent
        self.base_bonus = list(risk_map.continent_bonus)
        self.max_turns = rules.max_player_turns if rules.max_player_turns > 0 else 1 << 16
        self.can_run_out = self.P * 5 > self.N + 2                       # O/CardManager.java:65

    # ---- card sets (declared assumption) ---------------------------------------------------------------------------
    def symbol(self, card: int) -> int:
        return 3 if card == WILD else self.m.symbol(card)

    def is_a_set(self, c1: int, c2: int, c3: int) -> bool:
        s = (self.symbol(c1), self.symbol(c2), self.symbol(c3))
        if 3 in s:
            return True
        return s[0] == s[1] == s[2] or len(set(s)) == 3

    def valid_triples(self, hand: List[int]):
        n = len(hand)
        return [(i, j, k) for i in range(n) for j in range(i + 1, n) for k in range(j + 1, n) if self.is_a_set(hand[i], hand[j], hand[k])]

    def get_random_set(self, hand: List[int]):
        t = self.valid_triples(hand)
        if not t:
            raise JavaError("Card.getRandomSet returned null")
        i, j, k = t[self.rnd.nextInt(len(t))]
        return hand[i], hand[j], hand[k]

    # ---- CardManager, simulation half --------------------------------------------------------------------------------
    def sim_deal(self) -> Optional[int]:
        if len(self.deck) == 0:
            if self.can_run_out:
                return None
            raise JavaError("Invalidly run out of cards!")
        gen = self.rnd if self.deal_from_stream else self.card_random
        return self.deck.pop(gen.nextInt(len(self.deck)))

    # ---- setupState (:178-359) -----------------------------------------------------------------------------------------
    def setup_state(self, bs, stream: Stream):
        N, P = self.N, self.P
        self.rnd = InjectedRandom(stream)
        self.stream = stream
        self.flags = 0
        self.term_pos = None
        self.player_turns = 0
        # engine-level record validation (include/orx_state.h "Leaf validity")
        ok = (0 <= bs.currentPlayer < P and 0 <= bs.currentPhase <= FORTIFICATION and A_START <= bs.attackState <= A_PLACE
              and -1 <= bs.attackingCC < N and -1 <= bs.defendingCC < N and -1 <= bs.defendingPlayer < P
              and bs.cardProgressionPos >= 0 and len(bs.hand) + len(bs.deck) <= N + 2
              and -MAX_ARMIES <= bs.numberOfPlaceableArmies <= MAX_ARMIES)
        for t in range(N):
            o = bs.owner[t]
            if not (0 <= o < P or (o == -1 and bs.currentPhase == COUNTRY_SELECTION)):
                ok = False
            lo = 0 if (bs.currentPhase == ATTACK and bs.attackState == A_INVADE and t == bs.defendingCC) else 1
            if not (lo <= bs.armies[t] <= MAX_ARMIES):
                ok = False
        for c in list(bs.hand) + list(bs.deck):
            if not (0 <= c < N or c == WILD):
                ok = False
        if not ok:
            raise JavaError("invalid leaf record")

        self.sim_start_turn = bs.turnCount
        self.cp = bs.currentPlayer
        self.phase = bs.currentPhase
        self.attack_until_dead = bool(bs.attackUntilDead)               # the board's sticky bit
        self.owner = list(bs.owner)                                    # setupCountries (:192-203)
        self.armies = list(bs.armies)
        # setupCardState (:205-227) with CardManager.simReset (:240-248)
        self.deck = list(bs.deck)
        self.card_random = LcgRandom(bs.deal_seed(bs.deck))
        self.deal_from_stream = bool(bs.dealFromStream)
        self.card_pos = bs.cardProgressionPos
        self.hands = [[] for _ in range(P)]
        for player in range(P):
            if player == self.cp:
                self.hands[player] = list(bs.hand)
            else:
                for _ in range(bs.playerNumberOfCards[player]):
                    card = self.sim_deal()
                    if card is None:
                        raise JavaError("null card dealt into a hidden hand")
                    self.hands[player].append(card)
        # setupGeneralGameState (:229-276)
        self.remaining = 0
        self.sim_turn = 0
        self.out_on_turn = [0] * P
        self.count = [0] * P
        self.missing = [[0] * self.C for _ in range(P)]
        for cc in range(N):
            owner, cont = self.owner[cc], self.continent[cc]
            for p in range(P):
                if p != owner:
                    self.missing[p][cont] += 1
            if not (self.phase == COUNTRY_SELECTION and owner == -1):
                self.count[owner] += 1
        if self.phase == COUNTRY_SELECTION:
            self.remaining = P
        else:
            self.remaining = sum(1 for p in range(P) if self.count[p] > 0)
        # setupPhaseGameState (:278-359)
        self.placeable = 0
        self.has_invaded = False
        self.acc = self.dcc = self.ap = self.dp = -1
        self.astate = A_START
        if self.phase == COUNTRY_SELECTION:
            self.left = self.starting_armies()
            self.rem_countries = 0
            for cc in range(N):
                if self.owner[cc] == -1:
                    self.rem_countries += 1
                else:
                    self.left[self.owner[cc]] -= 1
        elif self.phase == INITIAL_PLACEMENT:
            tot = [0] * P
            for cc in range(N):
                tot[self.owner[cc]] += self.armies[cc]
            self.left = self.starting_armies()
            for i in range(P):
                self.left[i] -= tot[i]
            self.placeable = bs.numberOfPlaceableArmies
        elif self.phase in (CARDS, PLACEMENT):
            self.placeable = bs.numberOfPlaceableArmies
        elif self.phase == ATTACK:
            self.has_invaded = bool(bs.hasInvadedACountry)
            self.astate = bs.attackState
            self.acc, self.dcc = bs.attackingCC, bs.defendingCC
            if self.acc != -1:
                self.ap = self.owner[self.acc]
                self.dp = bs.defendingPlayer
            if bs.attackState == A_INVADE:
                if not (0 <= self.dp < P):
                    raise JavaError("playerCountries[-1]")
                if self.count[self.dp] == 0:
                    self.remaining += 1
            elif bs.attackState in (A_CASH, A_PLACE):
                self.placeable = bs.numberOfPlaceableArmies
        self.recalculate_bonuses()

    def starting_armies(self) -> List[int]:
        f1 = self.N / 42.0
        f1 -= 1.0
        f1 /= 2.0
        f1 += 1.0
        base = int(math.floor(f1 * (50 - self.P * 5) + 0.5))         # Math.round
        return [base + (i if i > 1 else 0) for i in range(self.P)]

    def recalculate_bonuses(self):
        e = max(self.sim_start_turn + self.sim_turn - 1, 0)
        if self.r.continent_increase != 0 and e >= 1024:
            raise JavaError("engine limit: ORX_BONUS_ROUNDS")
        factor = math.pow(1 + self.r.continent_increase / 100.0, e)
        self.bonus = [java_int_cast(math.floor(b * factor)) for b in self.base_bonus]

    # ---- the rules ---------------------------------------------------------------------------------------------------------
    def player_income(self, player: int) -> int:
        if self.count[player] == 0:
            return 0
        income = max(self.count[self.cp] // 3, 3)
        for i in range(self.C):
            if self.missing[self.cp][i] == 0:
                income = i32(income + self.bonus[i])
        return income

    def player_defeated(self, player: int, conquerer: int):
        self.remaining -= 1
        self.out_on_turn[player] = self.sim_turn
        if self.r.transfer_cards:
            if conquerer < 0:
                raise JavaError("playerCards.get(-1)")
            self.hands[conquerer].extend(list(self.hands[player]))   # addAll; the loser's list is left as it is
        else:
            self.deck.extend(self.hands[player])                     # simReturnToDeck
        if self.remaining == 1:
            self.out_on_turn[self.cp] = -1
            if self.term_pos is None:
                self.term_pos = self.stream.pos

    def cash_cards(self, c1: int, c2: int, c3: int) -> bool:
        if self.phase != CARDS and (self.phase != ATTACK or self.astate != A_CASH):
            return False
        if not self.is_a_set(c1, c2, c3):
            return False
        hand = self.hands[self.cp]
        if not (c1 in hand and c2 in hand and c3 in hand):
            return False
        hand.remove(c1); hand.remove(c2); hand.remove(c3)            # List.remove(Object): the first equal element
        self.deck.extend([c1, c2, c3])                               # CardManager.simCash
        if "%" in self.r.card_progression and self.card_pos >= 1024:
            raise JavaError("engine limit: ORX_CARD_VALUES")
        self.placeable += card_cash_value(self.r.card_progression, self.card_pos)
        self.card_pos += 1
        for card in (c1, c2, c3):
            if card != WILD and self.owner[card] == self.cp:
                self.armies[card] += 2
        return True

    def place_armies(self, n: int, cc: int):
        if self.phase not in (PLACEMENT, INITIAL_PLACEMENT) and (self.phase != ATTACK or self.astate != A_PLACE):
            raise JavaError("cannot place in this phase")
        if n > self.placeable or self.owner[cc] != self.cp or n < 0:
            raise JavaError("illegal placement")
        self.armies[cc] += n
        self.placeable -= n
        if self.phase == INITIAL_PLACEMENT:
            self.left[self.cp] -= n

    # BattleSim.simulateIndividualBattle (:152-162) / simulateBattle (:66-145)
    def individual_battle(self, attackers: int, defenders: int) -> int:
        if attackers > 3 or defenders > 2 or attackers < 1 or defenders < 1:
            raise JavaError("simulateIndividualBattle arguments")
        r = self.rnd.nextDouble()
        d = _SINGLE[(attackers - 1) * 2 + (defenders - 1)]
        return 0 if r <= d[0] else (1 if r <= d[1] else 2)

    def simulate_battle(self, attackers: int, defenders: int) -> Tuple[int, int]:
        al = dl = 0
        if attackers > 500 and defenders > 500:
            matched = java_int_cast(0.922 * defenders)
            if matched > attackers:
                al = attackers
                variance = java_int_cast(self.rnd.nextGaussian() * math.sqrt(attackers))
                dl = java_int_cast(min(variance + attackers / 0.922, float(defenders - 1)))
            else:
                dl = defenders
                variance = java_int_cast(self.rnd.nextGaussian() * math.sqrt(defenders))
                al = min(i32(variance + matched), attackers - 1)
        else:
            ra, rd = attackers, defenders
            if ra > 100 or rd > 100:
                while (ra > 100 or rd > 100) and not (ra == 0 or rd == 0):
                    pa, pd = min(ra, 3), min(rd, 2)
                    a_loss = self.individual_battle(pa, pd)
                    d_loss = min(pa, pd) - a_loss
                    al += a_loss; dl += d_loss; ra -= a_loss; rd -= d_loss
            if ra != 0 and rd != 0:
                if ra < 0 or rd < 0 or ra > 100 or rd > 100:
                    raise JavaError("generateProbabilities would throw")
                r = self.rnd.nextDouble()
                total = 0.0
                for prob, oal, odl in self.outcomes(ra, rd):
                    total += prob
                    if r <= total:
                        al += oal; dl += odl
                        break
        return al, dl

    def execute_attack(self):
        if self.acc < 0 or self.dcc < 0:
            raise JavaError("countries[-1]")
        a = self.armies[self.acc] - 1
        d = self.armies[self.dcc]
        if self.attack_until_dead:
            self.a_losses, self.d_losses = self.simulate_battle(a, d)
        else:
            pa, pd = min(a, 3), min(d, 2)
            self.a_losses = self.individual_battle(pa, pd)
            self.d_losses = min(pa, pd) - self.a_losses

    def finish_attack(self):
        self.armies[self.acc] -= self.a_losses
        self.armies[self.dcc] -= self.d_losses
        self.astate = A_INVADE if self.armies[self.dcc] == 0 else A_START

    def setup_invasion(self):
        cont = self.continent[self.dcc]
        self.ap, self.dp = self.owner[self.acc], self.owner[self.dcc]
        self.count[self.ap] += 1; self.missing[self.ap][cont] -= 1
        self.count[self.dp] -= 1; self.missing[self.dp][cont] += 1
        self.owner[self.dcc] = self.ap

    def execute_invasion(self, armies: int):
        a = self.armies[self.acc]
        lo, hi = min(a - 1, 3), a - 1
        inv = min(max(armies, lo), hi)
        self.armies[self.dcc] = inv
        self.armies[self.acc] = a - inv

    def finish_invasion(self):
        if not (0 <= self.dp < self.P):
            raise JavaError("playerCountries[-1]")
        if self.count[self.dp] == 0:
            self.player_defeated(self.dp, self.ap)
        self.has_invaded = True
        self.astate = A_START
        self.acc = self.dcc = self.ap = self.dp = -1

    def attack(self, acc: int, dcc: int, until_dead: bool) -> int:      # SimulationBoard.attack (:713-744)
        self.a_losses = self.d_losses = 0
        self.acc, self.dcc = acc, dcc
        self.ap, self.dp = self.owner[acc], self.owner[dcc]
        self.attack_until_dead = until_dead
        self.astate = A_EXECUTE
        self.execute_attack()
        self.finish_attack()
        if self.astate == A_INVADE:
            self.setup_invasion()
            self.execute_invasion(self.agent_move_armies_in(acc, dcc))
            self.finish_invasion()
            return DEFENDERS_DESTROYED
        return ATTACKERS_DESTROYED if self.armies[acc] == 1 else ATTACK_INCONCLUSIVE

    def check_for_overflow(self):
        totals = [0] * self.P
        for cc in range(self.N):
            if self.owner[cc] < 0:
                raise JavaError("armyTotals[-1]")
            totals[self.owner[cc]] += self.armies[cc]
        total = 0; max_armies = 0; max_player = -1
        for i in range(self.P):
            total += totals[i]
            if totals[i] > max_armies:
                max_player, max_armies = i, totals[i]
        if total > 100000:
            for i in range(self.P):
                if self.count[i] > 0 and i != max_player:
                    self.player_defeated(i, max_player)

    def tear_down_fortification(self):
        nxt = self.cp
        guard = 0
        while True:
            nxt = (nxt + 1) % self.P
            guard += 1
            if guard > self.P:
                raise JavaError("getNextPlayer would spin forever")
            if self.count[nxt] != 0:
                break
        if nxt < self.cp:
            self.sim_turn += 1
            self.recalculate_bonuses()
        self.cp = nxt
        self.phase = CARDS
        self.check_for_overflow()
        self.player_turns += 1

    def tear_down_initial_placement(self):
        nxt = (self.cp + 1) % self.P
        while nxt != self.cp and self.left[nxt] == 0:
            nxt = (nxt + 1) % self.P
        if nxt == self.cp:
            self.cp = 0
            self.phase = CARDS
            self.sim_turn += 1
        else:
            self.cp = nxt
        self.player_turns += 1

    def tear_down_cards(self):
        hand = self.hands[self.cp]
        if len(hand) > 4:
            while len(self.hands[self.cp]) > 4:
                c = self.get_random_set(list(self.hands[self.cp]))
                self.cash_cards(*c)
        self.phase = PLACEMENT

    def tear_down_attack(self):
        if self.has_invaded and self.r.use_cards:
            card = self.sim_deal()
            if card is not None:
                self.hands[self.cp].append(card)
        self.phase = FORTIFICATION

    # ---- SimRandom ------------------------------------------------------------------------------------------------------------
    def agent_pick_country(self) -> int:
        cand = [cc for cc in range(self.N) if self.owner[cc] == -1]
        return cand[self.rnd.nextInt(len(cand))]

    def agent_cards_phase(self):
        hand = list(self.hands[self.cp])
        if len(hand) < 3 or not self.valid_triples(hand):
            return
        r = self.rnd.nextDouble()
        if (len(hand) == 3 and r < 0.5) or (len(hand) == 4 and r < 0.75) or len(hand) >= 5:
            self.cash_cards(*self.get_random_set(hand))

    def agent_place(self, n: int):
        me = self.cp
        cand = [c for c in range(self.N) if self.owner[c] == me and any(self.owner[x] != me for x in self.adj[c])]
        while n > 0:
            armies = 1 + self.rnd.nextInt(n)
            country = cand[self.rnd.nextInt(len(cand))]
            self.place_armies(armies, country)
            n -= armies

    def agent_move_armies_in(self, a: int, d: int) -> int:
        available = self.armies[a] - 1
        if available == 0:
            raise JavaError("zero available attackers")
        mn = min(available, 3)
        return mn + (0 if mn == available else self.rnd.nextInt(available - mn))

    def agent_attack_phase(self):
        me = self.cp
        attackers = [c for c in range(self.N) if self.owner[c] == me and self.armies[c] > 1 and any(self.owner[x] != me for x in self.adj[c])]
        while len(attackers) > 0:
            ai = self.rnd.nextInt(len(attackers))
            attacker = attackers[ai]
            defenders = [x for x in self.adj[attacker] if self.owner[x] != me]
            if not defenders:
                attackers.pop(ai)
                continue
            defender = defenders[self.rnd.nextInt(len(defenders))]
            result = self.attack(attacker, defender, True)
            if result == ATTACKERS_DESTROYED:
                attackers.pop(ai)
            elif result == DEFENDERS_DESTROYED:
                if self.armies[attacker] == 1:
                    attackers.pop(ai)
                if self.armies[defender] > 1:
                    attackers.append(defender)

    def agent_fortify_phase(self):                                    # SimRandom.fortifyPhase: empty
        pass

    def setup_fortification_phase(self):                              # (:900-906) only observable through getMoveableArmies: ModellingBoard below
        pass

    # ---- simulate (:365-481) -----------------------------------------------------------------------------------------------
    def step_limit_hit(self):
        if self.player_turns >= self.max_turns:
            self.flags |= FLAG_STEP_LIMIT
            raise StepLimit()

    def simulate(self):
        ph = self.phase
        if ph == COUNTRY_SELECTION:
            if self.rem_countries == 0:
                self.phase = INITIAL_PLACEMENT; self.cp = 0
        elif ph == INITIAL_PLACEMENT:
            if self.placeable > 0:
                self.agent_place(self.placeable)
            self.tear_down_initial_placement()
        elif ph == CARDS:
            self.agent_cards_phase()
            self.tear_down_cards()
        elif ph == PLACEMENT:
            if self.placeable > 0:
                self.agent_place(self.placeable)
            self.phase = ATTACK
        elif ph == ATTACK:
            if self.astate == A_EXECUTE:
                self.execute_attack()
                self.finish_attack()
            if self.astate == A_INVADE:
                if self.acc < 0 or self.dcc < 0:
                    raise JavaError("countries[-1]")
                self.execute_invasion(self.agent_move_armies_in(self.acc, self.dcc))
                self.finish_invasion()
            if self.astate == A_CASH:
                self.placeable = 0
                self.agent_cards_phase()
                self.astate = A_PLACE
            if self.astate == A_PLACE:
                self.agent_place(self.placeable)
                self.astate = A_START
            self.agent_attack_phase()
            self.tear_down_attack()
        elif ph == FORTIFICATION:
            self.agent_fortify_phase()                                # (:425: no setupFortificationPhase on the catch-up path)
            self.tear_down_fortification()
        else:
            raise JavaError("NEXT_PHASE")
        while self.remaining > 1:
            ph = self.phase
            if ph == COUNTRY_SELECTION:
                cc = self.agent_pick_country()
                if cc < 0 or self.owner[cc] != -1:
                    raise JavaError("already owned")
                self.owner[cc] = self.cp
                self.count[self.cp] += 1; self.left[self.cp] -= 1; self.rem_countries -= 1
                self.missing[self.cp][self.continent[cc]] -= 1
                self.cp = (self.cp + 1) % self.P
                if self.rem_countries == 0:
                    self.phase = INITIAL_PLACEMENT; self.cp = 0
            elif ph == INITIAL_PLACEMENT:
                a = self.left[self.cp]
                self.placeable = 5 if a == 5 else min(4, a)
                self.agent_place(self.placeable)
                self.tear_down_initial_placement()
                self.step_limit_hit()
            elif ph == CARDS:
                self.placeable = 0
                self.agent_cards_phase()
                self.tear_down_cards()
            elif ph == PLACEMENT:
                self.placeable = i32(self.placeable + self.player_income(self.cp))
                self.agent_place(self.placeable)
                self.phase = ATTACK
            elif ph == ATTACK:
                self.has_invaded = False
                self.acc = self.dcc = self.ap = self.dp = -1
                self.astate = A_START
                self.agent_attack_phase()
                self.tear_down_attack()
            elif ph == FORTIFICATION:
                self.setup_fortification_phase()
                self.agent_fortify_phase()
                self.tear_down_fortification()
                self.step_limit_hit()
            else:
                raise JavaError("NEXT_PHASE")


class StepLimit(Exception):
    pass


AGENT_RANDOM, AGENT_ANGRY, AGENT_STINKY = 0, 1, 2
MAX_REDRAWS = 65536      # engine limit on SimStinky.placeArmies' rejection loop (include/orx_state.h ORX_MAX_REDRAWS)


class ModellingBoard(SimulationBoard):
    """O/sim/boards/ModellingBoard.java + the modelled policies O/sim/agents/SimAngry.java and SimStinky.java, written from
    the Java.  SDK semantics (Country.getWeakestEnemyNeighbor / getNumberEnemyNeighbors / getNumberPlayerNeighbors, the
    one-ahead CountryIterators, BoardHelper.getPlayerArmies) are the DECLARED ASSUMPTIONS of include/orx_state.h; engine
    conventions of the same header: the agents' private `rand` objects draw from the playout's stream, a FORTIFICATION leaf
    starts with setupFortificationPhase's moveable armies, SimStinky's redraw loop is capped."""

    def __init__(self, risk_map, rules, outcomes, sim_agents: Sequence[int], modelling_turn_limit: int):
        super().__init__(risk_map, rules, outcomes)
        self.initial_agents = list(sim_agents)
        self.limit = modelling_turn_limit

    # ModellingBoard.setupGeneralGameState (:21-27): every setupState starts from the modelled line-up again
    def setup_state(self, bs, stream: Stream):
        self.agents = list(self.initial_agents)
        super().setup_state(bs, stream)
        fort = self.phase == FORTIFICATION
        self.moveable = [self.armies[c] if fort and self.owner[c] == self.cp else 0 for c in range(self.N)]

    def agent(self) -> int:
        return self.agents[self.cp]

    # ModellingBoard.tearDownFortificationPhase (:29-46), then SimulationBoard's (:965-982: moveable armies back to 0 first)
    def tear_down_fortification(self):
        nxt = self.cp
        for _ in range(self.P + 1):
            nxt = (nxt + 1) % self.P
            if self.count[nxt] != 0:
                break
        else:
            raise JavaError("getNextPlayer would spin forever")
        if nxt < self.cp and self.sim_turn + 1 == self.limit:
            self.agents = [AGENT_RANDOM] * self.P
        self.moveable = [0] * self.N
        super().tear_down_fortification()

    def setup_fortification_phase(self):
        self.moveable = [self.armies[c] if self.owner[c] == self.cp else 0 for c in range(self.N)]

    # SimulationBoard.fortifyArmies (:908-954)
    def fortify_armies(self, n: int, src: int, dst: int) -> int:
        source_armies, moveable = self.armies[src], self.moveable[src]
        if self.phase != FORTIFICATION:
            raise JavaError("cannot fortify in this phase")
        if self.owner[src] != self.cp or self.owner[dst] != self.cp:
            return -1
        if n < 0:
            return -1
        k = min(n, min(moveable, source_armies - 1))
        if k == 0:
            return 0
        self.armies[src] = source_armies - k
        self.moveable[src] = moveable - k
        self.armies[dst] += k
        return 1

    # ---- Lux SDK (declared assumptions) ----------------------------------------------------------------------------------
    def enemy_neighbors(self, cc: int) -> int:
        return sum(1 for x in self.adj[cc] if self.owner[x] != self.owner[cc])

    def player_neighbors(self, cc: int, player: int) -> int:
        return sum(1 for x in self.adj[cc] if self.owner[x] == player)

    def weakest_enemy_neighbor(self, cc: int) -> Optional[int]:
        best = None
        for x in self.adj[cc]:
            if self.owner[x] != self.owner[cc] and (best is None or self.armies[x] < self.armies[best]):
                best = x
        return best

    def country_iterator(self, player: int, armies: int = 0):
        """PlayerIterator (armies == 0) / ArmiesIterator: code order, ONE element held ahead -- the element after `us` is looked
        for when `us` is handed out, on the board as it is at that moment"""
        def matches(cc):
            if self.owner[cc] != player:
                return False
            if armies > 0:
                return self.armies[cc] >= armies
            if armies < 0:
                return self.armies[cc] <= -armies
            return True

        def find(start):
            for cc in range(start, self.N):
                if matches(cc):
                    return cc
            return None

        held = find(0)
        while held is not None:
            us = held
            held = find(us + 1)
            yield us

    # ---- MapHelper (:132-191) -----------------------------------------------------------------------------------------------
    def smallest_positive_continent(self, need_all: bool) -> int:
        smallest_size, smallest = 2 ** 31 - 1, -1
        for cont in range(self.C):
            members = [cc for cc in range(self.N) if self.continent[cc] == cont]
            free = [cc for cc in members if self.owner[cc] == -1]
            ok = (len(free) == len(members)) if need_all else (len(free) > 0)
            if len(members) < smallest_size and ok and self.bonus[cont] > 0:
                smallest_size, smallest = len(members), cont
        return smallest

    # ---- SimAngry (SimAngry.java) ----------------------------------------------------------------------------------------------
    def angry_pick_country(self) -> int:
        me = self.cp
        goal = self.smallest_positive_continent(True)
        if goal == -1:
            goal = self.smallest_positive_continent(False)
        members = [cc for cc in range(self.N) if self.continent[cc] == goal]      # ContinentIterator: code order
        for cc in members:
            if self.owner[cc] == -1 and self.player_neighbors(cc, me) > 0:
                return cc
        best, fewest = -1, 1000000
        for cc in members:
            if self.owner[cc] == -1 and len(self.adj[cc]) < fewest:
                best, fewest = cc, len(self.adj[cc])
        return best

    def angry_place(self, n: int):
        most, place_on = -1, None
        for us in self.country_iterator(self.cp):
            sub = self.enemy_neighbors(us)
            if sub > most:
                most, place_on = sub, us
        if place_on is None:
            raise JavaError("placeOn.getCode() on null")
        self.place_armies(n, place_on)

    def angry_attack_phase(self):
        made_attack = True
        while made_attack:
            made_attack = False
            for us in self.country_iterator(self.cp, 2):
                weakest = self.weakest_enemy_neighbor(us)
                if weakest is not None and self.armies[us] > self.armies[weakest]:
                    self.attack(us, weakest, True)
                    made_attack = True

    def angry_move_armies_in(self, cca: int, ccd: int) -> int:
        if self.enemy_neighbors(cca) > self.enemy_neighbors(ccd):
            return 0
        return self.armies[cca] - 1

    def angry_fortify_phase(self):
        me = self.cp
        for i in range(self.N):
            if self.owner[i] == me and self.moveable[i] > 0:
                neighbors = self.adj[i]
                best_prospect, best_enemies = -1, 0
                for x in neighbors:
                    if self.owner[x] == me:
                        e = self.enemy_neighbors(x)
                        if e > best_enemies:
                            best_prospect, best_enemies = x, e
                e = self.enemy_neighbors(i)
                if best_enemies > e:
                    self.fortify_armies(self.moveable[i], i, best_prospect)
                elif e == 0:
                    rand_cc = self.rnd.nextInt(len(neighbors))
                    self.fortify_armies(self.moveable[i], i, neighbors[rand_cc])

    # ---- SimStinky (SimStinky.java) -----------------------------------------------------------------------------------------------
    def stinky_place(self, n: int):
        tries = 0
        while True:
            if tries >= MAX_REDRAWS:
                self.flags |= FLAG_STEP_LIMIT
                raise StepLimit()
            tries += 1
            test = self.rnd.nextInt(self.N)
            if not (self.owner[test] != self.cp or self.weakest_enemy_neighbor(test) is None):
                break
        self.place_armies(n, test)

    def stinky_attack_pass(self, care_about_odds: bool) -> bool:
        attacked = False
        for us in self.country_iterator(self.cp):
            weak = self.weakest_enemy_neighbor(us)
            if weak is not None and self.armies[us] > 1 and (self.armies[us] > self.armies[weak] * 1.5 or not care_about_odds):
                self.attack(us, weak, True)
                attacked = True
        return attacked

    def stinky_attack_phase(self):
        self.stinky_attack_pass(True)
        if sum(self.armies[cc] for cc in range(self.N) if self.owner[cc] == self.cp) > 300:
            while self.stinky_attack_pass(False):
                pass

    def stinky_move_armies_in(self, cca: int, ccd: int) -> int:
        attack_weak, defend_weak = self.weakest_enemy_neighbor(cca), self.weakest_enemy_neighbor(ccd)
        if attack_weak is None:
            return 1000000
        if defend_weak is None:
            return 0
        if self.armies[attack_weak] < self.armies[defend_weak]:
            return 0
        return 1000000

    # ---- simAgents[currentPlayer].xxx() ----------------------------------------------------------------------------------------
    def agent_pick_country(self) -> int:
        a = self.agent()
        if a == AGENT_ANGRY:
            return self.angry_pick_country()
        if a == AGENT_STINKY:
            return -1
        return super().agent_pick_country()

    def agent_cards_phase(self):
        if self.agent() == AGENT_RANDOM:
            super().agent_cards_phase()

    def agent_place(self, n: int):
        a = self.agent()
        if a == AGENT_ANGRY:
            self.angry_place(n)
        elif a == AGENT_STINKY:
            self.stinky_place(n)
        else:
            super().agent_place(n)

    def agent_move_armies_in(self, a: int, d: int) -> int:
        ag = self.agent()
        if ag == AGENT_ANGRY:
            return self.angry_move_armies_in(a, d)
        if ag == AGENT_STINKY:
            return self.stinky_move_armies_in(a, d)
        return super().agent_move_armies_in(a, d)

    def agent_attack_phase(self):
        a = self.agent()
        if a == AGENT_ANGRY:
            self.angry_attack_phase()
        elif a == AGENT_STINKY:
            self.stinky_attack_phase()
        else:
            super().agent_attack_phase()

    def agent_fortify_phase(self):
        if self.agent() == AGENT_ANGRY:
            self.angry_fortify_phase()


def play(risk_map, rules, bs, outcomes, seed: int = 0, game: int = 0, words: Optional[Sequence[int]] = None,
         sim_agents: Optional[Sequence[int]] = None, modelling_turn_limit: int = 2) -> dict:
    """one playout -> the fields of OrxGameResult"""
    b = SimulationBoard(risk_map, rules, outcomes) if sim_agents is None else ModellingBoard(risk_map, rules, outcomes, sim_agents, modelling_turn_limit)
    stream = Stream(seed, game, words)
    flags = 0
    try:
        b.setup_state(bs, stream)
        b.simulate()
    except JavaError:
        flags = FLAG_ERROR
    except (IndexError, ValueError):
        flags = FLAG_ERROR                 # ArrayIndexOutOfBounds / IllegalArgument analogues
    except StreamEnd:
        flags = FLAG_STREAM_END
    except StepLimit:
        flags = FLAG_STEP_LIMIT
    if flags & (FLAG_ERROR | FLAG_STREAM_END):
        return dict(player_out_on_turn=[0] * 8, sim_turn_count=0, flags=flags, stream_pos=0, board_hash=0)
    pot = list(b.out_on_turn) + [0] * (8 - b.P)
    h = sum(orx_mix(t, b.owner[t] & 0xFF, b.armies[t]) for t in range(b.N)) & M32
    return dict(player_out_on_turn=pot, sim_turn_count=b.sim_turn, flags=flags,
                stream_pos=b.term_pos if b.term_pos is not None else stream.pos, board_hash=h)
