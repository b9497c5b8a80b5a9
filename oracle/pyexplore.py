"""Pure-Python restatement of Oponn's ExplorationBoard (the leaf producer).  TEST INFRASTRUCTURE ONLY.

Checker for the native leaf producer (oponn_b200/csrc/orx_tree.cpp behind include/orx_tree.h): only tests/ may import
it.  Written from the Java source, independently of the C++:
    O/sim/boards/ExplorationBoard.java  (O/ = /root/reference/src/com/mld46/oponn/)
        setupPhaseGameState :77-114, updateState :116-283, getPossibleMoves :289-706, sparseIndices :755-779,
        getBoardState :816-835, FortificationAnalysis :838-902
    rule primitives it inherits: O/sim/boards/SimulationBoard.java :485-1015, O/CardManager.java :240-346
    chance outcomes: O/BattleSim.java :174-286 (numpy float32 sweep of tests/battlesim_ref.py is passed in)

PARITY PIN STATUS: the reference holds no test or fixture for ExplorationBoard => "parity unpinned"; this file and
the C++ share the DECLARED ASSUMPTIONS about the un-vendored Lux SDK listed at the top of orx_tree.cpp
(getNumberEnemyNeighbors, getFriendlyAdjoiningCodeList, Card.getBestSet, the hidden-hand seed).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple

WILD = 0xFF
(COUNTRY_SELECTION, INITIAL_PLACEMENT, CARDS, PLACEMENT, ATTACK, FORTIFICATION, NEXT_PHASE) = range(7)
(A_START, A_EXECUTE, A_INVADE, A_CASH, A_PLACE) = range(5)
(M_COUNTRY_SELECTION, M_INITIAL_PLACEMENT, M_CARD_CASH, M_PLACEMENT, M_ATTACK, M_ATTACK_OUTCOME, M_INVASION, M_FORTIFICATION,
 M_NEXT_PHASE) = range(9)

Move = Tuple  # (type, phase, a, b, c, d, probability, flag)


class JavaRandom:
    MASK = (1 << 48) - 1

    def __init__(self, seed: int):
        self.seed = (seed ^ 0x5DEECE66D) & self.MASK

    def next(self, bits: int) -> int:
        self.seed = (self.seed * 0x5DEECE66D + 0xB) & self.MASK
        return self.seed >> (48 - bits)

    def next_int(self, bound: int) -> int:
        r = self.next(31); m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if u - r + m < (1 << 31):
                return r
            u = self.next(31)


def card_cash_value(cp: str, pos: int) -> int:
    """CardManager.getCardCashValue (O/CardManager.java:283-341), the non-exponential families"""
    table = {"0": lambda p: 0, "5, 5, 5...": lambda p: 5,
             "4, 4, 6, 6, 6, 8, 8, 8, 8, 10...": lambda p: int(math.ceil((math.sqrt(8 * p + 9) - 1) / 2) * 2),
             "4, 5, 6...": lambda p: p + 3, "4, 6, 8...": lambda p: (p + 1) * 2, "3, 6, 9...": lambda p: p * 3,
             "4, 6, 8, 10, 15, 20...": lambda p: (p + 1) * 2 if p <= 4 else (p - 2) * 5, "5, 10, 15...": lambda p: p * 5,
             "4, 6, 8, 10, 15, 20, 25, 10, 10, 10...": lambda p: (p + 1) * 2 if p <= 4 else ((p - 2) * 5 if p <= 7 else 10)}
    return table[cp](pos) if cp in table else 2


class ExplorationBoard:
    def __init__(self, risk_map, rules, outcomes: Callable[[int, int], List[Tuple[float, int, int, bool]]],
                 placement_partial_order=True, attack_aggressive=True, attack_repeat_moves=True, attack_partial_order=False,
                 attack_until_dead=True, invasion=True, fortification_outwards=True):
        self.m, self.r, self.outcomes = risk_map, rules, outcomes
        self.N, self.P, self.C = risk_map.n_territories, risk_map.n_players, risk_map.n_continents
        self.opt = dict(ppo=placement_partial_order, agg=attack_aggressive, rep=attack_repeat_moves, apo=attack_partial_order,
                        aud=attack_until_dead, inv=invasion, fow=fortification_outwards)

    # ---- setupState (SimulationBoard.java:178-359 + ExplorationBoard.java:77-114) from an unpacked BoardState
    def setup(self, bs, moveable: Optional[List[int]] = None, fortified: Optional[List[int]] = None, explore_until_dead=False):
        N, P = self.N, self.P
        self.owner = list(bs.owner); self.armies = list(bs.armies)
        self.cp, self.phase, self.start_turn, self.sim_turn = bs.currentPlayer, bs.currentPhase, bs.turnCount, 0
        if moveable is None:
            moveable = [self.armies[t] if (self.phase == FORTIFICATION and self.owner[t] == self.cp) else 0 for t in range(N)]
        self.moveable = list(moveable); self.fortified = list(fortified) if fortified is not None else [0] * N
        self.real_hand, self.real_deck, self.real_pos = list(bs.hand), list(bs.deck), bs.cardProgressionPos
        self.deck = list(self.real_deck); self.card_pos = self.real_pos
        self.rnd = JavaRandom(bs.deal_seed(self.deck))   # new Random(simDeck.hashCode()) (CardManager.java:245)
        self.cards = [[] for _ in range(P)]
        for p in range(P):
            if p == self.cp:
                self.cards[p] = list(self.real_hand)
            else:
                for _ in range(bs.playerNumberOfCards[p]):
                    self.cards[p].append(self.deck.pop(self.rnd.next_int(len(self.deck))))
        self.count = [0] * P
        self.missing = [[0] * self.C for _ in range(P)]
        for cc in range(N):
            for p in range(P):
                if p != self.owner[cc]:
                    self.missing[p][self.m.continent[cc]] += 1
            if not (self.phase == COUNTRY_SELECTION and self.owner[cc] == -1):
                self.count[self.owner[cc]] += 1
        self.remaining = P if self.phase == COUNTRY_SELECTION else sum(1 for p in range(P) if self.count[p] > 0)
        self.out_on_turn = [0] * P
        self.placeable = 0
        self.acc = self.dcc = self.ap = self.dp = -1; self.astate = A_START; self.invaded = False
        self.min_place = self.min_src = self.min_dst = 0; self.init_inv = False
        self.made = set(); self.until_dead = False; self.al = self.dl = 0
        base = int(math.floor((((N / 42.0) - 1.0) / 2.0 + 1.0) * (50 - P * 5) + 0.5))
        start = [base + (i if i > 1 else 0) for i in range(P)]
        if self.phase == COUNTRY_SELECTION:
            self.left = list(start); self.rem_countries = 0
            for cc in range(N):
                if self.owner[cc] == -1:
                    self.rem_countries += 1
                else:
                    self.left[self.owner[cc]] -= 1
        elif self.phase == INITIAL_PLACEMENT:
            tot = [0] * P
            for cc in range(N):
                tot[self.owner[cc]] += self.armies[cc]
            self.left = [start[i] - tot[i] for i in range(P)]
            self.placeable = bs.numberOfPlaceableArmies
        elif self.phase in (CARDS, PLACEMENT):
            self.placeable = bs.numberOfPlaceableArmies
        elif self.phase == ATTACK:
            self.invaded = bool(bs.hasInvadedACountry); self.astate = bs.attackState
            self.acc, self.dcc = bs.attackingCC, bs.defendingCC
            if self.acc != -1:
                self.ap = self.owner[self.acc]; self.dp = bs.defendingPlayer
            if self.astate == A_INVADE and self.count[self.dp] == 0:
                self.remaining += 1
            elif self.astate in (A_CASH, A_PLACE):
                self.placeable = bs.numberOfPlaceableArmies
            self.init_inv = self.astate == A_INVADE
            self.until_dead = explore_until_dead
        self._bonuses()

    def _bonuses(self):
        e = max(self.start_turn + self.sim_turn - 1, 0)
        f = math.pow(1 + self.r.continent_increase / 100.0, e)
        self.bonus = [int(math.floor(b * f)) for b in self.m.continent_bonus]

    # ---- SDK-style helpers (declared assumptions)
    def _enemy_neighbours(self, cc):
        return sum(1 for x in self.m.adjacency[cc] if self.owner[x] != self.owner[cc])

    def _friendly(self, cc):
        return [x for x in self.m.adjacency[cc] if self.owner[x] == self.owner[cc]]

    def _is_set(self, a, b, c):
        if WILD in (a, b, c):
            return True
        s = {self.m.symbol(a), self.m.symbol(b), self.m.symbol(c)}
        return len(s) != 2

    # ---- rule primitives
    def _player_defeated(self, player, conq):
        self.remaining -= 1; self.out_on_turn[player] = self.sim_turn
        if self.r.transfer_cards:
            self.cards[conq].extend(self.cards[player])
        else:
            self.deck.extend(self.cards[player])
        if self.remaining == 1:
            self.out_on_turn[self.cp] = -1

    def _next_player(self):
        np_ = self.cp
        while True:
            np_ = (np_ + 1) % self.P
            if self.count[np_] != 0:
                return np_

    def _tear_down_fortification(self):
        self.moveable = [0] * self.N
        np_ = self._next_player()
        if np_ < self.cp:
            self.sim_turn += 1; self._bonuses()
        self.cp, self.phase = np_, CARDS
        tot = [0] * self.P
        for cc in range(self.N):
            tot[self.owner[cc]] += self.armies[cc]
        if sum(tot) > 100000:
            mx, mp = 0, -1
            for i in range(self.P):
                if tot[i] > mx:
                    mp, mx = i, tot[i]
            for i in range(self.P):
                if self.count[i] > 0 and i != mp:
                    self._player_defeated(i, mp)

    def _tear_down_initial_placement(self):
        np_ = (self.cp + 1) % self.P
        while np_ != self.cp and self.left[np_] == 0:
            np_ = (np_ + 1) % self.P
        if np_ == self.cp:
            self.cp, self.phase = 0, CARDS; self.sim_turn += 1
        else:
            self.cp = np_

    def _income(self):
        if self.count[self.cp] == 0:
            return 0
        inc = max(self.count[self.cp] // 3, 3)
        return inc + sum(self.bonus[c] for c in range(self.C) if self.missing[self.cp][c] == 0)

    # ---- updateState (ExplorationBoard.java:116-283)
    def update(self, mv: Move) -> bool:
        t, ph, a, b, c, d, prob, flag = mv
        assert ph == self.phase or t == M_NEXT_PHASE
        if t == M_COUNTRY_SELECTION:
            self.owner[a] = self.cp; self.count[self.cp] += 1; self.left[self.cp] -= 1; self.rem_countries -= 1
            self.missing[self.cp][self.m.continent[a]] -= 1
            self.cp = (self.cp + 1) % self.P
        elif t in (M_INITIAL_PLACEMENT, M_PLACEMENT):
            assert b <= self.placeable and self.owner[a] == self.cp and b >= 0
            self.armies[a] += b; self.placeable -= b
            if self.phase == INITIAL_PLACEMENT:
                self.left[self.cp] -= b
            if self.opt["ppo"]:
                self.min_place = a + 1
        elif t == M_CARD_CASH:
            hand = self.cards[self.cp]
            if self._is_set(a, b, c) and all(x in hand for x in (a, b, c)):
                for x in (a, b, c):
                    hand.remove(x)
                self.deck.extend([a, b, c])
                self.placeable += card_cash_value(self.r.card_progression, self.card_pos); self.card_pos += 1
                for x in (a, b, c):
                    if x != WILD and self.owner[x] == self.cp:
                        self.armies[x] += 2
        elif t == M_ATTACK:
            if self.opt["rep"] and self.acc != -1 and self.dcc != -1 and (self.acc != a or self.dcc != b):
                self.made.add((self.acc, self.dcc))
            self.min_src, self.min_dst = a, b
            self.al = self.dl = 0; self.acc, self.dcc = a, b; self.ap, self.dp = self.owner[a], self.owner[b]
            self.until_dead = bool(flag); self.astate = A_EXECUTE
        elif t == M_ATTACK_OUTCOME:
            self.armies[self.acc] -= a; self.armies[self.dcc] -= b
            self.astate = A_INVADE if self.armies[self.dcc] == 0 else A_START
            if flag:   # setupInvasion
                cont = self.m.continent[self.dcc]
                self.ap, self.dp = self.owner[self.acc], self.owner[self.dcc]
                self.count[self.ap] += 1; self.missing[self.ap][cont] -= 1
                self.count[self.dp] -= 1; self.missing[self.dp][cont] += 1
                self.owner[self.dcc] = self.ap
        elif t == M_INVASION:
            acc, dcc = self.acc, self.dcc
            arm = self.armies[acc]
            inv = min(max(c, min(arm - 1, 3)), arm - 1)
            self.armies[dcc] = inv; self.armies[acc] = arm - inv
            if self.count[self.dp] == 0:
                self._player_defeated(self.dp, self.ap)
            self.invaded = True; self.astate = A_START
            self.acc = self.dcc = self.ap = self.dp = -1
            if self.opt["apo"]:
                if self.init_inv:
                    self.init_inv = False
                else:
                    self.min_src = acc
                    for nc in self.m.adjacency[dcc]:
                        if self.owner[nc] != self.cp and nc < self.min_src:
                            self.min_src = nc
                    self.min_dst = dcc if self.min_src == acc else 0
        elif t == M_FORTIFICATION:
            if c == -1:
                self.fortified[b] = 1
            elif self.owner[b] == self.cp and self.owner[c] == self.cp and a >= 0:
                k = min(a, min(self.moveable[b], self.armies[b] - 1))
                if k != 0:
                    self.armies[b] -= k; self.moveable[b] -= k; self.armies[c] += k
        elif t == M_NEXT_PHASE:
            nxt, prev = a, b
            if nxt == INITIAL_PLACEMENT:
                if prev == COUNTRY_SELECTION:
                    if self.rem_countries == 0:
                        self.phase, self.cp = INITIAL_PLACEMENT, 0
                else:
                    self._tear_down_initial_placement()
                left = self.left[self.cp]
                self.placeable = 5 if left == 5 else min(4, left); self.min_place = 0
            elif nxt == CARDS:
                if prev == INITIAL_PLACEMENT:
                    self._tear_down_initial_placement()
                else:
                    self._tear_down_fortification()
                self.placeable = 0
            elif nxt == PLACEMENT:
                self.phase = PLACEMENT; self.placeable += self._income(); self.min_place = 0
            elif nxt == ATTACK:
                self.phase = ATTACK
                self.invaded = False; self.acc = self.dcc = self.ap = self.dp = -1; self.astate = A_START
                self.min_src = self.min_dst = 0
                if self.opt["rep"]:
                    self.made = set()
            elif nxt == FORTIFICATION:
                if self.invaded and self.r.use_cards and self.deck:
                    self.cards[self.cp].append(self.deck.pop(self.rnd.next_int(len(self.deck))))
                self.phase = FORTIFICATION
                self.moveable = [self.armies[t_] if self.owner[t_] == self.cp else 0 for t_ in range(self.N)]
        return self.remaining != 1

    # ---- getPossibleMoves (ExplorationBoard.java:289-706)
    @staticmethod
    def sparse_indices(n: int) -> List[int]:
        if n < 10:
            return [i + 1 for i in range(n)]
        inc = float((n - 1) // 9)
        return [1 + int(i * inc) for i in range(10)]

    def _card_moves(self, ph) -> List[Move]:
        out = []
        h = self.cards[self.cp]; n = len(h)
        if n >= 3:
            best, pick = -1, None
            for i in range(n):
                for j in range(i + 1, n):
                    for k in range(j + 1, n):
                        if not self._is_set(h[i], h[j], h[k]):
                            continue
                        trip = (h[i], h[j], h[k])
                        owned = sum(1 for x in trip if x != WILD and self.owner[x] == self.cp)
                        score = owned * 4 + (3 - sum(1 for x in trip if x == WILD))
                        if score > best:
                            best, pick = score, trip
            if pick is not None:
                out.append((M_CARD_CASH, ph, pick[0], pick[1], pick[2], 0, 0.0, 0))
        if n < 5:
            out.append((M_NEXT_PHASE, NEXT_PHASE, PLACEMENT, CARDS, self.cp, 0, 0.0, 0))
        return out

    def _placement_moves(self, typ, ph) -> List[Move]:
        start = self.min_place if self.opt["ppo"] else 0
        poss = [cc for cc in range(start, self.N) if self.owner[cc] == self.cp and self._enemy_neighbours(cc) > 0]
        amounts = self.sparse_indices(self.placeable)
        out = []
        if self.opt["ppo"]:
            for cc in poss[:-1]:
                out += [(typ, ph, cc, a, 0, 0, 0.0, 0) for a in amounts]
            out.append((typ, ph, poss[-1], self.placeable, 0, 0, 0.0, 0))
        else:
            for cc in poss:
                out += [(typ, ph, cc, a, 0, 0, 0.0, 0) for a in amounts]
        return out

    def possible_moves(self) -> List[Move]:
        NP = NEXT_PHASE
        ph = self.phase
        if ph == COUNTRY_SELECTION:
            out = [(M_COUNTRY_SELECTION, ph, cc, 0, 0, 0, 0.0, 0) for cc in range(self.N) if self.owner[cc] == -1]
            return out or [(M_NEXT_PHASE, NP, INITIAL_PLACEMENT, COUNTRY_SELECTION, 0, 0, 0.0, 0)]
        if ph == INITIAL_PLACEMENT:
            if self.placeable == 0:
                np_ = (self.cp + 1) % self.P
                while np_ != self.cp and self.left[np_] == 0:
                    np_ = (np_ + 1) % self.P
                if np_ == self.cp:
                    return [(M_NEXT_PHASE, NP, CARDS, INITIAL_PLACEMENT, 0, 0, 0.0, 0)]
                return [(M_NEXT_PHASE, NP, INITIAL_PLACEMENT, INITIAL_PLACEMENT, np_, 0, 0.0, 0)]
            return self._placement_moves(M_INITIAL_PLACEMENT, ph)
        if ph == CARDS:
            return self._card_moves(CARDS)
        if ph == PLACEMENT:
            if self.placeable == 0:
                return [(M_NEXT_PHASE, NP, ATTACK, PLACEMENT, self.cp, 0, 0.0, 0)]
            return self._placement_moves(M_PLACEMENT, PLACEMENT)
        if ph == ATTACK:
            st = self.astate
            if st == A_START:
                out = []
                for a in range(self.min_src if self.opt["apo"] else 0, self.N):
                    att = self.armies[a]
                    if self.owner[a] != self.cp or att <= 1:
                        continue
                    d0 = self.min_dst if self.opt["apo"] else 0
                    for d in self.m.adjacency[a]:
                        de = self.armies[d]
                        unfav = (att - 1) == 1 and de >= 2
                        if self.owner[d] != self.cp and d >= d0 and not (self.opt["rep"] and (a, d) in self.made) and not unfav:
                            out.append((M_ATTACK, ph, a, d, att, de, 0.0, 0))
                            if self.opt["aud"] and att > 2:
                                out.append((M_ATTACK, ph, a, d, att, de, 0.0, 1))
                if self.opt["agg"] and self.remaining == 2:
                    if not any(m[4] - 1 >= 3 for m in out):
                        out.append((M_NEXT_PHASE, NP, FORTIFICATION, ATTACK, self.cp, 0, 0.0, 0))
                else:
                    out.append((M_NEXT_PHASE, NP, FORTIFICATION, ATTACK, self.cp, 0, 0.0, 0))
                return out
            if st == A_EXECUTE:
                att, de = self.armies[self.acc], self.armies[self.dcc]
                if self.opt["aud"] and self.until_dead:
                    return [(M_ATTACK_OUTCOME, ph, al, dl, 0, 0, float(p), int(inv)) for (p, al, dl, inv) in self.outcomes(att - 1, de)]
                K = {(1, 1): (0.4167, 0.5833), (1, 2): (0.2546, 0.7454), (2, 1): (0.5787, 0.4213),
                     (2, 2): (0.2276, 0.3241, 0.4483), (3, 1): (0.6597, 0.3403), (3, 2): (0.3717, 0.3358, 0.2926)}
                pa, pd = min(att - 1, 3), min(de, 2)
                if pa < 1 or pd < 1:
                    return []
                n = min(pa, pd)
                return [(M_ATTACK_OUTCOME, ph, x, n - x, 0, 0, K[(pa, pd)][x], 1 if (x == 0 and de == n) else 0) for x in range(n + 1)]
            if st == A_INVADE:
                att = self.armies[self.acc]
                only = False
                if self.opt["inv"]:
                    ah = [x for x in self.m.adjacency[self.acc] if self.owner[x] != self.cp]
                    dh = [x for x in self.m.adjacency[self.dcc] if self.owner[x] != self.cp]
                    only = all(x in dh for x in ah)
                if only:
                    return [(M_INVASION, ph, self.acc, self.dcc, att - 1, 0, 0.0, 0)]
                mn, mx = min(att - 1, 3), att - 1
                return [(M_INVASION, ph, self.acc, self.dcc, e + mn - 1, 0, 0.0, 0) for e in self.sparse_indices(mx - mn + 1)]
            if st == A_CASH:
                return self._card_moves(ATTACK)
            if self.placeable == 0:
                return [(M_NEXT_PHASE, NP, ATTACK, PLACEMENT, self.cp, 0, 0.0, 0)]
            return self._placement_moves(M_PLACEMENT, ATTACK)
        # FORTIFICATION
        cc = 0
        while cc != self.N and (not (self.owner[cc] == self.cp and self.moveable[cc] > 0 and self.armies[cc] > 1)
                                or self.fortified[cc] or not self._friendly(cc)):
            cc += 1
        if cc == self.N:
            return [(M_NEXT_PHASE, NP, CARDS, FORTIFICATION, self._next_player(), 0, 0.0, 0)]
        w = None
        if self.opt["fow"]:
            w = [-1] * self.N
            cur = [t for t in range(self.N) if self.owner[t] == self.cp and self._enemy_neighbours(t) != 0]
            for t in cur:
                w[t] = 0
            while cur:
                nxt = []
                for t in cur:
                    for nb in self.m.adjacency[t]:
                        if self.owner[nb] == self.cp and w[nb] == -1:
                            w[nb] = w[t] + 1; nxt.append(nb)
                cur = nxt
        amounts = self.sparse_indices(min(self.moveable[cc], self.armies[cc] - 1))
        out = []
        for nc in self._friendly(cc):
            if w is None or w[nc] <= w[cc]:
                out += [(M_FORTIFICATION, ph, a, cc, nc, 0, 0.0, 0) for a in amounts]
        out.append((M_FORTIFICATION, ph, -1, cc, -1, 0, 0.0, 0))
        return out

    # ---- getBoardState (ExplorationBoard.java:816-835): the fields a successor record carries
    def state(self) -> dict:
        return dict(owner=list(self.owner), armies=[max(a, 1) for a in self.armies], currentPlayer=self.cp, currentPhase=self.phase,
                    turnCount=self.sim_turn + self.start_turn, playerNumberOfCards=[min(len(c), 255) for c in self.cards],
                    numberOfPlaceableArmies=self.placeable, hasInvadedACountry=self.invaded, attackState=self.astate,
                    attackingCC=self.acc, defendingCC=self.dcc, defendingPlayer=self.dp, moveable=list(self.moveable),
                    fortified=list(self.fortified), explore_until_dead=self.until_dead)
