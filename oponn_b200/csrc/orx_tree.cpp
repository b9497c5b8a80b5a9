// orx_tree.cpp -- native search side of the rollout seam (include/orx_tree.h): the leaf producer and the batched
// UCT driver.  Host code only; the playouts it schedules run in the CUDA engine through orx_rollout_batch.
//
// Restates (O/ = /root/reference/src/com/mld46/oponn/):
//   O/sim/boards/ExplorationBoard.java (whole) on top of the rule primitives of O/sim/boards/SimulationBoard.java
//   (:485-1015), O/CardManager.java:240-346, O/BattleSim.java:174-286, and O/MCTSTree.java:159-431,541-674.
//
// SDK semantics that are not in the reference repository (com.sillysoft.lux, jar unpinned) -- DECLARED ASSUMPTIONS:
//   Country.getNumberEnemyNeighbors()      = adjoining countries whose owner differs from the country's owner
//   Country.getFriendlyAdjoiningCodeList() = codes of adjoining countries with the same owner, adjacency order
//   Card.getBestSet(cards, player, countries) = among the valid sets (orx_state.h) the one with most cards whose
//       territory `player` owns (+2 armies each, SimulationBoard.java:605-616), then fewest wildcards, then the
//       first in lexicographic hand order
//   CardManager.simReset's `new Random(simDeck.hashCode())` = java.util.Random seeded with an FNV-1a hash of the deck
//       (List.hashCode of SDK objects cannot be reproduced; what matters is "same deck, same hidden hands").
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/orx_tree.h"

// thread-local error text of liborx.so (orx_api.cu); not part of the C ABI
__attribute__((visibility("hidden"))) void orx_internal_set_error(const char *text);

namespace {

enum { MAXP = ORX_MAX_PLAYERS, MAXC = ORX_MAX_CONTINENTS };

// ---- java.util.Random (documented LCG) ---------------------------------------------------------------------
struct JavaRandom {
    uint64_t seed = 0;
    void set_seed(uint64_t s) { seed = (s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
    int32_t next(int bits) { seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1); return (int32_t)((int64_t)seed >> (48 - bits)); }
    int32_t next_int(int32_t bound)
    {
        if (bound <= 0) return 0;
        int32_t r = next(31), m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        for (int32_t u = r; (int32_t)((uint32_t)u - (uint32_t)(r = u % bound) + (uint32_t)m) < 0; u = next(31)) {}
        return r;
    }
    double next_double() { return (double)(((int64_t)next(26) << 27) + next(27)) * 0x1.0p-53; }
};

struct SplitMix {  // expansion choices / tie-break noise (the reference uses an unseeded Random and Math.random())
    uint64_t s;
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
    int below(int n) { return n <= 1 ? 0 : (int)(next() % (uint64_t)n); }
    double unit() { return (double)(next() >> 11) * 0x1.0p-53; }
};

// ---- BattleSim outcome tables on the host (O/BattleSim.java:190-286), built lazily --------------------------
struct Outcome { float p; int al, dl; bool inv; };

const float kSingle[3][2][3] = { { { 0.4167f, 0.5833f, 0.f }, { 0.2546f, 0.7454f, 0.f } },
                                 { { 0.5787f, 0.4213f, 0.f }, { 0.2276f, 0.3241f, 0.4483f } },
                                 { { 0.6597f, 0.3403f, 0.f }, { 0.3717f, 0.3358f, 0.2926f } } };

std::vector<Outcome> generate_attack_outcomes(int attackers, int defenders)
{
    const int rl = defenders + 1, cl = attackers + 1;
    std::vector<float> last((size_t)cl * rl, 0.f), next((size_t)cl * rl);
    last[(size_t)attackers * rl + defenders] = 1.0f;
    bool absorbed = false;
    while (!absorbed) {
        std::fill(next.begin(), next.end(), 0.f);
        absorbed = true;
        for (int a = 0; a < cl; a++)
            for (int d = 0; d < rl; d++) {
                volatile float v = last[(size_t)a * rl + d];
                if (v == 0) continue;
                if (a == 0 || d == 0) { next[(size_t)a * rl + d] += v; continue; }
                int pa = a < 3 ? a : 3, pd = d < 2 ? d : 2, n = pa < pd ? pa : pd;
                for (int x = 0; x <= n; x++) {  // attacker loses x, defender n - x: the reference's scatter order
                    volatile float c = v * kSingle[pa - 1][pd - 1][x];
                    next[(size_t)(a - x) * rl + (d - (n - x))] += c;
                }
                absorbed = false;
            }
        last.swap(next);
    }
    std::vector<Outcome> out;
    for (int ra = 0; ra <= attackers; ra++)
        for (int rd = 0; rd <= defenders; rd++) {
            float p = last[(size_t)ra * rl + rd];
            if (p != 0.0f) out.push_back({ p, attackers - ra, defenders - rd, rd == 0 });
        }
    std::stable_sort(out.begin(), out.end(), [](const Outcome &x, const Outcome &y) { return x.p > y.p; });
    return out;
}

// ---- static game description --------------------------------------------------------------------------------
struct Static {
    int N = 0, C = 0, P = 0;
    std::vector<int> adj_off, adj, continent, base_bonus, symbol;
    int use_cards = 1, transfer_cards = 1, cont_inc = 0, max_turns = 1 << 16;
    std::string progression;
    bool can_run_out = false;
    OrxLeafLayout L;
    std::map<std::pair<int, int>, std::vector<Outcome>> tables;
    const std::vector<Outcome> &outcomes(int a, int d)
    {
        auto key = std::make_pair(a, d);
        auto it = tables.find(key);
        if (it == tables.end()) it = tables.emplace(key, generate_attack_outcomes(a, d)).first;
        return it->second;
    }
};

int32_t java_d2i(double d)
{
    if (d != d) return 0;
    if (d >= 2147483647.0) return 2147483647;
    if (d <= -2147483648.0) return (int32_t)(-2147483647 - 1);
    return (int32_t)d;
}

// CardManager.getCardCashValue (O/CardManager.java:283-341)
int card_cash_value(const std::string &cp, int pos)
{
    if (cp.find('%') != std::string::npos) {
        int base, offset;
        if (cp[0] == '5') { base = 5 * pos; offset = 6; } else { base = pos <= 4 ? 2 * (pos + 1) : 5 * (pos - 2); offset = 8; }
        size_t hat = cp.find('^');
        int pct = hat == std::string::npos ? 0 : atoi(cp.substr(hat + 1, cp.size() - hat - 2).c_str());
        double e = 1 + (pct / 100.0);
        return java_d2i(floor(base <= 25 ? (double)base : base * pow(e, (double)(pos - offset))));
    }
    if (cp == "0") return 0;
    if (cp == "5, 5, 5...") return 5;
    if (cp == "4, 4, 6, 6, 6, 8, 8, 8, 8, 10...") return java_d2i(ceil((sqrt((double)(8 * pos + 9)) - 1) / 2) * 2);
    if (cp == "4, 5, 6...") return pos + 3;
    if (cp == "4, 6, 8...") return (pos + 1) * 2;
    if (cp == "3, 6, 9...") return pos * 3;
    if (cp == "4, 6, 8, 10, 15, 20...") return pos <= 4 ? (pos + 1) * 2 : (pos - 2) * 5;
    if (cp == "5, 10, 15...") return pos * 5;
    if (cp == "4, 6, 8, 10, 15, 20, 25, 10, 10, 10...") return pos <= 4 ? (pos + 1) * 2 : (pos <= 7 ? (pos - 2) * 5 : 10);
    return 2;
}

OrxMove mk(int type, int phase, int a = 0, int b = 0, int c = 0, int d = 0, float p = 0.f, int flag = 0)
{
    OrxMove m; m.type = type; m.phase = phase; m.a = a; m.b = b; m.c = c; m.d = d; m.probability = p; m.flag = flag;
    return m;
}

// ---- the exploration board -----------------------------------------------------------------------------------
struct Board {
    Static &S;
    const OrxTreeOptions &O;
    std::vector<int> owner, armies, moveable;
    std::vector<std::vector<uint8_t>> playerCards;
    std::vector<uint8_t> simDeck, realHand, realDeck;
    int simCardPos = 0, realCardPos = 0;
    int playerCountries[MAXP], missing[MAXP][MAXC], bonus[MAXC], playerOutOnTurn[MAXP], initialArmiesLeft[MAXP];
    int remainingPlayers = 0, currentPlayer = 0, phase = 0, simStartTurn = 0, simTurnCount = 0, remainingCountries = 0;
    int placeable = 0, attackingCC = -1, defendingCC = -1, attackingPlayer = -1, defendingPlayer = -1;
    int attackerLosses = 0, defenderLosses = 0, attackState = ORX_ATTACK_START;
    bool hasInvaded = false, attackUntilDead = false, error = false;
    int minPlacementCC = 0, minAttackSourceCC = 0, minAttackDestinationCC = 0;
    bool initiatedWithInvasion = false;
    std::vector<uint8_t> attacksMade, fortified;
    JavaRandom rnd;

    Board(Static &s, const OrxTreeOptions &o) : S(s), O(o), owner(s.N), armies(s.N), moveable(s.N), playerCards(s.P),
                                                attacksMade((size_t)s.N * s.N, 0), fortified(s.N, 0) {}

    // -- SDK-style helpers
    int enemy_neighbours(int cc) const
    {
        int n = 0;
        for (int e = S.adj_off[cc]; e < S.adj_off[cc + 1]; e++) n += owner[S.adj[e]] != owner[cc];
        return n;
    }
    bool is_set(int c1, int c2, int c3) const
    {
        if (c1 == ORX_CARD_WILD || c2 == ORX_CARD_WILD || c3 == ORX_CARD_WILD) return true;
        int a = S.symbol[c1], b = S.symbol[c2], c = S.symbol[c3];
        return (a == b && b == c) || (a != b && b != c && a != c);
    }

    // -- CardManager simulation half
    int sim_deal()
    {
        if (simDeck.empty()) { if (!S.can_run_out) error = true; return -1; }
        int i = rnd.next_int((int)simDeck.size());
        int c = simDeck[i];
        simDeck.erase(simDeck.begin() + i);
        return c;
    }

    void recalc_bonuses()
    {
        int e = simStartTurn + simTurnCount - 1; if (e < 0) e = 0;
        double factor = pow(1 + S.cont_inc / 100.0, (double)e);
        for (int i = 0; i < S.C; i++) bonus[i] = java_d2i(floor(S.base_bonus[i] * factor));
    }

    int starting_armies(int i) const
    {
        double f1 = S.N / 42.0; f1 -= 1.0; f1 /= 2.0; f1 += 1.0;
        int base = (int)floor(f1 * (50 - S.P * 5) + 0.5);
        return base + (i > 1 ? i : 0);
    }

    // -- setupState (SimulationBoard.java:178-359 + ExplorationBoard.java:77-114)
    bool setup(const uint8_t *leaf, const uint8_t *extra)
    {
        const OrxLeafHeader *h = (const OrxLeafHeader *)leaf;
        const OrxLeafLayout &L = S.L;
        const uint8_t *ow = leaf + L.off_owner; const int32_t *ar = (const int32_t *)(leaf + L.off_armies);
        const uint8_t *nc = leaf + L.off_ncards; const uint8_t *cards = leaf + L.off_cards;
        const int N = S.N, P = S.P;
        error = false;
        bool bad = h->current_player < 0 || h->current_player >= P || h->phase < 0 || h->phase > ORX_PHASE_FORTIFICATION ||
                   h->attack_state < 0 || h->attack_state > ORX_ATTACK_PLACE || h->attacking_cc < -1 || h->attacking_cc >= N ||
                   h->defending_cc < -1 || h->defending_cc >= N || h->defending_player < -1 || h->defending_player >= P ||
                   h->card_pos < 0 || (int)h->n_hand + (int)h->n_deck > N + 2;
        for (int t = 0; t < N && !bad; t++) {
            int o = ow[t];
            if (!(o < P || (o == ORX_OWNER_NONE && h->phase == ORX_PHASE_COUNTRY_SELECTION))) bad = true;
            if (ar[t] < 1 || ar[t] > ORX_MAX_ARMIES) bad = true;
        }
        for (int i = 0; i < (int)h->n_hand + (int)h->n_deck && !bad; i++) if (!(cards[i] < N || cards[i] == ORX_CARD_WILD)) bad = true;
        if (bad) { error = true; return false; }

        simStartTurn = h->turn_count; currentPlayer = h->current_player; phase = h->phase;
        for (int t = 0; t < N; t++) { owner[t] = ow[t] == ORX_OWNER_NONE ? -1 : ow[t]; armies[t] = ar[t]; }
        if (extra) {
            const int32_t *mv = (const int32_t *)extra; const uint8_t *ft = (const uint8_t *)extra + 4 * N;
            for (int t = 0; t < N; t++) { moveable[t] = mv[t]; fortified[t] = ft[t] != 0; }
        } else {
            for (int t = 0; t < N; t++) { moveable[t] = (phase == ORX_PHASE_FORTIFICATION && owner[t] == currentPlayer) ? armies[t] : 0; fortified[t] = 0; }
        }
        // cards: CardManager.simReset (:240-248) then setupCardState (:205-227)
        realHand.assign(cards, cards + h->n_hand);
        realDeck.assign(cards + h->n_hand, cards + h->n_hand + h->n_deck);
        realCardPos = h->card_pos;
        simDeck = realDeck; simCardPos = realCardPos;
        // random = new Random(simDeck.hashCode()): the record's deal_seed, as the rollout engine uses it (orx_state.h)
        rnd.set_seed((uint64_t)(int64_t)h->deal_seed);
        for (int p = 0; p < P; p++) {
            playerCards[p].clear();
            if (p == currentPlayer) playerCards[p] = realHand;
            else for (int j = 0; j < nc[p]; j++) { int c = sim_deal(); if (c < 0) { error = true; return false; } playerCards[p].push_back((uint8_t)c); }
        }
        // setupGeneralGameState (:229-276)
        remainingPlayers = 0; simTurnCount = 0;
        for (int p = 0; p < P; p++) { playerOutOnTurn[p] = 0; playerCountries[p] = 0; for (int c = 0; c < S.C; c++) missing[p][c] = 0; }
        for (int cc = 0; cc < N; cc++) {
            for (int p = 0; p < P; p++) if (p != owner[cc]) missing[p][S.continent[cc]]++;
            if (!(phase == ORX_PHASE_COUNTRY_SELECTION && owner[cc] == -1)) playerCountries[owner[cc]]++;
        }
        if (phase == ORX_PHASE_COUNTRY_SELECTION) remainingPlayers = P;
        else for (int p = 0; p < P; p++) remainingPlayers += playerCountries[p] > 0;
        // setupPhaseGameState (:278-359) + the ExplorationBoard override (:77-114)
        placeable = 0;
        attackingCC = defendingCC = attackingPlayer = defendingPlayer = -1; attackState = ORX_ATTACK_START; hasInvaded = false;
        switch (phase) {
        case ORX_PHASE_COUNTRY_SELECTION:
            remainingCountries = 0;
            for (int i = 0; i < P; i++) initialArmiesLeft[i] = starting_armies(i);
            for (int cc = 0; cc < N; cc++) { if (owner[cc] == -1) remainingCountries++; else initialArmiesLeft[owner[cc]]--; }
            break;
        case ORX_PHASE_INITIAL_PLACEMENT: {
            int pa[MAXP] = { 0 };
            for (int cc = 0; cc < N; cc++) pa[owner[cc]] += armies[cc];
            for (int i = 0; i < P; i++) initialArmiesLeft[i] = starting_armies(i) - pa[i];
            placeable = h->placeable; minPlacementCC = 0;
            break;
        }
        case ORX_PHASE_CARDS: placeable = h->placeable; break;
        case ORX_PHASE_PLACEMENT: placeable = h->placeable; minPlacementCC = 0; break;
        case ORX_PHASE_ATTACK:
            std::fill(attacksMade.begin(), attacksMade.end(), 0);
            hasInvaded = h->has_invaded != 0; attackState = h->attack_state;
            attackingCC = h->attacking_cc; defendingCC = h->defending_cc;
            if (attackingCC != -1) { attackingPlayer = owner[attackingCC]; defendingPlayer = h->defending_player; }
            if (attackState == ORX_ATTACK_INVADE) {
                if (defendingPlayer < 0) { error = true; return false; }
                if (playerCountries[defendingPlayer] == 0) remainingPlayers++;
            } else if (attackState == ORX_ATTACK_CASH || attackState == ORX_ATTACK_PLACE) placeable = h->placeable;
            initiatedWithInvasion = attackState == ORX_ATTACK_INVADE;
            minPlacementCC = minAttackSourceCC = minAttackDestinationCC = 0;
            attackUntilDead = h->tree_attack_until_dead != 0;   // the exploration board's own flag (set by the Attack move), tree-only
            break;
        default: break;
        }
        recalc_bonuses();
        return true;
    }

    // -- rule primitives (SimulationBoard.java:485-1015)
    void select_country(int cc)
    {
        if (cc < 0 || owner[cc] != -1) { error = true; return; }
        owner[cc] = currentPlayer; playerCountries[currentPlayer]++; initialArmiesLeft[currentPlayer]--; remainingCountries--;
        missing[currentPlayer][S.continent[cc]]--;
        currentPlayer = (currentPlayer + 1) % S.P;
    }
    void tear_down_country_selection() { if (remainingCountries == 0) { phase = ORX_PHASE_INITIAL_PLACEMENT; currentPlayer = 0; } }
    void setup_initial_placement() { int a = initialArmiesLeft[currentPlayer]; placeable = a == 5 ? 5 : (a < 4 ? a : 4); minPlacementCC = 0; }
    void tear_down_initial_placement()
    {
        int np = (currentPlayer + 1) % S.P;
        while (np != currentPlayer && initialArmiesLeft[np] == 0) np = (np + 1) % S.P;
        if (np == currentPlayer) { currentPlayer = 0; phase = ORX_PHASE_CARDS; simTurnCount++; } else currentPlayer = np;
    }
    bool cash_cards(int c1, int c2, int c3)
    {
        if (phase != ORX_PHASE_CARDS && (phase != ORX_PHASE_ATTACK || attackState != ORX_ATTACK_CASH)) return false;
        if (!is_set(c1, c2, c3)) return false;
        std::vector<uint8_t> &hand = playerCards[currentPlayer];
        int cs[3] = { c1, c2, c3 };
        std::vector<uint8_t> tmp = hand;
        for (int c : cs) { auto it = std::find(tmp.begin(), tmp.end(), (uint8_t)c); if (it == tmp.end()) return false; tmp.erase(it); }
        hand = tmp;
        for (int c : cs) simDeck.push_back((uint8_t)c);
        placeable += card_cash_value(S.progression, simCardPos); simCardPos++;
        for (int c : cs) if (c != ORX_CARD_WILD && owner[c] == currentPlayer) armies[c] += 2;
        return true;
    }
    void tear_down_cards()
    {
        // forced cashing while the hand is above four cards uses Card.getRandomSet in the reference (:620-638); the tree
        // only reaches this through NextPhase moves, which getCardMoves offers with fewer than five cards
        phase = ORX_PHASE_PLACEMENT;
    }
    int income() const
    {
        if (playerCountries[currentPlayer] == 0) return 0;
        int inc = playerCountries[currentPlayer] / 3; if (inc < 3) inc = 3;
        for (int i = 0; i < S.C; i++) if (missing[currentPlayer][i] == 0) inc += bonus[i];
        return inc;
    }
    void place_armies(int n, int cc)
    {
        bool okphase = phase == ORX_PHASE_PLACEMENT || phase == ORX_PHASE_INITIAL_PLACEMENT || (phase == ORX_PHASE_ATTACK && attackState == ORX_ATTACK_PLACE);
        if (!okphase || n > placeable || cc < 0 || cc >= S.N || owner[cc] != currentPlayer || n < 0) { error = true; return; }
        armies[cc] += n; placeable -= n;
        if (phase == ORX_PHASE_INITIAL_PLACEMENT) initialArmiesLeft[currentPlayer] -= n;
    }
    void setup_attack_phase()
    {
        hasInvaded = false; attackingCC = defendingCC = attackingPlayer = defendingPlayer = -1; attackState = ORX_ATTACK_START;
        minAttackSourceCC = minAttackDestinationCC = 0;
        if (O.attack_repeat_moves) std::fill(attacksMade.begin(), attacksMade.end(), 0);
    }
    void setup_attack(int a, int d, bool until_dead)
    {
        attackerLosses = defenderLosses = 0; attackingCC = a; defendingCC = d;
        attackingPlayer = owner[a]; defendingPlayer = owner[d]; attackUntilDead = until_dead; attackState = ORX_ATTACK_EXECUTE;
    }
    void finish_attack()
    {
        if (attackingCC < 0 || defendingCC < 0) { error = true; return; }
        armies[attackingCC] -= attackerLosses; armies[defendingCC] -= defenderLosses;
        attackState = armies[defendingCC] == 0 ? ORX_ATTACK_INVADE : ORX_ATTACK_START;
    }
    void setup_invasion()
    {
        int cont = S.continent[defendingCC];
        attackingPlayer = owner[attackingCC]; defendingPlayer = owner[defendingCC];
        playerCountries[attackingPlayer]++; missing[attackingPlayer][cont]--;
        playerCountries[defendingPlayer]--; missing[defendingPlayer][cont]++;
        owner[defendingCC] = attackingPlayer;
    }
    void execute_invasion(int n)
    {
        int a = armies[attackingCC];
        int mn = a - 1 < 3 ? a - 1 : 3, mx = a - 1;
        int inv = n > mn ? n : mn; if (inv > mx) inv = mx;
        armies[defendingCC] = inv; armies[attackingCC] = a - inv;
    }
    void player_defeated(int player, int conq)
    {
        remainingPlayers--; playerOutOnTurn[player] = simTurnCount;
        if (S.transfer_cards) { if (conq < 0) { error = true; return; } if (conq != player) { auto &d = playerCards[conq]; auto &s = playerCards[player]; d.insert(d.end(), s.begin(), s.end()); } }
        else simDeck.insert(simDeck.end(), playerCards[player].begin(), playerCards[player].end());
        if (remainingPlayers == 1) playerOutOnTurn[currentPlayer] = -1;
    }
    void finish_invasion()
    {
        if (defendingPlayer < 0) { error = true; return; }
        if (playerCountries[defendingPlayer] == 0) player_defeated(defendingPlayer, attackingPlayer);
        hasInvaded = true; attackState = ORX_ATTACK_START;
        attackingCC = defendingCC = attackingPlayer = defendingPlayer = -1;
    }
    void tear_down_attack()
    {
        if (hasInvaded && S.use_cards) { int c = sim_deal(); if (c >= 0) playerCards[currentPlayer].push_back((uint8_t)c); }
        phase = ORX_PHASE_FORTIFICATION;
    }
    void setup_fortification() { for (int t = 0; t < S.N; t++) moveable[t] = owner[t] == currentPlayer ? armies[t] : 0; }
    int fortify(int n, int src, int dst)
    {
        if (phase != ORX_PHASE_FORTIFICATION) { error = true; return -1; }
        if (owner[src] != currentPlayer || owner[dst] != currentPlayer || n < 0) return -1;
        int k = std::min(n, std::min(moveable[src], armies[src] - 1));
        if (k == 0) return 0;
        armies[src] -= k; moveable[src] -= k; armies[dst] += k;
        return 1;
    }
    int next_player() const
    {
        int np = currentPlayer;
        for (int g = 0; g <= S.P; g++) { np = (np + 1) % S.P; if (playerCountries[np] != 0) return np; }
        return currentPlayer;
    }
    void check_overflow()
    {
        long long tot[MAXP] = { 0 }, total = 0, mx = 0; int mp = -1;
        for (int c = 0; c < S.N; c++) if (owner[c] >= 0) tot[owner[c]] += armies[c];
        for (int i = 0; i < S.P; i++) { total += tot[i]; if (tot[i] > mx) { mp = i; mx = tot[i]; } }
        if (total > 100000) for (int i = 0; i < S.P; i++) if (playerCountries[i] > 0 && i != mp) player_defeated(i, mp);
    }
    void tear_down_fortification()
    {
        for (int t = 0; t < S.N; t++) moveable[t] = 0;
        int np = next_player();
        if (np < currentPlayer) { simTurnCount++; recalc_bonuses(); }
        currentPlayer = np; phase = ORX_PHASE_CARDS;
        check_overflow();
    }

    // -- updateState (ExplorationBoard.java:116-283)
    bool update(const OrxMove &m)
    {
        if (m.phase != phase && m.type != ORX_MOVE_NEXT_PHASE) { error = true; return remainingPlayers != 1; }
        switch (m.type) {
        case ORX_MOVE_COUNTRY_SELECTION: select_country(m.a); break;
        case ORX_MOVE_INITIAL_PLACEMENT:
        case ORX_MOVE_PLACEMENT:
            place_armies(m.b, m.a);
            if (O.placement_partial_order) minPlacementCC = m.a + 1;
            break;
        case ORX_MOVE_CARD_CASH: cash_cards(m.a, m.b, m.c); break;
        case ORX_MOVE_ATTACK:
            if (m.a < 0 || m.a >= S.N || m.b < 0 || m.b >= S.N) { error = true; break; }
            if (O.attack_repeat_moves && attackingCC != -1 && defendingCC != -1 && (attackingCC != m.a || defendingCC != m.b))
                attacksMade[(size_t)attackingCC * S.N + defendingCC] = 1;
            minAttackSourceCC = m.a; minAttackDestinationCC = m.b;
            setup_attack(m.a, m.b, m.flag != 0);
            break;
        case ORX_MOVE_ATTACK_OUTCOME:
            attackerLosses = m.a; defenderLosses = m.b;
            finish_attack();
            if (!error && m.flag) setup_invasion();
            break;
        case ORX_MOVE_INVASION: {
            if (attackingCC < 0 || defendingCC < 0) { error = true; break; }
            int aCC = attackingCC, dCC = defendingCC;
            execute_invasion(m.c);
            finish_invasion();
            if (O.attack_partial_order) {
                if (initiatedWithInvasion) initiatedWithInvasion = false;
                else {
                    minAttackSourceCC = aCC;
                    for (int e = S.adj_off[dCC]; e < S.adj_off[dCC + 1]; e++) { int nc = S.adj[e]; if (owner[nc] != currentPlayer && nc < minAttackSourceCC) minAttackSourceCC = nc; }
                    minAttackDestinationCC = minAttackSourceCC == aCC ? dCC : 0;
                }
            }
            break;
        }
        case ORX_MOVE_FORTIFICATION:
            if (m.b < 0 || m.b >= S.N || m.c >= S.N) { error = true; break; }
            if (m.c == -1) fortified[m.b] = 1; else fortify(m.a, m.b, m.c);
            break;
        case ORX_MOVE_NEXT_PHASE:
            switch (m.a) {
            case ORX_PHASE_INITIAL_PLACEMENT:
                if (m.b == ORX_PHASE_COUNTRY_SELECTION) tear_down_country_selection(); else tear_down_initial_placement();
                setup_initial_placement();
                break;
            case ORX_PHASE_CARDS:
                if (m.b == ORX_PHASE_INITIAL_PLACEMENT) tear_down_initial_placement(); else tear_down_fortification();
                placeable = 0;
                break;
            case ORX_PHASE_PLACEMENT: tear_down_cards(); placeable += income(); minPlacementCC = 0; break;
            case ORX_PHASE_ATTACK: phase = ORX_PHASE_ATTACK; setup_attack_phase(); break;
            case ORX_PHASE_FORTIFICATION: tear_down_attack(); setup_fortification(); break;
            default: error = true;
            }
            break;
        default: error = true;
        }
        return remainingPlayers != 1;
    }

    // -- getPossibleMoves (ExplorationBoard.java:289-706)
    static std::vector<int> sparse_indices(int max_armies)
    {
        std::vector<int> idx;
        const int suggestions = 10;
        if (max_armies < suggestions) { for (int i = 0; i < max_armies; i++) idx.push_back(i + 1); }
        else { double inc = (double)((max_armies - 1) / (suggestions - 1)); for (int i = 0; i < suggestions; i++) idx.push_back(1 + (int)(i * inc)); }
        return idx;
    }
    void card_moves(int move_phase, std::vector<OrxMove> &out)
    {
        const std::vector<uint8_t> &h = playerCards[currentPlayer];
        int n = (int)h.size();
        if (n >= 3) {  // Card.getBestSet: declared assumption (file header)
            int best = -1, bi = 0, bj = 0, bk = 0;
            for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) for (int k = j + 1; k < n; k++) {
                if (!is_set(h[i], h[j], h[k])) continue;
                int owned = 0, wild = 0;
                for (int c : { (int)h[i], (int)h[j], (int)h[k] }) { if (c == ORX_CARD_WILD) wild++; else if (owner[c] == currentPlayer) owned++; }
                int score = owned * 4 + (3 - wild);
                if (score > best) { best = score; bi = i; bj = j; bk = k; }
            }
            if (best >= 0) out.push_back(mk(ORX_MOVE_CARD_CASH, move_phase, h[bi], h[bj], h[bk]));
        }
        if (n < 5) out.push_back(mk(ORX_MOVE_NEXT_PHASE, ORX_PHASE_NEXT_PHASE, ORX_PHASE_PLACEMENT, ORX_PHASE_CARDS, currentPlayer));
    }
    void placement_moves(int type, int move_phase, std::vector<OrxMove> &out)
    {
        std::vector<int> poss;
        for (int cc = O.placement_partial_order ? minPlacementCC : 0; cc < S.N; cc++)
            if (owner[cc] == currentPlayer && enemy_neighbours(cc) > 0) poss.push_back(cc);
        std::vector<int> amounts = sparse_indices(placeable);
        if (O.placement_partial_order) {
            if (poss.empty()) { error = true; return; }  // the reference indexes possiblePlacements.get(-1) here
            for (size_t i = 0; i + 1 < poss.size(); i++) for (int a : amounts) out.push_back(mk(type, move_phase, poss[i], a));
            out.push_back(mk(type, move_phase, poss.back(), placeable));
        } else {
            for (int cc : poss) for (int a : amounts) out.push_back(mk(type, move_phase, cc, a));
        }
    }
    std::vector<int> fortification_weights() const  // FortificationAnalysis.calculateWeights (:868-902)
    {
        std::vector<int> w(S.N, -1), cur, nxt;
        for (int cc = 0; cc < S.N; cc++) if (owner[cc] == currentPlayer && enemy_neighbours(cc) != 0) { w[cc] = 0; cur.push_back(cc); }
        while (!cur.empty()) {
            nxt.clear();
            for (int cc : cur) for (int e = S.adj_off[cc]; e < S.adj_off[cc + 1]; e++) { int nb = S.adj[e]; if (owner[nb] == currentPlayer && w[nb] == -1) { w[nb] = w[cc] + 1; nxt.push_back(nb); } }
            cur.swap(nxt);
        }
        return w;
    }
    std::vector<OrxMove> possible_moves()
    {
        std::vector<OrxMove> out;
        const int NP = ORX_PHASE_NEXT_PHASE;
        switch (phase) {
        case ORX_PHASE_COUNTRY_SELECTION:
            for (int cc = 0; cc < S.N; cc++) if (owner[cc] == -1) out.push_back(mk(ORX_MOVE_COUNTRY_SELECTION, phase, cc));
            if (out.empty()) out.push_back(mk(ORX_MOVE_NEXT_PHASE, NP, ORX_PHASE_INITIAL_PLACEMENT, ORX_PHASE_COUNTRY_SELECTION, 0));
            break;
        case ORX_PHASE_INITIAL_PLACEMENT:
            if (placeable == 0) {
                int np = (currentPlayer + 1) % S.P;
                while (np != currentPlayer && initialArmiesLeft[np] == 0) np = (np + 1) % S.P;
                if (np == currentPlayer) out.push_back(mk(ORX_MOVE_NEXT_PHASE, NP, ORX_PHASE_CARDS, ORX_PHASE_INITIAL_PLACEMENT, 0));
                else out.push_back(mk(ORX_MOVE_NEXT_PHASE, NP, ORX_PHASE_INITIAL_PLACEMENT, ORX_PHASE_INITIAL_PLACEMENT, np));
            } else placement_moves(ORX_MOVE_INITIAL_PLACEMENT, phase, out);
            break;
        case ORX_PHASE_CARDS: card_moves(ORX_PHASE_CARDS, out); break;
        case ORX_PHASE_PLACEMENT:
            if (placeable == 0) out.push_back(mk(ORX_MOVE_NEXT_PHASE, NP, ORX_PHASE_ATTACK, ORX_PHASE_PLACEMENT, currentPlayer));
            else placement_moves(ORX_MOVE_PLACEMENT, ORX_PHASE_PLACEMENT, out);
            break;
        case ORX_PHASE_ATTACK:
            switch (attackState) {
            case ORX_ATTACK_START: {
                for (int a = O.attack_partial_order ? minAttackSourceCC : 0; a < S.N; a++) {
                    int att = armies[a];
                    if (owner[a] != currentPlayer || att <= 1) continue;
                    int d0 = O.attack_partial_order ? minAttackDestinationCC : 0;
                    for (int e = S.adj_off[a]; e < S.adj_off[a + 1]; e++) {
                        int d = S.adj[e], def = armies[d];
                        bool unfav = (att - 1) == 1 && def >= 2;  // BattleSim.unfavourableForAttacker (:169-172)
                        if (owner[d] != currentPlayer && d >= d0 && !(O.attack_repeat_moves && attacksMade[(size_t)a * S.N + d]) && !unfav) {
                            out.push_back(mk(ORX_MOVE_ATTACK, phase, a, d, att, def, 0.f, 0));
                            if (O.attack_until_dead && att > 2) out.push_back(mk(ORX_MOVE_ATTACK, phase, a, d, att, def, 0.f, 1));
                        }
                    }
                }
                bool add_next = true;
                if (O.attack_aggressive && remainingPlayers == 2) {
                    bool exhausted = true;   // BattleSim.favourableForAttacker(free, d) = free >= 3 (:164-167)
                    for (const OrxMove &m : out) if (m.c - 1 >= 3) exhausted = false;
                    add_next = exhausted;
                }
                if (add_next) out.push_back(mk(ORX_MOVE_NEXT_PHASE, NP, ORX_PHASE_FORTIFICATION, ORX_PHASE_ATTACK, currentPlayer));
                break;
            }
            case ORX_ATTACK_EXECUTE: {
                if (attackingCC < 0 || defendingCC < 0) { error = true; break; }
                int att = armies[attackingCC], def = armies[defendingCC];
                if (O.attack_until_dead && attackUntilDead) {
                    for (const Outcome &o : S.outcomes(att - 1, def)) out.push_back(mk(ORX_MOVE_ATTACK_OUTCOME, phase, o.al, o.dl, 0, 0, o.p, o.inv));
                } else {
                    int pa = att - 1 < 3 ? att - 1 : 3, pd = def < 2 ? def : 2;
                    if (pa >= 1 && pd >= 1) {
                        int n = pa < pd ? pa : pd;
                        for (int x = 0; x <= n; x++)
                            out.push_back(mk(ORX_MOVE_ATTACK_OUTCOME, phase, x, n - x, 0, 0, kSingle[pa - 1][pd - 1][x], (x == 0 && def == n) ? 1 : 0));
                    }
                }
                break;
            }
            case ORX_ATTACK_INVADE: {
                if (attackingCC < 0 || defendingCC < 0) { error = true; break; }
                int att = armies[attackingCC];
                bool only_largest = false;
                if (O.invasion) {  // defender's hostile neighbours contain all of the attacker's
                    only_largest = true;
                    for (int e = S.adj_off[attackingCC]; e < S.adj_off[attackingCC + 1] && only_largest; e++) {
                        int nb = S.adj[e];
                        if (owner[nb] == currentPlayer) continue;
                        bool found = false;
                        for (int f = S.adj_off[defendingCC]; f < S.adj_off[defendingCC + 1]; f++) if (S.adj[f] == nb && owner[nb] != currentPlayer) found = true;
                        if (!found) only_largest = false;
                    }
                }
                if (only_largest) out.push_back(mk(ORX_MOVE_INVASION, phase, attackingCC, defendingCC, att - 1));
                else {
                    int mn = att - 1 < 3 ? att - 1 : 3, mx = att - 1;
                    for (int extra : sparse_indices(mx - mn + 1)) out.push_back(mk(ORX_MOVE_INVASION, phase, attackingCC, defendingCC, extra + mn - 1));
                }
                break;
            }
            case ORX_ATTACK_CASH: card_moves(ORX_PHASE_ATTACK, out); break;
            case ORX_ATTACK_PLACE:
                if (placeable == 0) out.push_back(mk(ORX_MOVE_NEXT_PHASE, NP, ORX_PHASE_ATTACK, ORX_PHASE_PLACEMENT, currentPlayer));
                else placement_moves(ORX_MOVE_PLACEMENT, ORX_PHASE_ATTACK, out);
                break;
            }
            break;
        case ORX_PHASE_FORTIFICATION: {
            auto friendly = [&](int cc) { std::vector<int> f; for (int e = S.adj_off[cc]; e < S.adj_off[cc + 1]; e++) if (owner[S.adj[e]] == owner[cc]) f.push_back(S.adj[e]); return f; };
            int cc = 0;
            while (cc != S.N && (!(owner[cc] == currentPlayer && moveable[cc] > 0 && armies[cc] > 1) || fortified[cc] || friendly(cc).empty())) cc++;
            if (cc == S.N) out.push_back(mk(ORX_MOVE_NEXT_PHASE, NP, ORX_PHASE_CARDS, ORX_PHASE_FORTIFICATION, next_player()));
            else {
                std::vector<int> w;
                if (O.fortification_outwards) w = fortification_weights();
                std::vector<int> amounts = sparse_indices(std::min(moveable[cc], armies[cc] - 1));
                for (int nc : friendly(cc))
                    if (!O.fortification_outwards || w[nc] <= w[cc]) for (int a : amounts) out.push_back(mk(ORX_MOVE_FORTIFICATION, phase, a, cc, nc));
                out.push_back(mk(ORX_MOVE_FORTIFICATION, phase, -1, cc, -1));
            }
            break;
        }
        default: error = true;
        }
        return out;
    }

    // -- getBoardState (ExplorationBoard.java:816-835) as a packed record + extras
    void pack(uint8_t *leaf, uint8_t *extra) const
    {
        const OrxLeafLayout &L = S.L;
        memset(leaf, 0, (size_t)L.stride);
        OrxLeafHeader *h = (OrxLeafHeader *)leaf;
        h->turn_count = simTurnCount + simStartTurn; h->placeable = placeable;
        h->attacking_cc = (int16_t)attackingCC; h->defending_cc = (int16_t)defendingCC; h->card_pos = (int16_t)realCardPos;
        h->current_player = (int8_t)currentPlayer; h->phase = (int8_t)phase; h->attack_state = (int8_t)attackState;
        h->has_invaded = hasInvaded; h->defending_player = (int8_t)defendingPlayer;
        h->attack_until_dead = 1;  // the simulation boards' sticky bit in steady state
        h->tree_attack_until_dead = attackUntilDead ? 1 : 0;   // tree-only: Attack.attackUntilDead of the pending attack
        h->deal_mode = ORX_DEAL_LEAF_LCG;
        h->deal_seed = orx_deck_hash(realDeck.data(), (int)realDeck.size());   // stand-in for simDeck.hashCode()
        h->n_hand = (uint8_t)realHand.size(); h->n_deck = (uint8_t)realDeck.size();
        uint8_t *ow = leaf + L.off_owner; int32_t *ar = (int32_t *)(leaf + L.off_armies); uint8_t *nc = leaf + L.off_ncards; uint8_t *cards = leaf + L.off_cards;
        for (int t = 0; t < S.N; t++) { ow[t] = owner[t] < 0 ? (uint8_t)ORX_OWNER_NONE : (uint8_t)owner[t]; ar[t] = armies[t] < 1 ? 1 : armies[t]; }
        for (int p = 0; p < S.P; p++) nc[p] = (uint8_t)std::min<size_t>(playerCards[p].size(), 255);
        size_t k = 0;
        for (uint8_t c : realHand) cards[k++] = c;
        for (uint8_t c : realDeck) cards[k++] = c;
        if (extra) {
            int32_t *mv = (int32_t *)extra; uint8_t *ft = extra + 4 * S.N;
            for (int t = 0; t < S.N; t++) { mv[t] = moveable[t]; ft[t] = fortified[t]; }
        }
    }
};

// ---- the search tree (MCTSTree.java) -------------------------------------------------------------------------
struct Node {
    OrxMove move{};
    int parent = -1, depth = 0, currentPlayer = 0, rootIndex = -1;
    long long visits = 0, inflight = 0;
    long long wins[MAXP] = { 0 }, losses[MAXP] = { 0 }, winLen[MAXP] = { 0 }, lossLen[MAXP] = { 0 };
    std::vector<int> children;
    std::vector<OrxMove> unvisited;
    std::vector<int> unvisitedRootIndex;
};

}  // namespace

struct OrxTree {
    Static S;
    OrxTreeOptions O;
    OrxCtx *ctx = nullptr;
    unsigned long long games = 0;   // global game counter: every playout of every search gets its own stream
    std::vector<Node> nodes;
};

extern "C" void orx_tree_default_options(OrxTreeOptions *o)
{
    memset(o, 0, sizeof *o);
    o->placement_partial_order = 1; o->attack_aggressive = 1; o->attack_repeat_moves = 1; o->attack_partial_order = 0;
    o->attack_until_dead = 1; o->invasion = 1; o->fortification_outwards = 1;
    o->agent_id = 0; o->playouts_per_leaf = 8; o->leaves_in_flight = 1024; o->max_child_policy = 0; o->seed = 1;
}

extern "C" int orx_tree_create(const OrxMap *map, const OrxRules *rules, const OrxTreeOptions *opt, OrxCtx *ctx, OrxTree **out)
{
    if (!map || !rules || !out) return ORX_E_INVALID;
    const int N = map->n_territories, C = map->n_continents, P = map->n_players;
    if (N < 1 || N > ORX_MAX_TERRITORIES || P < 2 || P > MAXP || C < 1 || C > MAXC) return ORX_E_INVALID;
    OrxTree *t = new OrxTree();
    Static &S = t->S;
    S.N = N; S.C = C; S.P = P;
    S.adj_off.assign(map->adj_offsets, map->adj_offsets + N + 1);
    S.adj.assign(map->adj, map->adj + map->adj_offsets[N]);
    S.continent.assign(map->continent, map->continent + N);
    S.base_bonus.assign(map->continent_bonus, map->continent_bonus + C);
    S.symbol.resize(N);
    for (int i = 0; i < N; i++) S.symbol[i] = map->card_symbol ? map->card_symbol[i] : i % 3;
    S.use_cards = rules->use_cards != 0; S.transfer_cards = rules->transfer_cards != 0; S.cont_inc = rules->continent_increase;
    S.progression = rules->card_progression ? rules->card_progression : "";
    S.can_run_out = P * 5 > N + 2;
    S.L = orx_leaf_layout(N, P);
    if (opt) t->O = *opt; else orx_tree_default_options(&t->O);
    if (t->O.playouts_per_leaf < 1) t->O.playouts_per_leaf = 1;
    if (t->O.leaves_in_flight < 1) t->O.leaves_in_flight = 1;
    t->ctx = ctx;
    *out = t;
    return ORX_OK;
}

extern "C" void orx_tree_destroy(OrxTree *t) { delete t; }

extern "C" int orx_tree_possible_moves(OrxTree *t, const void *leaf, const void *extra, OrxMove *out, int cap)
{
    if (!t || !leaf || (!out && cap > 0)) return ORX_E_INVALID;
    Board b(t->S, t->O);
    if (!b.setup((const uint8_t *)leaf, (const uint8_t *)extra)) return ORX_E_INVALID;
    std::vector<OrxMove> mv = b.possible_moves();
    if (b.error) return ORX_E_INVALID;
    for (int i = 0; i < (int)mv.size() && i < cap; i++) out[i] = mv[i];
    return (int)mv.size();
}

extern "C" int orx_tree_apply(OrxTree *t, const void *leaf, const void *extra, const OrxMove *move, void *leaf_out,
                              void *extra_out, int32_t *ongoing)
{
    if (!t || !leaf || !move || !leaf_out) return ORX_E_INVALID;
    Board b(t->S, t->O);
    if (!b.setup((const uint8_t *)leaf, (const uint8_t *)extra)) return ORX_E_INVALID;
    bool on = b.update(*move);
    if (b.error) return ORX_E_INVALID;
    b.pack((uint8_t *)leaf_out, (uint8_t *)extra_out);
    if (ongoing) *ongoing = on;
    return ORX_OK;
}

extern "C" int orx_new_game(const OrxMap *map, uint64_t seed, void *leaf_out)
{
    if (!map || !leaf_out) return ORX_E_INVALID;
    const int N = map->n_territories, P = map->n_players;
    if (N < 1 || N > ORX_MAX_TERRITORIES || P < 2 || P > MAXP) return ORX_E_INVALID;
    OrxLeafLayout L = orx_leaf_layout(N, P);
    uint8_t *leaf = (uint8_t *)leaf_out;
    memset(leaf, 0, (size_t)L.stride);
    std::vector<int> ccs(N);
    for (int i = 0; i < N; i++) ccs[i] = i;
    JavaRandom rnd; rnd.set_seed(seed);
    for (int i = N; i > 1; i--) std::swap(ccs[i - 1], ccs[rnd.next_int(i)]);   // Collections.shuffle(list, rnd)
    uint8_t *ow = leaf + L.off_owner; int32_t *ar = (int32_t *)(leaf + L.off_armies); uint8_t *cards = leaf + L.off_cards;
    int player = 0, owned0 = 0;
    for (int cc : ccs) { ow[cc] = (uint8_t)player; ar[cc] = 1; owned0 += player == 0; player = (player + 1) % P; }
    double f1 = N / 42.0; f1 -= 1.0; f1 /= 2.0; f1 += 1.0;
    int left0 = (int)floor(f1 * (50 - P * 5) + 0.5) - owned0;             // calculateStartingArmies()[0] - playerCountries[0]
    OrxLeafHeader *h = (OrxLeafHeader *)leaf;
    h->turn_count = 0; h->placeable = left0 == 5 ? 5 : (left0 < 4 ? left0 : 4);
    h->attacking_cc = -1; h->defending_cc = -1; h->card_pos = 0; h->current_player = 0; h->phase = ORX_PHASE_INITIAL_PLACEMENT;
    h->attack_state = ORX_ATTACK_START; h->defending_player = -1; h->attack_until_dead = 1;
    h->n_hand = 0; h->n_deck = (uint8_t)(N + 2);
    for (int i = 0; i < N; i++) cards[i] = (uint8_t)i;                      // CardManager's notInHand: every card, then two wildcards
    cards[N] = ORX_CARD_WILD; cards[N + 1] = ORX_CARD_WILD;
    h->deal_mode = ORX_DEAL_LEAF_LCG; h->deal_seed = orx_deck_hash(cards, N + 2);
    return ORX_OK;
}

extern "C" int orx_tree_choose(const OrxTree *t, const OrxRootChild *ch, int n, uint64_t tie_break)
{
    if (!t || !ch || n <= 0) return ORX_E_INVALID;
    const int id = t->O.agent_id;
    std::vector<int> best;
    if (!t->O.max_child_policy) {  // chooseRobustMove (MCTSTree.java:541-568)
        long long bv = -1; double bl = 2147483647.0;
        for (int i = 0; i < n; i++) {
            double wl = ch[i].wins[id] ? (double)ch[i].win_len_sum[id] / (double)ch[i].wins[id] : 0.0;
            if (ch[i].visits > bv || (ch[i].visits == bv && wl < bl)) { bv = ch[i].visits; bl = wl; best.clear(); best.push_back(i); }
            else if (ch[i].visits == bv && wl == bl) best.push_back(i);
        }
    } else {  // chooseMaxMove (:570-595)
        double br = 4.9e-324, bl = 2147483647.0;
        for (int i = 0; i < n; i++) {
            if (!ch[i].visits) continue;
            double r = (double)ch[i].wins[id] / (double)ch[i].visits;
            double wl = ch[i].wins[id] ? (double)ch[i].win_len_sum[id] / (double)ch[i].wins[id] : 0.0;
            if (r > br || (r == br && wl < bl)) { br = r; bl = wl; best.clear(); best.push_back(i); }
            else if (r == br && wl == bl) best.push_back(i);
        }
    }
    if (best.empty()) return 0;
    SplitMix sm{ tie_break };
    return best[sm.below((int)best.size())];
}

static bool same_move(const OrxMove &x, const OrxMove &y)
{
    return x.type == y.type && x.phase == y.phase && x.a == y.a && x.b == y.b && x.c == y.c && x.d == y.d && x.flag == y.flag;
}

extern "C" int orx_tree_search(OrxTree *t, const void *root_leaf, const void *extra, int iterations,
                               OrxRootChild *children, int cap, int32_t *n_children, OrxSearchStats *stats)
{
    if (!t || !root_leaf || iterations < 0) return ORX_E_INVALID;
    if (!t->ctx) return ORX_E_CUDA;  // no CPU path: the playouts run in the CUDA engine
    Static &S = t->S;
    const OrxTreeOptions &O = t->O;
    const int P = S.P, ppl = O.playouts_per_leaf, stride = S.L.stride;
    const double Cx = 1.0, A = 0.05, EPS = 1e-10;  // MCTSTree.java:23-27
    SplitMix rng{ O.seed ^ 0xA5A5A5A5DEADBEEFULL };
    OrxSearchStats st; memset(&st, 0, sizeof st);
    using clk = std::chrono::steady_clock;
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };

    std::vector<Node> &nodes = t->nodes;
    nodes.clear();
    Board board(S, O);
    if (!board.setup((const uint8_t *)root_leaf, (const uint8_t *)extra)) { orx_internal_set_error("orx_tree_search: the root leaf record is invalid"); return ORX_E_INVALID; }
    nodes.emplace_back();
    nodes[0].currentPlayer = O.agent_id;  // new SearchNode(agentID, numberOfPlayers) (:163)
    nodes[0].unvisited = board.possible_moves();
    if (board.error) return ORX_E_INVALID;
    const std::vector<OrxMove> rootMoves = nodes[0].unvisited;
    for (int i = 0; i < (int)rootMoves.size(); i++) nodes[0].unvisitedRootIndex.push_back(i);

    struct Pending { std::vector<int> path; bool terminal; int pot[MAXP]; int finalTurn; };
    // Two batches in flight (orx_rollout_begin / _end on two streams): while the GPU plays one, the host selects the
    // next under the virtual loss of both; a batch's tail (it ends with its longest game) overlaps the next one.
    struct Batch { std::vector<uint8_t> leaves; std::vector<Pending> pend; int nleaf = 0; bool outstanding = false; clk::time_point t_begin; };
    Batch batch[2];
    std::vector<OrxGameResult> games;
    const int slots = O.leaves_in_flight >= 2 ? 2 : 1;
    const int per_batch = std::max(1, O.leaves_in_flight / slots);
    // Every early exit below (an illegal board, a failed launch, a failed copy) must leave the context usable: batches
    // still in flight are ended (their results dropped) so that their slots are free for the next call.
    struct Drain {
        OrxCtx *ctx; Batch *b; bool armed = true;
        ~Drain() { if (armed) for (int k = 0; k < 2; k++) if (b[k].outstanding && b[k].nleaf) { orx_rollout_end(ctx, k, nullptr, nullptr); b[k].outstanding = false; } }
    } drain{ t->ctx, batch };
    auto fail = [&](int code, const char *why) { orx_internal_set_error(why); return code; };

    auto finish = [&](int sl) -> int {
        Batch &B = batch[sl];
        auto t1 = clk::now();
        if (B.nleaf) {
            games.resize((size_t)B.nleaf * ppl);
            B.outstanding = false;   // ended (successfully or not): nothing left to drain for this slot
            int rc = orx_rollout_end(t->ctx, sl, nullptr, games.data());
            if (rc != ORX_OK) return rc;
        }
        auto t2 = clk::now();
        // --- back-propagation (:399-431)
        int li = 0;
        for (Pending &pd : B.pend) {
            for (int j = 0; j < ppl; j++) {
                const int *pot; int finalTurn;
                if (pd.terminal) { pot = pd.pot; finalTurn = pd.finalTurn; }
                else {
                    const OrxGameResult &g = games[(size_t)li * ppl + j];
                    if (g.flags & ORX_FLAG_ERROR) continue;  // an illegal leaf: nothing to learn from it
                    pot = g.player_out_on_turn; finalTurn = g.sim_turn_count;
                }
                for (int n : pd.path) {
                    Node &nd = nodes[n];
                    for (int p = 0; p < P; p++) {
                        if (pot[p] == -1) { nd.wins[p]++; nd.winLen[p] += finalTurn; } else { nd.losses[p]++; nd.lossLen[p] += finalTurn; }
                    }
                    nd.visits++;
                }
            }
            for (int n : pd.path) nodes[n].inflight--;
            if (!pd.terminal) li++;
        }
        B.outstanding = false;
        auto t3 = clk::now();
        st.rollout_ms += ms(t1, t2); st.backprop_ms += ms(t2, t3);
        return ORX_OK;
    };

    int done = 0, scheduled = 0, sl = 0;
    for (;;) {
        if (batch[sl].outstanding) { int rc = finish(sl); if (rc != ORX_OK) return rc; done += (int)batch[sl].pend.size(); }
        const bool forced = nodes[0].children.size() == 1 && nodes[0].unvisited.empty();   // (:179)
        if (scheduled < iterations && !forced) {
        auto t0 = clk::now();
        Batch &B = batch[sl];
        int want = std::min(per_batch, iterations - scheduled);
        B.leaves.clear(); B.pend.clear();
        int nleaf = 0;
        std::vector<uint8_t> &leaves = B.leaves;
        std::vector<Pending> &pend = B.pend;
        for (int b = 0; b < want; b++) {
            // --- selection (:186-198) under virtual loss
            board.setup((const uint8_t *)root_leaf, (const uint8_t *)extra);
            Pending pd; pd.terminal = false; pd.path.push_back(0);
            int cur = 0; bool ongoing = true;
            while (ongoing) {
                Node &nd = nodes[cur];
                if (!nd.unvisited.empty() || nd.children.empty()) break;
                int pick = -1;
                long long nvis = nd.visits + nd.inflight * ppl;
                if (nd.move.type == ORX_MOVE_ATTACK && cur != 0) {  // stochasticSelection (:226-246)
                    double bs = -2147483648.0;
                    for (int c : nd.children) {
                        const Node &ch = nodes[c];
                        double sc = ch.move.probability;
                        if (nvis != 0) sc -= (double)(ch.visits + ch.inflight * ppl) / (double)nvis;
                        if (sc > bs) { bs = sc; pick = c; }
                    }
                } else {  // deterministicSelection (:248-331)
                    double logv = log((double)nvis + 1.0);
                    double mxW = -INFINITY, mnW = INFINITY, mxL = -INFINITY, mnL = INFINITY;
                    for (int c : nd.children) {
                        const Node &ch = nodes[c]; int cp = ch.currentPlayer;
                        double aw = ch.wins[cp] ? (double)ch.winLen[cp] / (double)ch.wins[cp] : 0.0;
                        double al = ch.losses[cp] ? (double)ch.lossLen[cp] / (double)ch.losses[cp] : 0.0;
                        mxW = std::max(mxW, aw); mnW = std::min(mnW, aw); mxL = std::max(mxL, al); mnL = std::min(mnL, al);
                    }
                    double dW = mxW - mnW, dL = mxL - mnL;
                    double a = A * std::min(std::max((int)nd.children.size() - 2, 0), 100) / 100.0;
                    double bs = -INFINITY;
                    for (int c : nd.children) {
                        const Node &ch = nodes[c]; int cp = ch.currentPlayer;
                        long long cv = ch.visits + ch.inflight * ppl;   // playouts in flight count as visits without wins
                        double wp = cv ? (double)ch.wins[cp] / (double)cv : 0.0;
                        double aw = ch.wins[cp] ? (double)ch.winLen[cp] / (double)ch.wins[cp] : 0.0;
                        double al = ch.losses[cp] ? (double)ch.lossLen[cp] / (double)ch.losses[cp] : 0.0;
                        double ws = (1 - a) * wp, wls = 0, lls = 0;
                        if (wp != 0.0) { wls = a * wp; if (dW != 0.0) wls *= (mxW - aw) / dW; }
                        if (wp != 1.0) { lls = a * (1 - wp); if (dL != 0.0) lls *= (al - mnL) / dL; }
                        double sc = Cx * sqrt(logv / ((double)cv + 1.0)) + ws + wls + lls + rng.unit() * EPS;
                        if (sc > bs) { bs = sc; pick = c; }
                    }
                }
                if (pick < 0) break;
                pd.path.push_back(pick);
                ongoing = board.update(nodes[pick].move);
                cur = pick;
            }
            if (board.error) return fail(ORX_E_INVALID, "orx_tree_search: a move of the selection path is illegal on its board");
            if (ongoing && !nodes[cur].unvisited.empty()) {
                // --- expansion (:335-355)
                int k = rng.below((int)nodes[cur].unvisited.size());
                OrxMove mv = nodes[cur].unvisited[k];
                int ri = cur == 0 ? nodes[cur].unvisitedRootIndex[k] : -1;
                nodes[cur].unvisited.erase(nodes[cur].unvisited.begin() + k);
                if (cur == 0) nodes[cur].unvisitedRootIndex.erase(nodes[cur].unvisitedRootIndex.begin() + k);
                Node nn; nn.move = mv; nn.parent = cur; nn.depth = nodes[cur].depth + 1; nn.rootIndex = ri;
                nn.currentPlayer = mv.type == ORX_MOVE_NEXT_PHASE ? mv.c : nodes[cur].currentPlayer;
                ongoing = board.update(mv);
                if (board.error) return fail(ORX_E_INVALID, "orx_tree_search: the expanded move is illegal on its board");
                if (ongoing) { nn.unvisited = board.possible_moves(); if (board.error || nn.unvisited.empty()) return fail(ORX_E_INVALID, "orx_tree_search: an ongoing position has no legal move"); }
                int id = (int)nodes.size();
                nodes.push_back(nn);
                nodes[cur].children.push_back(id);
                pd.path.push_back(id);
                st.max_depth = std::max<int64_t>(st.max_depth, nn.depth);
            }
            if (!ongoing) {
                // the game ended inside the tree: score the finished position itself (the reference re-reads the
                // previous iteration's boards here -- SURVEY 8.1 quirk 16 -- which is not reproduced)
                pd.terminal = true; pd.finalTurn = 0;  // lengths are counted from the leaf (getSimTurnCount), and this leaf is the end
                for (int p = 0; p < P; p++) pd.pot[p] = board.playerOutOnTurn[p];
                st.terminal_hits++;
            } else {
                leaves.resize((size_t)(nleaf + 1) * stride);
                board.pack(leaves.data() + (size_t)nleaf * stride, nullptr);
                nleaf++;
            }
            for (int n : pd.path) nodes[n].inflight++;
            pend.push_back(std::move(pd));
        }
        auto t1 = clk::now();
        st.select_ms += ms(t0, t1);
        // --- simulation (:359-395): every leaf of the batch x playouts_per_leaf playouts, one launch
        B.nleaf = nleaf;
        if (nleaf) {
            int rc = orx_rollout_begin(t->ctx, sl, leaves.data(), nleaf, ppl, O.seed, t->games, 1);
            if (rc != ORX_OK) return rc;
            t->games += (unsigned long long)nleaf * ppl;
            st.batches++;
        }
        B.outstanding = true;
        scheduled += (int)pend.size();
        } else if (!batch[sl ^ 1].outstanding) {
            break;
        }
        if (slots == 2) sl ^= 1;
    }
    drain.armed = false;
    st.iterations = done; st.playouts = (int64_t)done * ppl; st.nodes = (int64_t)nodes.size();
    if (stats) *stats = st;
    // root children in getPossibleMoves() order
    int n = (int)rootMoves.size();
    if (n_children) *n_children = n;
    for (int i = 0; i < n && i < cap; i++) { memset(&children[i], 0, sizeof children[i]); children[i].move = rootMoves[i]; }
    for (int c : nodes[0].children) {
        const Node &nd = nodes[c];
        int i = nd.rootIndex;
        if (i < 0 || i >= cap || !same_move(nd.move, rootMoves[i])) continue;
        children[i].visits = nd.visits;
        for (int p = 0; p < P; p++) { children[i].wins[p] = nd.wins[p]; children[i].losses[p] = nd.losses[p]; children[i].win_len_sum[p] = nd.winLen[p]; children[i].loss_len_sum[p] = nd.lossLen[p]; }
    }
    return ORX_OK;
}
