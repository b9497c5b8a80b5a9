// orx_game.cuh -- one complete playout (setupState + simulate + getters) by one warp.
//
// Restates, in warp-uniform control flow with lane-parallel scans (O/ = /root/reference/src/com/mld46/oponn/):
//   O/sim/boards/SimulationBoard.java:178-359 (setupState), :365-481 (simulate), :485-1015 (rules),
//   :1076-1111 (income, next player);  O/sim/boards/ModellingBoard.java:21-47 (no observable effect for an
//   all-SimRandom line-up);  O/sim/agents/SimRandom.java (whole: the playout policy);
//   O/CardManager.java:240-346 (sim deck);  Card set semantics: declared assumption, see orx_state.h.
//
// Layout: territory t lives in lane t%32 of scan word t/32.  owner[] / armies[] / the ordered Java lists
// (possibleAttackers, the hands, the sim deck) are byte / int arrays in this warp's shared-memory slice;
// per-player counters (playerCountries, playerOutOnTurn, list sizes, initialArmiesLeftToPlace) are
// lane-distributed registers (lane p <-> player p); everything else is a warp-uniform scalar.
#pragma once

#include "orx_device.cuh"

namespace orx {

struct RolloutArgs {
    const uint8_t *leaves;        // n_leaves records, dm.leaf_stride apart
    int32_t n_leaves, playouts_per_leaf;
    unsigned long long total;     // n_leaves * playouts_per_leaf (replay: n_games)
    unsigned long long seed, first_game;
    OrxLeafResult *agg;           // [n_leaves] or null
    OrxGameResult *games;         // [total] or null
    unsigned long long *ticket;   // work counter (zeroed before launch)
    const uint32_t *words;        // replay: explicit streams
    uint32_t words_per_game;
    uint32_t rk[20];              // Philox round keys of `seed`: rk[2r] = seed_lo + r * 0x9E3779B9, rk[2r+1] = seed_hi + r * 0xBB67AE85
};

// The hand positions cashCards really removes for the triple i0 < i1 < i2 (packed i0 | i1 << 8 | i2 << 16) of the n-card hand
// at h: List.remove(Object) takes the first element EQUAL to the card, and the two wildcards are one object
// (CardManager.java:14,62-63) -- a triple with exactly one wildcard that is not the hand's first wildcard removes the first
// one instead.  Out of line on purpose: a once-per-cash fix whose registers must not weigh on the kernel's hot loops
// (inlined it cost 7 % of the whole kernel).
static __device__ __noinline__ uint32_t cash_indices(saddr h, int n, uint32_t pk)
{
    const int lane = lane_id();
    int i0 = (int)(pk & 0xFFu), i1 = (int)((pk >> 8) & 0xFFu), i2 = (int)(pk >> 16);
    int fw = -1;   // position of the hand's first wildcard
    for (int base = 0; base < n && fw < 0; base += 32) {
        const int i = base + lane;
        const uint32_t m = __ballot_sync(ORX_FULL, i < n && lds8(h + (uint32_t)(i < n ? i : 0)) == ORX_CARD_WILD);
        if (m) fw = base + __ffs(m) - 1;
    }
    if (fw < 0 || fw == i0 || fw == i1 || fw == i2) return pk;
    // exactly one wildcard in the triple and it is the later one: fw replaces it (fw is smaller than the index it replaces)
    if (lds8(h + (uint32_t)i0) == ORX_CARD_WILD) i0 = fw; else if (lds8(h + (uint32_t)i1) == ORX_CARD_WILD) i1 = fw; else i2 = fw;
    int t;
    if (i1 < i0) { t = i0; i0 = i1; i1 = t; }
    if (i2 < i1) { t = i1; i1 = i2; i2 = t; }
    if (i1 < i0) { t = i0; i0 = i1; i1 = t; }
    return (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)i2 << 16);
}

// nextInt sites whose bound is a list length or a neighbour count (< ORX_MAGIC_N): the general division is compiled out there
#ifdef ORX_V_SMALLINT_OFF
#define ORX_NI_SMALL false
#else
#define ORX_NI_SMALL true
#endif

template <int NWT, bool REPLAY, bool MODEL = false>
struct Game {
    const DevMap &dm;
    // 32-bit shared-window addresses: the static map image (per block) and this warp's slice
    saddr s_adj, s_deg, s_sym, s_nbr, s_cont, s_bonus;
    saddr a_arm, a_own, a_alist, a_cards, a_hdr;   // (the warp's running OrxLeafResult sums sit right behind the header)
    saddr a_move;             // MODEL: Country.moveableArmies, int32 per territory
    uint32_t agents;          // MODEL: ORX_AGENT_* of player p in bits 3p..3p+2 (simAgents[], ModellingBoard.java)
    Stream<REPLAY> s;
    const uint32_t *rk;       // RolloutArgs::rk (constant bank)
    int lane;

    // warp-uniform scalars (SimulationBoard.java:62-96)
    int cp, phase, simTurn, remaining, placeable, cardpos, hasInvaded, playerTurns;
    int attackState, attackingCC, defendingCC, attackingPlayer, defendingPlayer, attackUntilDead;
    uint32_t termPos; bool termSet;
    // lane-distributed: lane p = player p (list index P = the sim deck for ncard)
    int pc, pot, ncard, ialp;

    __device__ __forceinline__ Game(const DevMap &d) : dm(d) {}

    __device__ __forceinline__ void fail() { s.flags |= ORX_FLAG_ERROR; }

    // ---- static map accessors --------------------------------------------------------
    __device__ __forceinline__ int adj(int t, int j) const { return (int)lds8(s_adj + (uint32_t)(t * dm.degp + j)); }
    __device__ __forceinline__ int deg(int t) const { return (int)lds8(s_deg + (uint32_t)t); }
    __device__ __forceinline__ uint32_t symbit(int card) const { return card == (int)ORX_CARD_WILD ? 8u : lds8(s_sym + (uint32_t)card); }
    __device__ __forceinline__ uint32_t nbr(int w, int t) const { return lds32(s_nbr + 4u * (uint32_t)(w * (32 * NWT) + t)); }
    __device__ __forceinline__ uint32_t contmask(int c, int w) const { return lds32(s_cont + 4u * (uint32_t)(c * NWT + w)); }

    // ---- shared-memory helpers (one writer lane, fenced on both sides) ------------------
    __device__ __forceinline__ int armies(int t) const { return (int)lds32(a_arm + 4u * (uint32_t)t); }
    __device__ __forceinline__ int owner(int t) const { return (int)lds8(a_own + (uint32_t)t); }
    __device__ __forceinline__ uint32_t hdr(int i) const { return lds32(a_hdr + 4u * (uint32_t)i); }
    __device__ __forceinline__ void set_hdr(int i, uint32_t v) { __syncwarp(); if (lane == 0) sts32(a_hdr + 4u * (uint32_t)i, v); __syncwarp(); }
    __device__ __forceinline__ void set_armies(int t, int v) { __syncwarp(); if (lane == 0) sts32(a_arm + 4u * (uint32_t)t, (uint32_t)v); __syncwarp(); }
    __device__ __forceinline__ void set_owner(int t, int v) { __syncwarp(); if (lane == 0) sts8(a_own + (uint32_t)t, (uint32_t)v); __syncwarp(); }
    template <typename T> __device__ __forceinline__ T from_lane(T v, int l) const { return __shfl_sync(ORX_FULL, v, l); }

    // ---- ordered lists (java.util.List<Card>): list li < P is player li's hand, li == P the sim deck ----
    __device__ __forceinline__ saddr list(int li) const { return a_cards + (uint32_t)(li * dm.cap); }
    __device__ __forceinline__ int list_len(int li) const { return from_lane(ncard, li); }
    __device__ __forceinline__ void list_add(int li, int card)
    {
        int n = list_len(li);
        if (n >= dm.cap) { fail(); return; }
        __syncwarp();
        if (lane == 0) sts8(list(li) + (uint32_t)n, (uint32_t)card);
        if (lane == li) ncard++;
        __syncwarp();
    }
    // remove(int index) on a byte list of length n held at p
    __device__ __forceinline__ int bytes_remove_at(saddr p, int n, int i)
    {
        int c = (int)lds8(p + (uint32_t)i);
        for (int base = i; base < n - 1; base += 32) {
            int idx = base + lane;
            uint32_t v = idx < n - 1 ? lds8(p + (uint32_t)idx + 1u) : 0u;
            __syncwarp();
            if (idx < n - 1) sts8(p + (uint32_t)idx, v);
            __syncwarp();
        }
        return c;
    }
    __device__ __forceinline__ int list_remove_at(int li, int i)
    {
        int n = list_len(li);
        int c = bytes_remove_at(list(li), n, i);
        if (lane == li) ncard--;
        return c;
    }

    // ---- board scans ------------------------------------------------------------------
    // bitmask of the territories `player` owns, one word per scan word
    __device__ __forceinline__ void owned_mask(int player, uint32_t (&m)[NWT]) const
    {
#pragma unroll
        for (int w = 0; w < NWT; w++) m[w] = __ballot_sync(ORX_FULL, owner(w * 32 + lane) == player);
    }
    // owned territories with at least one out-neighbour not owned (SimRandom.java:68-82,250-262)
    __device__ __forceinline__ void border_mask(const uint32_t (&mine)[NWT], bool need_two, uint32_t (&b)[NWT]) const
    {
#pragma unroll
        for (int w = 0; w < NWT; w++) {
            int t = w * 32 + lane;
            uint32_t enemy = 0;
#pragma unroll
            for (int w2 = 0; w2 < NWT; w2++) enemy |= nbr(w2, t) & ~mine[w2];
            bool f = ((mine[w] >> lane) & 1u) && enemy != 0u && (!need_two || armies(t) > 1);
            b[w] = __ballot_sync(ORX_FULL, f);
        }
    }
    // territory index of the n-th set bit (code order) of a multi-word mask
    __device__ __forceinline__ int nth_territory(const uint32_t (&m)[NWT], int n) const
    {
        int res = -1;
#pragma unroll
        for (int w = 0; w < NWT; w++) {
            int c = __popc(m[w]);
            if (res < 0 && n < c) res = w * 32 + nth_bit(m[w], n);
            n -= c;
        }
        return res;
    }
    __device__ __forceinline__ int count_bits(const uint32_t (&m)[NWT]) const
    {
        int c = 0;
#pragma unroll
        for (int w = 0; w < NWT; w++) c += __popc(m[w]);
        return c;
    }

    // ---- cards ------------------------------------------------------------------------
    // DECLARED ASSUMPTION (Lux SDK Card.isASet not vendored, orx_state.h): wildcard, or all equal, or all distinct
    static __device__ __forceinline__ bool is_set_bits(uint32_t m) { return (m & 8u) || __popc(m) != 2; }

    // number of valid triples i<j<k of the current player's hand; when pick >= 0 also returns, in idx[],
    // the pick-th one in lexicographic order (Card.containsASet / Card.getRandomSet)
    __device__ __forceinline__ int scan_sets(int n, int pick, int idx[3])
    {
        saddr h = list(cp);
        // symbol bits of the hand, staged in the (idle) attackers-list buffer
        __syncwarp();
        for (int base = 0; base < n; base += 32) { int i = base + lane; if (i < n) sts8(a_alist + (uint32_t)i, symbit((int)lds8(h + (uint32_t)i))); }
        __syncwarp();
        int count = 0;
        for (int i = 0; i < n - 2; i++) {
            uint32_t si = lds8(a_alist + (uint32_t)i);
            for (int j = i + 1; j < n - 1; j++) {
                uint32_t sij = si | lds8(a_alist + (uint32_t)j);
                for (int kb = j + 1; kb < n; kb += 32) {
                    int k = kb + lane;
                    bool ok = k < n && is_set_bits(sij | lds8(a_alist + (uint32_t)(k < n ? k : 0)));
                    uint32_t m = __ballot_sync(ORX_FULL, ok);
                    int c = __popc(m);
                    if (pick >= count && pick < count + c) { idx[0] = i; idx[1] = j; idx[2] = kb + nth_bit(m, pick - count); }
                    count += c;
                }
            }
        }
        return count;
    }

    // CardManager.simDeal (O/CardManager.java:250-263); returns -1 when the deck is legitimately empty
    __device__ __forceinline__ int sim_deal()
    {
        int n = list_len(dm.P);
        if (n == 0) { if (!dm.can_run_out) fail(); return -1; }
        return list_remove_at(dm.P, hdr(WH_DEAL_MODE) == (uint32_t)ORX_DEAL_GAME_STREAM ? s.next_int(n) : deal_next_int(n));
    }
    // CardManager.random = new Random(simDeck.hashCode()) (CardManager.java:245): java.util.Random's own LCG, private to
    // the playout, state in the warp header (one deal per player-turn at most)
    __device__ __forceinline__ int deal_next31()
    {
        uint64_t st = ((uint64_t)hdr(WH_DEAL_HI) << 32) | (uint64_t)hdr(WH_DEAL_LO);
        st = (st * 0x5DEECE66Dull + 0xBull) & ((1ull << 48) - 1ull);
        __syncwarp();
        if (lane == 0) { sts32(a_hdr + 4u * WH_DEAL_LO, (uint32_t)st); sts32(a_hdr + 4u * WH_DEAL_HI, (uint32_t)(st >> 32)); }
        __syncwarp();
        return (int)(st >> 17);
    }
    __device__ __forceinline__ int deal_next_int(int bound)   // bound > 0
    {
        int r = deal_next31();
        const int m = bound - 1;
        if ((bound & m) == 0) return (int)(((long long)bound * (long long)r) >> 31);
        for (int u = r;; u = deal_next31()) {
            r = u % bound;
            if ((int)((uint32_t)u - (uint32_t)r + (uint32_t)m) >= 0) return r;
        }
    }

    // SimulationBoard.cashCards (SimulationBoard.java:577-618) for hand positions i<j<k
    __device__ __forceinline__ void cash_cards(const int idx[3])
    {
        saddr h = list(cp);
        int c0 = (int)lds8(h + (uint32_t)idx[0]), c1 = (int)lds8(h + (uint32_t)idx[1]), c2 = (int)lds8(h + (uint32_t)idx[2]);
        // currentCards.remove(card) x3 (:596-598): List.remove(Object) takes out the FIRST element equal to the card.
        // Territory cards are unique; the two wildcards are one object, so a triple that holds the hand's SECOND wildcard
        // but not the first removes the first one (cash_indices, out of line); then remove by position from the back.
        uint32_t pk = (uint32_t)idx[0] | ((uint32_t)idx[1] << 8) | ((uint32_t)idx[2] << 16);
        if (c0 == (int)ORX_CARD_WILD || c1 == (int)ORX_CARD_WILD || c2 == (int)ORX_CARD_WILD)
            pk = __shfl_sync(ORX_FULL, cash_indices(h, list_len(cp), pk), 0);   // (the shuffle tells the PTX pass the value is warp-uniform)
        list_remove_at(cp, (int)(pk >> 16)); list_remove_at(cp, (int)((pk >> 8) & 0xFFu)); list_remove_at(cp, (int)(pk & 0xFFu));
        list_add(dm.P, c0); list_add(dm.P, c1); list_add(dm.P, c2);  // CardManager.simCash :265-276
        placeable += card_cash_value(dm, cardpos, s.flags);
        cardpos++;
        int cs[3] = { c0, c1, c2 };
#pragma unroll
        for (int q = 0; q < 3; q++)
            if (cs[q] != (int)ORX_CARD_WILD && owner(cs[q]) == cp) set_armies(cs[q], armies(cs[q]) + 2);
    }

    // SimRandom.cardsPhase (SimRandom.java:36-49)
    __device__ __forceinline__ void agent_cards_phase()
    {
        int n = list_len(cp);
        if (n < 3) return;
        int idx[3] = { 0, 0, 0 };
        int count = scan_sets(n, -1, idx);
        if (count == 0) return;
        uint64_t k = s.next_k53();  // r < 0.5 <=> k < 2^52 ; r < 0.75 <=> k < 3 * 2^51
        if ((n == 3 && k < (1ull << 52)) || (n == 4 && k < (3ull << 51)) || n >= 5) {
            int pick = s.next_int(count);
            scan_sets(n, pick, idx);
            cash_cards(idx);
        }
    }

    // SimulationBoard.tearDownCardsPhase (:620-638)
    __device__ __forceinline__ void tear_down_cards_phase()
    {
        for (;;) {
            int n = list_len(cp);
            if (n <= 4 || s.flags) break;
            int idx[3] = { 0, 0, 0 };
            int count = scan_sets(n, -1, idx);
            if (count == 0) { fail(); break; }
            int pick = s.next_int(count);
            scan_sets(n, pick, idx);
            cash_cards(idx);
        }
        phase = ORX_PHASE_PLACEMENT;
    }

    // ---- placement --------------------------------------------------------------------
    // SimRandom.place (SimRandom.java:247-277) + SimulationBoard.placeArmies (:647-684)
    __device__ __forceinline__ void agent_place(int n_armies)
    {
        uint32_t mine[NWT], cand[NWT];
        owned_mask(cp, mine);
        border_mask(mine, false, cand);
        int n = count_bits(cand);
        while (n_armies > 0 && !s.flags) {
            int k = 1 + s.next_int(n_armies);
            int i = s.template next_int<ORX_NI_SMALL>(n);
            if (s.flags) return;
            int t = nth_territory(cand, i);
            set_armies(t, armies(t) + k);
            placeable -= k;
            if (phase == ORX_PHASE_INITIAL_PLACEMENT && lane == cp) ialp -= k;
            n_armies -= k;
        }
    }

    // SimulationBoard.getPlayerIncome(currentPlayer) (:1076-1090, quirk 3 is moot: player == currentPlayer)
    __device__ __forceinline__ int player_income()
    {
        int mycount = from_lane(pc, cp);
        if (mycount == 0) return 0;
        uint32_t mine[NWT];
        owned_mask(cp, mine);
        int income = mycount / 3; if (income < 3) income = 3;
        int bonus = 0;
        if (lane < dm.C) {
            uint32_t missing = 0;
#pragma unroll
            for (int w = 0; w < NWT; w++) missing |= contmask(lane, w) & ~mine[w];
            if (missing == 0u) bonus = current_bonus(lane);
        }
        return (int)((uint32_t)income + __reduce_add_sync(ORX_FULL, (unsigned)bonus));
    }

    // currentContinentBonuses[c] (recalculateContinentBonuses :956-963), rows precomputed on the host
    __device__ __forceinline__ int current_bonus(int c) const
    {
        if (dm.cont_inc == 0) return (int)lds32(s_bonus + 4u * (uint32_t)c);
        int e = (int)hdr(WH_SIM_START) + simTurn - 1; if (e < 0) e = 0;
        return __ldg(dm.bonus_table + e * dm.C + c);
    }
    __device__ __forceinline__ void recalculate_bonuses()
    {
        int e = (int)hdr(WH_SIM_START) + simTurn - 1;
        if (dm.cont_inc != 0 && e >= ORX_BONUS_ROUNDS) fail();  // engine table limit
    }

    // ---- attack -----------------------------------------------------------------------
    // SimulationBoard.playerDefeated (:843-867)
    __device__ __forceinline__ void player_defeated(int player, int conquerer)
    {
        remaining--;
        if (lane == player) pot = simTurn;
        int dst = dm.transfer_cards ? conquerer : dm.P;  // hand -> conqueror, or simReturnToDeck
        if (dst < 0) { fail(); return; }
        if (dst != player) {
            int nsrc = list_len(player), ndst = list_len(dst);
            if (ndst + nsrc > dm.cap) { fail(); return; }
            saddr src = list(player), d = list(dst);
            __syncwarp();
            for (int base = 0; base < nsrc; base += 32) { int i = base + lane; if (i < nsrc) sts8(d + (uint32_t)(ndst + i), lds8(src + (uint32_t)i)); }
            __syncwarp();
            if (lane == dst) ncard += nsrc;
            if (lane == player) ncard = 0;  // the reference leaves a stale copy that nothing reads
        }
        if (remaining == 1) {
            if (lane == cp) pot = -1;
            if (!termSet) { termSet = true; termPos = s.pos(); }
        }
    }

    // SimulationBoard.executeAttack (:758-780): losses for the armies on attackingCC / defendingCC
    __device__ __forceinline__ void execute_attack(int A, int D, int &al, int &dl)
    {
        int a = armies(A) - 1, d = armies(D);
        if (attackUntilDead) {
#ifdef ORX_DICE_2SITE
            simulate_battle<REPLAY, true>(dm, s, rk, a, d, al, dl);
#else
            simulate_battle<REPLAY, false>(dm, s, rk, a, d, al, dl);   // the catch-up battle of an EXECUTE leaf: once per game
#endif
        } else {
            int pa = a < 3 ? a : 3, pd = d < 2 ? d : 2;
            if (pa < 1 || pd < 1) { fail(); al = dl = 0; return; }
            al = single_roll(dm, pa, pd, s.next_k53());
            dl = (pa < pd ? pa : pd) - al;
        }
    }

    // SimRandom.moveArmiesIn (SimRandom.java:144-155)
    __device__ __forceinline__ int random_move_in(int A)
    {
        int available = armies(A) - 1;
        if (available == 0) { fail(); return 0; }
        int mn = available < 3 ? available : 3;
        return mn + (mn == available ? 0 : s.next_int(available - mn));
    }
    // executeInvasion (:814-825)
    __device__ __forceinline__ void execute_invasion(int A, int D, int req)
    {
        int aArm = armies(A);
        int lo = aArm - 1 < 3 ? aArm - 1 : 3, hi = aArm - 1;
        int inv = req > lo ? req : lo; if (inv > hi) inv = hi;
        __syncwarp();
        if (lane == 0) { sts32(a_arm + 4u * (uint32_t)D, (uint32_t)inv); sts32(a_arm + 4u * (uint32_t)A, (uint32_t)(aArm - inv)); }
        __syncwarp();
    }
    // simAgents[currentPlayer].moveArmiesIn + executeInvasion
    __device__ __forceinline__ void invade(int A, int D)
    {
        if (MODEL) {
            int req = sim_move_in(A, D);
            if (s.flags) return;
            execute_invasion(A, D, req);
            return;
        }
        // all-SimRandom instantiation: SimRandom.moveArmiesIn and executeInvasion fused, as the kernel was tuned
        int aArm = armies(A);
        int available = aArm - 1;
        if (available == 0) { fail(); return; }
        int mn = available < 3 ? available : 3;
        int req = mn + (mn == available ? 0 : s.next_int(available - mn));
        int lo = aArm - 1 < 3 ? aArm - 1 : 3, hi = aArm - 1;
        int inv = req > lo ? req : lo; if (inv > hi) inv = hi;
        __syncwarp();
        if (lane == 0) { sts32(a_arm + 4u * (uint32_t)D, (uint32_t)inv); sts32(a_arm + 4u * (uint32_t)A, (uint32_t)(aArm - inv)); }
        __syncwarp();
    }

    // finishInvasion (:827-841)
    __device__ __forceinline__ void finish_invasion()
    {
        if (defendingPlayer < 0 || defendingPlayer >= dm.P) { fail(); return; }
        if (from_lane(pc, defendingPlayer) == 0) player_defeated(defendingPlayer, attackingPlayer);
        hasInvaded = 1;
        attackState = ORX_ATTACK_START;
        attackingCC = defendingCC = attackingPlayer = defendingPlayer = -1;
    }

    // possibleAttackers (SimRandom.java:61,88-140) as a register-resident ordered list when it fits 64 entries
    // (NWT <= 2: lane l holds entries l and l + 32); shared memory otherwise.  get / remove(index) / add(tail).
    struct AttackerList {
        int lo, hi, n;
    };
    __device__ __forceinline__ int alist_get(const AttackerList &L, int i) const
    {
        if (NWT <= 2) {
            int a = __shfl_sync(ORX_FULL, L.lo, i & 31), b = __shfl_sync(ORX_FULL, L.hi, i & 31);
            return i < 32 ? a : b;
        }
        return (int)lds8(a_alist + (uint32_t)i);
    }
    __device__ __forceinline__ void alist_remove(AttackerList &L, int i)
    {
        if (NWT <= 2) {
            int dlo = __shfl_down_sync(ORX_FULL, L.lo, 1);
            if (L.n > 32) {
                int h0 = __shfl_sync(ORX_FULL, L.hi, 0), dhi = __shfl_down_sync(ORX_FULL, L.hi, 1);
                if (lane == 31) dlo = h0;
                if (lane + 32 >= i) L.hi = dhi;
            }
            if (lane >= i) L.lo = dlo;
        } else {
            bytes_remove_at(a_alist, L.n, i);
        }
        L.n--;
    }
    __device__ __forceinline__ void alist_add(AttackerList &L, int t)
    {
        if (NWT <= 2) {
            if (lane == L.n) L.lo = t;
            if (lane + 32 == L.n) L.hi = t;
        } else {
            __syncwarp(); if (lane == 0) sts8(a_alist + (uint32_t)L.n, (uint32_t)t); __syncwarp();
        }
        L.n++;
    }

    // SimRandom.attackPhase (SimRandom.java:58-141) with SimulationBoard.attack (:713-744) inlined
    __device__ __forceinline__ void agent_attack_phase()
    {
        uint32_t mine[NWT], att[NWT];
        owned_mask(cp, mine);
        border_mask(mine, true, att);
        // possibleAttackers: code order
        AttackerList L;
        L.n = 0; L.lo = 0; L.hi = 0;
        __syncwarp();
#pragma unroll
        for (int w = 0; w < NWT; w++) {
            if ((att[w] >> lane) & 1u) sts8(a_alist + (uint32_t)(L.n + __popc(att[w] & lanemask_lt())), (uint32_t)(w * 32 + lane));
            L.n += __popc(att[w]);
        }
        __syncwarp();
        if (NWT <= 2) { L.lo = (int)lds8(a_alist + (uint32_t)lane); L.hi = (int)lds8(a_alist + 32u + (uint32_t)lane); }
        attackUntilDead = 1;
        const int me = cp;
        while (L.n > 0 && !s.flags) {
            int ai = s.template next_int<ORX_NI_SMALL>(L.n);
            int A = alist_get(L, ai);
            // possibleDefenders: non-owned out-neighbours in getAdjoiningList() order
            int dg = deg(A);
            int nb = 0;
            bool enemy = false;
            if (lane < dg) { nb = adj(A, lane); enemy = owner(nb) != me; }
            uint32_t dmask = __ballot_sync(ORX_FULL, enemy);
            if (dmask == 0u) { alist_remove(L, ai); continue; }
            int D = from_lane(nb, nth_bit(dmask, s.template next_int<ORX_NI_SMALL>(__popc(dmask))));
            // SimulationBoard.attack: setupAttack, executeAttack, finishAttack
            saddr pA = a_arm + 4u * (uint32_t)A, pD = a_arm + 4u * (uint32_t)D;
            int aArm = (int)lds32(pA), dArm = (int)lds32(pD);
            int al, dl;
            simulate_battle(dm, s, rk, aArm - 1, dArm, al, dl);
            aArm -= al; dArm -= dl;
            if (dArm == 0) {
                // setupInvasion (:797-812)
                int dp = owner(D);
                if (lane == me) pc++;
                if (lane == dp) pc--;
                // SimRandom.moveArmiesIn (SimRandom.java:144-155) + executeInvasion (:814-825)
                int available = aArm - 1;
                if (available == 0) { fail(); return; }
                int inv = available;
                if (available > 3) inv = 3 + s.next_int(available - 3);  // already within [min(avail,3), avail]
                aArm -= inv; dArm = inv;
                __syncwarp();
                if (lane == 0) { sts32(pA, (uint32_t)aArm); sts32(pD, (uint32_t)dArm); sts8(a_own + (uint32_t)D, (uint32_t)me); }
                __syncwarp();
                // finishInvasion (:827-841)
                if (from_lane(pc, dp) == 0) player_defeated(dp, me);
                hasInvaded = 1;
                // result == DEFENDERS_DESTROYED (SimRandom.java:127-137)
                if (aArm == 1) alist_remove(L, ai);
                if (dArm > 1) {
                    if (L.n >= dm.cap) { fail(); return; }
                    alist_add(L, D);
                }
            } else {
                __syncwarp();
                if (lane == 0) { sts32(pA, (uint32_t)aArm); sts32(pD, (uint32_t)dArm); }
                __syncwarp();
                if (aArm == 1) alist_remove(L, ai);  // ATTACKERS_DESTROYED
            }
        }
        attackState = ORX_ATTACK_START;
        attackingCC = defendingCC = attackingPlayer = defendingPlayer = -1;
    }

    // tearDownAttackPhase (:884-896)
    __device__ __forceinline__ void tear_down_attack_phase()
    {
        if (hasInvaded && dm.use_cards) {
            int card = sim_deal();
            if (card >= 0) list_add(cp, card);
        }
        phase = ORX_PHASE_FORTIFICATION;
    }


    // ==================================================================================================
    // Modelled opponents: SimAngry and SimStinky behind ModellingBoard.  Only the MODEL
    // instantiations carry this code; the SDK semantics are the DECLARED ASSUMPTIONS of include/orx_state.h.
    // ==================================================================================================
    __device__ __forceinline__ int agent_of(int p) const { return MODEL ? (int)((agents >> (3 * p)) & 7u) : (int)ORX_AGENT_RANDOM; }
    __device__ __forceinline__ int moveable(int t) const { return (int)lds32(a_move + 4u * (uint32_t)t); }

    // Country.getNumberEnemyNeighbors(): adjoining countries whose owner differs from t's owner (one per lane)
    __device__ __forceinline__ int num_enemy_neighbors(int t) const
    {
        const int dg = deg(t), o = owner(t);
        bool e = false;
        if (lane < dg) e = owner(adj(t, lane)) != o;
        return __popc(__ballot_sync(ORX_FULL, e));
    }
    // Country.getWeakestEnemyNeighbor(): fewest armies, first in getAdjoiningList() order; -1 = null
    __device__ __forceinline__ int weakest_enemy_neighbor(int t) const
    {
        const int dg = deg(t), o = owner(t);
        uint32_t key = 0xFFFFFFFFu;
        int nb = 0;
        if (lane < dg) { nb = adj(t, lane); if (owner(nb) != o) key = ((uint32_t)armies(nb) << 5) | (uint32_t)lane; }
        uint32_t best = __reduce_min_sync(ORX_FULL, key);
        if (best == 0xFFFFFFFFu) return -1;
        return from_lane(nb, (int)(best & 31u));
    }
    // the element a PlayerIterator (arm == 0) / ArmiesIterator(player, arm) holds after looking from `from` on
    __device__ __forceinline__ int iter_find(int from, int player, int arm) const
    {
        int res = -1;
#pragma unroll
        for (int w = 0; w < NWT; w++) {
            int t = w * 32 + lane;
            bool ok = t >= from && t < dm.N && owner(t) == player;
            if (ok && arm != 0) { int a = armies(t); ok = arm > 0 ? a >= arm : a <= -arm; }
            uint32_t m = __ballot_sync(ORX_FULL, ok);
            if (res < 0 && m) res = w * 32 + __ffs(m) - 1;
        }
        return res;
    }

    // SimulationBoard.placeArmies (:647-684)
    __device__ __forceinline__ void place_armies(int n, int t)
    {
        bool okphase = phase == ORX_PHASE_PLACEMENT || phase == ORX_PHASE_INITIAL_PLACEMENT ||
                       (phase == ORX_PHASE_ATTACK && attackState == ORX_ATTACK_PLACE);
        if (!okphase || n > placeable || n < 0 || owner(t) != cp) { fail(); return; }
        set_armies(t, armies(t) + n);
        placeable -= n;
        if (phase == ORX_PHASE_INITIAL_PLACEMENT && lane == cp) ialp -= n;
    }

    // SimulationBoard.fortifyArmies (:908-954)
    __device__ __forceinline__ int fortify_armies(int n, int src, int dst)
    {
        if (phase != ORX_PHASE_FORTIFICATION) { fail(); return -1; }
        if (owner(src) != cp || owner(dst) != cp) return -1;
        if (n < 0) return -1;
        int sa = armies(src), mv = moveable(src);
        int k = mv < sa - 1 ? mv : sa - 1;
        if (n < k) k = n;
        if (k == 0) return 0;
        int da = dst == src ? sa - k : armies(dst);
        __syncwarp();
        if (lane == 0) {
            sts32(a_arm + 4u * (uint32_t)src, (uint32_t)(sa - k)); sts32(a_move + 4u * (uint32_t)src, (uint32_t)(mv - k));
            sts32(a_arm + 4u * (uint32_t)dst, (uint32_t)(da + k));
        }
        __syncwarp();
        return 1;
    }
    // setupFortificationPhase (:900-906)
    __device__ __forceinline__ void setup_fortification_phase()
    {
        __syncwarp();
#pragma unroll
        for (int w = 0; w < NWT; w++) { int t = w * 32 + lane; sts32(a_move + 4u * (uint32_t)t, owner(t) == cp ? (uint32_t)armies(t) : 0u); }
        __syncwarp();
    }

    // SimAngry / SimStinky .moveArmiesIn (SimAngry.java:199-213, SimStinky.java:111-125), else SimRandom's
    __device__ __forceinline__ int sim_move_in(int A, int D)
    {
        const int ag = agent_of(cp);
        if (ag == ORX_AGENT_ANGRY) {
            int ae = num_enemy_neighbors(A), de = num_enemy_neighbors(D);
            return ae > de ? 0 : armies(A) - 1;
        }
        if (ag == ORX_AGENT_STINKY) {
            int aw = weakest_enemy_neighbor(A), dw = weakest_enemy_neighbor(D);
            if (aw < 0) return 1000000;
            if (dw < 0) return 0;
            return armies(aw) < armies(dw) ? 0 : 1000000;
        }
        return random_move_in(A);
    }

    // SimulationBoard.attack(attackingCC, defendingCC, true) (:713-744): 7 conquered, 13 attackers spent, else 0
    __device__ __forceinline__ int board_attack(int A, int D)
    {
        attackUntilDead = 1;
        int aArm = armies(A), dArm = armies(D), al, dl;
        simulate_battle(dm, s, rk, aArm - 1, dArm, al, dl);
        if (s.flags) return 0;
        aArm -= al; dArm -= dl;
        __syncwarp();
        if (lane == 0) { sts32(a_arm + 4u * (uint32_t)A, (uint32_t)aArm); sts32(a_arm + 4u * (uint32_t)D, (uint32_t)dArm); }
        __syncwarp();
        if (dArm == 0) {
            const int ap = owner(A), dp = owner(D);  // setupInvasion (:797-812)
            if (lane == ap) pc++;
            if (lane == dp) pc--;
            set_owner(D, ap);
            attackingCC = A; defendingCC = D; attackingPlayer = ap; defendingPlayer = dp;
            int req = sim_move_in(A, D);
            if (s.flags) return 0;
            execute_invasion(A, D, req);
            finish_invasion();
            return 7;
        }
        attackState = ORX_ATTACK_START;
        return aArm == 1 ? 13 : 0;
    }

    // MapHelper.getSmallestPositiveEmptyContinent / ...OpenContinent (MapHelper.java:132-191)
    __device__ __forceinline__ int smallest_positive_continent(const uint32_t (&free_)[NWT], bool need_all) const
    {
        int smallestSize = 0x7FFFFFFF, smallest = -1;
        for (int c = 0; c < dm.C; c++) {
            int size = 0, un = 0;
#pragma unroll
            for (int w = 0; w < NWT; w++) { uint32_t cm = contmask(c, w); size += __popc(cm); un += __popc(cm & free_[w]); }
            bool ok = need_all ? un == size : un > 0;
            if (size < smallestSize && ok && current_bonus(c) > 0) { smallestSize = size; smallest = c; }
        }
        return smallest;
    }
    // SimAngry.pickCountry + pickCountryInContinent (SimAngry.java:49-105)
    __device__ __forceinline__ int angry_pick_country()
    {
        uint32_t free_[NWT], mine[NWT];
        owned_mask((int)ORX_OWNER_NONE, free_);
        owned_mask(cp, mine);
        int goal = smallest_positive_continent(free_, true);
        if (goal == -1) goal = smallest_positive_continent(free_, false);
        if (goal == -1) return -1;
        int res = -1;
        // an unowned country of the continent that touches one of ours, first in code order
#pragma unroll
        for (int w = 0; w < NWT; w++) {
            int t = w * 32 + lane;
            uint32_t touch = 0;
#pragma unroll
            for (int w2 = 0; w2 < NWT; w2++) touch |= nbr(w2, t) & mine[w2];
            bool ok = ((contmask(goal, w) & free_[w]) >> lane) & 1u;
            uint32_t m = __ballot_sync(ORX_FULL, ok && touch != 0u);
            if (res < 0 && m) res = w * 32 + __ffs(m) - 1;
        }
        if (res >= 0) return res;
        // else the unowned one with the fewest neighbours, first in code order
        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
        for (int w = 0; w < NWT; w++) {
            int t = w * 32 + lane;
            bool ok = ((contmask(goal, w) & free_[w]) >> lane) & 1u;
            uint32_t key = ok ? ((uint32_t)deg(t) << 8) | (uint32_t)t : 0xFFFFFFFFu;
            uint32_t k = __reduce_min_sync(ORX_FULL, key);
            if (k < best) best = k;
        }
        return best == 0xFFFFFFFFu ? -1 : (int)(best & 0xFFu);
    }
    // SimAngry.placeArmies (SimAngry.java:126-151)
    __device__ __forceinline__ void angry_place(int n)
    {
        int most = -1, placeOn = -1;
        for (int t = iter_find(0, cp, 0); t >= 0; t = iter_find(t + 1, cp, 0)) {
            int e = num_enemy_neighbors(t);
            if (e > most) { most = e; placeOn = t; }
        }
        if (placeOn < 0) { fail(); return; }
        place_armies(n, placeOn);
    }
    // SimAngry.attackPhase (SimAngry.java:157-195)
    __device__ __forceinline__ void angry_attack_phase()
    {
        const int me = cp;
        bool made = true;
        while (made && !s.flags) {
            made = false;
            int held = iter_find(0, me, 2);
            while (held >= 0 && !s.flags) {
                int us = held;
                held = iter_find(us + 1, me, 2);
                int weak = weakest_enemy_neighbor(us);
                if (weak >= 0 && armies(us) > armies(weak)) { board_attack(us, weak); made = true; }
            }
        }
    }
    // SimAngry.fortifyPhase (SimAngry.java:215-279)
    __device__ __forceinline__ void angry_fortify_phase()
    {
        for (int i = 0; i < dm.N && !s.flags; i++) {
            if (owner(i) != cp || moveable(i) <= 0) continue;
            const int dg = deg(i);
            int bestT = -1, bestE = 0;
            for (int j = 0; j < dg; j++) {
                int nb = adj(i, j);
                if (owner(nb) == cp) { int e = num_enemy_neighbors(nb); if (e > bestE) { bestT = nb; bestE = e; } }
            }
            int e = num_enemy_neighbors(i);
            if (bestE > e) {
                fortify_armies(moveable(i), i, bestT);
            } else if (e == 0) {
                int r = s.next_int(dg);  // nextInt(0) throws
                if (s.flags) return;
                fortify_armies(moveable(i), i, adj(i, r));
            }
        }
    }
    // SimStinky.placeArmies (SimStinky.java:66-78)
    __device__ __forceinline__ void stinky_place(int n)
    {
        int test, tries = 0;
        for (;;) {
            if (tries++ >= ORX_MAX_REDRAWS) { s.flags |= ORX_FLAG_STEP_LIMIT; return; }  // engine limit
            test = s.next_int(dm.N);
            if (s.flags) return;
            if (owner(test) == cp && num_enemy_neighbors(test) > 0) break;
        }
        place_armies(n, test);
    }
    // SimStinky.attackPhase(boolean) (SimStinky.java:91-109); armies > weak * 1.5 <=> 2 * armies > 3 * weak
    __device__ __forceinline__ bool stinky_attack_pass(bool care)
    {
        const int me = cp;
        bool attacked = false;
        int held = iter_find(0, me, 0);
        while (held >= 0 && !s.flags) {
            int us = held;
            held = iter_find(us + 1, me, 0);
            int weak = weakest_enemy_neighbor(us);
            if (weak >= 0) {
                long long a = armies(us), w = armies(weak);
                if (a > 1 && (!care || 2 * a > 3 * w)) { board_attack(us, weak); attacked = true; }
            }
        }
        return attacked;
    }
    // SimStinky.attackPhase() (SimStinky.java:80-89)
    __device__ __forceinline__ void stinky_attack_phase()
    {
        stinky_attack_pass(true);
        if (s.flags) return;
        int mysum = 0;  // BoardHelper.getPlayerArmies
#pragma unroll
        for (int w = 0; w < NWT; w++) { int t = w * 32 + lane; if (t < dm.N && owner(t) == cp) mysum += armies(t); }
        int total = (int)__reduce_add_sync(ORX_FULL, (unsigned)mysum);
        if (total > 300)
            while (!s.flags && stinky_attack_pass(false)) { }
    }

    // simAgents[currentPlayer].xxx()
    __device__ __forceinline__ void sim_place(int n)
    {
        if (MODEL) {
            const int ag = agent_of(cp);
            if (ag == ORX_AGENT_ANGRY) { angry_place(n); return; }
            if (ag == ORX_AGENT_STINKY) { stinky_place(n); return; }
        }
        agent_place(n);
    }
    __device__ __forceinline__ void sim_cards_phase()
    {
        if (MODEL && agent_of(cp) != ORX_AGENT_RANDOM) return;  // SimAngry / SimStinky: empty
        agent_cards_phase();
    }
    __device__ __forceinline__ void sim_attack_phase()
    {
        if (MODEL) {
            const int ag = agent_of(cp);
            if (ag == ORX_AGENT_ANGRY) { angry_attack_phase(); return; }
            if (ag == ORX_AGENT_STINKY) { stinky_attack_phase(); return; }
        }
        agent_attack_phase();
    }
    __device__ __forceinline__ void sim_fortify_phase()
    {
        if (MODEL && agent_of(cp) == ORX_AGENT_ANGRY) angry_fortify_phase();  // SimRandom, SimStinky: empty
    }

    // ---- fortification / turn advance -------------------------------------------------
    // checkForOverflow (:984-1015)
    __device__ __forceinline__ void check_for_overflow()
    {
        int mysum = 0;
#pragma unroll
        for (int w = 0; w < NWT; w++) { int t = w * 32 + lane; if (t < dm.N) mysum += armies(t); }
        int total = (int)__reduce_add_sync(ORX_FULL, (unsigned)mysum);
        if (total <= 100000) return;
        int maxArmies = 0, maxPlayer = -1;
        for (int p = 0; p < dm.P; p++) {
            int ps = 0;
#pragma unroll
            for (int w = 0; w < NWT; w++) { int t = w * 32 + lane; if (t < dm.N && owner(t) == p) ps += armies(t); }
            int tot = (int)__reduce_add_sync(ORX_FULL, (unsigned)ps);
            if (tot > maxArmies) { maxPlayer = p; maxArmies = tot; }
        }
        for (int p = 0; p < dm.P; p++)
            if (from_lane(pc, p) > 0 && p != maxPlayer) { player_defeated(p, maxPlayer); if (s.flags) return; }
    }

    // tearDownFortificationPhase (:965-982) with getNextPlayer (:1102-1111)
    __device__ __forceinline__ void tear_down_fortification_phase()
    {
        uint32_t alive = __ballot_sync(ORX_FULL, lane < dm.P && pc > 0);
        if (alive == 0u) { fail(); return; }  // the reference would spin forever
        uint32_t above = alive & ~((2u << cp) - 1u);
        int np = above ? __ffs(above) - 1 : __ffs(alive) - 1;
        if (MODEL && np < cp && simTurn + 1 == dm.model_limit) agents = 0u;  // ModellingBoard.java:29-47: all SimRandom
        if (np < cp) { simTurn++; recalculate_bonuses(); }
        cp = np;
        phase = ORX_PHASE_CARDS;
        if (!s.flags) check_for_overflow();
        playerTurns++;
    }

    // ---- pre-game phases (SimulationBoard.java:485-568, SimRandom.java:16-27) ----------
    __device__ __forceinline__ void setup_initial_placement_phase()
    {
        int a = from_lane(ialp, cp);
        placeable = a == 5 ? 5 : (a < 4 ? a : 4);
    }
    __device__ __forceinline__ void tear_down_country_selection_phase()
    {
        if ((int)hdr(WH_REM_COUNTRIES) == 0) { phase = ORX_PHASE_INITIAL_PLACEMENT; cp = 0; }
    }
    __device__ __forceinline__ void select_country()
    {
        int cc;
        if (MODEL && agent_of(cp) == ORX_AGENT_ANGRY) {
            cc = angry_pick_country();
        } else if (MODEL && agent_of(cp) == ORX_AGENT_STINKY) {
            cc = -1;  // SimStinky.pickCountry (SimStinky.java:51-54); selectCountry then indexes countries[-1]
        } else {
            uint32_t free_[NWT];
            owned_mask((int)ORX_OWNER_NONE, free_);
            int n = count_bits(free_);
            int i = s.next_int(n);  // nextInt(0) throws
            if (s.flags) return;
            cc = nth_territory(free_, i);
        }
        if (MODEL && (cc < 0 || owner(cc) != (int)ORX_OWNER_NONE)) { fail(); return; }  // selectCountry (:485-494); a SimRandom pick is always free
        set_owner(cc, cp);
        if (lane == cp) { pc++; ialp--; }
        set_hdr(WH_REM_COUNTRIES, hdr(WH_REM_COUNTRIES) - 1u);
        cp = (cp + 1) % dm.P;
    }
    __device__ __forceinline__ void tear_down_initial_placement_phase()
    {
        // next player (cyclic, excluding the current one) who still has armies to place
        uint32_t todo = __ballot_sync(ORX_FULL, lane < dm.P && ialp != 0) & ~(1u << cp);
        if (todo == 0u) {
            cp = 0; phase = ORX_PHASE_CARDS; simTurn++;
        } else {
            uint32_t above = todo & ~((2u << cp) - 1u);
            cp = above ? __ffs(above) - 1 : __ffs(todo) - 1;
        }
        playerTurns++;
    }

    __device__ __forceinline__ void step_limit_hit()
    {
        if (playerTurns >= dm.max_turns) s.flags |= ORX_FLAG_STEP_LIMIT;
    }

    // ---- setupState (:178-359) from a packed leaf record --------------------------------
    __device__ __forceinline__ void setup_state(const uint8_t *leaf)
    {
        const uint32_t *hw = (const uint32_t *)leaf;  // header words (orx_state.h OrxLeafHeader)
        uint32_t h2 = __ldg(hw + 2), h3 = __ldg(hw + 3), h4 = __ldg(hw + 4), h5 = __ldg(hw + 5);
        int turn_count = (int)__ldg(hw + 0);
        int h_placeable = (int)__ldg(hw + 1);
        int h_acc = (int)(int16_t)(h2 & 0xFFFFu), h_dcc = (int)(int16_t)(h2 >> 16);
        int h_cardpos = (int)(int16_t)(h3 & 0xFFFFu);
        int h_cp = (int)(int8_t)((h3 >> 16) & 0xFFu), h_phase = (int)(int8_t)(h3 >> 24);
        int h_astate = (int)(int8_t)(h4 & 0xFFu), h_inv = (int)(int8_t)((h4 >> 8) & 0xFFu);
        int h_dp = (int)(int8_t)((h4 >> 16) & 0xFFu), h_aud = (int)(int8_t)(h4 >> 24);
        int n_hand = (int)(h5 & 0xFFu), n_deck = (int)((h5 >> 8) & 0xFFu), h_deal_mode = (int)(h5 >> 24);
        int h_deal_seed = (int)__ldg(hw + 6);
        const int N = dm.N, P = dm.P;

        // engine-level record validation (orx_state.h "Leaf validity"): the same rules as the oracle
        bool bad = h_cp < 0 || h_cp >= P || h_phase < 0 || h_phase > ORX_PHASE_FORTIFICATION ||
                   h_astate < 0 || h_astate > ORX_ATTACK_PLACE || h_acc < -1 || h_acc >= N || h_dcc < -1 || h_dcc >= N ||
                   h_dp < -1 || h_dp >= P || h_cardpos < 0 || n_hand + n_deck > N + 2 ||
                   h_placeable < -ORX_MAX_ARMIES || h_placeable > ORX_MAX_ARMIES ||
                   (h_deal_mode != (int)ORX_DEAL_LEAF_LCG && h_deal_mode != (int)ORX_DEAL_GAME_STREAM);
        // an INVADE leaf may carry the conquered country at 0 armies (executeInvasion overwrites it)
        const int zero_ok = (h_phase == ORX_PHASE_ATTACK && h_astate == ORX_ATTACK_INVADE) ? h_dcc : -1;

        // setupCountries (:192-203)
        __syncwarp();
        bool badt = false;
#pragma unroll
        for (int w = 0; w < NWT; w++) {
            int t = w * 32 + lane;
            int o = 0xFE, a = 0;
            if (t < N) {
                o = __ldg(leaf + dm.off_owner + t);
                a = __ldg((const int32_t *)(leaf + dm.off_armies) + t);
                if (!(o < P || (o == (int)ORX_OWNER_NONE && h_phase == ORX_PHASE_COUNTRY_SELECTION))) badt = true;
                if (a < (t == zero_ok ? 0 : 1) || a > ORX_MAX_ARMIES) badt = true;
            }
            sts8(a_own + (uint32_t)t, (uint32_t)o); sts32(a_arm + 4u * (uint32_t)t, (uint32_t)a);
        }
        // cards: hand then deck (CardManager.simReset :240-248)
        const uint8_t *lc = leaf + dm.off_cards;
        for (int base = 0; base < n_hand + n_deck && base < N + 2; base += 32) {
            int i = base + lane;
            if (i < n_hand + n_deck && i < N + 2) {
                int c = __ldg(lc + i);
                if (!(c < N || c == (int)ORX_CARD_WILD)) badt = true;
                if (!bad) { if (i < n_hand) sts8(list(h_cp) + (uint32_t)i, (uint32_t)c); else sts8(list(P) + (uint32_t)(i - n_hand), (uint32_t)c); }
            }
        }
        int my_ncards = lane < P ? (int)__ldg(leaf + dm.off_ncards + lane) : 0;
        __syncwarp();
        if (bad || __any_sync(ORX_FULL, badt)) { fail(); return; }

        if (MODEL) {
            // simAgents = initialSimAgents (ModellingBoard.java:21-27); moveableArmies are not in the record: a
            // FORTIFICATION leaf starts from setupFortificationPhase's values (orx_state.h)
            agents = dm.agents;
#pragma unroll
            for (int w = 0; w < NWT; w++) {
                int t = w * 32 + lane;
                sts32(a_move + 4u * (uint32_t)t, (h_phase == ORX_PHASE_FORTIFICATION && owner(t) == h_cp) ? (uint32_t)armies(t) : 0u);
            }
            __syncwarp();
        }
        cp = h_cp; phase = h_phase; simTurn = 0; cardpos = h_cardpos;
        placeable = 0; hasInvaded = 0; playerTurns = 0;
        attackState = ORX_ATTACK_START; attackingCC = defendingCC = attackingPlayer = defendingPlayer = -1;
        attackUntilDead = h_aud != 0;
        termSet = false; termPos = 0;
        pot = 0; ialp = 0;
        ncard = lane == cp ? n_hand : (lane == P ? n_deck : 0);
        if (lane == 0) {
            sts32(a_hdr + 4u * WH_SIM_START, (uint32_t)turn_count); sts32(a_hdr + 4u * WH_REM_COUNTRIES, 0u);
            const uint64_t st = ((uint64_t)(long long)h_deal_seed ^ 0x5DEECE66Dull) & ((1ull << 48) - 1ull);   // Random(long seed)
            sts32(a_hdr + 4u * WH_DEAL_LO, (uint32_t)st); sts32(a_hdr + 4u * WH_DEAL_HI, (uint32_t)(st >> 32));
            sts32(a_hdr + 4u * WH_DEAL_MODE, (uint32_t)h_deal_mode);
        }
        __syncwarp();

        // setupCardState (:205-227): hidden hands are dealt at random, in player order
        for (int p = 0; p < P; p++) {
            if (p == cp) continue;
            int want = from_lane(my_ncards, p);
            for (int j = 0; j < want; j++) {
                int card = sim_deal();
                if (card < 0) { fail(); return; }  // the reference would store null and fail later
                if (s.flags) return;
                list_add(p, card);
            }
        }

        // setupGeneralGameState (:229-276)
        {
            int cnt = 0;
            for (int p = 0; p < P; p++) {
                uint32_t m[NWT];
                owned_mask(p, m);
                int c = count_bits(m);
                if (lane == p) cnt = c;
            }
            pc = cnt;
            remaining = phase == ORX_PHASE_COUNTRY_SELECTION ? P : __popc(__ballot_sync(ORX_FULL, lane < P && pc > 0));
        }

        // setupPhaseGameState (:278-359)
        int start = dm.start_armies + (lane > 1 ? lane : 0);  // calculateStartingArmies (:518-533)
        switch (phase) {
        case ORX_PHASE_COUNTRY_SELECTION: {
            uint32_t free_[NWT];
            owned_mask((int)ORX_OWNER_NONE, free_);
            ialp = start - pc;
            set_hdr(WH_REM_COUNTRIES, (uint32_t)count_bits(free_));
            break;
        }
        case ORX_PHASE_INITIAL_PLACEMENT: {
            int mine_total = 0;
            for (int p = 0; p < P; p++) {
                int ps = 0;
#pragma unroll
                for (int w = 0; w < NWT; w++) { int t = w * 32 + lane; if (owner(t) == p) ps += armies(t); }
                int tot = (int)__reduce_add_sync(ORX_FULL, (unsigned)ps);
                if (lane == p) mine_total = tot;
            }
            ialp = start - mine_total;
            placeable = h_placeable;
            break;
        }
        case ORX_PHASE_CARDS:
        case ORX_PHASE_PLACEMENT:
            placeable = h_placeable;
            break;
        case ORX_PHASE_ATTACK:
            hasInvaded = h_inv != 0;
            attackState = h_astate;
            attackingCC = h_acc; defendingCC = h_dcc;
            if (attackingCC != -1) { attackingPlayer = owner(attackingCC); defendingPlayer = h_dp; }
            if (h_astate == ORX_ATTACK_INVADE) {
                if (defendingPlayer < 0 || defendingPlayer >= P) { fail(); return; }
                if (from_lane(pc, defendingPlayer) == 0) remaining++;
            } else if (h_astate == ORX_ATTACK_CASH || h_astate == ORX_ATTACK_PLACE) {
                placeable = h_placeable;
            }
            break;
        default:  // FORTIFICATION
            break;
        }
        recalculate_bonuses();
    }

    // ---- simulate (:365-481): the catch-up pass is the first trip through the same switch ----
    __device__ __forceinline__ void simulate()
    {
        bool first = true;
        for (;;) {
            if (s.flags) break;
            if (!first && remaining <= 1) break;
            switch (phase) {
            case ORX_PHASE_COUNTRY_SELECTION:
                if (!first) select_country();
                if (!s.flags) tear_down_country_selection_phase();
                break;
            case ORX_PHASE_INITIAL_PLACEMENT:
                if (!first) setup_initial_placement_phase();
                if (!(MODEL && first && placeable <= 0)) sim_place(placeable);  // catch-up: only when there is something to place
                tear_down_initial_placement_phase();
                if (!first) step_limit_hit();
                break;
            case ORX_PHASE_CARDS:
                if (!first) placeable = 0;  // setupCardsPhase (:572-575)
                sim_cards_phase();
                tear_down_cards_phase();
                break;
            case ORX_PHASE_PLACEMENT:
                if (!first) placeable = (int)((uint32_t)placeable + (uint32_t)player_income());  // setupPlacementPhase
                if (!(MODEL && first && placeable <= 0)) sim_place(placeable);
                phase = ORX_PHASE_ATTACK;  // tearDownPlacementPhase (:686-694)
                break;
            case ORX_PHASE_ATTACK:
                if (!first) {
                    hasInvaded = 0;  // setupAttackPhase (:698-707)
                    attackingCC = defendingCC = attackingPlayer = defendingPlayer = -1;
                    attackState = ORX_ATTACK_START;
                } else {
                    if (attackState == ORX_ATTACK_EXECUTE) {
                        if (attackingCC < 0 || defendingCC < 0) { fail(); break; }
                        int al, dl;
                        execute_attack(attackingCC, defendingCC, al, dl);
                        if (s.flags) break;
                        int aArm = armies(attackingCC) - al, dArm = armies(defendingCC) - dl;  // finishAttack
                        set_armies(attackingCC, aArm); set_armies(defendingCC, dArm);
                        attackState = armies(defendingCC) == 0 ? ORX_ATTACK_INVADE : ORX_ATTACK_START;
                    }
                    if (attackState == ORX_ATTACK_INVADE) {
                        // no setupInvasion() on this path: SURVEY 8.1 quirk 1
                        if (attackingCC < 0 || defendingCC < 0) { fail(); break; }
                        invade(attackingCC, defendingCC);
                        if (s.flags) break;
                        finish_invasion();
                        if (s.flags) break;
                    }
                    if (attackState == ORX_ATTACK_CASH) {  // executeAttackCash (:869-876)
                        placeable = 0;
                        sim_cards_phase();
                        attackState = ORX_ATTACK_PLACE;
                        if (s.flags) break;
                    }
                    if (attackState == ORX_ATTACK_PLACE) {  // executeAttackPlacement (:878-882)
                        sim_place(placeable);
                        attackState = ORX_ATTACK_START;
                        if (s.flags) break;
                    }
                }
                sim_attack_phase();
                if (!s.flags) tear_down_attack_phase();
                break;
            default:  // FORTIFICATION: SimRandom.fortifyPhase is a no-op (SimRandom.java:158-160)
                if (MODEL) {
                    if (!first) setup_fortification_phase();
                    sim_fortify_phase();
                    if (s.flags) break;
                }
                tear_down_fortification_phase();
                if (!first) step_limit_hit();
                break;
            }
            first = false;
        }
    }

    // ---- result record (SimulationBoard.java:1031-1044 + the parity pins of orx_state.h) ----
    __device__ __forceinline__ void write_result(OrxGameResult *out, OrxLeafResult *agg)
    {
        uint32_t flags = s.flags;
        bool dead = (flags & (ORX_FLAG_ERROR | ORX_FLAG_STREAM_END)) != 0u;
        if (dead) flags = (flags & ORX_FLAG_STREAM_END) ? ORX_FLAG_STREAM_END : ORX_FLAG_ERROR;  // the first fault only
        uint32_t hsum = 0;
        if (!dead) {
#pragma unroll
            for (int w = 0; w < NWT; w++) {
                int t = w * 32 + lane;
                if (t < dm.N) hsum += orx_mix((uint32_t)t, (uint32_t)owner(t), (uint32_t)armies(t));
            }
            hsum = __reduce_add_sync(ORX_FULL, hsum);
        }
        int my_pot = (!dead && lane < dm.P) ? pot : 0;
        int turns = dead ? 0 : simTurn;
        uint32_t spos = dead ? 0u : (termSet ? termPos : s.pos());
        if (out) {
            uint32_t v = lane < ORX_MAX_PLAYERS ? (uint32_t)my_pot
                       : lane == 8 ? (uint32_t)turns : lane == 9 ? flags : lane == 10 ? spos : hsum;
            if (lane < 12) ((uint32_t *)out)[lane] = v;
        }
        if (agg) {
            // exact integer sums of this warp's playouts of the current leaf, kept in the warp's slice (lane l owns words
            // 2l, 2l+1 of the OrxLeafResult image) and flushed with one atomic per non-zero field when the leaf changes
            // or the warp retires (flush_leaf_sums): ~20 global atomics per warp instead of ~5 per playout
#ifndef ORX_ACC_PER_WARP   // default: ~5 global atomics per playout on the leaf's record (the per-warp sums below cost 5 %: DESIGN.md 8)
            unsigned long long *a = (unsigned long long *)agg;
            if (lane < ORX_MAX_PLAYERS && my_pot == -1) { atomicAdd(a + 2 + lane, 1ull); atomicAdd(a + 10 + lane, (unsigned long long)(long long)turns); }
            if (lane == 8) { atomicAdd(a + 0, 1ull); atomicAdd(a + 1, (unsigned long long)(long long)turns); }
            if (lane == 9 && flags) atomicAdd(a + 18, 1ull);
#else
            if (lane < ORX_MAX_PLAYERS && my_pot == -1) { acc_add(2 + lane, 1u); acc_add(10 + lane, (uint32_t)turns); }
            if (lane == 8) { acc_add(0, 1u); acc_add(1, (uint32_t)turns); }
            if (lane == 9 && flags) acc_add(18, 1u);
#endif
        }
    }
    // 64-bit add into field f of the warp's OrxLeafResult image (each field has exactly one writer lane)
    __device__ __forceinline__ void acc_add(int f, uint32_t v)
    {
        const saddr p = a_hdr + (uint32_t)ORX_WARP_HDR_BYTES + 8u * (uint32_t)f;
        uint32_t lo = lds32(p), hi = lds32(p + 4u);
        const uint32_t nlo = lo + v;
        hi += nlo < lo;
        sts32(p, nlo); sts32(p + 4u, hi);
    }
    __device__ __forceinline__ void flush_leaf_sums(OrxLeafResult *agg)
    {
        __syncwarp();
        if (lane < 20) {
            const saddr p = a_hdr + (uint32_t)ORX_WARP_HDR_BYTES + 8u * (uint32_t)lane;
            const unsigned long long v = ((unsigned long long)lds32(p + 4u) << 32) | (unsigned long long)lds32(p);
            if (v) { atomicAdd((unsigned long long *)agg + lane, v); sts32(p, 0u); sts32(p + 4u, 0u); }
        }
        __syncwarp();
    }
};

// The rollout kernel: persistent warps pull game indices from a ticket counter.
template <int NWT, bool REPLAY, bool MODEL>
__global__ void __launch_bounds__(ORX_BLOCK_THREADS, ORX_MIN_BLOCKS) rollout_kernel(const __grid_constant__ DevMap dm, const __grid_constant__ RolloutArgs ar)
{
    extern __shared__ __align__(16) uint8_t smem[];
#ifdef ORX_PAD   // tuning knob: shifts the code that follows by ORX_PAD instructions (fetch alignment of the hot loops)
#pragma unroll
    for (int i = 0; i < ORX_PAD; i++) asm volatile("bar.warp.sync 0xffffffff;");
#endif
    // stage the static map image
    {
        const uint4 *src = (const uint4 *)dm.blob; uint4 *dst = (uint4 *)smem;
        for (int i = threadIdx.x; i < dm.blob_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    // warp index through a shuffle: a warp-uniform value the PTX uniformity pass can see (ptx_uniform.py)
    const int wid = __shfl_sync(ORX_FULL, (int)(threadIdx.x >> 5), 0);
    uint8_t *slice = smem + dm.blob_bytes + wid * dm.w_bytes;

    Game<NWT, REPLAY, MODEL> g(dm);
    // pinned in registers: without the barrier ptxas re-derives the shared window base inside the hot loops
    saddr sbase = to_saddr(smem), wbase = to_saddr(slice);
    asm volatile("" : "+r"(sbase), "+r"(wbase));
    g.s_adj = sbase + (uint32_t)dm.adj_off; g.s_deg = sbase + (uint32_t)dm.deg_off; g.s_sym = sbase + (uint32_t)dm.sym_off;
    g.s_nbr = sbase + (uint32_t)dm.nbr_off; g.s_cont = sbase + (uint32_t)dm.cont_off; g.s_bonus = sbase + (uint32_t)dm.bonus_off;
    g.lane = lane_id();
    g.rk = ar.rk;
    g.a_hdr = wbase;
    g.a_arm = wbase + (uint32_t)dm.w_armies; g.a_own = wbase + (uint32_t)dm.w_owner;
    g.a_alist = wbase + (uint32_t)dm.w_alist; g.a_cards = wbase + (uint32_t)dm.w_cards;
    g.a_move = wbase + (uint32_t)dm.w_move; g.agents = 0u;
    const saddr rng = wbase + (uint32_t)dm.w_rng;
#ifdef ORX_ACC_PER_WARP
    if (g.lane < 20) { sts32(wbase + (uint32_t)ORX_WARP_HDR_BYTES + 8u * (uint32_t)g.lane, 0u); sts32(wbase + (uint32_t)ORX_WARP_HDR_BYTES + 8u * (uint32_t)g.lane + 4u, 0u); }
    if (g.lane == 0) sts32(wbase + 4u * WH_CUR_LEAF, 0xFFFFFFFFu);   // the leaf the warp's running sums belong to: none yet
    __syncwarp();
#endif

    for (;;) {
        unsigned long long idx = 0;
        if (g.lane == 0) idx = atomicAdd(ar.ticket, 1ull);
        idx = __shfl_sync(ORX_FULL, idx, 0);
        if (idx >= ar.total) break;
        unsigned long long leaf_i = REPLAY ? idx : idx / (unsigned long long)ar.playouts_per_leaf;
#ifdef ORX_ACC_PER_WARP
        if (ar.agg) {
            const uint32_t cur = lds32(g.a_hdr + 4u * WH_CUR_LEAF);
            if (cur != (uint32_t)leaf_i) {
                if (cur != 0xFFFFFFFFu) g.flush_leaf_sums(ar.agg + cur);
                __syncwarp();
                if (g.lane == 0) sts32(g.a_hdr + 4u * WH_CUR_LEAF, (uint32_t)leaf_i);
                __syncwarp();
            }
        }
#endif
        const uint8_t *leaf = ar.leaves + leaf_i * (unsigned long long)dm.leaf_stride;
        g.s.init(rng, g.a_hdr, ar.seed, ar.first_game + idx,
                 REPLAY ? ar.words + idx * (unsigned long long)ar.words_per_game : nullptr, ar.words_per_game);
        g.simTurn = 0; g.pot = 0; g.termSet = false; g.termPos = 0;
        g.setup_state(leaf);
        if (!g.s.flags) g.simulate();
        g.write_result(ar.games ? ar.games + idx : nullptr, ar.agg ? ar.agg + leaf_i : nullptr);
        __syncwarp();
    }
#ifdef ORX_ACC_PER_WARP
    if (ar.agg) {
        const uint32_t cur = lds32(g.a_hdr + 4u * WH_CUR_LEAF);
        if (cur != 0xFFFFFFFFu) g.flush_leaf_sums(ar.agg + cur);
    }
#endif
}

}  // namespace orx
