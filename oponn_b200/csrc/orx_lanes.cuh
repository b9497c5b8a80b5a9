// orx_lanes.cuh -- the lane-per-game rollout kernel: 32 games per warp, one lane each.
//
// Why a second kernel: the warp-per-game kernel (orx_game.cuh) runs a game's scalar decisions
// redundantly in 32 lanes (5.6 M warp instructions for ~13 M algorithmic lane operations per playout).  Here every
// lane plays its own game with plain scalar code, and the warp re-converges through a flat loop of SECTIONS
// (NEW / TURN / SEL / DICE / FIN, below): a lane waits in its state until the warp runs that section, so the lanes
// that are in the same part of the game execute it together.  Long dice battles (the dominant work: ~360 k single
// rolls per playout) are rolled by the whole warp for one lane at a time (COOP): 32 Philox counters = 64 rolls per
// step, exactly the doubles the reference's loop would draw.
//
// Scope: all-SimRandom line-ups, counter-mode streams, maps up to 64 territories.  Everything else (replay streams,
// modelled opponents, larger maps, small batches) runs on the warp-per-game kernel; both produce bit-identical
// records (same oracle, same goldens), so the choice is invisible to the caller (orx_api.cu: use_lane_kernel).
//
// Data layout: a lane's mutable board lives in shared memory, INTERLEAVED by lane (word w of lane l at
// slice + (32 w + l) * 4), so any per-lane index pattern is bank-conflict free; the ordered card lists
// (java.util.List semantics, touched about once per player-turn) live in thread-local memory.
//
// Reference restated (O/ = /root/reference/src/com/mld46/oponn/): O/sim/boards/SimulationBoard.java:178-359
// (setupState), :365-481 (simulate), :485-1015 (rules), :1076-1111; O/sim/agents/SimRandom.java (whole);
// O/BattleSim.java:66-162; O/CardManager.java:240-346.  Same quirks as orx_game.cuh / the oracle.
//
// This header is written so that the per-lane logic also compiles with a host compiler (ORX_LANES_HOST): the CPU
// emulation harness under tests/tools/lane_emulator runs the same sections lane after lane and is compared with the
// oracle on the CPU.  That harness is test infrastructure; the product only ever runs the CUDA kernel at the bottom.
#pragma once

#include "orx_devmap.h"

#ifdef ORX_LANES_HOST
#include <math.h>
#include <string.h>
#define LN_M inline
#define LN_NOUNROLL
#define LN_COLD inline
#define LN_NOINLINE static inline
#define LN_LDG(p) (*(p))
extern "C" double orc_fdlibm_log(double);   // the oracle's fdlibm restatement (host emulation only)
#else
#include "orx_device.cuh"
#define LN_M __device__ __forceinline__
#define LN_NOUNROLL _Pragma("unroll 1")
#define LN_COLD __device__ __noinline__       // once-per-turn code: out of line, called on a memory copy of the lane (see the kernel)
#define LN_NOINLINE static __device__ __noinline__
#define LN_LDG(p) __ldg(p)
#endif

namespace orx {

// lane states: the section a lane waits for
enum { LS_NEW = 0, LS_TURN = 1, LS_SEL = 2, LS_DICE = 3, LS_FIN = 4, LS_COOP = 5, LS_EXIT = 6, LS_GAUSS = 7, LS_ELIM = 8 };

// bits of Lane::fl above the ORX_FLAG_* bits
#define LF_INVADED   0x100u   // hasInvadedACountry
#define LF_TERMSET   0x200u   // stream position of the terminal move recorded
#define LF_GAUSS     0x400u   // java.util.Random.haveNextNextGaussian
#define LF_DEALSTRM  0x800u   // ORX_DEAL_GAME_STREAM
#define LF_HASGAME   0x1000u  // the lane holds a game whose result has not been written yet
#define LF_NOTEAR    0x2000u  // the next TURN section starts at the phase loop (no attack policy has just returned)
#define LF_ERRMASK   (ORX_FLAG_ERROR | ORX_FLAG_STEP_LIMIT | ORX_FLAG_STREAM_END)

#define LN_MAXN 64
#define LN_MAXP ORX_MAX_PLAYERS
#define LN_CAP  (LN_MAXN + 4)           // longest card list (N + 2), rounded
#define LN_LISTS (LN_MAXP + 1)          // hands + sim deck
#define LN_MAX_SEL_ITERS (1 << 22)      // watchdog: attack-loop trips per attack phase

// tuning of the section scheduler (kernel parameter; ORX_LANE_TUNE overrides the defaults at launch)
struct LaneTune {
    int32_t thr_new, thr_turn, thr_sel, thr_fin;   // run the section when at least this many lanes wait for it ...
    int32_t age_new, age_turn, age_sel, age_fin;   // ... or when somebody has waited this many loop trips
    int32_t coop_min;                              // dice battles expected to last at least this many rolls go to the warp
    int32_t reserved_[3];
};

struct LaneArgs {
    const uint8_t *leaves;
    int32_t n_leaves, playouts_per_leaf;
    unsigned long long total, seed, first_game;
    OrxLeafResult *agg;
    OrxGameResult *games;
    unsigned long long *ticket;
    uint32_t rk[20];            // Philox round keys: rk[2r] = seed_lo + r * 0x9E3779B9, rk[2r+1] = seed_hi + r * 0xBB67AE85
    LaneTune tune;
};

// word offsets of a lane's interleaved slice
struct LaneLayout { int ARM, OWN, AL, POT, PC, NC, RING, GAUSS, words; };
static inline ORX_HD LaneLayout lane_layout(int N, int P)
{
    LaneLayout L;
    int w = 0;
    L.ARM = w;   w += N;                          // int32 armies[N]
    L.OWN = w;   w += (N + 3) / 4;                // uint8 owner[N]
    const int alw = (N + 3) / 4 > P ? (N + 3) / 4 : P;
    L.AL = w;    w += alw;                        // uint8 possibleAttackers[N]; aliased by int32 initialArmiesLeftToPlace[P] (pre-game only)
    L.POT = w;   w += P;                          // int32 playerOutOnTurn[P]
    L.PC = w;    w += 2;                          // uint8 playerCountries[P]
    L.NC = w;    w += 3;                          // uint8 list lengths [P + 1]
    L.RING = w;  w += 16;                         // the next (at most 11) stream words, word i in slot i % 16
    L.GAUSS = w; w += 2;                          // nextNextGaussian
    L.words = w;
    return L;
}

#ifdef ORX_LANES_HOST
static inline uint32_t ln_umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int ln_popc(uint32_t x) { return __builtin_popcount(x); }
static inline int ln_ffs(uint32_t x) { return __builtin_ffs((int)x); }
#define LN_DMUL(a, b) ((a) * (b))
#define LN_DADD(a, b) ((a) + (b))
#define LN_DSUB(a, b) ((a) - (b))
#define LN_DDIV(a, b) ((a) / (b))
#define LN_DSQRT(a) sqrt(a)
static inline int ln_d2i(double d) { if (d != d) return 0; if (d >= 2147483647.0) return 2147483647; if (d <= -2147483648.0) return (int)(-2147483647 - 1); return (int)d; }
#else
#define ln_umulhi __umulhi
#define ln_popc __popc
#define ln_ffs __ffs
#define LN_DMUL(a, b) __dmul_rn((a), (b))
#define LN_DADD(a, b) __dadd_rn((a), (b))
#define LN_DSUB(a, b) __dsub_rn((a), (b))
#define LN_DDIV(a, b) __ddiv_rn((a), (b))
#define LN_DSQRT(a) __dsqrt_rn(a)
#define ln_d2i __double2int_rz
#endif

// Philox4x32-10 with the round keys precomputed on the host (constant-bank operands on the device)
LN_M void ln_philox(uint32_t c0, uint32_t c2, uint32_t c3, const uint32_t *rk, uint32_t o[4])
{
    uint32_t c1 = 0u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = ln_umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = ln_umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

struct LnWords { uint32_t w0, w1, w2, w3; };
LN_NOINLINE LnWords ln_philox_call(uint32_t c0, uint32_t c2, uint32_t c3, const uint32_t *rk)
{
    uint32_t o[4];
    ln_philox(c0, c2, c3, rk, o);
    LnWords r; r.w0 = o[0]; r.w1 = o[1]; r.w2 = o[2]; r.w3 = o[3];
    return r;
}

// index of the i-th (0-based) set bit of m; i < popc(m)
LN_M int ln_nth32(uint32_t m, int i)
{
    int r = 0, c;
    c = ln_popc(m & 0xFFFFu); if (i >= c) { i -= c; r += 16; m >>= 16; }
    c = ln_popc(m & 0xFFu);   if (i >= c) { i -= c; r += 8;  m >>= 8; }
    c = ln_popc(m & 0xFu);    if (i >= c) { i -= c; r += 4;  m >>= 4; }
    c = ln_popc(m & 0x3u);    if (i >= c) { i -= c; r += 2;  m >>= 2; }
    c = (int)(m & 1u);        if (i >= c) { r += 1; }
    return r;
}
LN_M int ln_nth64(uint32_t lo, uint32_t hi, int i)
{
    int c = ln_popc(lo);
    return i < c ? ln_nth32(lo, i) : 32 + ln_nth32(hi, i - c);
}

// What a lane needs to reach its data: immutable for the life of the kernel.
struct LaneEnv {
    const DevMap *dm; const LaneArgs *ar;
#ifdef ORX_LANES_HOST
    uint32_t *mem; const uint8_t *blob;
#else
    saddr lbase, mbase;
#endif
    int oARM, oOWN, oAL, oPOT, oPC, oNC, oRING, oGAUSS;
    uint8_t *cards;             // [LN_LISTS][LN_CAP], thread-local memory
};

// One lane's game: registers on the device.  (The once-per-turn code runs out of line on a memory copy of the STATE;
// the environment is not copied back, so the hot sections keep seeing kernel parameters as constant-bank operands.)
struct LaneState {
    int st;
    uint32_t fl;
    int cp, phase, simTurn, remaining, placeable, cardpos, nAtt, playerTurns, simStart;
    uint32_t pos, lim;          // stream: next word, end of the buffered range (a multiple of 4)
    uint32_t mine_lo, mine_hi;  // territories of the current player
    uint32_t g0, g1;            // game id (Philox counter words 2, 3)
    unsigned long long gidx;    // index of the game in this launch
    int bA, bD, bai, ra, rd;    // battle in progress: attacking / defending country, list index, armies left to fight
    uint32_t termPos;
    uint64_t deal;              // CardManager.random (48-bit java.util.Random state)
    int selIters;               // engine watchdog: trips of the attack loop in this attack phase
};

struct Lane : LaneEnv, LaneState {
    // ---- memory -----------------------------------------------------------------------
#ifdef ORX_LANES_HOST
    LN_M uint32_t ldw(int w) const { return mem[w]; }
    LN_M void stw(int w, uint32_t v) { mem[w] = v; }
    LN_M uint32_t ldb(int w0, int b) const { return ((const uint8_t *)(mem + w0))[b]; }
    LN_M void stb(int w0, int b, uint32_t v) { ((uint8_t *)(mem + w0))[b] = (uint8_t)v; }
    LN_M uint32_t mapb(int off) const { return blob[off]; }
    LN_M uint32_t mapw(int off) const { uint32_t v; memcpy(&v, blob + off, 4); return v; }
#else
    LN_M uint32_t ldw(int w) const { return lds32(lbase + ((uint32_t)w << 7)); }
    LN_M void stw(int w, uint32_t v) { sts32(lbase + ((uint32_t)w << 7), v); }
    LN_M uint32_t ldb(int w0, int b) const { return lds8(lbase + ((uint32_t)(w0 + (b >> 2)) << 7) + (uint32_t)(b & 3)); }
    LN_M void stb(int w0, int b, uint32_t v) { sts8(lbase + ((uint32_t)(w0 + (b >> 2)) << 7) + (uint32_t)(b & 3), v); }
    LN_M uint32_t mapb(int off) const { return lds8(mbase + (uint32_t)off); }
    LN_M uint32_t mapw(int off) const { return lds32(mbase + (uint32_t)off); }
#endif
    LN_M void fail() { fl |= ORX_FLAG_ERROR; }
    LN_M bool bad() const { return (fl & LF_ERRMASK) != 0u; }

    LN_M int armies(int t) const { return (int)ldw(oARM + t); }
    LN_M void set_armies(int t, int v) { stw(oARM + t, (uint32_t)v); }
    LN_M int owner(int t) const { return (int)ldb(oOWN, t); }
    LN_M void set_owner(int t, int v) { stb(oOWN, t, (uint32_t)v); }
    LN_M int pc(int p) const { return (int)ldb(oPC, p); }
    LN_M void set_pc(int p, int v) { stb(oPC, p, (uint32_t)v); }
    LN_M int ncard(int li) const { return (int)ldb(oNC, li); }
    LN_M void set_ncard(int li, int v) { stb(oNC, li, (uint32_t)v); }
    LN_M int pot(int p) const { return (int)ldw(oPOT + p); }
    LN_M void set_pot(int p, int v) { stw(oPOT + p, (uint32_t)v); }
    LN_M int ialp(int p) const { return (int)ldw(oAL + p); }
    LN_M void set_ialp(int p, int v) { stw(oAL + p, (uint32_t)v); }
    LN_M int alist(int i) const { return (int)ldb(oAL, i); }
    LN_M void set_alist(int i, int v) { stb(oAL, i, (uint32_t)v); }

    // static map (staged once per block)
    LN_M uint32_t nbr_lo(int t) const { return mapw(dm->nbr_off + 4 * t); }
    LN_M uint32_t nbr_hi(int t) const { return mapw(dm->nbr_off + 4 * (dm->NP + t)); }
    LN_M uint32_t cont_lo(int c) const { return mapw(dm->cont_off + 8 * c); }
    LN_M uint32_t cont_hi(int c) const { return mapw(dm->cont_off + 8 * c + 4); }
    LN_M uint32_t symbit(int card) const { return card == (int)ORX_CARD_WILD ? 8u : mapb(dm->sym_off + card); }
    LN_M bool is_mine(int t) const { return (((t & 32) ? mine_hi : mine_lo) >> (t & 31)) & 1u; }
    LN_M void add_mine(int t) { if (t & 32) mine_hi |= 1u << (t & 31); else mine_lo |= 1u << (t & 31); }

    // ---- ordered lists: list li < P is player li's hand, li == P the sim deck ------------
    LN_M uint8_t *list(int li) const { return cards + li * LN_CAP; }
    LN_COLD void list_add(int li, int card)
    {
        int n = ncard(li);
        if (n >= dm->N + 2) { fail(); return; }
        list(li)[n] = (uint8_t)card;
        set_ncard(li, n + 1);
    }
    LN_COLD int list_remove_at(int li, int i)
    {
        uint8_t *p = list(li);
        int n = ncard(li), c = p[i];
        LN_NOUNROLL
        for (int k = i; k < n - 1; k++) p[k] = p[k + 1];
        set_ncard(li, n - 1);
        return c;
    }

    // ---- the word stream (orx_state.h): word i = Philox(ctr = {i / 4, 0, game}, key = seed)[i % 4] ------------
    // The ring holds words [pos, lim); lim is a multiple of 4 and lim - pos <= 11.
    LN_M void gen()
    {
        uint32_t o[4];
        ln_philox(lim >> 2, g0, g1, ar->rk, o);
        const int s = oRING + (int)(lim & 12u);
        stw(s, o[0]); stw(s + 1, o[1]); stw(s + 2, o[2]); stw(s + 3, o[3]);
        lim += 4u;
    }
    LN_M void gen_call()   // the same through an out-of-line Philox (refills outside the hot sections' top-up)
    {
        const LnWords o = ln_philox_call(lim >> 2, g0, g1, ar->rk);
        const int s = oRING + (int)(lim & 12u);
        stw(s, o.w0); stw(s + 1, o.w1); stw(s + 2, o.w2); stw(s + 3, o.w3);
        lim += 4u;
    }
    // afterwards at least 8 words are buffered: a SEL trip (2 words) followed by a DICE trip (4) never runs dry, and the
    // FIN section checks (usually one Philox per loop trip; up to three right after a cooperative battle)
    LN_M void topup() { while ((int)(lim - pos) < 8) gen(); }
    LN_M uint32_t word_nc() { uint32_t v = ldw(oRING + (int)(pos & 15u)); pos++; return v; }   // caller knows pos < lim
    LN_M uint32_t word() { if ((int)(lim - pos) <= 0) gen_call(); return word_nc(); }

    // java.util.Random.nextInt(bound) for a bound the caller knows to be in 1..255: u % bound by multiplication with
    // floor(2^32 / bound) + 1 from the map image (the estimate is q or q + 1); CHECKED = false right after topup()
    template <bool CHECKED>
    LN_M int next_int_small(int bound)
    {
        uint32_t u = (CHECKED ? word() : word_nc()) >> 1;
        const uint32_t b = (uint32_t)bound, m = b - 1u;
        if ((b & m) == 0u) return (int)(((unsigned long long)b * u) >> 31);
        LN_NOUNROLL
        for (;;) {
            uint32_t q = ln_umulhi(u, mapw(dm->magic_off + 4 * (int)b));
            uint32_t r = u - q * b;
            if ((int)r < 0) r += b;
            if ((int)(u - r + m) >= 0) return (int)r;
            u = word() >> 1;
        }
    }
    // java.util.Random.nextInt(bound), any bound
    LN_COLD int next_int(int bound)
    {
        if (bound <= 0) { fail(); return 0; }
        uint32_t u = word() >> 1;
        const uint32_t b = (uint32_t)bound, m = b - 1u;
        if ((b & m) == 0u) return (int)(((unsigned long long)b * u) >> 31);
        LN_NOUNROLL
        for (;;) {
            uint32_t r = u % b;
            if ((int)(u - r + m) >= 0) return (int)r;
            u = word() >> 1;
        }
    }
    // the same, inline, for the hot sections (bound > 0)
    LN_M int next_int_hot(int bound)
    {
        uint32_t u = word() >> 1;
        const uint32_t b = (uint32_t)bound, m = b - 1u;
        if ((b & m) == 0u) return (int)(((unsigned long long)b * u) >> 31);
        LN_NOUNROLL
        for (;;) {
            uint32_t r = u % b;
            if ((int)(u - r + m) >= 0) return (int)r;
            u = word() >> 1;
        }
    }
    // the 53-bit integer k of nextDouble() == k * 2^-53
    template <bool CHECKED>
    LN_M uint64_t next_k53()
    {
        uint32_t w0 = CHECKED ? word() : word_nc(), w1 = CHECKED ? word() : word_nc();
        return ((uint64_t)(w0 >> 6) << 27) + (uint64_t)(w1 >> 5);
    }
    LN_M double next_double() { return (double)(long long)next_k53<true>() * 0x1.0p-53; }
    // java.util.Random.nextGaussian(): polar method, second value cached per playout
    LN_COLD double next_gaussian()
    {
        if (fl & LF_GAUSS) {
            fl &= ~LF_GAUSS;
            uint32_t lo = ldw(oGAUSS), hi = ldw(oGAUSS + 1);
            uint64_t b = ((uint64_t)hi << 32) | lo; double g;
            memcpy(&g, &b, 8);
            return g;
        }
        double v1, v2, q;
        LN_NOUNROLL
        do {
            v1 = LN_DSUB(LN_DMUL(2.0, next_double()), 1.0);
            v2 = LN_DSUB(LN_DMUL(2.0, next_double()), 1.0);
            q = LN_DADD(LN_DMUL(v1, v1), LN_DMUL(v2, v2));
        } while (q >= 1.0 || q == 0.0);
        double second, first;
#ifdef ORX_LANES_HOST
        double mult = sqrt((-2.0 * orc_fdlibm_log(q)) / q);
        first = v1 * mult; second = v2 * mult;
#else
        first = gaussian_from_pair(v1, v2, q, &second);
#endif
        uint64_t b; memcpy(&b, &second, 8);
        stw(oGAUSS, (uint32_t)b); stw(oGAUSS + 1, (uint32_t)(b >> 32));
        fl |= LF_GAUSS;
        return first;
    }

    // ---- CardManager.random = new Random(simDeck.hashCode()) (O/CardManager.java:245) ----------------------
    LN_M int deal_next31() { deal = (deal * 0x5DEECE66Dull + 0xBull) & ((1ull << 48) - 1ull); return (int)(deal >> 17); }
    LN_COLD int deal_next_int(int bound)   // bound > 0
    {
        int r = deal_next31();
        const int m = bound - 1;
        if ((bound & m) == 0) return (int)(((long long)bound * (long long)r) >> 31);
        LN_NOUNROLL
        for (int u = r;; u = deal_next31()) {
            r = u % bound;
            if ((int)((uint32_t)u - (uint32_t)r + (uint32_t)m) >= 0) return r;
        }
    }
    // CardManager.simDeal (:250-263); -1 when the deck is legitimately empty
    LN_COLD int sim_deal()
    {
        const int n = ncard(dm->P);
        if (n == 0) { if (!dm->can_run_out) fail(); return -1; }
        return list_remove_at(dm->P, (fl & LF_DEALSTRM) ? next_int(n) : deal_next_int(n));
    }

    // ---- BattleSim ---------------------------------------------------------------------------------------
    // one uniform against the (a, d) outcome table: first row with k <= floor(cdf_i * 2^53); none => no losses
    LN_M void table_battle(int a, int d, uint64_t k, int &al, int &dl) const
    {
        const uint32_t meta = LN_LDG(dm->bt_meta + (a - 1) * ORX_TABLE_LIMIT + (d - 1));
        const uint64_t *rows = dm->bt_rows + (meta >> 8);
        const int n = (int)(meta & 0xFFu);
        const uint64_t key = k << 11;
        int lo = 0, hi = n;   // first index whose row takes the key
        LN_NOUNROLL
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (key <= LN_LDG(rows + mid)) hi = mid; else lo = mid + 1;
        }
        al = 0; dl = 0;
        if (lo < n) {
            const uint32_t code = (uint32_t)LN_LDG(rows + lo) & 0xFFu;
            const int rem = (int)(code & 0x7Fu);
            if (code & 0x80u) { al = a - rem; dl = d; } else { al = a; dl = d - rem; }
        }
    }
    // BattleSim.simulateIndividualBattle: attacker losses of one roll of pa v pd dice
    LN_M int single_roll(int pa, int pd, uint64_t k) const
    {
        const int row = (pa - 1) * 2 + (pd - 1);
        return k <= dm->single_thr[row][0] ? 0 : (k <= dm->single_thr[row][1] ? 1 : 2);
    }
    // the >500 v >500 branch of simulateBattle (O/BattleSim.java:71-93): losses from one gaussian
    LN_COLD void gaussian_battle(int attackers, int defenders, int &al, int &dl)
    {
        int matched = ln_d2i(LN_DMUL(0.922, (double)defenders));
        if (matched > attackers) {
            al = attackers;
            int variance = ln_d2i(LN_DMUL(next_gaussian(), LN_DSQRT((double)attackers)));
            double v = LN_DADD((double)variance, LN_DDIV((double)attackers, 0.922));
            double lim_ = (double)(defenders - 1);
            dl = ln_d2i(v < lim_ ? v : lim_);   // Math.min; v is never NaN here
        } else {
            dl = defenders;
            int variance = ln_d2i(LN_DMUL(next_gaussian(), LN_DSQRT((double)defenders)));
            int v = (int)((uint32_t)variance + (uint32_t)matched);
            al = v < attackers - 1 ? v : attackers - 1;
        }
    }
    // BattleSim.simulateBattle, whole, in this lane (catch-up of an EXECUTE leaf only; the attack loop goes through the
    // SEL / DICE / FIN sections)
    LN_COLD void battle_inline(int attackers, int defenders, int &alOut, int &dlOut)
    {
        if (attackers > 500 && defenders > 500) { gaussian_battle(attackers, defenders, alOut, dlOut); return; }
        int a = attackers, d = defenders;
        LN_NOUNROLL
        while ((a > 100 || d > 100) && a > 0 && d > 0) {
            const int pa = a < 3 ? a : 3, pd = d < 2 ? d : 2, n = pa < pd ? pa : pd;
            const int x = single_roll(pa, pd, next_k53<true>());
            a -= x; d -= n - x;
        }
        if (a != 0 && d != 0) {
            if (a < 0 || d < 0 || a > ORX_TABLE_LIMIT || d > ORX_TABLE_LIMIT) { fail(); alOut = dlOut = 0; return; }
            int al, dl;
            table_battle(a, d, next_k53<true>(), al, dl);
            a -= al; d -= dl;
        }
        alOut = attackers - a; dlOut = defenders - d;
    }

    // ---- cards -------------------------------------------------------------------------------------------
    // DECLARED ASSUMPTION (orx_state.h): wildcard, or all equal, or all distinct
    static LN_M bool is_set_bits(uint32_t m) { return (m & 8u) || ln_popc(m) != 2; }
    // number of valid triples i<j<k of the current player's hand; with pick >= 0 also the pick-th one (lexicographic)
    LN_COLD int scan_sets(int n, int pick, int idx[3]) const
    {
        const uint8_t *h = list(cp);
        int count = 0;
        LN_NOUNROLL
        for (int i = 0; i < n - 2; i++) {
            const uint32_t si = symbit(h[i]);
            LN_NOUNROLL
            for (int j = i + 1; j < n - 1; j++) {
                const uint32_t sij = si | symbit(h[j]);
                LN_NOUNROLL
                for (int k = j + 1; k < n; k++)
                    if (is_set_bits(sij | symbit(h[k]))) {
                        if (count == pick) { idx[0] = i; idx[1] = j; idx[2] = k; }
                        count++;
                    }
            }
        }
        return count;
    }
    LN_COLD int card_cash_value(int pos_)
    {
        switch (dm->prog_family) {
        case PROG_0: return 0;
        case PROG_555: return 5;
        case PROG_4466688: {
            long long v = 8ll * pos_ + 9ll;
            if (v <= 0) return 0;
            long long k = (long long)((sqrt((double)v) - 1.0) * 0.5);
            LN_NOUNROLL
            while ((2 * k + 1) * (2 * k + 1) >= v && k > 0) k--;
            LN_NOUNROLL
            while ((2 * k + 1) * (2 * k + 1) < v) k++;
            return (int)(2 * k);
        }
        case PROG_456: return pos_ + 3;
        case PROG_468: return (pos_ + 1) * 2;
        case PROG_369: return pos_ * 3;
        case PROG_46810: return pos_ <= 4 ? (pos_ + 1) * 2 : (pos_ - 2) * 5;
        case PROG_51015: return pos_ * 5;
        case PROG_4681015202510: return pos_ <= 4 ? (pos_ + 1) * 2 : (pos_ <= 7 ? (pos_ - 2) * 5 : 10);
        case PROG_TABLE:
            if (pos_ < 0 || pos_ >= ORX_CARD_VALUES) { fail(); return 0; }
            return LN_LDG(dm->card_table + pos_);
        default: return 2;
        }
    }
    // SimulationBoard.cashCards (:577-618) for hand positions i<j<k
    LN_COLD void cash_cards(const int idx[3])
    {
        const uint8_t *h = list(cp);
        const int c0 = h[idx[0]], c1 = h[idx[1]], c2 = h[idx[2]];
        // currentCards.remove(card) x3 (:596-598): the FIRST element equal to the card (the two wildcards are one object)
        const int cv[3] = { c0, c1, c2 };
        LN_NOUNROLL
        for (int q = 0; q < 3; q++) {
            const uint8_t *hh = list(cp);
            const int n = ncard(cp);
            int at = -1;
            LN_NOUNROLL
            for (int i = 0; i < n; i++) if (hh[i] == cv[q]) { at = i; break; }
            if (at < 0) { fail(); return; }
            list_remove_at(cp, at);
        }
        list_add(dm->P, c0); list_add(dm->P, c1); list_add(dm->P, c2);   // CardManager.simCash
        placeable += card_cash_value(cardpos);
        cardpos++;
        const int cs[3] = { c0, c1, c2 };
        LN_NOUNROLL
        for (int q = 0; q < 3; q++)
            if (cs[q] != (int)ORX_CARD_WILD && owner(cs[q]) == cp) set_armies(cs[q], armies(cs[q]) + 2);
    }
    // SimRandom.cardsPhase (SimRandom.java:36-49)
    LN_COLD void agent_cards_phase()
    {
        const int n = ncard(cp);
        if (n < 3) return;
        int idx[3] = { 0, 0, 0 };
        const int count = scan_sets(n, -1, idx);
        if (count == 0) return;
        const uint64_t k = next_k53<true>();
        if ((n == 3 && k < (1ull << 52)) || (n == 4 && k < (3ull << 51)) || n >= 5) {
            const int pick = next_int(count);
            scan_sets(n, pick, idx);
            cash_cards(idx);
        }
    }
    // SimulationBoard.tearDownCardsPhase (:620-638)
    LN_COLD void tear_down_cards_phase()
    {
        LN_NOUNROLL
        for (;;) {
            const int n = ncard(cp);
            if (n <= 4 || bad()) break;
            int idx[3] = { 0, 0, 0 };
            const int count = scan_sets(n, -1, idx);
            if (count == 0) { fail(); break; }
            const int pick = next_int(count);
            scan_sets(n, pick, idx);
            cash_cards(idx);
        }
        phase = ORX_PHASE_PLACEMENT;
    }

    // ---- board scans ---------------------------------------------------------------------------------------
    LN_COLD void compute_mine()
    {
        uint32_t lo = 0, hi = 0;
        const int N = dm->N;
        LN_NOUNROLL
        for (int t = 0; t < N && t < 32; t++) lo |= (uint32_t)(owner(t) == cp) << t;
        LN_NOUNROLL
        for (int t = 32; t < N; t++) hi |= (uint32_t)(owner(t) == cp) << (t - 32);
        mine_lo = lo; mine_hi = hi;
    }
    // owned territories with at least one out-neighbour not owned (SimRandom.java:68-82,250-262)
    LN_COLD void border_mask(uint32_t &blo, uint32_t &bhi) const
    {
        blo = 0; bhi = 0;
        LN_NOUNROLL
        for (uint32_t m = mine_lo; m; m &= m - 1u) {
            const int t = ln_ffs(m) - 1;
            if ((nbr_lo(t) & ~mine_lo) | (nbr_hi(t) & ~mine_hi)) blo |= 1u << t;
        }
        LN_NOUNROLL
        for (uint32_t m = mine_hi; m; m &= m - 1u) {
            const int t = ln_ffs(m) - 1;
            if ((nbr_lo(32 + t) & ~mine_lo) | (nbr_hi(32 + t) & ~mine_hi)) bhi |= 1u << t;
        }
    }

    // currentContinentBonuses[c] (recalculateContinentBonuses :956-963)
    LN_M int current_bonus(int c) const
    {
        if (dm->cont_inc == 0) return (int)mapw(dm->bonus_off + 4 * c);
        int e = simStart + simTurn - 1; if (e < 0) e = 0;
        return LN_LDG(dm->bonus_table + e * dm->C + c);
    }
    LN_M void recalculate_bonuses()
    {
        const int e = simStart + simTurn - 1;
        if (dm->cont_inc != 0 && e >= ORX_BONUS_ROUNDS) fail();
    }
    // getPlayerIncome(currentPlayer) (:1076-1090)
    LN_COLD int player_income() const
    {
        const int mycount = pc(cp);
        if (mycount == 0) return 0;
        uint32_t income = (uint32_t)(mycount / 3 < 3 ? 3 : mycount / 3);
        LN_NOUNROLL
        for (int c = 0; c < dm->C; c++)
            if (((cont_lo(c) & ~mine_lo) | (cont_hi(c) & ~mine_hi)) == 0u) income += (uint32_t)current_bonus(c);
        return (int)income;
    }
    // SimRandom.place (SimRandom.java:247-277) + placeArmies (:647-684)
    LN_COLD void agent_place(int n_armies)
    {
        uint32_t clo, chi;
        border_mask(clo, chi);
        const int n = ln_popc(clo) + ln_popc(chi);
        LN_NOUNROLL
        while (n_armies > 0 && !bad()) {
            const int k = 1 + next_int(n_armies);
            const int i = next_int(n);
            if (bad()) return;
            const int t = ln_nth64(clo, chi, i);
            set_armies(t, armies(t) + k);
            placeable -= k;
            if (phase == ORX_PHASE_INITIAL_PLACEMENT) set_ialp(cp, ialp(cp) - k);
            n_armies -= k;
        }
    }

    // SimulationBoard.playerDefeated (:843-867)
    LN_COLD void player_defeated(int player, int conquerer)
    {
        remaining--;
        set_pot(player, simTurn);
        const int dst = dm->transfer_cards ? conquerer : dm->P;
        if (dst < 0) { fail(); return; }
        if (dst != player) {
            const int nsrc = ncard(player), ndst = ncard(dst);
            if (ndst + nsrc > dm->N + 2) { fail(); return; }
            const uint8_t *src = list(player); uint8_t *d = list(dst);
            LN_NOUNROLL
            for (int i = 0; i < nsrc; i++) d[ndst + i] = src[i];
            set_ncard(dst, ndst + nsrc);
            set_ncard(player, 0);
        }
        if (remaining == 1) {
            set_pot(cp, -1);
            if (!(fl & LF_TERMSET)) { fl |= LF_TERMSET; termPos = pos; }
        }
    }
    // executeInvasion (:814-825) with SimRandom.moveArmiesIn (SimRandom.java:144-155); armies[A] is read from the board
    LN_COLD void invade_catchup(int A, int D)
    {
        const int aArm = armies(A), available = aArm - 1;
        if (available == 0) { fail(); return; }
        const int mn = available < 3 ? available : 3;
        const int req = mn + (mn == available ? 0 : next_int(available - mn));
        const int lo = aArm - 1 < 3 ? aArm - 1 : 3, hi = aArm - 1;
        int inv = req > lo ? req : lo; if (inv > hi) inv = hi;
        set_armies(D, inv); set_armies(A, aArm - inv);
    }

    // checkForOverflow (:984-1015)
    LN_COLD void check_for_overflow()
    {
        const int N = dm->N, P = dm->P;
        int total = 0;
        LN_NOUNROLL
        for (int t = 0; t < N; t++) total += armies(t);
        if (total <= 100000) return;
        int maxArmies = 0, maxPlayer = -1;
        LN_NOUNROLL
        for (int p = 0; p < P; p++) {
            int ps = 0;
            LN_NOUNROLL
            for (int t = 0; t < N; t++) if (owner(t) == p) ps += armies(t);
            if (ps > maxArmies) { maxPlayer = p; maxArmies = ps; }
        }
        LN_NOUNROLL
        for (int p = 0; p < P; p++)
            if (pc(p) > 0 && p != maxPlayer) { player_defeated(p, maxPlayer); if (bad()) return; }
    }
    // tearDownFortificationPhase (:965-982) with getNextPlayer (:1102-1111)
    LN_COLD void tear_down_fortification_phase()
    {
        const int P = dm->P;
        int np = cp, guard = 0;
        LN_NOUNROLL
        do { np = np + 1 == P ? 0 : np + 1; if (++guard > P) { fail(); return; } } while (pc(np) == 0);
        if (np < cp) { simTurn++; recalculate_bonuses(); }
        cp = np;
        phase = ORX_PHASE_CARDS;
        if (!bad()) check_for_overflow();
        playerTurns++;
    }
    LN_M void step_limit_hit() { if (playerTurns >= dm->max_turns) fl |= ORX_FLAG_STEP_LIMIT; }

    // ---- pre-game phases (:485-568, SimRandom.java:16-27) -----------------------------------------------------
    LN_M int count_free() const { int n = 0; for (int t = 0; t < dm->N; t++) n += owner(t) == (int)ORX_OWNER_NONE; return n; }
    LN_COLD void select_country()
    {
        const int n = count_free();
        int i = next_int(n);   // nextInt(0) throws
        if (bad()) return;
        int cc = -1;
        LN_NOUNROLL
        for (int t = 0; t < dm->N; t++) if (owner(t) == (int)ORX_OWNER_NONE && i-- == 0) { cc = t; break; }
        set_owner(cc, cp);
        set_pc(cp, pc(cp) + 1); set_ialp(cp, ialp(cp) - 1);
        cp = cp + 1 == dm->P ? 0 : cp + 1;
        if (n - 1 == 0) { phase = ORX_PHASE_INITIAL_PLACEMENT; cp = 0; }   // tearDownCountrySelectionPhase
    }
    LN_M void setup_initial_placement_phase() { const int a = ialp(cp); placeable = a == 5 ? 5 : (a < 4 ? a : 4); }
    LN_COLD void tear_down_initial_placement_phase()
    {
        const int P = dm->P;
        int np = cp + 1 == P ? 0 : cp + 1;
        LN_NOUNROLL
        while (np != cp && ialp(np) == 0) np = np + 1 == P ? 0 : np + 1;
        if (np == cp) { cp = 0; phase = ORX_PHASE_CARDS; simTurn++; } else cp = np;
        playerTurns++;
    }

    // ---- the attack phase's list (SimRandom.java:61-87): owned, armies > 1, an enemy neighbour; code order ----
    LN_COLD void build_attackers()
    {
        uint32_t blo, bhi;
        border_mask(blo, bhi);
        int n = 0;
        LN_NOUNROLL
        for (uint32_t m = blo; m; m &= m - 1u) { const int t = ln_ffs(m) - 1; if (armies(t) > 1) set_alist(n++, t); }
        LN_NOUNROLL
        for (uint32_t m = bhi; m; m &= m - 1u) { const int t = 32 + ln_ffs(m) - 1; if (armies(t) > 1) set_alist(n++, t); }
        nAtt = n; selIters = 0;
    }
    LN_M void alist_remove(int i)
    {
        LN_NOUNROLL
        for (int k = i; k < nAtt - 1; k++) set_alist(k, alist(k + 1));
        nAtt--;
    }

    // ---- setupState (:178-359) ------------------------------------------------------------------------------
    // returns the leaf's attack sub-state fields through the out parameters (only the catch-up needs them)
    LN_COLD void setup_state(const uint8_t *leaf, int &attackState, int &attackingCC, int &defendingCC, int &defendingPlayer,
                          int &attackUntilDead)
    {
        const uint32_t *hw = (const uint32_t *)leaf;
        const uint32_t h2 = LN_LDG(hw + 2), h3 = LN_LDG(hw + 3), h4 = LN_LDG(hw + 4), h5 = LN_LDG(hw + 5);
        const int turn_count = (int)LN_LDG(hw + 0), h_placeable = (int)LN_LDG(hw + 1), h_deal_seed = (int)LN_LDG(hw + 6);
        const int h_acc = (int)(int16_t)(h2 & 0xFFFFu), h_dcc = (int)(int16_t)(h2 >> 16);
        const int h_cardpos = (int)(int16_t)(h3 & 0xFFFFu);
        const int h_cp = (int)(int8_t)((h3 >> 16) & 0xFFu), h_phase = (int)(int8_t)(h3 >> 24);
        const int h_astate = (int)(int8_t)(h4 & 0xFFu), h_inv = (int)(int8_t)((h4 >> 8) & 0xFFu);
        const int h_dp = (int)(int8_t)((h4 >> 16) & 0xFFu), h_aud = (int)(int8_t)(h4 >> 24);
        const int n_hand = (int)(h5 & 0xFFu), n_deck = (int)((h5 >> 8) & 0xFFu), h_deal_mode = (int)(h5 >> 24);
        const int N = dm->N, P = dm->P;

        bool bad_ = h_cp < 0 || h_cp >= P || h_phase < 0 || h_phase > ORX_PHASE_FORTIFICATION ||
                    h_astate < 0 || h_astate > ORX_ATTACK_PLACE || h_acc < -1 || h_acc >= N || h_dcc < -1 || h_dcc >= N ||
                    h_dp < -1 || h_dp >= P || h_cardpos < 0 || n_hand + n_deck > N + 2 ||
                    h_placeable < -ORX_MAX_ARMIES || h_placeable > ORX_MAX_ARMIES ||
                    (h_deal_mode != (int)ORX_DEAL_LEAF_LCG && h_deal_mode != (int)ORX_DEAL_GAME_STREAM);
        const int zero_ok = (h_phase == ORX_PHASE_ATTACK && h_astate == ORX_ATTACK_INVADE) ? h_dcc : -1;

        // setupCountries (:192-203)
        LN_NOUNROLL
        for (int t = 0; t < N; t++) {
            const int o = LN_LDG(leaf + dm->off_owner + t);
            const int a = LN_LDG((const int32_t *)(leaf + dm->off_armies) + t);
            if (!(o < P || (o == (int)ORX_OWNER_NONE && h_phase == ORX_PHASE_COUNTRY_SELECTION))) bad_ = true;
            if (a < (t == zero_ok ? 0 : 1) || a > ORX_MAX_ARMIES) bad_ = true;
            set_owner(t, o); set_armies(t, a);
        }
        LN_NOUNROLL
        for (int li = 0; li <= P; li++) set_ncard(li, 0);
        LN_NOUNROLL
        for (int p = 0; p < P; p++) { set_pot(p, 0); set_pc(p, 0); }
        // cards: the real hand, then the unseen deck (CardManager.simReset :240-248)
        if (!bad_) {
            const uint8_t *lc = leaf + dm->off_cards;
            LN_NOUNROLL
            for (int i = 0; i < n_hand + n_deck; i++) {
                const int c = LN_LDG(lc + i);
                if (!(c < N || c == (int)ORX_CARD_WILD)) bad_ = true;
                if (i < n_hand) list(h_cp)[i] = (uint8_t)c; else list(P)[i - n_hand] = (uint8_t)c;
            }
        }
        cp = h_cp >= 0 && h_cp < P ? h_cp : 0; phase = h_phase; simTurn = 0; cardpos = h_cardpos; simStart = turn_count;
        placeable = 0; playerTurns = 0; nAtt = 0; termPos = 0;
        attackState = ORX_ATTACK_START; attackingCC = defendingCC = defendingPlayer = -1;
        attackUntilDead = h_aud != 0;
        mine_lo = mine_hi = 0;
        if (bad_) { fail(); return; }
        set_ncard(cp, n_hand); set_ncard(P, n_deck);
        if (h_deal_mode == (int)ORX_DEAL_GAME_STREAM) fl |= LF_DEALSTRM;
        deal = ((uint64_t)(long long)h_deal_seed ^ 0x5DEECE66Dull) & ((1ull << 48) - 1ull);   // new Random(simDeck.hashCode())

        // setupCardState (:205-227): hidden hands, in player order
        LN_NOUNROLL
        for (int p = 0; p < P; p++) {
            if (p == cp) continue;
            const int want = LN_LDG(leaf + dm->off_ncards + p);
            LN_NOUNROLL
            for (int j = 0; j < want; j++) {
                const int card = sim_deal();
                if (card < 0) { fail(); return; }
                if (bad()) return;
                list_add(p, card);
            }
        }
        // setupGeneralGameState (:229-276)
        LN_NOUNROLL
        for (int t = 0; t < N; t++) { const int o = owner(t); if (o < P) set_pc(o, pc(o) + 1); }
        if (phase == ORX_PHASE_COUNTRY_SELECTION) remaining = P;
        else { remaining = 0; for (int p = 0; p < P; p++) remaining += pc(p) > 0; }

        // setupPhaseGameState (:278-359)
        switch (phase) {
        case ORX_PHASE_COUNTRY_SELECTION:
            LN_NOUNROLL
            for (int p = 0; p < P; p++) set_ialp(p, dm->start_armies + (p > 1 ? p : 0) - pc(p));
            break;
        case ORX_PHASE_INITIAL_PLACEMENT:
            LN_NOUNROLL
            for (int p = 0; p < P; p++) {
                int tot = 0;
                LN_NOUNROLL
                for (int t = 0; t < N; t++) if (owner(t) == p) tot += armies(t);
                set_ialp(p, dm->start_armies + (p > 1 ? p : 0) - tot);
            }
            placeable = h_placeable;
            break;
        case ORX_PHASE_CARDS:
        case ORX_PHASE_PLACEMENT:
            placeable = h_placeable;
            break;
        case ORX_PHASE_ATTACK:
            if (h_inv != 0) fl |= LF_INVADED;
            attackState = h_astate; attackingCC = h_acc; defendingCC = h_dcc;
            if (attackingCC != -1) defendingPlayer = h_dp;
            if (h_astate == ORX_ATTACK_INVADE) {
                if (defendingPlayer < 0 || defendingPlayer >= P) { fail(); return; }
                if (pc(defendingPlayer) == 0) remaining++;
            } else if (h_astate == ORX_ATTACK_CASH || h_astate == ORX_ATTACK_PLACE) {
                placeable = h_placeable;
            }
            break;
        default: break;
        }
        recalculate_bonuses();
    }

    // ---- one playout: start (setupState + the catch-up pass of simulate :373-431), then the phase loop ----------
    LN_COLD void start_game(unsigned long long idx)
    {
        gidx = idx;
        const unsigned long long game = ar->first_game + idx;
        g0 = (uint32_t)game; g1 = (uint32_t)(game >> 32);
        pos = 0; lim = 0; fl = LF_HASGAME;
        const unsigned long long leaf_i = idx / (unsigned long long)ar->playouts_per_leaf;
        const uint8_t *leaf = ar->leaves + leaf_i * (unsigned long long)dm->leaf_stride;
        int attackState, attackingCC, defendingCC, defendingPlayer, attackUntilDead;
        setup_state(leaf, attackState, attackingCC, defendingCC, defendingPlayer, attackUntilDead);
        st = LS_TURN; fl |= LF_NOTEAR;   // the phase loop (advance) follows in the TURN section
        if (bad()) return;
        switch (phase) {   // catch up to the start of the next phase
        case ORX_PHASE_COUNTRY_SELECTION:
            if (count_free() == 0) { phase = ORX_PHASE_INITIAL_PLACEMENT; cp = 0; }
            break;
        case ORX_PHASE_INITIAL_PLACEMENT:
            compute_mine();
            agent_place(placeable);
            tear_down_initial_placement_phase();
            break;
        case ORX_PHASE_CARDS:
            agent_cards_phase();
            tear_down_cards_phase();
            break;
        case ORX_PHASE_PLACEMENT:
            compute_mine();
            agent_place(placeable);
            phase = ORX_PHASE_ATTACK;
            break;
        case ORX_PHASE_ATTACK:
            if (attackState == ORX_ATTACK_EXECUTE) {
                if (attackingCC < 0 || defendingCC < 0) { fail(); break; }
                int al, dl;
                const int a = armies(attackingCC) - 1, d = armies(defendingCC);
                if (attackUntilDead) battle_inline(a, d, al, dl);
                else {
                    const int pa = a < 3 ? a : 3, pd = d < 2 ? d : 2;
                    if (pa < 1 || pd < 1) { fail(); break; }
                    al = single_roll(pa, pd, next_k53<true>());
                    dl = (pa < pd ? pa : pd) - al;
                }
                if (bad()) break;
                set_armies(attackingCC, armies(attackingCC) - al); set_armies(defendingCC, armies(defendingCC) - dl);   // finishAttack
                attackState = armies(defendingCC) == 0 ? ORX_ATTACK_INVADE : ORX_ATTACK_START;
            }
            if (attackState == ORX_ATTACK_INVADE) {   // no setupInvasion() on this path: SURVEY 8.1 quirk 1
                if (attackingCC < 0 || defendingCC < 0) { fail(); break; }
                invade_catchup(attackingCC, defendingCC);
                if (bad()) break;
                if (defendingPlayer < 0 || defendingPlayer >= dm->P) { fail(); break; }   // finishInvasion (:827-841)
                if (pc(defendingPlayer) == 0) player_defeated(defendingPlayer, owner(attackingCC));
                fl |= LF_INVADED;
                attackState = ORX_ATTACK_START;
                if (bad()) break;
            }
            if (attackState == ORX_ATTACK_CASH) {   // executeAttackCash (:869-876)
                placeable = 0;
                agent_cards_phase();
                attackState = ORX_ATTACK_PLACE;
                if (bad()) break;
            }
            compute_mine();
            if (attackState == ORX_ATTACK_PLACE) {   // executeAttackPlacement (:878-882)
                agent_place(placeable);
                if (bad()) break;
            }
            build_attackers();   // SimRandom.attackPhase; tearDownAttackPhase follows in the TURN section
            fl &= ~LF_NOTEAR;
            if (nAtt > 0) st = LS_SEL;
            break;
        default:   // FORTIFICATION
            tear_down_fortification_phase();
            break;
        }
    }

    // tearDownAttackPhase (:884-896)
    LN_COLD void tear_down_attack_phase()
    {
        if ((fl & LF_INVADED) && dm->use_cards) {
            const int card = sim_deal();
            if (card >= 0) list_add(cp, card);
        }
        phase = ORX_PHASE_FORTIFICATION;
    }

    // the phase loop of simulate (:435-479) up to the point where an attack phase has somebody to attack with
    LN_COLD void advance()
    {
        LN_NOUNROLL
        for (;;) {
            if (bad() || remaining <= 1) { st = LS_NEW; return; }
            switch (phase) {
            case ORX_PHASE_COUNTRY_SELECTION:
                select_country();
                break;
            case ORX_PHASE_INITIAL_PLACEMENT:
                setup_initial_placement_phase();
                compute_mine();
                agent_place(placeable);
                tear_down_initial_placement_phase();
                step_limit_hit();
                break;
            case ORX_PHASE_CARDS:
                placeable = 0;
                agent_cards_phase();
                tear_down_cards_phase();
                break;
            case ORX_PHASE_PLACEMENT:
                compute_mine();
                placeable = (int)((uint32_t)placeable + (uint32_t)player_income());
                agent_place(placeable);
                phase = ORX_PHASE_ATTACK;
                break;
            case ORX_PHASE_ATTACK:
                fl &= ~LF_INVADED;   // setupAttackPhase (:698-707)
                build_attackers();
                if (nAtt > 0) { st = LS_SEL; return; }
                tear_down_attack_phase();
                break;
            default:   // FORTIFICATION: SimRandom.fortifyPhase is a no-op
                tear_down_fortification_phase();
                step_limit_hit();
                break;
            }
        }
    }

    // ---- sections ---------------------------------------------------------------------------------------------
    // TURN: the attack policy has returned
    LN_COLD void sec_turn()
    {
        if (!bad() && !(fl & LF_NOTEAR)) tear_down_attack_phase();
        fl &= ~LF_NOTEAR;
        advance();
    }
    // GAUSS: the > 500 v > 500 battle of the pair SEL picked (rare: out of line)
    LN_COLD void sec_gauss()
    {
        int al, dl;
        gaussian_battle(ra, rd, al, dl);
        ra -= al; rd -= dl;
        st = LS_FIN;
    }

    // SEL: one trip of SimRandom.attackPhase's loop (SimRandom.java:88-140) up to the battle
    LN_M void sec_sel()
    {
        // engine watchdog (the reference would spin forever): an attack phase cannot need this many trips on a valid board
        if (++selIters > LN_MAX_SEL_ITERS) { fail(); st = LS_TURN; return; }
        const int ai = next_int_small<false>(nAtt);
        const int A = alist(ai);
        const uint32_t elo = nbr_lo(A) & ~mine_lo, ehi = nbr_hi(A) & ~mine_hi;
        const int nd = ln_popc(elo) + ln_popc(ehi);
        if (nd == 0) {
            alist_remove(ai);
            if (nAtt == 0) st = LS_TURN;
            return;
        }
        int j = next_int_small<false>(nd);
        // the j-th non-owned out-neighbour in getAdjoiningList() order
        int D = -1;
        const int row = dm->adj_off + A * dm->degp, dg = (int)mapb(dm->deg_off + A);
        LN_NOUNROLL
        for (int q = 0; q < dg; q++) {
            const int nb = (int)mapb(row + q);
            if (!is_mine(nb)) { if (j == 0) { D = nb; break; } j--; }
        }
        const int a = armies(A) - 1, d = armies(D);
        bA = A; bD = D; bai = ai;
        ra = a; rd = d;
        if (a > 500 && d > 500) st = LS_GAUSS;
        else if (a > 100 || d > 100) {
            // rolls until one side is out or both are within the tables, roughly: long battles go to the whole warp
            const int over = (a > 100 ? a - 100 : 0) + (d > 100 ? d - 100 : 0);
            const int est = (a < d ? a : d) < (over + 1) / 2 ? (a < d ? a : d) : (over + 1) / 2;
            st = (est >= ar->tune.coop_min && a >= 3 && d >= 2) ? LS_COOP : LS_DICE;   // the warp only rolls 3 v 2 dice
        } else st = LS_FIN;
    }

    // DICE: up to two single rolls (one Philox counter) of the "either side above 100" loop (O/BattleSim.java:99-116)
    LN_M void sec_dice()
    {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int pa = ra < 3 ? ra : 3, pd = rd < 2 ? rd : 2, n = pa < pd ? pa : pd;
            const int x = single_roll(pa, pd, next_k53<false>());
            ra -= x; rd -= n - x;
            if (!((ra > 100 || rd > 100) && ra > 0 && rd > 0)) { st = LS_FIN; break; }
        }
    }

#ifdef ORX_LANES_HOST
    // COOP on the host emulation: the rolls the warp would make together (3 v 2 dice only), one after the other
    LN_M void sec_coop_host()
    {
        LN_NOUNROLL
        while ((ra > 100 || rd > 100) && ra >= 3 && rd >= 2) {
            const uint64_t k = next_k53<true>();
            const int x = single_roll(3, 2, k);
            ra -= x; rd -= 2 - x;
        }
        st = ((ra > 100 || rd > 100) && ra > 0 && rd > 0) ? LS_DICE : LS_FIN;
    }
#endif

    // FIN: the outcome-table draw if both sides are left, then SimulationBoard.attack's bookkeeping (:713-744,
    // :797-841) and the list edits of SimRandom.attackPhase (SimRandom.java:121-137)
    LN_M void sec_fin()
    {
        if (ra > 0 && rd > 0) {
            int al, dl;
            table_battle(ra, rd, next_k53<true>(), al, dl);
            ra -= al; rd -= dl;
        }
        const int A = bA, D = bD, me = cp;
        int aArm = ra + 1, dArm = rd;
        if (dArm == 0) {
            const int dp = owner(D);   // setupInvasion (:797-812)
            set_owner(D, me); add_mine(D);
            set_pc(me, pc(me) + 1);
            const int left = pc(dp) - 1;
            set_pc(dp, left);
            // SimRandom.moveArmiesIn + executeInvasion
            const int available = aArm - 1;
            if (available == 0) { fail(); st = LS_TURN; return; }
            int inv = available;
            if (available > 3) inv = 3 + next_int_hot(available - 3);
            aArm -= inv; dArm = inv;
            fl |= LF_INVADED;
            if (left == 0) { rd = dp; ra = aArm; bD |= dArm > 1 ? 0x100 : 0; st = LS_ELIM; }   // finishInvasion: the defender is out (rare: out of line)
        }
        set_armies(A, aArm); set_armies(D, dArm);
        if (st == LS_ELIM) return;
        if (aArm == 1) alist_remove(bai);              // ATTACKERS_DESTROYED, or a conquest that left one army behind
        if (rd == 0 && dArm > 1) { set_alist(nAtt, D); nAtt++; }
        st = (nAtt == 0 || bad()) ? LS_TURN : LS_SEL;
    }
    // ELIM: playerDefeated for the conquest FIN has just booked, then the list edits FIN skipped
    LN_COLD void sec_elim()
    {
        const int D = bD & 0xFF, dp = rd, aArm = ra;
        player_defeated(dp, cp);
        if (aArm == 1) alist_remove(bai);
        if (bD & 0x100) { set_alist(nAtt, D); nAtt++; }
        st = (nAtt == 0 || bad()) ? LS_TURN : LS_SEL;
    }

    // ---- result record (SimulationBoard.java:1031-1044 + the parity pins of orx_state.h) ---------------------
    LN_COLD void game_result(OrxGameResult &r) const
    {
        uint32_t flags = fl & LF_ERRMASK;
        const bool dead = (flags & (ORX_FLAG_ERROR | ORX_FLAG_STREAM_END)) != 0u;
        if (dead) flags = (flags & ORX_FLAG_STREAM_END) ? ORX_FLAG_STREAM_END : ORX_FLAG_ERROR;
        LN_NOUNROLL
        for (int p = 0; p < ORX_MAX_PLAYERS; p++) r.player_out_on_turn[p] = (!dead && p < dm->P) ? pot(p) : 0;
        r.sim_turn_count = dead ? 0 : simTurn;
        r.flags = flags;
        r.stream_pos = dead ? 0u : ((fl & LF_TERMSET) ? termPos : pos);
        uint32_t hsum = 0;
        if (!dead) for (int t = 0; t < dm->N; t++) hsum += orx_mix((uint32_t)t, (uint32_t)owner(t), (uint32_t)armies(t));
        r.board_hash = hsum;
    }
};

#ifndef ORX_LANES_HOST
// ---------------------------------------------------------------------------------------------------------------
// The kernel: persistent warps; every lane pulls game indices from a ticket counter.
// ---------------------------------------------------------------------------------------------------------------
#ifndef ORX_LANE_THREADS
#define ORX_LANE_THREADS 64
#endif
#ifndef ORX_LANE_MIN_BLOCKS
#define ORX_LANE_MIN_BLOCKS 9
#endif

__device__ __noinline__ void lane_write_result(const Lane *Lp, const LaneArgs *arp)
{
    const Lane &L = *Lp; const LaneArgs &ar = *arp;
    OrxGameResult r;
    L.game_result(r);
    if (ar.games) {
        uint4 *o = (uint4 *)(ar.games + L.gidx);
        const uint32_t *w = (const uint32_t *)&r;
        o[0] = make_uint4(w[0], w[1], w[2], w[3]); o[1] = make_uint4(w[4], w[5], w[6], w[7]); o[2] = make_uint4(w[8], w[9], w[10], w[11]);
    }
    if (ar.agg) {
        unsigned long long *a = (unsigned long long *)(ar.agg + L.gidx / (unsigned long long)ar.playouts_per_leaf);
        atomicAdd(a + 0, 1ull); atomicAdd(a + 1, (unsigned long long)(long long)r.sim_turn_count);
#pragma unroll 1
        for (int p = 0; p < ORX_MAX_PLAYERS; p++)
            if (r.player_out_on_turn[p] == -1) { atomicAdd(a + 2 + p, 1ull); atomicAdd(a + 10 + p, (unsigned long long)(long long)r.sim_turn_count); }
        if (r.flags) atomicAdd(a + 18, 1ull);
    }
}

__global__ void __launch_bounds__(ORX_LANE_THREADS, ORX_LANE_MIN_BLOCKS)
rollout_lanes_kernel(const __grid_constant__ DevMap dm, const __grid_constant__ LaneArgs ar)
{
    extern __shared__ __align__(16) uint8_t smem[];
    {
        const uint4 *src = (const uint4 *)dm.blob; uint4 *dst = (uint4 *)smem;
        for (int i = threadIdx.x; i < dm.blob_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int lane = lane_id(), wid = (int)(threadIdx.x >> 5);
    const LaneLayout LL = lane_layout(dm.N, dm.P);
    uint8_t cards[LN_LISTS * LN_CAP];

    Lane L;
    L.dm = &dm; L.ar = &ar;
    L.mbase = to_saddr(smem);
    L.lbase = to_saddr(smem + dm.blob_bytes + wid * (LL.words * 128)) + 4u * (uint32_t)lane;
    L.oARM = LL.ARM; L.oOWN = LL.OWN; L.oAL = LL.AL; L.oPOT = LL.POT; L.oPC = LL.PC; L.oNC = LL.NC; L.oRING = LL.RING; L.oGAUSS = LL.GAUSS;
    L.cards = cards;
    L.st = LS_NEW; L.fl = 0u;
    L.cp = 0; L.phase = 0; L.simTurn = 0; L.remaining = 0; L.placeable = 0; L.cardpos = 0; L.nAtt = 0; L.playerTurns = 0; L.simStart = 0;
    L.pos = 0; L.lim = 0; L.mine_lo = L.mine_hi = 0; L.g0 = L.g1 = 0; L.gidx = 0; L.bA = L.bD = L.bai = L.ra = L.rd = 0; L.termPos = 0; L.deal = 0; L.selIters = 0;

    // the once-per-turn code runs out of line on a memory copy of the lane's state (LN_COLD)
#define LANE_COLD(call) do { Lane t_; (LaneEnv &)t_ = (const LaneEnv &)L; (LaneState &)t_ = (const LaneState &)L; t_.call; (LaneState &)L = (const LaneState &)t_; } while (0)

    const LaneTune tn = ar.tune;
    int aN = 0, aT = 0, aS = 0, aF = 0;
    bool force = false;
    for (;;) {
        // how many lanes wait for each section: one REDUX over 6-bit fields (NEW, TURN, SEL, DICE, FIN), one ballot for the rare states
        const uint32_t cnt = __reduce_add_sync(ORX_FULL, L.st <= LS_FIN ? 1u << (6 * L.st) : 0u);
        const uint32_t rare = __ballot_sync(ORX_FULL, L.st == LS_COOP || L.st == LS_GAUSS);   // (LS_ELIM never survives a trip)
        if ((cnt | rare) == 0u) break;   // every lane is LS_EXIT
        const int nN = (int)(cnt & 63u), nT = (int)((cnt >> 6) & 63u), nS = (int)((cnt >> 12) & 63u), nF = (int)((cnt >> 24) & 63u);
        bool ran = (cnt >> 18 & 63u) != 0u || rare != 0u;
        // NEW: write the finished game's record, take the next game, set it up
        if (nN) {
            if (force || nN >= tn.thr_new || aN >= tn.age_new) {
                aN = 0; ran = true;
                const uint32_t m = __ballot_sync(ORX_FULL, L.st == LS_NEW);
                const int leader = __ffs(m) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(ar.ticket, (unsigned long long)__popc(m));
                base = __shfl_sync(ORX_FULL, base, leader);
                if (L.st == LS_NEW) {
                    const unsigned long long idx = base + (unsigned long long)__popc(m & lanemask_lt());
                    Lane t_; (LaneEnv &)t_ = (const LaneEnv &)L; (LaneState &)t_ = (const LaneState &)L;
                    if (t_.fl & LF_HASGAME) lane_write_result(&t_, &ar);
                    t_.fl = 0u;
                    if (idx < ar.total) t_.start_game(idx); else t_.st = LS_EXIT;
                    (LaneState &)L = (const LaneState &)t_;
                }
            } else aN++;
        }
        if (nT) {
            if (force || nT >= tn.thr_turn || aT >= tn.age_turn) { aT = 0; ran = true; if (L.st == LS_TURN) LANE_COLD(sec_turn()); }
            else aT++;
        }
        // the one place the hot sections' stream words come from: at least 8 buffered for every lane about to run one
        if (L.st >= LS_SEL && L.st <= LS_FIN) L.topup();
        if (nS) {
            if (force || nS >= tn.thr_sel || aS >= tn.age_sel) { aS = 0; ran = true; if (L.st == LS_SEL) L.sec_sel(); }
            else aS++;
        }
        if (__any_sync(ORX_FULL, L.st == LS_GAUSS)) { ran = true; if (L.st == LS_GAUSS) LANE_COLD(sec_gauss()); }
        if (__any_sync(ORX_FULL, L.st == LS_DICE)) { ran = true; if (L.st == LS_DICE) L.sec_dice(); }
        // COOP: long dice battles, one lane after the other, rolled by the whole warp.  After DICE: a lane handed back to
        // DICE has restarted its ring and must pass the next trip's top-up first (FIN checks its reads)
        for (uint32_t mc = __ballot_sync(ORX_FULL, L.st == LS_COOP); mc; mc &= mc - 1u) {
            const int src = __ffs(mc) - 1;
            int cra = __shfl_sync(ORX_FULL, L.ra, src), crd = __shfl_sync(ORX_FULL, L.rd, src);
            uint32_t cpos = __shfl_sync(ORX_FULL, L.pos, src);
            const uint32_t cg0 = __shfl_sync(ORX_FULL, L.g0, src), cg1 = __shfl_sync(ORX_FULL, L.g1, src);
            warp_dice_3v2(dm, ar.rk, cra, crd, cpos, cg0, cg1);
            if (lane == src) {
                L.ra = cra; L.rd = crd; L.pos = cpos; L.lim = cpos & ~3u;   // the ring restarts at the counter that holds pos
                L.st = ((cra > 100 || crd > 100) && cra > 0 && crd > 0) ? LS_DICE : LS_FIN;
            }
            ran = true;
        }
        if (nF) {
            if (force || nF >= tn.thr_fin || aF >= tn.age_fin) {
                aF = 0; ran = true;
                if (L.st == LS_FIN) L.sec_fin();
                if (__any_sync(ORX_FULL, L.st == LS_ELIM)) { if (L.st == LS_ELIM) LANE_COLD(sec_elim()); }
            } else aF++;
        }
        force = !ran;
    }
#undef LANE_COLD
}
#endif  // !ORX_LANES_HOST

}  // namespace orx
