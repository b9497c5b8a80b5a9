// orx_device.cuh -- device-side building blocks of the rollout engine (sm_100a).
//
// One warp plays one game.  Everything a game mutates lives in that warp's slice of shared
// memory or in lane-distributed registers; all control flow is warp-uniform; lane-parallel
// work (board scans, defender selection, CDF search, dice batches, Philox) uses ballot /
// shuffle / redux.  Randomness is the injected word stream of include/orx_state.h consumed
// through java.util.Random's derivations, so a Java harness can replay a playout bit for bit.
//
// Shared memory is addressed through explicit 32-bit shared-window addresses (ld.shared /
// st.shared): the generic-pointer form made ptxas rebuild the window base (S2R SR_CgaCtaId + LEA)
// inside the attack loop (profiles/r1a).
//
// Reference files restated here (O/ = /root/reference/src/com/mld46/oponn/):
//   O/BattleSim.java (whole), O/CardManager.java:283-341 ; the game itself is in orx_game.cuh
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "orx_devmap.h"

namespace orx {

// floor(2^32 / b) + 1 for b < ORX_MAGIC_N: u % b for u < 2^31 in four instructions (next_int)
__constant__ uint32_t c_magic[ORX_MAGIC_N];

// per-warp header words (uint32 indices into the first 32 bytes of the slice)
enum { WH_GAUSS_LO = 0, WH_GAUSS_HI = 1, WH_HAVE_GAUSS = 2, WH_G0 = 3, WH_G1 = 4, WH_SIM_START = 5, WH_REM_COUNTRIES = 6,
       WH_DEAL_LO = 7, WH_DEAL_HI = 8, WH_DEAL_MODE = 9, WH_CUR_LEAF = 10 };   // CardManager.random (48-bit java.util.Random state) + ORX_DEAL_*
#define ORX_WARP_HDR_BYTES 48
#ifdef ORX_ACC_PER_WARP
#define ORX_WARP_ACC_BYTES 160   // 20 x u64: the image of one OrxLeafResult (rollout kernel: per-warp running sums)
#else
#define ORX_WARP_ACC_BYTES 0
#endif

__device__ __forceinline__ int lane_id() { int l; asm("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint32_t lanemask_le() { uint32_t m; asm("mov.u32 %0, %%lanemask_le;" : "=r"(m)); return m; }

// ---- shared memory through 32-bit shared-window addresses -------------------------------------
typedef uint32_t saddr;
__device__ __forceinline__ saddr to_saddr(const void *p) { return (saddr)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(saddr a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds8(saddr a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(saddr a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(saddr a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// index of the n-th (0-based) set bit of a warp-uniform mask; n < popc(mask)
__device__ __forceinline__ int nth_bit(uint32_t mask, int n)
{
    uint32_t hit = __ballot_sync(ORX_FULL, ((mask >> lane_id()) & 1u) && __popc(mask & lanemask_lt()) == n);
    return __ffs(hit) - 1;
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// all 32 lanes evaluate one Philox counter each: 128 consecutive words, lane l owning words 4l .. 4l+3 (orx_state.h)
#ifndef ORX_V_REFILL_OLD   // the refill reads the game id from the warp header itself: two arguments fewer at every word() site (+4 %)
__device__ __noinline__ void philox_refill_h(saddr buf, saddr hdr, uint32_t blk, uint32_t k0, uint32_t k1)
{
    uint32_t o[4];
    int l = lane_id();
    philox4x32_10(blk * 32u + (uint32_t)l, 0u, lds32(hdr + 12u), lds32(hdr + 16u), k0, k1, o);
    __syncwarp();
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(buf + 16u * (uint32_t)l), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
    __syncwarp();
}
#define ORX_REFILL(buf, hdr, blk, k0, k1) philox_refill_h(buf, hdr, blk, k0, k1)
#else
#define ORX_REFILL(buf, hdr, blk, k0, k1) philox_refill(buf, blk, lds32(hdr + 4 * WH_G0), lds32(hdr + 4 * WH_G1), k0, k1)
#endif
__device__ __noinline__ void philox_refill(saddr buf, uint32_t blk, uint32_t g0, uint32_t g1, uint32_t k0, uint32_t k1)
{
    uint32_t o[4];
    int l = lane_id();
    philox4x32_10(blk * 32u + (uint32_t)l, 0u, g0, g1, k0, k1, o);
    __syncwarp();
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(buf + 16u * (uint32_t)l), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// The injected word stream of one game + java.util.Random's derivations
// ---------------------------------------------------------------------------------------
template <bool REPLAY>
struct Stream {
    saddr     buf;          // this warp's 128-word block in shared memory
    saddr     hdr;          // this warp's header words (game id, cached gaussian)
    uint32_t  wi;           // counter mode: next unread word of the block (128 = block used up)
    uint32_t  base;         // counter mode: stream position of buf[0]; replay: words consumed
    uint32_t  k0, k1;       // Philox key = seed (kernel parameter, warp-uniform)
    const uint32_t *words;  // REPLAY: explicit words
    uint32_t  n_words;
    uint32_t  flags;        // ORX_FLAG_* raised by the stream (and by the game)

    __device__ __forceinline__ void init(saddr b, saddr h, uint64_t seed, uint64_t game, const uint32_t *w, uint32_t nw)
    {
        buf = b; hdr = h; wi = 128u; base = REPLAY ? 0u : (uint32_t)-128; k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        words = w; n_words = nw; flags = 0;
        __syncwarp();
        if (lane_id() == 0) {
            sts32(h + 4 * WH_G0, (uint32_t)game); sts32(h + 4 * WH_G1, (uint32_t)(game >> 32)); sts32(h + 4 * WH_HAVE_GAUSS, 0u);
        }
        __syncwarp();
    }

    __device__ __forceinline__ uint32_t pos() const { return REPLAY ? base : base + wi; }

    __device__ __forceinline__ void refill()
    {
        base += 128u;
        ORX_REFILL(buf, hdr, base >> 7, k0, k1);
        wi = 0u;
    }

    // counter mode: continue at stream position p (>= pos()): inside the buffered block, or with the block that holds p
    __device__ __forceinline__ void seek(uint32_t p)
    {
        if (p - base < 128u) { wi = p - base; return; }
        base = p & ~127u;
        ORX_REFILL(buf, hdr, base >> 7, k0, k1);
        wi = p & 127u;
    }

    __device__ __forceinline__ uint32_t word()
    {
        if (REPLAY) {
            uint32_t i = base++;
            if (i >= n_words) { if (!flags) flags |= ORX_FLAG_STREAM_END; return 0u; }  // only the first fault is recorded
            return __ldg(words + i);
        }
        if (wi == 128u) refill();
        uint32_t v = lds32(buf + 4u * wi);
        wi++;
        return v;
    }

    // java.util.Random.nextInt(int bound).  SMALL: the caller knows bound < ORX_MAGIC_N (list lengths, neighbour counts)
    template <bool SMALL = false>
    __device__ __forceinline__ int next_int(int bound)
    {
        if (bound <= 0) { flags |= ORX_FLAG_ERROR; return 0; }
        uint32_t u = word() >> 1;
        uint32_t b = (uint32_t)bound, m = b - 1u;
        if ((b & m) == 0u) return (int)(((unsigned long long)b * u) >> 31);
        for (;;) {
            uint32_t r;
            if (SMALL || b < (uint32_t)ORX_MAGIC_N) {  // u % b by multiplication: the estimate is q or q + 1
                uint32_t q = __umulhi(u, c_magic[b]);
                r = u - q * b;
                if ((int)r < 0) r += b;
            } else {
                r = u % b;
            }
            if ((int)(u - r + m) >= 0) return (int)r;
            if (flags) return 0;
            u = word() >> 1;
        }
    }

    // the 53-bit integer k of java.util.Random.nextDouble() == k * 2^-53
    __device__ __forceinline__ uint64_t next_k53()
    {
        uint32_t w0, w1;
        if (!REPLAY && wi <= 126u) {
            w0 = lds32(buf + 4u * wi); w1 = lds32(buf + 4u * wi + 4u); wi += 2u;
        } else {
            w0 = word(); w1 = word();
        }
        return ((uint64_t)(w0 >> 6) << 27) + (uint64_t)(w1 >> 5);
    }
    __device__ __forceinline__ double next_double() { return (double)(long long)next_k53() * 0x1.0p-53; }
};

// StrictMath.log == fdlibm __ieee754_log (e_log.c), restated with explicitly unfused IEEE operations.  The algorithm
// and its constants are fdlibm's:
//   Copyright (C) 1993 by Sun Microsystems, Inc. All rights reserved.
//   Developed at SunPro, a Sun Microsystems, Inc. business.  Permission to use, copy, modify, and distribute this
//   software is freely granted, provided that this notice is preserved.
__device__ __noinline__ double fdlibm_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                 Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
#define M(a, b) __dmul_rn((a), (b))
#define A(a, b) __dadd_rn((a), (b))
#define S(a, b) __dsub_rn((a), (b))
    int hx = __double2hiint(x);
    uint32_t lx = (uint32_t)__double2loint(x);
    int k = 0, i, j;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return __longlong_as_double(0xFFF0000000000000ll);
        if (hx < 0) return __longlong_as_double(0x7FF8000000000000ll);
        k -= 54; x = M(x, two54); hx = __double2hiint(x);
    }
    if (hx >= 0x7ff00000) return A(x, x);
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = __hiloint2double(hx | (i ^ 0x3ff00000), __double2loint(x));
    k += (i >> 20);
    double f = S(x, 1.0), dk, R;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) { if (k == 0) return 0.0; dk = (double)k; return A(M(dk, ln2_hi), M(dk, ln2_lo)); }
        R = M(M(f, f), S(0.5, M(0.33333333333333333, f)));
        if (k == 0) return S(f, R);
        dk = (double)k; return S(M(dk, ln2_hi), S(S(R, M(dk, ln2_lo)), f));
    }
    double s = __ddiv_rn(f, A(2.0, f));
    dk = (double)k;
    double z = M(s, s);
    i = hx - 0x6147a;
    double w = M(z, z);
    j = 0x6b851 - hx;
    double t1 = M(w, A(Lg2, M(w, A(Lg4, M(w, Lg6)))));
    double t2 = M(z, A(Lg1, M(w, A(Lg3, M(w, A(Lg5, M(w, Lg7)))))));
    i |= j;
    R = A(t2, t1);
    if (i > 0) {
        double hfsq = M(M(0.5, f), f);
        if (k == 0) return S(f, S(hfsq, M(s, A(hfsq, R))));
        return S(M(dk, ln2_hi), S(S(hfsq, A(M(s, A(hfsq, R)), M(dk, ln2_lo))), f));
    }
    if (k == 0) return S(f, M(s, S(f, R)));
    return S(M(dk, ln2_hi), S(S(M(s, S(f, R)), M(dk, ln2_lo)), f));
#undef M
#undef A
#undef S
}

// The arithmetic half of java.util.Random.nextGaussian() (Marsaglia polar method): from an
// accepted pair (v1, v2, q) to (first value, cached second value).
__device__ __noinline__ double gaussian_from_pair(double v1, double v2, double q, double *second)
{
    double multiplier = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, fdlibm_log(q)), q));
    *second = __dmul_rn(v2, multiplier);
    return __dmul_rn(v1, multiplier);
}

// java.util.Random.nextGaussian(): second value cached per playout (in the warp header)
template <bool REPLAY>
__device__ __forceinline__ double next_gaussian(Stream<REPLAY> &s)
{
    if (lds32(s.hdr + 4 * WH_HAVE_GAUSS)) {
        double g = __hiloint2double((int)lds32(s.hdr + 4 * WH_GAUSS_HI), (int)lds32(s.hdr + 4 * WH_GAUSS_LO));
        __syncwarp();
        if (lane_id() == 0) sts32(s.hdr + 4 * WH_HAVE_GAUSS, 0u);
        __syncwarp();
        return g;
    }
    double v1, v2, q;
    do {
        v1 = __dsub_rn(__dmul_rn(2.0, s.next_double()), 1.0);
        v2 = __dsub_rn(__dmul_rn(2.0, s.next_double()), 1.0);
        q = __dadd_rn(__dmul_rn(v1, v1), __dmul_rn(v2, v2));
        if (s.flags) return 0.0;
    } while (q >= 1.0 || q == 0.0);
    double second;
    double first = gaussian_from_pair(v1, v2, q, &second);
    // every lane computed the same pair; the shuffles only make that visible to the PTX uniformity pass
    first = __shfl_sync(ORX_FULL, first, 0); second = __shfl_sync(ORX_FULL, second, 0);
    __syncwarp();
    if (lane_id() == 0) {
        sts32(s.hdr + 4 * WH_GAUSS_LO, (uint32_t)__double2loint(second)); sts32(s.hdr + 4 * WH_GAUSS_HI, (uint32_t)__double2hiint(second));
        sts32(s.hdr + 4 * WH_HAVE_GAUSS, 1u);
    }
    __syncwarp();
    return first;
}

// Java (int) cast of a double: truncate toward zero, saturate, NaN -> 0  (== cvt.rzi.s32.f64)
__device__ __forceinline__ int java_d2i(double d) { return __double2int_rz(d); }

// Philox4x32-10 with the round keys precomputed on the host: rk[2r] = seed_lo + r * 0x9E3779B9, rk[2r+1] = seed_hi + r * 0xBB67AE85
__device__ __forceinline__ void philox4x32_10_rk(uint32_t c0, uint32_t c2, uint32_t c3, const uint32_t *rk, uint32_t o[4])
{
    uint32_t c1 = 0u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// The whole warp rolls one game's dice (both rollout kernels).  The "either side above 100" loop of BattleSim.simulateBattle
// (O/BattleSim.java:99-116) while the dice stay 3 v 2: 32 lanes evaluate 32 consecutive Philox counters = 128 stream
// words = up to 64 nextDouble()s, consumed in order from word `pos`; lane l holds the words 4 (c0 + l) .. + 3.  Two ways
// through a block, both consuming exactly the doubles the reference's loop would:
//   sum   no roll of the block can end the loop or change the dice: one REDUX over the attacker losses;
//   scan  otherwise the first roll after which the loop stops (or a side drops below 3 v 2) is found from a warp prefix
//         sum of the losses (three ballots) and two more ballots.
// All arguments are warp-uniform; rk = the 20 Philox round keys (a kernel parameter: constant-bank operands).
// Returns with ra, rd, pos advanced to just after the roll that ended the 3 v 2 regime.
__device__ __forceinline__ void warp_dice_3v2(const DevMap &dm, const uint32_t *rk, int &ra, int &rd, uint32_t &pos, uint32_t g0, uint32_t g1)
{
    const int lane = lane_id();
    const uint64_t t0 = dm.single_thr[5][0], t1 = dm.single_thr[5][1];
    const uint32_t lt = lanemask_lt();
    bool more = true;
#pragma unroll 1
    while (more) {
        const uint32_t e = pos & 3u, c0 = pos >> 2;
        uint32_t o[4];
        philox4x32_10_rk(c0 + (uint32_t)lane, g0, g1, rk, o);
        // doubles straddle the counters when pos is odd: (o1, o2) and (o3, next lane's o0)
        const uint32_t n0 = __shfl_down_sync(ORX_FULL, o[0], 1);
        const bool odd = (e & 1u) != 0u;
        const uint32_t a0 = odd ? o[1] : o[0], a1 = odd ? o[2] : o[1], b0 = odd ? o[3] : o[2], b1 = odd ? n0 : o[3];
        const bool vB = !(odd && lane == 31);
        const int rA = 2 * lane - (int)(e >> 1);      // index of this lane's first double among the block's rolls (-1: before pos)
        const bool vA = rA >= 0;
        const int nb = 64 - (int)(e & 1u) - (int)(e >> 1);
        const uint64_t kA = ((uint64_t)(a0 >> 6) << 27) + (uint64_t)(a1 >> 5), kB = ((uint64_t)(b0 >> 6) << 27) + (uint64_t)(b1 >> 5);
#ifdef ORX_DICE_HI32
        // compare the high 26 bits first: k > t  <=>  hi(k) > hi(t), unless they tie (2^-26 per compare: exact path then)
        const uint32_t hA = a0 >> 6, hB = b0 >> 6, t0h = (uint32_t)(t0 >> 27), t1h = (uint32_t)(t1 >> 27);
        int xA = (int)(hA > t0h) + (int)(hA > t1h), xB = (int)(hB > t0h) + (int)(hB > t1h);
        if (__any_sync(ORX_FULL, hA == t0h || hA == t1h || hB == t0h || hB == t1h)) { xA = (int)(kA > t0) + (int)(kA > t1); xB = (int)(kB > t0) + (int)(kB > t1); }
        if (!vA) xA = 0;
        if (!vB) xB = 0;
#else
        const int xA = vA ? (int)(kA > t0) + (int)(kA > t1) : 0, xB = vB ? (int)(kB > t0) + (int)(kB > t1) : 0;
#endif
        const int hi = ra > rd ? ra : rd;
        int la, played = nb;   // attacker losses and rolls consumed by this block
        if (ra - 2 * nb >= 3 && rd - 2 * nb >= 2 && (hi - 2 * nb > 100 || ra + rd - 2 * nb > 200)) {
            la = (int)__reduce_add_sync(ORX_FULL, (unsigned)(xA + xB));
        } else {
            const int sx = xA + xB;   // 0..4
            const uint32_t m0 = __ballot_sync(ORX_FULL, sx & 1), m1 = __ballot_sync(ORX_FULL, sx & 2), m2 = __ballot_sync(ORX_FULL, sx & 4);
            const int pre = __popc(m0 & lt) + 2 * __popc(m1 & lt) + 4 * __popc(m2 & lt);
            const int cA = pre + xA, cB = cA + xB;                  // attacker losses up to and including this lane's doubles
            const int raA = ra - cA, rdA = rd - (2 * (rA + 1) - cA), raB = ra - cB, rdB = rd - (2 * (rA + 2) - cB);
            const bool goA = (raA > 100 || rdA > 100) && raA >= 3 && rdA >= 2, goB = (raB > 100 || rdB > 100) && raB >= 3 && rdB >= 2;
            const uint32_t sA = __ballot_sync(ORX_FULL, vA && !goA), sB = __ballot_sync(ORX_FULL, vB && !goB);
            la = __popc(m0) + 2 * __popc(m1) + 4 * __popc(m2);      // the block ends before the loop does
            if (sA | sB) {
                const int fa = sA ? __ffs(sA) - 1 : 32, fb = sB ? __ffs(sB) - 1 : 32;
                const bool first = fa <= fb;                          // the stopping roll is lane fa's first / lane fb's second double
                la = __shfl_sync(ORX_FULL, first ? cA : cB, first ? fa : fb);
                played = (first ? 2 * fa + 1 : 2 * fb + 2) - (int)(e >> 1);
                more = false;
            }
        }
        ra -= la; rd -= 2 * played - la; pos += 2u * (uint32_t)played;
    }
}

// ---------------------------------------------------------------------------------------
// BattleSim  (O/BattleSim.java)
// ---------------------------------------------------------------------------------------

// One uniform against the (a,d) outcome table: first row with k <= floor(cdf_i * 2^53)
// (== "r <= total", O/BattleSim.java:129-140); no row => no losses.
// 32 rows are compared per step, one per lane.
__device__ __forceinline__ void table_battle(const DevMap &dm, int a, int d, uint64_t k, int &al, int &dl)
{
    // Row block of (a, d): tables are laid out (a, d) row-major with capacity a + d each, unused rows zero (a zero
    // threshold can only be reached by k = 0, which the first real row takes first), so neither the offset nor the
    // row count needs the meta table: one dependent global load per battle instead of two.
    const uint32_t off = (uint32_t)(50 * a * (a - 1) + 5050 * (a - 1) + a * (d - 1) + ((d * (d - 1)) >> 1));
    const uint64_t *rows = dm.bt_rows + off;
    const int cnt = a + d;
    const uint64_t key = k << 11;
    al = 0; dl = 0;
    for (int base = 0; base < cnt; base += 32) {
        int i = base + lane_id();
        uint64_t t = i < cnt ? __ldg(rows + i) : 0ull;
        uint32_t hit = __ballot_sync(ORX_FULL, key <= t);
        if (hit) {
            uint32_t code = __shfl_sync(ORX_FULL, (uint32_t)t & 0xFFu, __ffs(hit) - 1);
            int rem = (int)(code & 0x7Fu);
            if (code & 0x80u) { al = a - rem; dl = d; } else { al = a; dl = d - rem; }
            return;
        }
        if (__ballot_sync(ORX_FULL, i < cnt && t == 0ull)) return;   // past the last real row: r above the total mass
    }
}

// BattleSim.simulateIndividualBattle: attacker losses for one roll of pa v pd dice given k
__device__ __forceinline__ int single_roll(const DevMap &dm, int pa, int pd, uint64_t k)
{
    int row = (pa - 1) * 2 + (pd - 1);
    return k <= dm.single_thr[row][0] ? 0 : (k <= dm.single_thr[row][1] ? 1 : 2);
}

// BattleSim.simulateBattle (O/BattleSim.java:66-145); rk = Philox round keys of the launch (counter mode).
// WARP_DICE = false keeps the single rolls one at a time (the once-per-game catch-up battle: the dice block is inlined
// at one site only -- the kernel's instruction footprint matters, profiles/r2e).
template <bool REPLAY, bool WARP_DICE = true>
__device__ __forceinline__ void simulate_battle(const DevMap &dm, Stream<REPLAY> &s, const uint32_t *rk, int attackers, int defenders,
                                                int &alOut, int &dlOut)
{
    int attackerLosses = 0, defenderLosses = 0;
    if (attackers > 500 && defenders > 500) {
        int matched = java_d2i(__dmul_rn(0.922, (double)defenders));
        if (matched > attackers) {
            attackerLosses = attackers;
            int variance = java_d2i(__dmul_rn(next_gaussian(s), __dsqrt_rn((double)attackers)));
            double v = __dadd_rn((double)variance, __ddiv_rn((double)attackers, 0.922));
            defenderLosses = java_d2i(fmin(v, (double)(defenders - 1)));
        } else {
            defenderLosses = defenders;
            int variance = java_d2i(__dmul_rn(next_gaussian(s), __dsqrt_rn((double)defenders)));
            int v = (int)((uint32_t)variance + (uint32_t)matched);
            attackerLosses = v < attackers - 1 ? v : attackers - 1;
        }
    } else {
        int ra = attackers, rd = defenders;
#ifndef ORX_DICE_REGS   // default: 64-roll / 32-roll steps over the doubles left in the shared-memory word block.  (ORX_DICE_REGS: the
                       // register-resident block below, warp_dice_3v2 -- fewer instructions, measured slower in this kernel: DESIGN.md 8)
        // Single rolls while either side is above 100 (O/BattleSim.java:99-116).  Three ways through, all
        // consuming exactly the doubles the reference's loop would:
        //   scalar  one roll per trip when a side is about to run out (<= 2 armies), at a block boundary, or on replay;
        //   block   up to 64 rolls (2 per lane) when no roll of the step can end the loop or change the dice;
        //   scan    up to 32 rolls (1 per lane): the first roll after which the loop would stop is found by ballot.
        while ((ra > 100 || rd > 100) && ra > 0 && rd > 0 && !s.flags) {
            if (!REPLAY && ra > 2 && rd > 2) {   // 3 v 2 dice: lane-parallel steps over the doubles left in the block
                if (s.wi == 128u) s.refill();
                const int avail = (int)(128u - s.wi) >> 1;
                if (avail > 0) {
                    const uint64_t t0 = dm.single_thr[5][0], t1 = dm.single_thr[5][1];
                    const int l = lane_id();
                    const saddr p0 = s.buf + 4u * s.wi + 8u * (uint32_t)l;
                    const int hi = ra < rd ? rd : ra;
#ifdef ORX_V_SPEC
                    // Speculative block: play ALL the doubles left in the block at once and look at the state after the last
                    // of them.  Armies only fall, so if that state still has 3 v 2 dice and a side above 100, every state before
                    // it had too: the rolls were exactly the reference's.  Otherwise nothing is committed and the scan below
                    // finds the stopping roll.  Tried when the expected losses (about one army per side per roll) leave room.
                    if (ra - avail >= 3 && rd - avail >= 2 && (hi - avail > 100 || ra + rd - 2 * avail > 200)) {
                        int x = 0;
                        if (l < avail) {
                            uint64_t k = ((uint64_t)(lds32(p0) >> 6) << 27) + (uint64_t)(lds32(p0 + 4u) >> 5);
                            x = (k > t0) + (k > t1);
                        }
                        if (l + 32 < avail) {
                            uint64_t k = ((uint64_t)(lds32(p0 + 256u) >> 6) << 27) + (uint64_t)(lds32(p0 + 260u) >> 5);
                            x += (k > t0) + (k > t1);
                        }
                        const int la = (int)__reduce_add_sync(ORX_FULL, (unsigned)x), ld = 2 * avail - la;
                        const int ra2 = ra - la, rd2 = rd - ld;
                        if ((ra2 > 100 || rd2 > 100) && ra2 >= 3 && rd2 >= 2) {
                            ra = ra2; rd = rd2;
                            s.wi += 2u * (uint32_t)avail;
                            continue;
                        }
                    }
#else
                    const int slack = 2 * (avail - 1);  // armies lost before the last roll of the step, at most
                    if (avail >= 32 && ra - slack >= 3 && rd - slack >= 2 && hi - slack > 100) {
                        int x = 0;
                        {
                            uint64_t k = ((uint64_t)(lds32(p0) >> 6) << 27) + (uint64_t)(lds32(p0 + 4u) >> 5);
                            x = (k > t0) + (k > t1);
                        }
                        if (l + 32 < avail) {
                            uint64_t k = ((uint64_t)(lds32(p0 + 256u) >> 6) << 27) + (uint64_t)(lds32(p0 + 260u) >> 5);
                            x += (k > t0) + (k > t1);
                        }
                        int la = (int)__reduce_add_sync(ORX_FULL, (unsigned)x), ld = 2 * avail - la;
                        attackerLosses += la; defenderLosses += ld; ra -= la; rd -= ld;
                        s.wi += 2u * (uint32_t)avail;
                        continue;
                    }
#endif
                    const int np = avail > 32 ? 32 : avail;
                    int x = 0;
                    if (l < np) {
                        uint64_t k = ((uint64_t)(lds32(p0) >> 6) << 27) + (uint64_t)(lds32(p0 + 4u) >> 5);
                        x = (k > t0) + (k > t1);
                    }
                    // attacker losses up to and including roll l, from two ballots (x is 0, 1 or 2)
                    uint32_t b1 = __ballot_sync(ORX_FULL, x >= 1), b2 = __ballot_sync(ORX_FULL, x == 2);
                    uint32_t le = lanemask_le();
                    int cumA = __popc(b1 & le) + __popc(b2 & le);
                    int ra_l = ra - cumA, rd_l = rd - (2 * (l + 1) - cumA);
                    bool go_on = (ra_l > 100 || rd_l > 100) && ra_l >= 3 && rd_l >= 2;
                    uint32_t stop = __ballot_sync(ORX_FULL, !go_on || l >= np - 1);
                    int last = __ffs(stop) - 1;  // rolls 0..last are played
                    uint32_t upto = 0xFFFFFFFFu >> (31 - last);
                    int la = __popc(b1 & upto) + __popc(b2 & upto), ld = 2 * (last + 1) - la;
                    attackerLosses += la; defenderLosses += ld; ra -= la; rd -= ld;
                    s.wi += 2u * (uint32_t)(last + 1);
                    continue;
                }
            }
            // one roll: a side is down to its last armies, the next double straddles two blocks, or replay mode
            const int pa = ra < 3 ? ra : 3, pd = rd < 2 ? rd : 2, n = pa < pd ? pa : pd;
            const int x = single_roll(dm, pa, pd, s.next_k53());
            attackerLosses += x; defenderLosses += n - x; ra -= x; rd -= n - x;
        }
#else
        // Single rolls while either side is above 100 (O/BattleSim.java:99-116).  While the dice are 3 v 2 the whole warp
        // rolls up to 64 of them per step straight from Philox counters in registers (warp_dice_3v2: exactly the doubles the
        // reference's loop would draw); the few rolls with other dice (a side down to its last armies) and replayed streams
        // go one at a time through the word buffer.
        while ((ra > 100 || rd > 100) && ra > 0 && rd > 0 && !s.flags) {
            if (!REPLAY && WARP_DICE && ra >= 3 && rd >= 2) {
                uint32_t p = s.pos();
                warp_dice_3v2(dm, rk, ra, rd, p, lds32(s.hdr + 4 * WH_G0), lds32(s.hdr + 4 * WH_G1));
                s.seek(p);
                continue;
            }
            const int pa = ra < 3 ? ra : 3, pd = rd < 2 ? rd : 2, n = pa < pd ? pa : pd;
            const int x = single_roll(dm, pa, pd, s.next_k53());
            ra -= x; rd -= n - x;
        }
#endif
        attackerLosses = attackers - ra; defenderLosses = defenders - rd;
        if (ra != 0 && rd != 0 && !s.flags) {
            if (ra < 0 || rd < 0 || ra > ORX_TABLE_LIMIT || rd > ORX_TABLE_LIMIT) {
                s.flags |= ORX_FLAG_ERROR;
            } else {
                int al, dl;
                table_battle(dm, ra, rd, s.next_k53(), al, dl);
                attackerLosses += al; defenderLosses += dl;
            }
        }
    }
    alOut = attackerLosses; dlOut = defenderLosses;
}

// CardManager.getCardCashValue as integer closed forms (O/CardManager.java:283-341)
__device__ __forceinline__ int card_cash_value(const DevMap &dm, int pos, uint32_t &flags)
{
    switch (dm.prog_family) {
    case PROG_0: return 0;
    case PROG_555: return 5;
    case PROG_4466688: {  // (int)(ceil((sqrt(8*pos+9)-1)/2)*2): smallest k with (2k+1)^2 >= 8*pos+9
        long long v = 8ll * pos + 9ll;
        if (v <= 0) return 0;
        long long k = (long long)((__dsqrt_rn((double)v) - 1.0) * 0.5);
        while ((2 * k + 1) * (2 * k + 1) >= v && k > 0) k--;
        while ((2 * k + 1) * (2 * k + 1) < v) k++;
        return (int)(2 * k);
    }
    case PROG_456: return pos + 3;
    case PROG_468: return (pos + 1) * 2;
    case PROG_369: return pos * 3;
    case PROG_46810: return pos <= 4 ? (pos + 1) * 2 : (pos - 2) * 5;
    case PROG_51015: return pos * 5;
    case PROG_4681015202510: return pos <= 4 ? (pos + 1) * 2 : (pos <= 7 ? (pos - 2) * 5 : 10);
    case PROG_TABLE:
        if (pos < 0 || pos >= ORX_CARD_VALUES) { flags |= ORX_FLAG_ERROR; return 0; }
        return __ldg(dm.card_table + pos);
    default: return 2;  // the catch block (O/CardManager.java:333-339)
    }
}

}  // namespace orx
