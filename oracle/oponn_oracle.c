/*
 * oponn_oracle.c -- CPU restatement of Oponn's rollout path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA engine, never the thing shipped or measured:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (oponn_b200/) does not link, import or call it.
 *
 * It restates, function by function and in the reference's own control flow (ordered
 * Java Lists, linear CDF walks, scatter-order float sweeps), the code under
 * /root/reference/src/com/mld46/oponn/ (cited below as O/...).  The Java reference cannot
 * be compiled or run in this image (no JDK, Lux SDK not vendored -- SURVEY.md 8c), so:
 *
 *   PARITY PIN STATUS
 *   - BattleSim (tables + samplers): pinned by the reference's known-answer constants
 *     (O/BattleSim.java:15-44) and by ports of its three JUnit tests
 *     (O/tests/BattleSimTests.java:15-118) -- see tests/test_oracle_battlesim.py.
 *   - Whole-game results (SimulationBoard.simulate, SimRandom, CardManager): the
 *     reference holds no test, fixture or golden vector for them => "parity unpinned":
 *     pinned only by this line-by-line restatement and the rule invariants in tests/.
 *   - Card-set semantics come from the un-vendored Lux SDK (com.sillysoft.lux.Card,
 *     version unpinned).  DECLARED ASSUMPTION: a set is three equal symbols, three
 *     distinct symbols, or any triple holding a wildcard; Card.getRandomSet picks
 *     uniformly among the valid triples i<j<k of the hand in lexicographic order with
 *     one nextInt(count) draw.
 *
 * Randomness: one injected stream of 32-bit words per playout, consumed through
 * java.util.Random's documented derivations (include/orx_state.h, bottom).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: Java never fuses a*b+c).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/orx_state.h"

#define MAXN 256
#define MAXP ORX_MAX_PLAYERS
#define MAXC ORX_MAX_CONTINENTS

/* ------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3")   */
/* ------------------------------------------------------------------------------------ */

static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out)
{
    philox4x32_10(ctr, key, out);
}

/* ------------------------------------------------------------------------------------ */
/* The injected word stream + java.util.Random derivations                              */
/* ------------------------------------------------------------------------------------ */

typedef struct Stream {
    uint64_t seed, game;
    uint32_t pos;             /* words consumed                                       */
    uint32_t buf[128];        /* current 128-word block (counter mode)                */
    int64_t  buf_block;       /* index of the block in buf, -1 none                   */
    const uint32_t *inj;      /* replay mode: explicit words                          */
    uint32_t inj_len;
    int      have_gauss;      /* java.util.Random.haveNextNextGaussian (per playout)  */
    double   next_gauss;
    uint32_t *flags;          /* where to raise ORX_FLAG_STREAM_END / ERROR           */
} Stream;

static uint32_t stream_word(Stream *s)
{
    uint32_t i = s->pos++;
    if (s->inj) {
        if (i >= s->inj_len) { if (!*s->flags) *s->flags |= ORX_FLAG_STREAM_END; return 0; } /* only the first fault is recorded */
        return s->inj[i];
    }
    int64_t blk = i >> 7;
    if (blk != s->buf_block) {
        uint32_t key[2] = { (uint32_t)s->seed, (uint32_t)(s->seed >> 32) };
        for (uint32_t l = 0; l < 32; l++) { /* words 4j .. 4j+3 are the four outputs of counter j (orx_state.h) */
            uint32_t ctr[4] = { (uint32_t)(blk * 32 + l), 0u, (uint32_t)s->game, (uint32_t)(s->game >> 32) };
            uint32_t o[4];
            philox4x32_10(ctr, key, o);
            for (int c = 0; c < 4; c++) s->buf[4 * l + c] = o[c];
        }
        s->buf_block = blk;
    }
    return s->buf[i & 127];
}

/* java.util.Random.next(bits) with the LCG replaced by the injected word source */
static int32_t jr_next(Stream *s, int bits) { return (int32_t)(stream_word(s) >> (32 - bits)); }

/* java.util.Random.nextInt(int bound) */
static int32_t jr_next_int(Stream *s, int32_t bound)
{
    if (bound <= 0) { *s->flags |= ORX_FLAG_ERROR; return 0; } /* IllegalArgumentException */
    int32_t r = jr_next(s, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) {
        r = (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
    } else {
        for (int32_t u = r;; u = jr_next(s, 31)) {
            r = u % bound;
            /* "u - r + m < 0" in Java int arithmetic (wraps) */
            int32_t t = (int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m);
            if (t >= 0) break;
            if (*s->flags) break;
        }
    }
    return r;
}

/* java.util.Random.nextDouble(): (((long)next(26) << 27) + next(27)) * 0x1.0p-53 */
static double jr_next_double(Stream *s)
{
    int64_t hi = jr_next(s, 26);
    int64_t lo = jr_next(s, 27);
    return (double)((hi << 27) + lo) * 0x1.0p-53;
}

/* StrictMath.log == fdlibm __ieee754_log (e_log.c 1.3 95/01/18), restated.  Algorithm and constants are fdlibm's:
 *   Copyright (C) 1993 by Sun Microsystems, Inc. All rights reserved.
 *   Developed at SunPro, a Sun Microsystems, Inc. business.  Permission to use, copy, modify, and distribute this
 *   software is freely granted, provided that this notice is preserved. */
static double fdlibm_log(double x)
{
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                        Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                        Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
                        Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    uint64_t bits; memcpy(&bits, &x, 8);
    int32_t hx = (int32_t)(bits >> 32);
    uint32_t lx = (uint32_t)bits;
    int32_t k = 0, i, j;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -two54 / 0.0;
        if (hx < 0) return (x - x) / 0.0;
        k -= 54; x *= two54;
        memcpy(&bits, &x, 8); hx = (int32_t)(bits >> 32);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    memcpy(&bits, &x, 8);
    bits = (bits & 0xffffffffull) | ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32);
    memcpy(&x, &bits, 8);
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) { if (k == 0) return 0.0; dk = (double)k; return dk * ln2_hi + dk * ln2_lo; }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k; return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

double orc_fdlibm_log(double x) { return fdlibm_log(x); }

/* java.util.Random.nextGaussian(): Marsaglia polar method, second value cached */
static double jr_next_gaussian(Stream *s)
{
    if (s->have_gauss) { s->have_gauss = 0; return s->next_gauss; }
    double v1, v2, q;
    do {
        v1 = 2 * jr_next_double(s) - 1;
        v2 = 2 * jr_next_double(s) - 1;
        q = v1 * v1 + v2 * v2;
        if (*s->flags) return 0.0;
    } while (q >= 1 || q == 0);
    double multiplier = sqrt(-2 * fdlibm_log(q) / q); /* StrictMath.sqrt is correctly rounded */
    s->next_gauss = v2 * multiplier;
    s->have_gauss = 1;
    return v1 * multiplier;
}

/* Java (int) cast of a double: truncates, saturates, NaN -> 0 */
static int32_t java_d2i(double d)
{
    if (d != d) return 0;
    if (d >= 2147483647.0) return 2147483647;
    if (d <= -2147483648.0) return (int32_t)(-2147483647 - 1);
    return (int32_t)d;
}

/* test hooks for the derivations */
void orc_stream_draws(uint64_t seed, uint64_t game, const uint32_t *inj, uint32_t inj_len,
                      const int32_t *ops, int n_ops, double *out, uint32_t *pos_out, uint32_t *flags_out)
{
    /* ops[i]: >0 nextInt(ops[i]); 0 nextDouble; -1 nextGaussian; -2 raw word */
    uint32_t flags = 0;
    Stream s; memset(&s, 0, sizeof s);
    s.seed = seed; s.game = game; s.buf_block = -1; s.inj = inj; s.inj_len = inj_len; s.flags = &flags;
    for (int i = 0; i < n_ops; i++) {
        if (ops[i] > 0) out[i] = (double)jr_next_int(&s, ops[i]);
        else if (ops[i] == 0) out[i] = jr_next_double(&s);
        else if (ops[i] == -1) out[i] = jr_next_gaussian(&s);
        else out[i] = (double)stream_word(&s);
    }
    *pos_out = s.pos; *flags_out = flags;
}

/* ------------------------------------------------------------------------------------ */
/* BattleSim  (O/BattleSim.java)                                                        */
/* ------------------------------------------------------------------------------------ */

/* O/BattleSim.java:15-33 */
static const float a1_d1_al0 = 0.4167f, a1_d1_al1 = 0.5833f;
static const float a1_d2_al0 = 0.2546f, a1_d2_al1 = 0.7454f;
static const float a2_d1_al0 = 0.5787f, a2_d1_al1 = 0.4213f;
static const float a2_d2_al0 = 0.2276f, a2_d2_al1 = 0.3241f, a2_d2_al2 = 0.4483f;
static const float a3_d1_al0 = 0.6597f, a3_d1_al1 = 0.3403f;
static const float a3_d2_al0 = 0.3717f, a3_d2_al1 = 0.3358f, a3_d2_al2 = 0.2926f;

/* O/BattleSim.java:36-44 -- float sums widened to double */
static double single_cdf[6][3];
static void init_single_cdf(void)
{
    const double rows[6][3] = {
        { a1_d1_al0, 1.0, 1.0 },
        { a1_d2_al0, 1.0, 1.0 },
        { a2_d1_al0, 1.0, 1.0 },
        { a2_d2_al0, (float)(a2_d2_al0 + a2_d2_al1), 1.0 },
        { a3_d1_al0, 1.0, 1.0 },
        { a3_d2_al0, (float)(a3_d2_al0 + a3_d2_al1), 1.0 },
    };
    memcpy(single_cdf, rows, sizeof rows);
}

typedef struct AttackOutcome { /* O/moves/AttackOutcome.java:7-10 */
    float   probability;
    int32_t attackerLosses, defenderLosses;
    int32_t invading;
} AttackOutcome;

typedef struct OutcomeTable { AttackOutcome *rows; int n; } OutcomeTable;

#define TABLE_LIMIT 100 /* simulateBattle only looks tables up once both sides are <= 100
                           (O/BattleSim.java:99-118) */
static OutcomeTable g_tables[TABLE_LIMIT + 1][TABLE_LIMIT + 1];
static int g_tables_ready = 0;

/* O/BattleSim.java:190-262 */
static float *generateProbabilities(int attackers, int defenders)
{
    int rowLength = defenders + 1, colLength = attackers + 1;
    size_t cells = (size_t)colLength * rowLength;
    float *lastRow = calloc(cells, sizeof(float));
    float *newRow = malloc(cells * sizeof(float));
    lastRow[attackers * rowLength + defenders] = 1.0f;
    int allAbsorbed = 0;
    while (!allAbsorbed) {
        memset(newRow, 0, cells * sizeof(float));
        allAbsorbed = 1;
        for (int a = 0; a < colLength; a++) {
            for (int d = 0; d < rowLength; d++) {
                float v = lastRow[a * rowLength + d];
                if (v != 0) {
                    if (a == 0 || d == 0) {
                        newRow[a * rowLength + d] += v;
                    } else if (a == 1 && d == 1) {
                        newRow[1 * rowLength + 0] += v * a1_d1_al0;
                        newRow[0 * rowLength + 1] += v * a1_d1_al1;
                        allAbsorbed = 0;
                    } else if (a == 2 && d == 1) {
                        newRow[2 * rowLength + 0] += v * a2_d1_al0;
                        newRow[1 * rowLength + 1] += v * a2_d1_al1;
                        allAbsorbed = 0;
                    } else if (a == 1) {
                        newRow[1 * rowLength + (d - 1)] += v * a1_d2_al0;
                        newRow[0 * rowLength + (d - 0)] += v * a1_d2_al1;
                        allAbsorbed = 0;
                    } else if (a == 2) {
                        newRow[2 * rowLength + (d - 2)] += v * a2_d2_al0;
                        newRow[1 * rowLength + (d - 1)] += v * a2_d2_al1;
                        newRow[0 * rowLength + (d - 0)] += v * a2_d2_al2;
                        allAbsorbed = 0;
                    } else if (d == 1) {
                        newRow[(a - 0) * rowLength + (1 - 1)] += v * a3_d1_al0;
                        newRow[(a - 1) * rowLength + (1 - 0)] += v * a3_d1_al1;
                        allAbsorbed = 0;
                    } else {
                        newRow[(a - 0) * rowLength + (d - 2)] += v * a3_d2_al0;
                        newRow[(a - 1) * rowLength + (d - 1)] += v * a3_d2_al1;
                        newRow[(a - 2) * rowLength + (d - 0)] += v * a3_d2_al2;
                        allAbsorbed = 0;
                    }
                }
            }
        }
        float *t = lastRow; lastRow = newRow; newRow = t;
    }
    free(newRow);
    return lastRow;
}

/* O/BattleSim.java:264-286.  Collections.sort(list, reverseOrder()) is a stable merge
 * sort on AttackOutcome.compareTo (O/moves/AttackOutcome.java:42-45): descending
 * probability, ties keep insertion (remA asc, remD asc) order.  Stable insertion sort. */
static OutcomeTable generateAttackOutcomes(int a, int d)
{
    OutcomeTable t; t.rows = malloc(sizeof(AttackOutcome) * (size_t)(a * d + a + d + 1)); t.n = 0;
    float *probabilities = generateProbabilities(a, d);
    for (int remA = 0; remA <= a; remA++) {
        for (int remD = 0; remD <= d; remD++) {
            float probability = probabilities[remA * (d + 1) + remD];
            if (probability != 0.0) {
                AttackOutcome o = { probability, a - remA, d - remD, remD == 0 };
                int i = t.n++;
                while (i > 0 && t.rows[i - 1].probability < o.probability) { t.rows[i] = t.rows[i - 1]; i--; }
                t.rows[i] = o;
            }
        }
    }
    free(probabilities);
    t.rows = realloc(t.rows, sizeof(AttackOutcome) * (size_t)(t.n ? t.n : 1));
    return t;
}

typedef struct TableJob { int tid, nthreads; } TableJob;
static void *table_worker(void *arg)
{
    TableJob *j = arg;
    int idx = 0;
    /* big tables first so the threads stay balanced */
    for (int s = 2 * TABLE_LIMIT; s >= 2; s--)
        for (int a = 1; a <= TABLE_LIMIT; a++) {
            int d = s - a;
            if (d < 1 || d > TABLE_LIMIT) continue;
            if ((idx++ % j->nthreads) == j->tid) g_tables[a][d] = generateAttackOutcomes(a, d);
        }
    return NULL;
}

/* The reference pre-builds 1..49 x 1..49 at class load and the rest lazily
 * (O/BattleSim.java:46-59,122-127); contents do not depend on when a table is built. */
void orc_init(int nthreads)
{
    if (g_tables_ready) return;
    init_single_cdf();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    pthread_t th[64]; TableJob jobs[64];
    for (int i = 0; i < nthreads; i++) { jobs[i].tid = i; jobs[i].nthreads = nthreads; pthread_create(&th[i], NULL, table_worker, &jobs[i]); }
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    g_tables_ready = 1;
}

/* BattleSim.getAttackOutcomes (O/BattleSim.java:174-184): rows copied out for tests.
 * Works for any a,d >= 0 (tests draw up to 199); tables above the limit are built ad hoc. */
int orc_get_attack_outcomes(int a, int d, float *prob, int32_t *al, int32_t *dl, int32_t *inv, int cap)
{
    OutcomeTable t, *tp;
    int adhoc = !(a >= 1 && d >= 1 && a <= TABLE_LIMIT && d <= TABLE_LIMIT && g_tables_ready);
    if (adhoc) { t = generateAttackOutcomes(a, d); tp = &t; } else tp = &g_tables[a][d];
    int n = tp->n;
    for (int i = 0; i < n && i < cap; i++) {
        prob[i] = tp->rows[i].probability; al[i] = tp->rows[i].attackerLosses;
        dl[i] = tp->rows[i].defenderLosses; inv[i] = tp->rows[i].invading;
    }
    if (adhoc) free(t.rows);
    return n;
}

typedef struct Counters { /* algorithmic work, SURVEY.md 8(d) */
    uint64_t tv, ev, r_words, battles, cdf_steps, writes, player_turns, attacks, conquests, single_rolls, gauss_battles, table_battles;
    uint64_t cls[8]; /* battle classes by size: see simulateBattle */
} Counters;

static int ceil_log2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

/* O/BattleSim.java:152-162 */
static int simulateIndividualBattle(Stream *s, int attackers, int defenders)
{
    if (attackers > 3 || defenders > 2 || attackers < 1 || defenders < 1) { *s->flags |= ORX_FLAG_ERROR; return 0; }
    double r = jr_next_double(s);
    const double *distribution = single_cdf[(attackers - 1) * 2 + (defenders - 1)];
    return (r <= distribution[0] ? 0 : (r <= distribution[1] ? 1 : 2));
}

/* O/BattleSim.java:66-145 */
static void simulateBattle(Stream *s, Counters *ct, int attackers, int defenders, int *alOut, int *dlOut)
{
    int attackerLosses = 0, defenderLosses = 0;
    if (ct) { /* workload statistics only */
        int c = attackers > 100 ? (defenders == 1 ? 0 : defenders == 2 ? 1 : defenders <= 100 ? 2 : 3)
                                : (defenders > 100 ? (attackers == 1 ? 4 : attackers == 2 ? 5 : 6) : 7);
        ct->cls[c]++;
    }
    if (attackers > 500 && defenders > 500) {
        int matchedAttackers = java_d2i(0.922 * defenders);
        if (ct) ct->gauss_battles++;
        if (matchedAttackers > attackers) {
            attackerLosses = attackers;
            int variance = java_d2i(jr_next_gaussian(s) * sqrt((double)attackers));
            defenderLosses = java_d2i(fmin(variance + attackers / 0.922, (double)(defenders - 1)));
        } else {
            defenderLosses = defenders;
            int variance = java_d2i(jr_next_gaussian(s) * sqrt((double)defenders));
            int v = (int32_t)((uint32_t)variance + (uint32_t)matchedAttackers);
            attackerLosses = v < attackers - 1 ? v : attackers - 1;
        }
    } else {
        int remainingAttackers = attackers, remainingDefenders = defenders;
        if (remainingAttackers > 100 || remainingDefenders > 100) {
            while ((remainingAttackers > 100 || remainingDefenders > 100) && !(remainingAttackers == 0 || remainingDefenders == 0)) {
                int pa = remainingAttackers < 3 ? remainingAttackers : 3;
                int pd = remainingDefenders < 2 ? remainingDefenders : 2;
                int attackerLoss = simulateIndividualBattle(s, pa, pd);
                int defenderLoss = (pa < pd ? pa : pd) - attackerLoss;
                if (ct) ct->single_rolls++;
                attackerLosses += attackerLoss; defenderLosses += defenderLoss;
                remainingAttackers -= attackerLoss; remainingDefenders -= defenderLoss;
                if (*s->flags) break;
            }
        }
        if (remainingAttackers != 0 && remainingDefenders != 0 && !*s->flags) {
            if (remainingAttackers < 0 || remainingDefenders < 0 || remainingAttackers > TABLE_LIMIT || remainingDefenders > TABLE_LIMIT) {
                *s->flags |= ORX_FLAG_ERROR; /* negative sizes: generateProbabilities would throw */
            } else {
                OutcomeTable *t = &g_tables[remainingAttackers][remainingDefenders];
                double r = jr_next_double(s);
                double total = 0;
                for (int i = 0; i < t->n; i++) {
                    total += t->rows[i].probability;
                    if (r <= total) {
                        attackerLosses += t->rows[i].attackerLosses;
                        defenderLosses += t->rows[i].defenderLosses;
                        break;
                    }
                }
                if (ct) { ct->cdf_steps += (uint64_t)ceil_log2(t->n); ct->table_battles++; }
            }
        }
    }
    if (ct) ct->battles++;
    *alOut = attackerLosses; *dlOut = defenderLosses;
}

/* batched sampler hook for the BattleSimTests ports */
void orc_simulate_battles(uint64_t seed, uint64_t first_game, int a, int d, int n, int32_t *al, int32_t *dl)
{
    for (int i = 0; i < n; i++) {
        uint32_t flags = 0;
        Stream s; memset(&s, 0, sizeof s);
        s.seed = seed; s.game = first_game + (uint64_t)i; s.buf_block = -1; s.flags = &flags;
        int x = 0, y = 0;
        if (a >= 1 && d >= 1) simulateBattle(&s, NULL, a, d, &x, &y);
        al[i] = x; dl[i] = y;
    }
}

void orc_simulate_individual(uint64_t seed, uint64_t first_game, int a, int d, int n, int32_t *al)
{
    init_single_cdf();
    for (int i = 0; i < n; i++) {
        uint32_t flags = 0;
        Stream s; memset(&s, 0, sizeof s);
        s.seed = seed; s.game = first_game + (uint64_t)i; s.buf_block = -1; s.flags = &flags;
        al[i] = simulateIndividualBattle(&s, a, d);
    }
}

/* ------------------------------------------------------------------------------------ */
/* CardManager, simulation half  (O/CardManager.java:240-346)                           */
/* ------------------------------------------------------------------------------------ */

typedef struct CardList { uint8_t v[MAXN + 8]; int n; } CardList; /* a java.util.List<Card> */

static void list_add(CardList *l, uint8_t c) { if (l->n < MAXN + 8) l->v[l->n++] = c; }
static uint8_t list_remove_at(CardList *l, int i)
{
    uint8_t c = l->v[i];
    memmove(&l->v[i], &l->v[i + 1], (size_t)(l->n - i - 1));
    l->n--;
    return c;
}

enum { PROG_UNKNOWN = 0, PROG_0, PROG_555, PROG_4466688, PROG_456, PROG_468, PROG_369, PROG_46810, PROG_51015, PROG_4681015202510, PROG_EXP5, PROG_EXP4 };

typedef struct Progression { int family; double exp; } Progression;

static Progression parse_progression(const char *cp)
{
    Progression p = { PROG_UNKNOWN, 1.0 };
    if (!cp) return p;
    if (strchr(cp, '%')) { /* O/CardManager.java:287-305 */
        const char *hat = strchr(cp, '^');
        int pct = 0;
        if (hat) { /* Integer.valueOf(cp.substring(indexOf('^')+1, length-1)) */
            char tmp[32]; size_t len = strlen(hat + 1);
            if (len >= 1 && len - 1 < sizeof tmp) { memcpy(tmp, hat + 1, len - 1); tmp[len - 1] = 0; pct = atoi(tmp); }
        }
        p.family = cp[0] == '5' ? PROG_EXP5 : PROG_EXP4;
        p.exp = 1 + (pct / 100.0);
        return p;
    }
    if (!strcmp(cp, "0")) p.family = PROG_0;
    else if (!strcmp(cp, "5, 5, 5...")) p.family = PROG_555;
    else if (!strcmp(cp, "4, 4, 6, 6, 6, 8, 8, 8, 8, 10...")) p.family = PROG_4466688;
    else if (!strcmp(cp, "4, 5, 6...")) p.family = PROG_456;
    else if (!strcmp(cp, "4, 6, 8...")) p.family = PROG_468;
    else if (!strcmp(cp, "3, 6, 9...")) p.family = PROG_369;
    else if (!strcmp(cp, "4, 6, 8, 10, 15, 20...")) p.family = PROG_46810;
    else if (!strcmp(cp, "5, 10, 15...")) p.family = PROG_51015;
    else if (!strcmp(cp, "4, 6, 8, 10, 15, 20, 25, 10, 10, 10...")) p.family = PROG_4681015202510;
    return p;
}

/* O/CardManager.java:283-341 */
static int getCardCashValue(const Progression *pr, int pos)
{
    int base, offset;
    switch (pr->family) {
    case PROG_EXP5: base = 5 * pos; offset = 6; break;
    case PROG_EXP4: base = pos <= 4 ? 2 * (pos + 1) : 5 * (pos - 2); offset = 8; break;
    case PROG_0: return 0;
    case PROG_555: return 5;
    case PROG_4466688: return java_d2i(ceil((sqrt((double)(8 * pos + 9)) - 1) / 2) * 2);
    case PROG_456: return pos + 3;
    case PROG_468: return (pos + 1) * 2;
    case PROG_369: return pos * 3;
    case PROG_46810: return pos <= 4 ? (pos + 1) * 2 : (pos - 2) * 5;
    case PROG_51015: return pos * 5;
    case PROG_4681015202510: return pos <= 4 ? (pos + 1) * 2 : (pos <= 7 ? (pos - 2) * 5 : 10);
    default: return 2; /* the catch block (O/CardManager.java:333-339) */
    }
    return java_d2i(floor(base <= 25 ? (double)base : base * pow(pr->exp, (double)(pos - offset))));
}

int orc_card_cash_value(const char *progression, int pos)
{
    Progression p = parse_progression(progression);
    return getCardCashValue(&p, pos);
}

/* ------------------------------------------------------------------------------------ */
/* The board  (O/sim/boards/SimulationBoard.java, ModellingBoard.java)                  */
/* ------------------------------------------------------------------------------------ */

typedef struct Board {
    /* static */
    int N, C, P;
    const int32_t *adj_off, *adj, *continent, *base_bonus;
    int card_symbol[MAXN];
    int useCards, transferCards, immediateCash, continentIncrease;
    Progression progression;
    int maxPlayerTurns;
    int canRunOutOfCards;

    /* countries */
    int owner[MAXN], armies[MAXN];

    /* game state (SimulationBoard.java:62-96) */
    int playerCountries[MAXP];
    int missing[MAXP][MAXC]; /* countriesPlayerMissingInContinent */
    int currentContinentBonuses[MAXC];
    CardList playerCards[MAXP];
    CardList simDeck;
    int simCardProgressionPos;
    int remainingPlayers, currentPlayer, currentPhase, simStartTurn, simTurnCount;
    int playerOutOnTurn[MAXP];
    int remainingCountries;
    int initialArmiesLeftToPlace[MAXP];
    int numberOfPlaceableArmies;
    int hasInvadedACountry, attackingCC, defendingCC, attackingPlayer, defendingPlayer;
    int attackUntilDead, attackerLosses, defenderLosses, attackState;

    /* opponent modelling (ModellingBoard.java, SimAgent.java:248-307): policy id per player */
    int simAgentInitial[MAXP], simAgent[MAXP];
    int modellingTurnLimit;
    int moveable[MAXN];       /* Country.getMoveableArmies */

    /* CardManager.random: re-created by simReset on every setupState (O/CardManager.java:245) */
    int dealMode;             /* ORX_DEAL_* of the leaf record                          */
    uint64_t dealLcg;         /* java.util.Random's 48-bit state                        */

    /* engine bookkeeping */
    Stream rng;
    uint32_t flags;
    uint32_t playerTurns;
    int stopAtSimTurn;        /* fixture generation only: stop once simTurnCount reaches this */
    uint32_t termPos; int termSet;
    Counters ct;
} Board;

#define FAIL(b) do { (b)->flags |= ORX_FLAG_ERROR; } while (0)

static int isASet(const Board *b, uint8_t c1, uint8_t c2, uint8_t c3)
{
    /* DECLARED ASSUMPTION (Lux SDK Card.isASet not vendored): see file header */
    if (c1 == ORX_CARD_WILD || c2 == ORX_CARD_WILD || c3 == ORX_CARD_WILD) return 1;
    int s1 = b->card_symbol[c1], s2 = b->card_symbol[c2], s3 = b->card_symbol[c3];
    if (s1 == s2 && s2 == s3) return 1;
    if (s1 != s2 && s2 != s3 && s1 != s3) return 1;
    return 0;
}

static int containsASet(const Board *b, const CardList *h)
{
    for (int i = 0; i < h->n; i++)
        for (int j = i + 1; j < h->n; j++)
            for (int k = j + 1; k < h->n; k++)
                if (isASet(b, h->v[i], h->v[j], h->v[k])) return 1;
    return 0;
}

/* Card.getRandomSet (declared assumption): uniform over valid triples, lexicographic */
static int getRandomSet(Board *b, const CardList *h, int idx[3])
{
    int count = 0;
    for (int i = 0; i < h->n; i++)
        for (int j = i + 1; j < h->n; j++)
            for (int k = j + 1; k < h->n; k++)
                if (isASet(b, h->v[i], h->v[j], h->v[k])) count++;
    if (count == 0) { FAIL(b); return 0; } /* null -> NullPointerException */
    int pick = jr_next_int(&b->rng, count);
    int seen = 0;
    for (int i = 0; i < h->n; i++)
        for (int j = i + 1; j < h->n; j++)
            for (int k = j + 1; k < h->n; k++)
                if (isASet(b, h->v[i], h->v[j], h->v[k])) {
                    if (seen == pick) { idx[0] = i; idx[1] = j; idx[2] = k; return 1; }
                    seen++;
                }
    return 0;
}

/* java.util.Random itself (the LCG of the JDK documentation): CardManager.random after simReset's
 * `new Random(simDeck.hashCode())` (O/CardManager.java:245) */
static void lcg_set_seed(uint64_t *st, int64_t seed) { *st = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
static int32_t lcg_next(uint64_t *st, int bits)
{
    *st = (*st * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)(int64_t)(*st >> (48 - bits));
}
static int32_t lcg_next_int(uint64_t *st, int32_t bound) /* bound > 0: the deck is not empty */
{
    int32_t r = lcg_next(st, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
    for (int32_t u = r;; u = lcg_next(st, 31)) {
        r = u % bound;
        if ((int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m) >= 0) return r;
    }
}

/* O/CardManager.java:250-263 ; returns 0 when the deck is legitimately empty */
static int simDeal(Board *b, uint8_t *card)
{
    if (b->simDeck.n == 0) {
        if (!b->canRunOutOfCards) FAIL(b);
        return 0;
    }
    *card = list_remove_at(&b->simDeck, b->dealMode == ORX_DEAL_GAME_STREAM ? jr_next_int(&b->rng, b->simDeck.n)
                                                                           : lcg_next_int(&b->dealLcg, b->simDeck.n));
    b->ct.writes++;
    return 1;
}

/* SimulationBoard.java:956-963 */
static void recalculateContinentBonuses(Board *b)
{
    int e = b->simStartTurn + b->simTurnCount - 1; if (e < 0) e = 0;
    if (b->continentIncrease != 0 && e >= ORX_BONUS_ROUNDS) { FAIL(b); return; } /* engine table limit */
    double factor = pow(1 + b->continentIncrease / 100.0, (double)e);
    for (int i = 0; i < b->C; i++)
        b->currentContinentBonuses[i] = java_d2i(floor(b->base_bonus[i] * factor));
}

/* SimulationBoard.java:229-276 (+ ModellingBoard.java:21-27, no observable effect) */
static void setupGeneralGameState(Board *b)
{
    b->remainingPlayers = 0;
    b->simTurnCount = 0;
    for (int p = 0; p < b->P; p++) b->simAgent[p] = b->simAgentInitial[p]; /* ModellingBoard.java:21-27 */
    for (int p = 0; p < b->P; p++) {
        b->playerOutOnTurn[p] = 0;
        b->playerCountries[p] = 0;
        for (int c = 0; c < b->C; c++) b->missing[p][c] = 0;
    }
    for (int cc = 0; cc < b->N; cc++) {
        int owner = b->owner[cc], continent = b->continent[cc];
        for (int p = 0; p < b->P; p++) if (p != owner) b->missing[p][continent]++;
        if (!(b->currentPhase == ORX_PHASE_COUNTRY_SELECTION && owner == -1)) {
            if (owner < 0 || owner >= b->P) { FAIL(b); return; } /* ArrayIndexOutOfBounds */
            b->playerCountries[owner]++;
        }
    }
    if (b->currentPhase == ORX_PHASE_COUNTRY_SELECTION) b->remainingPlayers = b->P;
    else for (int p = 0; p < b->P; p++) if (b->playerCountries[p] > 0) b->remainingPlayers++;
}

/* SimulationBoard.java:518-533 */
static void calculateStartingArmies(const Board *b, int *armies)
{
    double f1 = b->N / 42.0;
    f1 -= 1.0; f1 /= 2.0; f1 += 1.0;
    int base = (int)floor(f1 * (50 - b->P * 5) + 0.5); /* Math.round */
    for (int i = 0; i < b->P; i++) armies[i] = base + (i > 1 ? i : 0);
}

int orc_starting_armies(int n_territories, int n_players, int32_t *out)
{
    Board b; b.N = n_territories; b.P = n_players;
    int a[MAXP]; calculateStartingArmies(&b, a);
    for (int i = 0; i < n_players; i++) out[i] = a[i];
    return n_players;
}

/* SimulationBoard.java:535-539 */
static void setupInitialPlacementPhase(Board *b)
{
    int armies = b->initialArmiesLeftToPlace[b->currentPlayer];
    b->numberOfPlaceableArmies = armies == 5 ? 5 : (armies < 4 ? armies : 4);
}

/* SimulationBoard.java:698-707 */
static void setupAttackPhase(Board *b)
{
    b->hasInvadedACountry = 0;
    b->attackingCC = -1; b->defendingCC = -1; b->attackingPlayer = -1; b->defendingPlayer = -1;
    b->attackState = ORX_ATTACK_START;
}

/* SimulationBoard.java:1076-1090 -- note: counts and continents of currentPlayer (quirk 3) */
static int getPlayerIncome(const Board *b, int player)
{
    if (b->playerCountries[player] == 0) return 0;
    int cp = b->currentPlayer;
    int income = b->playerCountries[cp] / 3; if (income < 3) income = 3;
    for (int i = 0; i < b->C; i++)
        if (b->missing[cp][i] == 0) income = (int32_t)((uint32_t)income + (uint32_t)b->currentContinentBonuses[i]);
    return income;
}

/* SimulationBoard.java:843-867 */
static void playerDefeated(Board *b, int player, int conquerer)
{
    b->remainingPlayers--;
    b->playerOutOnTurn[player] = b->simTurnCount;
    if (b->transferCards) {
        if (conquerer < 0) { FAIL(b); return; }
        CardList *dst = &b->playerCards[conquerer], *src = &b->playerCards[player];
        int n = src->n; /* addAll of a list onto itself (conquerer == player) doubles it; keep Java's behaviour */
        for (int i = 0; i < n; i++) list_add(dst, src->v[i]);
    } else {
        CardList *src = &b->playerCards[player];
        for (int i = 0; i < src->n; i++) list_add(&b->simDeck, src->v[i]); /* simReturnToDeck */
    }
    b->ct.writes += 2;
    if (b->remainingPlayers == 1) {
        b->playerOutOnTurn[b->currentPlayer] = -1;
        if (!b->termSet) { b->termSet = 1; b->termPos = b->rng.pos; }
    }
}

/* SimulationBoard.java:577-618 ; cards given by hand positions i<j<k */
static int cashCards(Board *b, const int idx[3])
{
    if (b->currentPhase != ORX_PHASE_CARDS && (b->currentPhase != ORX_PHASE_ATTACK || b->attackState != ORX_ATTACK_CASH)) return 0;
    CardList *hand = &b->playerCards[b->currentPlayer];
    uint8_t c[3] = { hand->v[idx[0]], hand->v[idx[1]], hand->v[idx[2]] };
    if (!isASet(b, c[0], c[1], c[2])) return 0;
    /* currentCards.remove(card1); remove(card2); remove(card3) (:596-598): List.remove(Object) takes out the FIRST
     * element equal to the card.  Territory cards are unique, but the two wildcards are one static object
     * (CardManager.java:14,62-63): a triple that holds the hand's second wildcard removes the first one. */
    for (int q = 0; q < 3; q++) {
        int at = -1;
        for (int i = 0; i < hand->n && at < 0; i++) if (hand->v[i] == c[q]) at = i;
        if (at < 0) { FAIL(b); return 0; }
        list_remove_at(hand, at);
    }
    /* CardManager.simCash (O/CardManager.java:265-276) */
    list_add(&b->simDeck, c[0]); list_add(&b->simDeck, c[1]); list_add(&b->simDeck, c[2]);
    /* engine limit: only the exponential ("%") progressions are table driven on the device */
    if ((b->progression.family == PROG_EXP4 || b->progression.family == PROG_EXP5) && b->simCardProgressionPos >= ORX_CARD_VALUES) { FAIL(b); return 0; }
    int armies = getCardCashValue(&b->progression, b->simCardProgressionPos);
    b->simCardProgressionPos++;
    b->numberOfPlaceableArmies += armies;
    for (int i = 0; i < 3; i++)
        if (c[i] != ORX_CARD_WILD && b->owner[c[i]] == b->currentPlayer) { b->armies[c[i]] += 2; b->ct.writes++; }
    b->ct.writes += 8;
    return 1;
}

/* SimRandom.cardsPhase (O/sim/agents/SimRandom.java:36-49) */
static void agent_cardsPhase(Board *b)
{
    CardList *cards = &b->playerCards[b->currentPlayer];
    if (cards->n < 3 || !containsASet(b, cards)) return;
    double r = jr_next_double(&b->rng);
    if ((cards->n == 3 && r < 0.5) || (cards->n == 4 && r < 0.75) || cards->n >= 5) {
        int idx[3];
        if (getRandomSet(b, cards, idx)) cashCards(b, idx);
    }
}

/* SimulationBoard.java:620-638 */
static void tearDownCardsPhase(Board *b)
{
    CardList *cards = &b->playerCards[b->currentPlayer];
    while (cards->n > 4 && !b->flags) {
        int idx[3];
        if (!getRandomSet(b, cards, idx)) break;
        cashCards(b, idx);
    }
    b->currentPhase = ORX_PHASE_PLACEMENT;
}

/* SimulationBoard.java:647-684 */
static void placeArmies(Board *b, int numberOfArmies, int cc)
{
    if (b->currentPhase != ORX_PHASE_PLACEMENT && b->currentPhase != ORX_PHASE_INITIAL_PLACEMENT &&
        (b->currentPhase != ORX_PHASE_ATTACK || b->attackState != ORX_ATTACK_PLACE)) { FAIL(b); return; }
    if (numberOfArmies > b->numberOfPlaceableArmies || b->owner[cc] != b->currentPlayer || numberOfArmies < 0) { FAIL(b); return; }
    b->armies[cc] += numberOfArmies;
    b->numberOfPlaceableArmies -= numberOfArmies;
    if (b->currentPhase == ORX_PHASE_INITIAL_PLACEMENT) b->initialArmiesLeftToPlace[b->currentPlayer] -= numberOfArmies;
    b->ct.writes += 2;
}

/* SimRandom.place (O/sim/agents/SimRandom.java:247-277) */
static void agent_place(Board *b, int numberOfAvailableArmies)
{
    int ID = b->currentPlayer;
    int cand[MAXN], n = 0;
    for (int c = 0; c < b->N; c++) {
        b->ct.tv++;
        if (b->owner[c] == ID)
            for (int e = b->adj_off[c]; e < b->adj_off[c + 1]; e++) {
                b->ct.ev++;
                if (b->owner[b->adj[e]] != ID) { cand[n++] = c; break; }
            }
    }
    while (numberOfAvailableArmies > 0 && !b->flags) {
        int armies = 1 + jr_next_int(&b->rng, numberOfAvailableArmies);
        int i = jr_next_int(&b->rng, n); /* nextInt(0) throws when there is no border */
        if (b->flags) return;
        placeArmies(b, armies, cand[i]);
        numberOfAvailableArmies -= armies;
    }
}

/* SimRandom.moveArmiesIn (O/sim/agents/SimRandom.java:144-155) */
static int agent_moveArmiesIn(Board *b, int attackerCC, int defenderCC)
{
    (void)defenderCC;
    int available = b->armies[attackerCC] - 1;
    if (available == 0) { FAIL(b); return 0; }
    int min = available < 3 ? available : 3;
    return min + (min == available ? 0 : jr_next_int(&b->rng, available - min));
}

/* SimulationBoard.java:758-780 */
static void executeAttack(Board *b)
{
    if (b->attackingCC < 0 || b->defendingCC < 0) { FAIL(b); return; }
    int aArmies = b->armies[b->attackingCC] - 1;
    int dArmies = b->armies[b->defendingCC];
    if (b->attackUntilDead) {
        simulateBattle(&b->rng, &b->ct, aArmies, dArmies, &b->attackerLosses, &b->defenderLosses);
    } else {
        int pa = aArmies < 3 ? aArmies : 3, pd = dArmies < 2 ? dArmies : 2;
        b->attackerLosses = simulateIndividualBattle(&b->rng, pa, pd);
        b->defenderLosses = (pa < pd ? pa : pd) - b->attackerLosses;
        b->ct.battles++;
    }
}

/* SimulationBoard.java:782-795 */
static void finishAttack(Board *b)
{
    b->armies[b->attackingCC] -= b->attackerLosses;
    b->armies[b->defendingCC] -= b->defenderLosses;
    b->ct.writes += 2;
    b->attackState = b->armies[b->defendingCC] == 0 ? ORX_ATTACK_INVADE : ORX_ATTACK_START;
}

/* SimulationBoard.java:797-812 */
static void setupInvasion(Board *b)
{
    int continent = b->continent[b->defendingCC];
    b->attackingPlayer = b->owner[b->attackingCC];
    b->defendingPlayer = b->owner[b->defendingCC];
    b->playerCountries[b->attackingPlayer]++;
    b->missing[b->attackingPlayer][continent]--;
    b->playerCountries[b->defendingPlayer]--;
    b->missing[b->defendingPlayer][continent]++;
    b->owner[b->defendingCC] = b->attackingPlayer;
    b->ct.writes += 5; b->ct.conquests++;
}

/* SimulationBoard.java:814-825 */
static void executeInvasion(Board *b, int armies)
{
    int a = b->armies[b->attackingCC];
    int minInvadingArmies = a - 1 < 3 ? a - 1 : 3;
    int maxInvadingArmies = a - 1;
    int inv = armies > minInvadingArmies ? armies : minInvadingArmies;
    if (inv > maxInvadingArmies) inv = maxInvadingArmies;
    b->armies[b->defendingCC] = inv;
    b->armies[b->attackingCC] = a - inv;
    b->ct.writes += 2;
}

/* SimulationBoard.java:827-841 */
static void finishInvasion(Board *b)
{
    if (b->defendingPlayer < 0 || b->defendingPlayer >= b->P) { FAIL(b); return; }
    if (b->playerCountries[b->defendingPlayer] == 0) playerDefeated(b, b->defendingPlayer, b->attackingPlayer);
    b->hasInvadedACountry = 1;
    b->attackState = ORX_ATTACK_START;
    b->attackingCC = -1; b->defendingCC = -1; b->attackingPlayer = -1; b->defendingPlayer = -1;
}

static int sim_moveArmiesIn(Board *b, int attackerCC, int defenderCC);

/* SimulationBoard.attack (SimulationBoard.java:713-744), attackState never becomes CASH
 * because the immediateCash branch is commented out (:852-855) */
static int board_attack(Board *b, int attackingCC, int defendingCC, int attackUntilDead)
{
    /* setupAttack :746-756 */
    b->attackerLosses = 0; b->defenderLosses = 0;
    b->attackingCC = attackingCC; b->defendingCC = defendingCC;
    b->attackingPlayer = b->owner[attackingCC]; b->defendingPlayer = b->owner[defendingCC];
    b->attackUntilDead = attackUntilDead;
    b->attackState = ORX_ATTACK_EXECUTE;
    executeAttack(b);
    if (b->flags) return 0;
    finishAttack(b);
    b->ct.attacks++;
    if (b->attackState == ORX_ATTACK_INVADE) {
        setupInvasion(b);
        executeInvasion(b, sim_moveArmiesIn(b, attackingCC, defendingCC));
        finishInvasion(b);
        return 7;
    }
    return b->armies[attackingCC] == 1 ? 13 : 0;
}

/* SimRandom.attackPhase (O/sim/agents/SimRandom.java:58-141) */
static void agent_attackPhase(Board *b)
{
    int ID = b->currentPlayer;
    int possibleAttackers[2 * MAXN], nAtt = 0;
    int possibleDefenders[MAXN], nDef;
    for (int c = 0; c < b->N; c++) {
        b->ct.tv++;
        if (b->owner[c] == ID && b->armies[c] > 1)
            for (int e = b->adj_off[c]; e < b->adj_off[c + 1]; e++) {
                b->ct.ev++;
                if (b->owner[b->adj[e]] != ID) { possibleAttackers[nAtt++] = c; break; }
            }
    }
    while (nAtt > 0 && !b->flags) {
        int attackerIndex = jr_next_int(&b->rng, nAtt);
        int attacker = possibleAttackers[attackerIndex];
        nDef = 0;
        for (int e = b->adj_off[attacker]; e < b->adj_off[attacker + 1]; e++) {
            b->ct.ev++;
            if (b->owner[b->adj[e]] != ID) possibleDefenders[nDef++] = b->adj[e];
        }
        if (nDef == 0) {
            memmove(&possibleAttackers[attackerIndex], &possibleAttackers[attackerIndex + 1], sizeof(int) * (size_t)(nAtt - attackerIndex - 1));
            nAtt--;
        } else {
            int defender = possibleDefenders[jr_next_int(&b->rng, nDef)];
            int result = board_attack(b, attacker, defender, 1);
            if (b->flags) return;
            if (result == 13) {
                memmove(&possibleAttackers[attackerIndex], &possibleAttackers[attackerIndex + 1], sizeof(int) * (size_t)(nAtt - attackerIndex - 1));
                nAtt--;
            } else if (result == 7) {
                if (b->armies[attacker] == 1) {
                    memmove(&possibleAttackers[attackerIndex], &possibleAttackers[attackerIndex + 1], sizeof(int) * (size_t)(nAtt - attackerIndex - 1));
                    nAtt--;
                }
                if (b->armies[defender] > 1) possibleAttackers[nAtt++] = defender;
            }
        }
    }
}

/* SimulationBoard.java:884-896 */
static void tearDownAttackPhase(Board *b)
{
    if (b->hasInvadedACountry && b->useCards) {
        uint8_t card;
        if (simDeal(b, &card)) list_add(&b->playerCards[b->currentPlayer], card);
    }
    b->currentPhase = ORX_PHASE_FORTIFICATION;
}

/* SimulationBoard.java:984-1015 */
static void checkForOverflow(Board *b)
{
    int armyTotals[MAXP] = { 0 };
    for (int c = 0; c < b->N; c++) {
        b->ct.tv++;
        if (b->owner[c] < 0) { FAIL(b); return; }
        armyTotals[b->owner[c]] += b->armies[c];
    }
    int totalArmies = 0, maxArmies = 0, maxPlayer = -1;
    for (int i = 0; i < b->P; i++) {
        totalArmies += armyTotals[i];
        if (armyTotals[i] > maxArmies) { maxPlayer = i; maxArmies = armyTotals[i]; }
    }
    if (totalArmies > 100000)
        for (int i = 0; i < b->P; i++)
            if (b->playerCountries[i] > 0 && i != maxPlayer) playerDefeated(b, i, maxPlayer);
}

/* SimulationBoard.java:965-982 (+ ModellingBoard.java:29-47: swapping in fresh SimRandom
 * agents changes nothing observable for an all-SimRandom line-up) */
static void tearDownFortificationPhase(Board *b)
{
    b->ct.tv += (uint64_t)b->N; /* moveable <- 0 sweep :967-970 */
    int nextPlayer = b->currentPlayer; /* getNextPlayer :1102-1111 */
    int guard = 0;
    do { nextPlayer = (nextPlayer + 1) % b->P; if (++guard > b->P) { FAIL(b); return; } } while (b->playerCountries[nextPlayer] == 0);
    /* ModellingBoard.tearDownFortificationPhase (ModellingBoard.java:29-47): everybody becomes SimRandom when the
     * round that is about to start is round `modellingTurnLimit` */
    if (nextPlayer < b->currentPlayer && b->simTurnCount + 1 == b->modellingTurnLimit)
        for (int p = 0; p < b->P; p++) b->simAgent[p] = ORX_AGENT_RANDOM;
    for (int c = 0; c < b->N; c++) b->moveable[c] = 0;
    if (nextPlayer < b->currentPlayer) { b->simTurnCount++; recalculateContinentBonuses(b); }
    b->currentPlayer = nextPlayer;
    b->currentPhase = ORX_PHASE_CARDS;
    b->ct.writes += 2;
    checkForOverflow(b);
    b->playerTurns++; b->ct.player_turns++;
}

/* SimulationBoard.java:485-505 */
static void selectCountry(Board *b, int cc)
{
    if (cc < 0 || b->owner[cc] != -1) { FAIL(b); return; }
    b->owner[cc] = b->currentPlayer;
    b->playerCountries[b->currentPlayer]++;
    b->initialArmiesLeftToPlace[b->currentPlayer]--;
    b->remainingCountries--;
    b->missing[b->currentPlayer][b->continent[cc]]--;
    b->currentPlayer = (b->currentPlayer + 1) % b->P;
    b->ct.writes += 5;
}

/* SimulationBoard.java:507-514 */
static void tearDownCountrySelectionPhase(Board *b)
{
    if (b->remainingCountries == 0) { b->currentPhase = ORX_PHASE_INITIAL_PLACEMENT; b->currentPlayer = 0; }
}

/* SimulationBoard.java:541-568 */
static void tearDownInitialPlacementPhase(Board *b)
{
    int nextPlayer = (b->currentPlayer + 1) % b->P;
    while (nextPlayer != b->currentPlayer && b->initialArmiesLeftToPlace[nextPlayer] == 0) nextPlayer = (nextPlayer + 1) % b->P;
    if (nextPlayer == b->currentPlayer) {
        b->currentPlayer = 0;
        b->currentPhase = ORX_PHASE_CARDS;
        b->simTurnCount++;
    } else b->currentPlayer = nextPlayer;
    b->playerTurns++;
}

/* SimRandom.pickCountry (O/sim/agents/SimRandom.java:16-27) */
static int agent_pickCountry(Board *b)
{
    int cand[MAXN], n = 0;
    for (int cc = 0; cc < b->N; cc++) { b->ct.tv++; if (b->owner[cc] == -1) cand[n++] = cc; }
    int i = jr_next_int(&b->rng, n);
    if (b->flags) return -1;
    return cand[i];
}


/* ------------------------------------------------------------------------------------ */
/* Modelled opponents: SimAngry, SimStinky behind ModellingBoard.  */
/* SDK semantics are DECLARED ASSUMPTIONS (include/orx_state.h): parity unpinned.       */
/* ------------------------------------------------------------------------------------ */

/* Country.getNumberEnemyNeighbors / getNumberPlayerNeighbors / getWeakestEnemyNeighbor (declared assumptions) */
static int numEnemyNeighbors(Board *b, int cc)
{
    int n = 0;
    for (int e = b->adj_off[cc]; e < b->adj_off[cc + 1]; e++) { b->ct.ev++; if (b->owner[b->adj[e]] != b->owner[cc]) n++; }
    return n;
}
static int numPlayerNeighbors(Board *b, int cc, int player)
{
    int n = 0;
    for (int e = b->adj_off[cc]; e < b->adj_off[cc + 1]; e++) { b->ct.ev++; if (b->owner[b->adj[e]] == player) n++; }
    return n;
}
static int weakestEnemyNeighbor(Board *b, int cc)
{
    int weakest = -1, weakestArmies = 0;
    for (int e = b->adj_off[cc]; e < b->adj_off[cc + 1]; e++) {
        int nb = b->adj[e];
        b->ct.ev++;
        if (b->owner[nb] != b->owner[cc] && (weakest < 0 || b->armies[nb] < weakestArmies)) { weakest = nb; weakestArmies = b->armies[nb]; }
    }
    return weakest;
}
/* the element a PlayerIterator (armies == 0) / ArmiesIterator holds after looking from `from` on */
static int iterFind(Board *b, int from, int player, int armies)
{
    for (int cc = from; cc < b->N; cc++) {
        b->ct.tv++;
        if (b->owner[cc] != player) continue;
        if (armies > 0 && b->armies[cc] < armies) continue;
        if (armies < 0 && b->armies[cc] > -armies) continue;
        return cc;
    }
    return -1;
}

/* SimulationBoard.fortifyArmies (SimulationBoard.java:908-954) */
static int fortifyArmies(Board *b, int numberOfArmies, int sourceCC, int destinationCC)
{
    if (b->currentPhase != ORX_PHASE_FORTIFICATION) { FAIL(b); return -1; }
    if (b->owner[sourceCC] != b->currentPlayer || b->owner[destinationCC] != b->currentPlayer) return -1;
    if (numberOfArmies < 0) return -1;
    int sourceArmies = b->armies[sourceCC], moveableArmies = b->moveable[sourceCC];
    int armiesToFortify = moveableArmies < sourceArmies - 1 ? moveableArmies : sourceArmies - 1;
    if (numberOfArmies < armiesToFortify) armiesToFortify = numberOfArmies;
    if (armiesToFortify == 0) return 0;
    b->armies[sourceCC] = sourceArmies - armiesToFortify;
    b->moveable[sourceCC] = moveableArmies - armiesToFortify;
    b->armies[destinationCC] += armiesToFortify;
    b->ct.writes += 3;
    return 1;
}

/* SimulationBoard.setupFortificationPhase (:900-906) */
static void setupFortificationPhase(Board *b)
{
    for (int c = 0; c < b->N; c++) b->moveable[c] = b->owner[c] == b->currentPlayer ? b->armies[c] : 0;
}

/* MapHelper.getSmallestPositiveEmptyContinent / getSmallestPositiveOpenContinent (MapHelper.java:132-191):
 * empty = every country of the continent unowned, open = at least one unowned */
static int smallestPositiveContinent(Board *b, int needAll)
{
    int smallestSize = 2147483647, smallestContinent = -1;
    for (int continent = 0; continent < b->C; continent++) {
        int size = 0, unowned = 0;
        for (int cc = 0; cc < b->N; cc++) { b->ct.tv++; if (b->continent[cc] == continent) { size++; if (b->owner[cc] == -1) unowned++; } }
        int ok = needAll ? unowned == size : unowned > 0;
        if (size < smallestSize && ok && b->currentContinentBonuses[continent] > 0) { smallestSize = size; smallestContinent = continent; }
    }
    return smallestContinent;
}

/* SimAngry.pickCountry + pickCountryInContinent (SimAngry.java:49-105) */
static int angry_pickCountry(Board *b)
{
    int ID = b->currentPlayer;
    int goalCont = smallestPositiveContinent(b, 1);
    if (goalCont == -1) goalCont = smallestPositiveContinent(b, 0);
    for (int cc = 0; cc < b->N; cc++) {
        b->ct.tv++;
        if (b->continent[cc] == goalCont && b->owner[cc] == -1 && numPlayerNeighbors(b, cc, ID) > 0) return cc;
    }
    int bestCode = -1, fewestNeib = 1000000;
    for (int cc = 0; cc < b->N; cc++) {
        b->ct.tv++;
        int nn = b->adj_off[cc + 1] - b->adj_off[cc];
        if (b->continent[cc] == goalCont && b->owner[cc] == -1 && nn < fewestNeib) { bestCode = cc; fewestNeib = nn; }
    }
    return bestCode; /* -1: selectCountry indexes countries[-1] and throws */
}

/* SimAngry.placeArmies (SimAngry.java:126-151): everything on the owned country with the most enemy neighbours */
static void angry_placeArmies(Board *b, int numberOfArmies)
{
    int ID = b->currentPlayer, mostEnemies = -1, placeOn = -1;
    for (int cc = iterFind(b, 0, ID, 0); cc >= 0; cc = iterFind(b, cc + 1, ID, 0)) {
        int subTotalEnemies = numEnemyNeighbors(b, cc);
        if (subTotalEnemies > mostEnemies) { mostEnemies = subTotalEnemies; placeOn = cc; }
    }
    if (placeOn < 0) { FAIL(b); return; } /* placeOn.getCode() on null */
    placeArmies(b, numberOfArmies, placeOn);
}

/* SimAngry.attackPhase (SimAngry.java:157-195) */
static void angry_attackPhase(Board *b)
{
    int ID = b->currentPlayer, madeAttack = 1;
    while (madeAttack && !b->flags) {
        madeAttack = 0;
        int held = iterFind(b, 0, ID, 2);
        while (held >= 0 && !b->flags) {
            int us = held;
            held = iterFind(b, us + 1, ID, 2);
            int weakestNeighbor = weakestEnemyNeighbor(b, us);
            if (weakestNeighbor >= 0 && b->armies[us] > b->armies[weakestNeighbor]) {
                board_attack(b, us, weakestNeighbor, 1);
                madeAttack = 1;
            }
        }
    }
}

/* SimAngry.moveArmiesIn (SimAngry.java:199-213) */
static int angry_moveArmiesIn(Board *b, int cca, int ccd)
{
    int Aenemies = numEnemyNeighbors(b, cca), Denemies = numEnemyNeighbors(b, ccd);
    if (Aenemies > Denemies) return 0;
    return b->armies[cca] - 1;
}

/* SimAngry.fortifyPhase (SimAngry.java:215-279) */
static void angry_fortifyPhase(Board *b)
{
    int ID = b->currentPlayer;
    for (int i = 0; i < b->N && !b->flags; i++) {
        b->ct.tv++;
        if (b->owner[i] != ID || b->moveable[i] <= 0) continue;
        int countryCodeBestProspect = -1, bestEnemyNeighbors = 0;
        for (int e = b->adj_off[i]; e < b->adj_off[i + 1]; e++) {
            int nb = b->adj[e];
            if (b->owner[nb] == ID) {
                int enemyNeighbors = numEnemyNeighbors(b, nb);
                if (enemyNeighbors > bestEnemyNeighbors) { countryCodeBestProspect = nb; bestEnemyNeighbors = enemyNeighbors; }
            }
        }
        int enemyNeighbors = numEnemyNeighbors(b, i);
        if (bestEnemyNeighbors > enemyNeighbors) {
            fortifyArmies(b, b->moveable[i], i, countryCodeBestProspect);
        } else if (enemyNeighbors == 0) {
            int nn = b->adj_off[i + 1] - b->adj_off[i];
            int randCC = jr_next_int(&b->rng, nn); /* nextInt(0) throws */
            if (b->flags) return;
            fortifyArmies(b, b->moveable[i], i, b->adj[b->adj_off[i] + randCC]);
        }
    }
}

/* SimStinky.placeArmies (SimStinky.java:66-78) */
static void stinky_placeArmies(Board *b, int numberOfArmies)
{
    int ID = b->currentPlayer, test, tries = 0;
    do {
        if (tries++ >= ORX_MAX_REDRAWS) { b->flags |= ORX_FLAG_STEP_LIMIT; return; } /* engine limit */
        test = jr_next_int(&b->rng, b->N);
        if (b->flags) return;
    } while (b->owner[test] != ID || weakestEnemyNeighbor(b, test) < 0);
    placeArmies(b, numberOfArmies, test);
}

/* SimStinky.attackPhase(boolean) (SimStinky.java:91-109) */
static int stinky_attackPass(Board *b, int careAboutOdds)
{
    int ID = b->currentPlayer, attacked = 0;
    int held = iterFind(b, 0, ID, 0);
    while (held >= 0 && !b->flags) {
        int us = held;
        held = iterFind(b, us + 1, ID, 0);
        int weak = weakestEnemyNeighbor(b, us);
        if (weak >= 0 && b->armies[us] > 1 && ((double)b->armies[us] > b->armies[weak] * 1.5 || !careAboutOdds)) {
            board_attack(b, us, weak, 1);
            attacked = 1;
        }
    }
    return attacked;
}

/* SimStinky.attackPhase() (SimStinky.java:80-89) */
static void stinky_attackPhase(Board *b)
{
    stinky_attackPass(b, 1);
    if (b->flags) return;
    int ID = b->currentPlayer, total = 0; /* BoardHelper.getPlayerArmies */
    for (int cc = 0; cc < b->N; cc++) { b->ct.tv++; if (b->owner[cc] == ID) total += b->armies[cc]; }
    if (total > 300)
        while (!b->flags && stinky_attackPass(b, 0)) { }
}

/* SimStinky.moveArmiesIn (SimStinky.java:111-125) */
static int stinky_moveArmiesIn(Board *b, int cca, int ccd)
{
    int attackWeak = weakestEnemyNeighbor(b, cca), defendWeak = weakestEnemyNeighbor(b, ccd);
    if (attackWeak < 0) return 1000000;
    if (defendWeak < 0) return 0;
    if (b->armies[attackWeak] < b->armies[defendWeak]) return 0;
    return 1000000;
}

/* ---- simAgents[currentPlayer].xxx() ------------------------------------------------ */
static int sim_agent(const Board *b) { return b->simAgent[b->currentPlayer]; }

static int sim_pickCountry(Board *b)
{
    switch (sim_agent(b)) {
    case ORX_AGENT_ANGRY: return angry_pickCountry(b);
    case ORX_AGENT_STINKY: return -1; /* "the world will give us a random unowned country" -- this board throws instead */
    default: return agent_pickCountry(b);
    }
}
static void sim_placeArmies(Board *b, int n) /* placeInitialArmies delegates to it in all three agents */
{
    switch (sim_agent(b)) {
    case ORX_AGENT_ANGRY: angry_placeArmies(b, n); break;
    case ORX_AGENT_STINKY: stinky_placeArmies(b, n); break;
    default: agent_place(b, n);
    }
}
static void sim_cardsPhase(Board *b)
{
    if (sim_agent(b) == ORX_AGENT_RANDOM) agent_cardsPhase(b); /* SimAngry / SimStinky: empty (tearDownCardsPhase cashes above 4) */
}
static void sim_attackPhase(Board *b)
{
    switch (sim_agent(b)) {
    case ORX_AGENT_ANGRY: angry_attackPhase(b); break;
    case ORX_AGENT_STINKY: stinky_attackPhase(b); break;
    default: agent_attackPhase(b);
    }
}
static int sim_moveArmiesIn(Board *b, int attackerCC, int defenderCC)
{
    switch (sim_agent(b)) {
    case ORX_AGENT_ANGRY: return angry_moveArmiesIn(b, attackerCC, defenderCC);
    case ORX_AGENT_STINKY: return stinky_moveArmiesIn(b, attackerCC, defenderCC);
    default: return agent_moveArmiesIn(b, attackerCC, defenderCC);
    }
}
static void sim_fortifyPhase(Board *b)
{
    if (sim_agent(b) == ORX_AGENT_ANGRY) angry_fortifyPhase(b); /* SimRandom, SimStinky: empty */
}

/* SimulationBoard.setupState (SimulationBoard.java:178-190) from a packed leaf record */
static void setupState(Board *b, const uint8_t *leaf)
{
    const OrxLeafHeader *h = (const OrxLeafHeader *)leaf;
    OrxLeafLayout L = orx_leaf_layout(b->N, b->P);
    const uint8_t *owner = leaf + L.off_owner;
    const int32_t *armies = (const int32_t *)(leaf + L.off_armies);
    const uint8_t *ncards = leaf + L.off_ncards;
    const uint8_t *cards = leaf + L.off_cards;

    /* engine-level record validation: "Leaf validity" in include/orx_state.h */
    {
        int N = b->N, P = b->P;
        int bad = h->current_player < 0 || h->current_player >= P || h->phase < 0 || h->phase > ORX_PHASE_FORTIFICATION ||
                  h->attack_state < 0 || h->attack_state > ORX_ATTACK_PLACE || h->attacking_cc < -1 || h->attacking_cc >= N ||
                  h->defending_cc < -1 || h->defending_cc >= N || h->defending_player < -1 || h->defending_player >= P ||
                  h->card_pos < 0 || (int)h->n_hand + (int)h->n_deck > N + 2 ||
                  h->placeable < -ORX_MAX_ARMIES || h->placeable > ORX_MAX_ARMIES;
        for (int t = 0; t < N && !bad; t++) {
            int o = owner[t];
            if (!(o < P || (o == ORX_OWNER_NONE && h->phase == ORX_PHASE_COUNTRY_SELECTION))) bad = 1;
            /* an INVADE leaf may carry the conquered country at 0 armies (executeInvasion overwrites it) */
            int lo = (h->phase == ORX_PHASE_ATTACK && h->attack_state == ORX_ATTACK_INVADE && t == h->defending_cc) ? 0 : 1;
            if (armies[t] < lo || armies[t] > ORX_MAX_ARMIES) bad = 1;
        }
        for (int i = 0; i < (int)h->n_hand + (int)h->n_deck && !bad; i++)
            if (!(cards[i] < N || cards[i] == ORX_CARD_WILD)) bad = 1;
        if (h->deal_mode != ORX_DEAL_LEAF_LCG && h->deal_mode != ORX_DEAL_GAME_STREAM) bad = 1;
        if (bad) { FAIL(b); return; }
    }

    b->simStartTurn = h->turn_count;
    b->currentPlayer = h->current_player;
    b->currentPhase = h->phase;

    /* setupCountries :192-203 */
    for (int cc = 0; cc < b->N; cc++) { b->owner[cc] = owner[cc] == ORX_OWNER_NONE ? -1 : owner[cc]; b->armies[cc] = armies[cc]; }
    /* moveableArmies are not in the record: setupFortificationPhase's values for a FORTIFICATION leaf, else 0 (orx_state.h) */
    for (int cc = 0; cc < b->N; cc++)
        b->moveable[cc] = (h->phase == ORX_PHASE_FORTIFICATION && b->owner[cc] == h->current_player) ? b->armies[cc] : 0;

    /* setupCardState :205-227 ; CardManager.simReset :240-248 */
    b->simDeck.n = 0;
    for (int i = 0; i < h->n_deck; i++) list_add(&b->simDeck, cards[h->n_hand + i]);
    b->simCardProgressionPos = h->card_pos;
    b->dealMode = h->deal_mode;
    lcg_set_seed(&b->dealLcg, (int64_t)h->deal_seed); /* random = new Random(simDeck.hashCode()) :245 */
    for (int p = 0; p < b->P; p++) {
        b->playerCards[p].n = 0;
        if (p == b->currentPlayer) {
            for (int i = 0; i < h->n_hand; i++) list_add(&b->playerCards[p], cards[i]);
        } else {
            for (int j = 0; j < ncards[p]; j++) {
                uint8_t card;
                if (simDeal(b, &card)) list_add(&b->playerCards[p], card);
                else { FAIL(b); return; } /* the reference would add null and fail later */
            }
        }
    }

    setupGeneralGameState(b);
    if (b->flags) return;

    /* setupPhaseGameState :278-359 */
    switch (b->currentPhase) {
    case ORX_PHASE_COUNTRY_SELECTION:
        calculateStartingArmies(b, b->initialArmiesLeftToPlace);
        b->remainingCountries = 0;
        for (int cc = 0; cc < b->N; cc++) {
            if (b->owner[cc] == -1) b->remainingCountries++;
            else b->initialArmiesLeftToPlace[b->owner[cc]]--;
        }
        break;
    case ORX_PHASE_INITIAL_PLACEMENT: {
        int playerArmies[MAXP] = { 0 };
        for (int cc = 0; cc < b->N; cc++) playerArmies[b->owner[cc]] += b->armies[cc];
        calculateStartingArmies(b, b->initialArmiesLeftToPlace);
        for (int i = 0; i < b->P; i++) b->initialArmiesLeftToPlace[i] -= playerArmies[i];
        setupInitialPlacementPhase(b);
        b->numberOfPlaceableArmies = h->placeable;
        break;
    }
    case ORX_PHASE_CARDS:
        b->numberOfPlaceableArmies = 0; /* setupCardsPhase */
        b->numberOfPlaceableArmies = h->placeable;
        break;
    case ORX_PHASE_PLACEMENT:
        /* setupPlacementPhase() adds income to a stale field, then the leaf value overwrites it */
        b->numberOfPlaceableArmies = h->placeable;
        break;
    case ORX_PHASE_ATTACK:
        setupAttackPhase(b);
        b->hasInvadedACountry = h->has_invaded;
        b->attackState = h->attack_state;
        b->attackingCC = h->attacking_cc;
        b->defendingCC = h->defending_cc;
        if (b->attackingCC != -1) {
            if (b->attackingCC < 0 || b->attackingCC >= b->N) { FAIL(b); return; }
            b->attackingPlayer = b->owner[b->attackingCC];
            b->defendingPlayer = h->defending_player;
        }
        if (h->attack_state == ORX_ATTACK_INVADE) {
            if (b->defendingPlayer < 0 || b->defendingPlayer >= b->P) { FAIL(b); return; }
            if (b->playerCountries[b->defendingPlayer] == 0) b->remainingPlayers++;
        } else if (h->attack_state == ORX_ATTACK_CASH || h->attack_state == ORX_ATTACK_PLACE) {
            b->numberOfPlaceableArmies = h->placeable;
        }
        break;
    case ORX_PHASE_FORTIFICATION:
        break;
    default:
        FAIL(b); return;
    }
    recalculateContinentBonuses(b);
}

static int step_limit_hit(Board *b)
{
    if (b->playerTurns >= (uint32_t)b->maxPlayerTurns) { b->flags |= ORX_FLAG_STEP_LIMIT; return 1; }
    if (b->stopAtSimTurn > 0 && b->simTurnCount >= b->stopAtSimTurn && b->currentPhase == ORX_PHASE_CARDS) { b->flags |= ORX_FLAG_STEP_LIMIT; return 1; }
    return 0;
}

/* SimulationBoard.simulate (SimulationBoard.java:365-481) */
static void simulate(Board *b)
{
    /* catch up to the start of the next phase :373-431 */
    switch (b->currentPhase) {
    case ORX_PHASE_COUNTRY_SELECTION:
        tearDownCountrySelectionPhase(b);
        break;
    case ORX_PHASE_INITIAL_PLACEMENT:
        if (b->numberOfPlaceableArmies > 0) sim_placeArmies(b, b->numberOfPlaceableArmies);
        tearDownInitialPlacementPhase(b);
        break;
    case ORX_PHASE_CARDS:
        sim_cardsPhase(b);
        tearDownCardsPhase(b);
        break;
    case ORX_PHASE_PLACEMENT:
        if (b->numberOfPlaceableArmies > 0) sim_placeArmies(b, b->numberOfPlaceableArmies);
        b->currentPhase = ORX_PHASE_ATTACK; /* tearDownPlacementPhase :686-694 */
        break;
    case ORX_PHASE_ATTACK:
        if (b->attackState == ORX_ATTACK_EXECUTE) { executeAttack(b); if (!b->flags) finishAttack(b); }
        if (!b->flags && b->attackState == ORX_ATTACK_INVADE) {
            /* no setupInvasion() here: SURVEY 8.1 quirk 1 */
            if (b->attackingCC < 0 || b->defendingCC < 0) { FAIL(b); break; }
            executeInvasion(b, sim_moveArmiesIn(b, b->attackingCC, b->defendingCC));
            if (!b->flags) finishInvasion(b);
        }
        if (!b->flags && b->attackState == ORX_ATTACK_CASH) { /* executeAttackCash :869-876 */
            b->numberOfPlaceableArmies = 0;
            sim_cardsPhase(b);
            b->attackState = ORX_ATTACK_PLACE;
        }
        if (!b->flags && b->attackState == ORX_ATTACK_PLACE) { /* executeAttackPlacement :878-882 */
            sim_placeArmies(b, b->numberOfPlaceableArmies);
            b->attackState = ORX_ATTACK_START;
        }
        if (!b->flags) { sim_attackPhase(b); if (!b->flags) tearDownAttackPhase(b); }
        break;
    case ORX_PHASE_FORTIFICATION:
        sim_fortifyPhase(b); /* SimRandom.fortifyPhase is a no-op (SimRandom.java:158-160) */
        if (!b->flags) tearDownFortificationPhase(b);
        break;
    default:
        FAIL(b);
    }

    /* play out :435-479 */
    while (b->remainingPlayers > 1 && !b->flags) {
        switch (b->currentPhase) {
        case ORX_PHASE_COUNTRY_SELECTION:
            selectCountry(b, sim_pickCountry(b));
            tearDownCountrySelectionPhase(b);
            break;
        case ORX_PHASE_INITIAL_PLACEMENT:
            setupInitialPlacementPhase(b);
            sim_placeArmies(b, b->numberOfPlaceableArmies);
            tearDownInitialPlacementPhase(b);
            step_limit_hit(b);
            break;
        case ORX_PHASE_CARDS:
            b->numberOfPlaceableArmies = 0; /* setupCardsPhase :572-575 */
            sim_cardsPhase(b);
            tearDownCardsPhase(b);
            break;
        case ORX_PHASE_PLACEMENT:
            b->numberOfPlaceableArmies += getPlayerIncome(b, b->currentPlayer); /* setupPlacementPhase :642-645 */
            sim_placeArmies(b, b->numberOfPlaceableArmies);
            b->currentPhase = ORX_PHASE_ATTACK;
            break;
        case ORX_PHASE_ATTACK:
            setupAttackPhase(b);
            sim_attackPhase(b);
            if (!b->flags) tearDownAttackPhase(b);
            break;
        case ORX_PHASE_FORTIFICATION:
            b->ct.tv += (uint64_t)b->N; /* setupFortificationPhase sweep :900-906 (dead stores for SimRandom) */
            setupFortificationPhase(b);
            sim_fortifyPhase(b);
            if (!b->flags) tearDownFortificationPhase(b);
            step_limit_hit(b);
            break;
        default:
            FAIL(b);
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Entry points                                                                          */
/* ------------------------------------------------------------------------------------ */

typedef struct OrcCtx {
    OrxMap map; OrxRules rules;
    int32_t *adj_off, *adj, *continent, *bonus, *symbol;
    char *progression;
    int agents[MAXP], modelling_turn_limit; /* ModellingBoard's constructor arguments */
} OrcCtx;

OrcCtx *orc_create(const OrxMap *map, const OrxRules *rules)
{
    if (map->n_territories < 1 || map->n_territories > ORX_MAX_TERRITORIES || map->n_players < 2 || map->n_players > MAXP ||
        map->n_continents < 1 || map->n_continents > MAXC) return NULL;
    OrcCtx *c = calloc(1, sizeof *c);
    int N = map->n_territories, E = map->adj_offsets[N];
    c->adj_off = malloc(sizeof(int32_t) * (size_t)(N + 1)); memcpy(c->adj_off, map->adj_offsets, sizeof(int32_t) * (size_t)(N + 1));
    c->adj = malloc(sizeof(int32_t) * (size_t)(E + 1)); memcpy(c->adj, map->adj, sizeof(int32_t) * (size_t)E);
    c->continent = malloc(sizeof(int32_t) * (size_t)N); memcpy(c->continent, map->continent, sizeof(int32_t) * (size_t)N);
    c->bonus = malloc(sizeof(int32_t) * (size_t)map->n_continents); memcpy(c->bonus, map->continent_bonus, sizeof(int32_t) * (size_t)map->n_continents);
    c->symbol = malloc(sizeof(int32_t) * (size_t)N);
    for (int i = 0; i < N; i++) c->symbol[i] = map->card_symbol ? map->card_symbol[i] : i % 3;
    c->progression = strdup(rules->card_progression ? rules->card_progression : "");
    c->map = *map; c->rules = *rules;
    c->map.adj_offsets = c->adj_off; c->map.adj = c->adj; c->map.continent = c->continent;
    c->map.continent_bonus = c->bonus; c->map.card_symbol = c->symbol;
    c->rules.card_progression = c->progression;
    orc_init(8);
    return c;
}

/* new ModellingBoard(bs, sas, cm, cid, modellingTurnLimit): see orx_set_sim_agents in include/orx.h */
int orc_set_sim_agents(OrcCtx *c, const int32_t *agent_ids, int modelling_turn_limit)
{
    if (!c) return -1;
    for (int p = 0; p < MAXP; p++) c->agents[p] = ORX_AGENT_RANDOM;
    c->modelling_turn_limit = 0;
    if (!agent_ids || modelling_turn_limit <= 0) return 0;
    for (int p = 0; p < c->map.n_players; p++) if (agent_ids[p] < 0 || agent_ids[p] >= ORX_AGENT_COUNT) return -1;
    for (int p = 0; p < c->map.n_players; p++) c->agents[p] = agent_ids[p];
    c->modelling_turn_limit = modelling_turn_limit;
    return 0;
}

void orc_destroy(OrcCtx *c)
{
    if (!c) return;
    free(c->adj_off); free(c->adj); free(c->continent); free(c->bonus); free(c->symbol); free(c->progression); free(c);
}

static void board_init_static(Board *b, const OrcCtx *c)
{
    memset(b, 0, sizeof *b);
    b->N = c->map.n_territories; b->C = c->map.n_continents; b->P = c->map.n_players;
    b->adj_off = c->adj_off; b->adj = c->adj; b->continent = c->continent; b->base_bonus = c->bonus;
    for (int i = 0; i < b->N; i++) b->card_symbol[i] = c->symbol[i];
    b->useCards = c->rules.use_cards; b->transferCards = c->rules.transfer_cards;
    b->immediateCash = c->rules.immediate_cash; b->continentIncrease = c->rules.continent_increase;
    b->progression = parse_progression(c->progression);
    b->maxPlayerTurns = c->rules.max_player_turns > 0 ? c->rules.max_player_turns : (1 << 16);
    b->canRunOutOfCards = b->P * 5 > b->N + 2; /* O/CardManager.java:65 */
    for (int p = 0; p < MAXP; p++) b->simAgentInitial[p] = c->agents[p];
    b->modellingTurnLimit = c->modelling_turn_limit;
}

static void run_one(Board *b, const OrcCtx *c, const uint8_t *leaf, uint64_t seed, uint64_t game,
                    const uint32_t *inj, uint32_t inj_len, OrxGameResult *out, Counters *ct_acc)
{
    /* dynamic part only; static part initialised once per thread */
    b->flags = 0; b->playerTurns = 0; b->termPos = 0; b->termSet = 0;
    memset(&b->ct, 0, sizeof b->ct);
    memset(&b->rng, 0, sizeof b->rng);
    b->rng.seed = seed; b->rng.game = game; b->rng.buf_block = -1; b->rng.inj = inj; b->rng.inj_len = inj_len; b->rng.flags = &b->flags;
    b->attackUntilDead = ((const OrxLeafHeader *)leaf)->attack_until_dead;
    b->numberOfPlaceableArmies = 0;
    (void)c;
    setupState(b, leaf);
    if (!b->flags) simulate(b);
    memset(out, 0, sizeof *out);
    if (b->flags & (ORX_FLAG_ERROR | ORX_FLAG_STREAM_END)) {
        /* a flagged playout reports only its first fault (include/orx_state.h): the stream ending can only be
         * recorded while no other flag is up, and whatever it provokes afterwards is not reported */
        out->flags = (b->flags & ORX_FLAG_STREAM_END) ? ORX_FLAG_STREAM_END : ORX_FLAG_ERROR;
    } else {
        for (int p = 0; p < b->P; p++) out->player_out_on_turn[p] = b->playerOutOnTurn[p];
        out->sim_turn_count = b->simTurnCount;
        out->flags = b->flags;
        out->stream_pos = b->termSet ? b->termPos : b->rng.pos;
        uint32_t hsum = 0;
        for (int t = 0; t < b->N; t++) hsum += orx_mix((uint32_t)t, (uint32_t)(b->owner[t] & 0xFF), (uint32_t)b->armies[t]);
        out->board_hash = hsum;
    }
    b->ct.r_words = b->rng.pos;
    if (ct_acc) {
        ct_acc->tv += b->ct.tv; ct_acc->ev += b->ct.ev; ct_acc->r_words += b->ct.r_words; ct_acc->battles += b->ct.battles;
        ct_acc->cdf_steps += b->ct.cdf_steps; ct_acc->writes += b->ct.writes; ct_acc->player_turns += b->ct.player_turns;
        ct_acc->attacks += b->ct.attacks; ct_acc->conquests += b->ct.conquests;
        ct_acc->single_rolls += b->ct.single_rolls; ct_acc->gauss_battles += b->ct.gauss_battles; ct_acc->table_battles += b->ct.table_battles;
        for (int k = 0; k < 8; k++) ct_acc->cls[k] += b->ct.cls[k];
    }
}

static void accumulate(OrxLeafResult *agg, const OrxGameResult *r, int P)
{
    agg->n++;
    agg->len_sum += r->sim_turn_count;
    for (int p = 0; p < P; p++)
        if (r->player_out_on_turn[p] == -1) { agg->wins[p]++; agg->win_len_sum[p] += r->sim_turn_count; }
    if (r->flags) agg->flagged++;
}

typedef struct BatchJob {
    const OrcCtx *ctx; const uint8_t *leaves; int n_leaves, ppl; uint64_t seed, first_game;
    OrxGameResult *games; OrxLeafResult *agg; Counters ct;
    int tid, nthreads; int stride;
    int64_t *next; /* shared ticket counter: games are handed out one at a time so long games balance */
} BatchJob;

static void *batch_worker(void *arg)
{
    BatchJob *j = arg;
    Board *b = malloc(sizeof(Board));
    board_init_static(b, j->ctx);
    int64_t total = (int64_t)j->n_leaves * j->ppl;
    for (;;) {
        int64_t g = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (g >= total) break;
        int leaf = (int)(g / j->ppl);
        OrxGameResult r;
        run_one(b, j->ctx, j->leaves + (size_t)leaf * (size_t)j->stride, j->seed, j->first_game + (uint64_t)g, NULL, 0, &r, &j->ct);
        if (j->games) j->games[g] = r;
        accumulate(&j->agg[leaf], &r, b->P);
    }
    free(b);
    return NULL;
}

/* setupState + simulate + getters for n_leaves x playouts_per_leaf playouts.
 * Game ids are first_game + leaf*playouts_per_leaf + k, so any sharding gives the same
 * per-game results.  agg (n_leaves) and counters (20 x u64) may be NULL; games may be NULL. */
int orc_rollout_batch(const OrcCtx *ctx, const uint8_t *leaves, int n_leaves, int playouts_per_leaf,
                      uint64_t seed, uint64_t first_game, OrxLeafResult *agg, OrxGameResult *games,
                      uint64_t *counters, int nthreads)
{
    if (!ctx || n_leaves < 0 || playouts_per_leaf < 0) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    OrxLeafLayout L = orx_leaf_layout(ctx->map.n_territories, ctx->map.n_players);
    BatchJob *jobs = calloc((size_t)nthreads, sizeof *jobs);
    pthread_t *th = calloc((size_t)nthreads, sizeof *th);
    OrxLeafResult *aggs = calloc((size_t)nthreads * (size_t)(n_leaves ? n_leaves : 1), sizeof *aggs);
    int64_t next = 0;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (BatchJob){ ctx, leaves, n_leaves, playouts_per_leaf, seed, first_game, games, aggs + (size_t)t * (size_t)n_leaves, { 0 }, t, nthreads, L.stride, &next };
        if (nthreads > 1) pthread_create(&th[t], NULL, batch_worker, &jobs[t]); else batch_worker(&jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    if (agg) {
        memset(agg, 0, sizeof(OrxLeafResult) * (size_t)n_leaves);
        for (int t = 0; t < nthreads; t++)
            for (int l = 0; l < n_leaves; l++) {
                const OrxLeafResult *s = &jobs[t].agg[l]; OrxLeafResult *d = &agg[l];
                d->n += s->n; d->len_sum += s->len_sum; d->flagged += s->flagged;
                for (int p = 0; p < MAXP; p++) { d->wins[p] += s->wins[p]; d->win_len_sum[p] += s->win_len_sum[p]; }
            }
    }
    if (counters) {
        memset(counters, 0, 20 * sizeof(uint64_t));
        for (int t = 0; t < nthreads; t++) {
            const Counters *c = &jobs[t].ct;
            counters[0] += c->tv; counters[1] += c->ev; counters[2] += c->r_words; counters[3] += c->battles; counters[4] += c->cdf_steps;
            counters[5] += c->writes; counters[6] += c->player_turns; counters[7] += c->attacks; counters[8] += c->conquests;
            counters[9] += c->single_rolls; counters[10] += c->gauss_battles; counters[11] += c->table_battles;
            for (int k = 0; k < 8; k++) counters[12 + k] += c->cls[k];
        }
    }
    free(aggs); free(th); free(jobs);
    return 0;
}

/* one playout per (leaf i, explicit word array i): replay of an injected stream */
int orc_rollout_replay(const OrcCtx *ctx, const uint8_t *leaves, int n_games, const uint32_t *words,
                       uint32_t words_per_game, OrxGameResult *games)
{
    if (!ctx) return -1;
    OrxLeafLayout L = orx_leaf_layout(ctx->map.n_territories, ctx->map.n_players);
    Board *b = malloc(sizeof(Board));
    board_init_static(b, ctx);
    for (int g = 0; g < n_games; g++)
        run_one(b, ctx, leaves + (size_t)g * (size_t)L.stride, 0, 0, words + (size_t)g * words_per_game, words_per_game, &games[g], NULL);
    free(b);
    return 0;
}

/* Run ONE playout until simTurnCount reaches stop_at_sim_turn and export the board as a new leaf
 * record (fixture generation: "3 full rounds under SimRandom -> snapshot").
 * The snapshot is taken at the start of a CARDS phase; returns the number of living players,
 * or -1 on error.  The current player's hand becomes the leaf's real hand; the other hands
 * are returned to the deck (their sizes are kept in ncards[]), which is exactly the
 * information a BoardState + CardManager pair carries. */
int orc_advance_to_leaf(const OrcCtx *ctx, const uint8_t *leaf_in, uint64_t seed, uint64_t game,
                        int stop_at_sim_turn, uint8_t *leaf_out)
{
    Board *b = malloc(sizeof(Board));
    board_init_static(b, ctx);
    b->stopAtSimTurn = stop_at_sim_turn;
    OrxGameResult r;
    run_one(b, ctx, leaf_in, seed, game, NULL, 0, &r, NULL);
    int ret = -1;
    if ((b->flags & ~ORX_FLAG_STEP_LIMIT) == 0 && b->currentPhase == ORX_PHASE_CARDS) {
        OrxLeafLayout L = orx_leaf_layout(b->N, b->P);
        memset(leaf_out, 0, (size_t)L.stride);
        OrxLeafHeader *h = (OrxLeafHeader *)leaf_out;
        h->turn_count = b->simStartTurn + b->simTurnCount;
        h->placeable = 0; h->attacking_cc = -1; h->defending_cc = -1;
        h->card_pos = (int16_t)b->simCardProgressionPos;
        h->current_player = (int8_t)b->currentPlayer; h->phase = ORX_PHASE_CARDS;
        h->attack_state = ORX_ATTACK_START; h->has_invaded = 0; h->defending_player = -1; h->attack_until_dead = 1;
        uint8_t *owner = leaf_out + L.off_owner; int32_t *armies = (int32_t *)(leaf_out + L.off_armies);
        uint8_t *ncards = leaf_out + L.off_ncards; uint8_t *cards = leaf_out + L.off_cards;
        for (int t = 0; t < b->N; t++) { owner[t] = (uint8_t)(b->owner[t] < 0 ? 0xFF : b->owner[t]); armies[t] = b->armies[t]; }
        int k = 0;
        CardList *mine = &b->playerCards[b->currentPlayer];
        for (int i = 0; i < mine->n; i++) cards[k++] = mine->v[i];
        h->n_hand = (uint8_t)mine->n;
        int nd = 0;
        for (int i = 0; i < b->simDeck.n; i++) { cards[k++] = b->simDeck.v[i]; nd++; }
        for (int p = 0; p < b->P; p++) {
            /* dead players' lists are stale copies (their cards moved on at playerDefeated) */
            if (p == b->currentPlayer || b->playerCountries[p] == 0) { ncards[p] = 0; continue; }
            ncards[p] = (uint8_t)b->playerCards[p].n;
            for (int i = 0; i < b->playerCards[p].n; i++) { cards[k++] = b->playerCards[p].v[i]; nd++; }
        }
        h->n_deck = (uint8_t)nd;
        ret = b->remainingPlayers;
    }
    free(b);
    return ret;
}
