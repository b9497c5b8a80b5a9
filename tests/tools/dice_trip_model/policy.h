#ifndef ENUM
#define ENUM 8
#endif
#ifndef SPEC_MIN
#define SPEC_MIN 1
#endif
#include <stdio.h>
#include <string.h>
#define NPOL 16
static unsigned long long P_specfail[NPOL];
static unsigned long long P_instr[NPOL], P_refill[NPOL], P_block[NPOL], P_scan[NPOL], P_scalar[NPOL], P_rolls, P_runs;
static __thread unsigned char XS[1<<20]; static __thread int XN; static __thread int A0, D0; static __thread unsigned POS0;
static int cont(int ra, int rd) { return (ra > 100 || rd > 100) && ra > 0 && rd > 0; }
/* policy p: bit0 = sum criterion in block test; bit1 = 64-wide scan; bit2 = register philox (fresh 64 per trip, no refill, philox cost in trip) */
static unsigned long long H_np[3], H_played[3], H_n[3], H_stop[3];
static void run_policy(int p)
{
    int ra = A0, rd = D0, i = 0; unsigned wi = POS0 & 127u;
    const int C_SPEC = 62, C_REFILL = 75, C_BLOCK = 70, C_SCAN = 90, C_SCAN64 = 110, C_SCALAR = 36, C_PHILOX = 50;   /* executed SASS instructions per trip, ncu SASS page of the shipped kernel (profiles/r2_sass_hot.txt capture) */
    while (cont(ra, rd)) {
        int three = (p & 4) ? (ra >= 3 && rd >= 2) : (ra > 2 && rd > 2);
        if (three) {
            int avail;
            if (p & 4) { avail = 64; P_instr[p] += C_PHILOX; }
            else { if (wi == 128) { P_refill[p]++; P_instr[p] += C_REFILL; wi = 0; } avail = (int)(128 - wi) >> 1; }
            if (avail > 0) {
                int hi = ra > rd ? ra : rd, slack = 2 * (avail - 1);
                if (p & 8) {
                    /* heuristic: expected losses 0.92 (attacker) / 1.08 (defender) per roll plus a margin; verification is exact */
                    int hi2 = ra > rd ? ra : rd; int e = (avail * ENUM) >> 3;
                    int likely = avail >= SPEC_MIN && ra - e >= 3 && rd - e >= 2 && (hi2 - e > 100 || ra + rd - 2 * avail > 200);
                    if (likely) {
                        int ra2 = ra, rd2 = rd;
                        for (int k = 0; k < avail; k++) { int x = XS[i + k]; ra2 -= x; rd2 -= 2 - x; }
                        P_instr[p] += C_SPEC;
                        if ((ra2 > 100 || rd2 > 100) && ra2 >= 3 && rd2 >= 2) { ra = ra2; rd = rd2; i += avail; wi += 2 * avail; P_block[p]++; continue; }
                        P_specfail[p]++;
                    }
                    { int np = avail > 32 ? 32 : avail, played = 0;
                      for (int k = 0; k < np; k++) { int x = XS[i++]; ra -= x; rd -= 2 - x; played++; if (!((ra > 100 || rd > 100) && ra >= 3 && rd >= 2)) break; }
                      P_scan[p]++; P_instr[p] += C_SCAN; wi += 2 * played; continue; }
                }
                int safe = avail >= 32 && ra - slack >= 3 && rd - slack >= 2 && (hi - slack > 100 || ((p & 1) && ra + rd - slack > 200));
                if (safe) {
                    for (int k = 0; k < avail; k++) { int x = XS[i++]; ra -= x; rd -= 2 - x; }
                    P_block[p]++; P_instr[p] += C_BLOCK; wi += 2 * avail; continue;
                }
                int np = (p & 2) ? avail : (avail > 32 ? 32 : avail), played = 0;
                for (int k = 0; k < np; k++) { int x = XS[i++]; ra -= x; rd -= 2 - x; played++; if (!((ra > 100 || rd > 100) && ra >= 3 && rd >= 2)) break; }
                if (p == 0) { int c = avail < 32 ? 0 : (avail < 64 ? 1 : 2); H_n[c]++; H_np[c] += np; H_played[c] += played; H_stop[c] += (played < np || !((ra > 100 || rd > 100) && ra >= 3 && rd >= 2)); } P_scan[p]++; P_instr[p] += (p & 2) ? C_SCAN64 : C_SCAN; wi += 2 * played; continue;
            }
        }
        { int pa = ra < 3 ? ra : 3, pd = rd < 2 ? rd : 2, n = pa < pd ? pa : pd; int x = XS[i++]; ra -= x; rd -= n - x;
          if (wi > 126) { P_refill[p]++; P_instr[p] += C_REFILL; wi = (wi + 2) & 127u; } else wi += 2;
          P_scalar[p]++; P_instr[p] += C_SCALAR; }
    }
    if (i != XN) { fprintf(stderr, "policy %d mismatch %d %d (a %d d %d)\n", p, i, XN, A0, D0); }
}
static void analyze(void) { if (!XN) return; __sync_fetch_and_add(&P_rolls, XN); __sync_fetch_and_add(&P_runs, 1); for (int p = 0; p < NPOL; p++) run_policy(p); }
void orc_policy_report(int games) {
 for (int c = 0; c < 3; c++) printf("scan trips class %d (avail <32 / 32..63 / 64): %.0f per game, mean np %.1f, mean played %.1f, ended the 3v2 regime %.0f per game\n", c, (double)H_n[c]/games, (double)H_np[c]/(H_n[c]+1), (double)H_played[c]/(H_n[c]+1), (double)H_stop[c]/games);
    printf("rolls/game %.0f runs/game %.0f\n", (double)P_rolls / games, (double)P_runs / games);
    for (int p = 0; p < NPOL; p++) if (p < 8 || p == 8 || p == 12) printf("policy %d spec %d (sum %d scan64 %d regs %d) specfail %.0f: instr/game %.0f = %.2f/roll; refills %.0f blocks %.0f scans %.0f scalars %.0f\n", p, (p >> 3) & 1, p & 1, (p >> 1) & 1, (p >> 2) & 1, (double)P_specfail[p] / games,
        (double)P_instr[p] / games, (double)P_instr[p] / P_rolls, (double)P_refill[p] / games, (double)P_block[p] / games, (double)P_scan[p] / games, (double)P_scalar[p] / games);
}
