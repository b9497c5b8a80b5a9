#define ORX_LANES_HOST 1
#include "../../../oponn_b200/csrc/orx_lanes.cuh"
#include <stdio.h>
#include <stdlib.h>
using namespace orx;
extern "C" double orc_fdlibm_log(double x) { return 0; }
static uint64_t T0, T1; static uint32_t rk[20];
static int popc(uint32_t x){return __builtin_popcount(x);} static int ffs_(uint32_t x){return __builtin_ffs(x);}
// transliteration of lane_coop_dice with explicit lanes
static void coop(int &ra, int &rd, uint32_t &pos, uint32_t g0, uint32_t g1)
{
    for (;;) {
        const uint32_t e = pos & 3u, c0 = pos >> 2;
        uint32_t o[32][4];
        for (int l = 0; l < 32; l++) ln_philox(c0 + l, g0, g1, rk, o[l]);
        int xA[32], xB[32], rA[32]; bool vA[32], vB[32];
        const int nb = 64 - (int)(e & 1u) - (int)(e >> 1);
        for (int l = 0; l < 32; l++) {
            uint32_t a0,a1,b0,b1; vB[l] = true;
            if (e & 1u) { uint32_t n0 = l < 31 ? o[l+1][0] : o[l][0]; a0=o[l][1]; a1=o[l][2]; b0=o[l][3]; b1=n0; vB[l] = l < 31; }
            else { a0=o[l][0]; a1=o[l][1]; b0=o[l][2]; b1=o[l][3]; }
            rA[l] = 2*l - (int)(e>>1); vA[l] = rA[l] >= 0;
            uint64_t kA = ((uint64_t)(a0>>6)<<27) + (a1>>5), kB = ((uint64_t)(b0>>6)<<27) + (b1>>5);
            xA[l] = vA[l] ? (kA>T0)+(kA>T1) : 0; xB[l] = vB[l] ? (kB>T0)+(kB>T1) : 0;
        }
        const int hi = ra > rd ? ra : rd;
        if (ra - 2*nb >= 3 && rd - 2*nb >= 2 && (hi - 2*nb > 100 || ra + rd - 2*nb > 200)) {
            int la = 0; for (int l = 0; l < 32; l++) la += xA[l]+xB[l];
            ra -= la; rd -= 2*nb - la; pos += 2u*nb; continue;
        }
        uint32_t m0=0,m1=0,m2=0;
        for (int l = 0; l < 32; l++) { int sx = xA[l]+xB[l]; if (sx&1) m0|=1u<<l; if (sx&2) m1|=1u<<l; if (sx&4) m2|=1u<<l; }
        uint32_t sA=0,sB=0; int cA[32], cB[32];
        for (int l = 0; l < 32; l++) {
            uint32_t lt = (1u<<l)-1;
            int pre = popc(m0&lt)+2*popc(m1&lt)+4*popc(m2&lt);
            cA[l] = pre + xA[l]; cB[l] = cA[l] + xB[l];
            int raA = ra - cA[l], rdA = rd - (2*(rA[l]+1) - cA[l]), raB = ra - cB[l], rdB = rd - (2*(rA[l]+2) - cB[l]);
            bool goA = (raA>100||rdA>100) && raA>=3 && rdA>=2, goB = (raB>100||rdB>100) && raB>=3 && rdB>=2;
            if (vA[l] && !goA) sA |= 1u<<l; if (vB[l] && !goB) sB |= 1u<<l;
        }
        if ((sA|sB)==0) { int la = popc(m0)+2*popc(m1)+4*popc(m2); ra -= la; rd -= 2*nb-la; pos += 2u*nb; continue; }
        int fa = sA ? ffs_(sA)-1 : 32, fb = sB ? ffs_(sB)-1 : 32; int la, played;
        if (fa <= fb) { la = cA[fa]; played = 2*fa - (int)(e>>1) + 1; } else { la = cB[fb]; played = 2*fb - (int)(e>>1) + 2; }
        ra -= la; rd -= 2*played - la; pos += 2u*played; return;
    }
}
int main()
{

    const float a3_d2_al0 = 0.3717f, a3_d2_al1 = 0.3358f;
    T0 = (uint64_t)((double)a3_d2_al0 * 9007199254740992.0); T1 = (uint64_t)((double)(float)(a3_d2_al0 + a3_d2_al1) * 9007199254740992.0);
    uint64_t seed = 20261019; for (int r = 0; r < 10; r++) { rk[2*r] = (uint32_t)seed + r*0x9E3779B9u; rk[2*r+1] = (uint32_t)(seed>>32) + r*0xBB67AE85u; }
    srand(1); long bad = 0, n = 0;
    for (int it = 0; it < 200000; it++) {
        int ra = 3 + rand() % (it % 3 == 0 ? 3000 : 300), rd = 2 + rand() % (it % 5 == 0 ? 3000 : 250);
        if (!(ra > 100 || rd > 100)) continue;
        uint32_t pos = rand() % 100000, g0 = rand(), g1 = rand() % 3;
        int a = ra, d = rd; uint32_t p = pos;
        // sequential reference
        while ((a > 100 || d > 100) && a >= 3 && d >= 2) {
            uint32_t w[2]; for (int q = 0; q < 2; q++) { uint32_t o[4]; ln_philox((p+q)>>2, g0, g1, rk, o); w[q] = o[(p+q)&3]; }
            uint64_t k = ((uint64_t)(w[0]>>6)<<27) + (w[1]>>5); int x = (k>T0)+(k>T1); a -= x; d -= 2-x; p += 2;
        }
        int ca = ra, cd = rd; uint32_t cp = pos; coop(ca, cd, cp, g0, g1);
        n++; if (ca != a || cd != d || cp != p) { if (bad < 5) printf("MISMATCH ra %d rd %d pos %u: seq (%d,%d,%u) coop (%d,%d,%u)\n", ra, rd, pos, a, d, p, ca, cd, cp); bad++; }
    }
    printf("%ld cases, %ld bad\n", n, bad);
}
