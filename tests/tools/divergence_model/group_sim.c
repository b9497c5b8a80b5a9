#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
enum { S_NEW, S_PHASE, S_SELECT, S_ROLL, S_TABLE, S_GAUSS, NS };
static double cost[NS] = { 2000, 640, 150, 140, 100, 400 };
static const char *nm[NS] = {"NEW","PHASE","SELECT","ROLL","TABLE","GAUSS"};
int main(int argc, char **argv) {
    int G = atoi(argv[1]); int rolls_per_step = 2*(32/G); int policy = atoi(argv[2]); int ngames_max = atoi(argv[3]);
    if (argc > 5) cost[S_ROLL] = atof(argv[5]);
    int thrP = argc > 6 ? atoi(argv[6]) : 1;
    FILE *f = fopen(argv[4], "rb"); if (!f) { perror(argv[4]); return 1; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *t = malloc(sz); fread(t, 1, sz, f); fclose(f);
    /* per-game step arrays */
    uint8_t *steps = malloc(sz); long ns = 0;
    long *gstart = malloc(sizeof(long) * 100000); int ng = 0;
    for (long i = 0; i < sz;) {
        char c = t[i];
        if (c == 'N') { if (ng >= ngames_max) break; gstart[ng++] = ns; steps[ns++] = S_NEW; i++; }
        else if (c == 'P') { steps[ns++] = S_PHASE; i++; }
        else if (c == 'S') { steps[ns++] = S_SELECT; i++; }
        else if (c == 'T') { steps[ns++] = S_TABLE; i++; }
        else if (c == 'G') { steps[ns++] = S_GAUSS; i++; }
        else if (c == 'R') { long j = i; while (j < sz && t[j] == 'R') j++; long n = j - i; long k = (n + rolls_per_step - 1) / rolls_per_step; for (long q = 0; q < k; q++) steps[ns++] = S_ROLL; i = j; }
        else i++;
    }
    gstart[ng] = ns;
    fprintf(stderr, "games %d steps %ld (%.0f per game)\n", ng, ns, (double)ns / ng);
    /* warp simulation */
    long pos[32], end[32]; for(int l=0;l<32;l++){pos[l]=end[l]=0;} int next = 0; int alive = 0;
    for (int l = 0; l < G; l++) { if (next < ng) { pos[l] = gstart[next]; end[l] = gstart[next + 1]; next++; alive++; } else pos[l] = end[l] = 0; }
    double total = 0, useful = 0; long iters = 0; double bystate[NS] = {0}; long lanes_served[NS] = {0}, execs[NS] = {0};
    int age[NS] = {0};
    while (alive > 0) {
        int cnt[NS] = {0};
        for (int l = 0; l < 32; l++) if (pos[l] < end[l]) cnt[steps[pos[l]]]++;
        int run[NS] = {0};
        if (policy == 0) { for (int s = 0; s < NS; s++) run[s] = cnt[s] > 0; }
        else if (policy == 1) { int b = 0; for (int s = 1; s < NS; s++) if (cnt[s] > cnt[b]) b = s; run[b] = 1; }
        else if (policy == 2) { /* rare states async with threshold, else max among frequent */
            int any = 0;
            if (cnt[S_NEW] >= 1 && (cnt[S_NEW] >= thrP || age[S_NEW] > 50)) { run[S_NEW] = 1; any = 1; }
            if (cnt[S_PHASE] >= 1 && (cnt[S_PHASE] >= thrP || age[S_PHASE] > 50)) { run[S_PHASE] = 1; any = 1; }
            if (cnt[S_GAUSS] >= 1 && (cnt[S_GAUSS] >= thrP || age[S_GAUSS] > 50)) { run[S_GAUSS] = 1; any = 1; }
            int b = S_SELECT; if (cnt[S_ROLL] > cnt[b]) b = S_ROLL; if (cnt[S_TABLE] > cnt[b]) b = S_TABLE;
            if (cnt[b] > 0) run[b] = 1; else if (!any) { for (int s = 0; s < NS; s++) if (cnt[s]) { run[s] = 1; break; } }
        }
        else if (policy == 3) { /* max of count/cost efficiency-weighted: serve state maximizing cnt*? */
            int b = -1; double bv = -1; for (int s = 0; s < NS; s++) if (cnt[s]) { double v = cnt[s] + age[s] * 0.25; if (v > bv) { bv = v; b = s; } } run[b] = 1;
        }
        for (int s = 0; s < NS; s++) { if (cnt[s] && !run[s]) age[s]++; else age[s] = 0; }
        for (int s = 0; s < NS; s++) if (run[s]) { total += cost[s] + 12; bystate[s] += cost[s]; useful += cost[s] * cnt[s] / (double)G; lanes_served[s] += cnt[s]; execs[s]++; }
        for (int l = 0; l < 32; l++) if (pos[l] < end[l] && run[steps[pos[l]]]) {
            pos[l]++;
            if (pos[l] == end[l]) { if (next < ng) { pos[l] = gstart[next]; end[l] = gstart[next + 1]; next++; } else alive--; }
        }
        iters++;
    }
    printf("policy %d G %d: warp instr per playout %.3fM  lane efficiency %.1f%%  iters/game %.0f\n", policy, G, total / ng / 1e6, 100 * useful / total, (double)iters / ng);
    for (int s = 0; s < NS; s++) printf("  %-7s execs/game %.0f avg lanes %.1f cost share %.1f%%\n", nm[s], (double)execs[s] / ng, execs[s] ? (double)lanes_served[s] / execs[s] : 0, 100 * bystate[s] / total);
    return 0;
}
