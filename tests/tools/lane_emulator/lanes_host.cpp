// lanes_host.cpp -- CPU emulation of the lane-per-game rollout kernel.  TEST INFRASTRUCTURE ONLY.
//
// Compiles the per-lane game logic of oponn_b200/csrc/orx_lanes.cuh with a host compiler (ORX_LANES_HOST) and runs
// the kernel's section loop over emulated warps of 32 lanes, one lane after the other.  Purpose: (1) debug the
// kernel's logic against the oracle without a GPU (tests/test_lane_emulator.py, -m "not gpu"); (2) count how often
// each section runs and with how many lanes, which is the input of the cost model used to pick the scheduler
// thresholds (LaneTune) before spending GPU time.  Never built into, loaded by, or shipped with oponn_b200/.
//
// Battle tables come from the caller (the test builds them from the oracle's float tables the way
// build_tables_kernel does); the map image is the product's own builder (orx_mapimage.h).
#define ORX_LANES_HOST 1
#include "../../../oponn_b200/csrc/orx_lanes.cuh"
#include "../../../oponn_b200/csrc/orx_mapimage.h"

#include <vector>

using namespace orx;

extern "C" {

// stats[5][34]: per section (NEW, TURN, SEL, DICE, FIN): [0] executions, [1 + k] executions with k lanes active
int lanes_emulate(const OrxMap *map, const OrxRules *rules, const uint8_t *leaves, int n_leaves, int ppl,
                  unsigned long long seed, unsigned long long first_game, const uint32_t *meta, const uint64_t *rows,
                  const int32_t *tune9, int n_warps, OrxGameResult *games, unsigned long long *stats)
{
    MapImage im;
    if (build_map_image(map, rules, im)) return -1;
    if (im.dm.N > LN_MAXN) return -2;
    DevMap dm = im.dm;
    dm.blob = im.blob.data();
    dm.bonus_table = im.bonus_table.empty() ? nullptr : im.bonus_table.data();
    dm.card_table = im.card_table.empty() ? nullptr : im.card_table.data();
    dm.bt_meta = meta; dm.bt_rows = rows;
    LaneArgs ar;
    memset(&ar, 0, sizeof ar);
    ar.leaves = leaves; ar.n_leaves = n_leaves; ar.playouts_per_leaf = ppl; ar.total = (unsigned long long)n_leaves * ppl;
    ar.seed = seed; ar.first_game = first_game; ar.games = games;
    for (int r = 0; r < 10; r++) { ar.rk[2 * r] = (uint32_t)seed + r * 0x9E3779B9u; ar.rk[2 * r + 1] = (uint32_t)(seed >> 32) + r * 0xBB67AE85u; }
    LaneTune &tn = ar.tune;
    tn.thr_new = tune9[0]; tn.thr_turn = tune9[1]; tn.thr_sel = tune9[2]; tn.thr_fin = tune9[3];
    tn.age_new = tune9[4]; tn.age_turn = tune9[5]; tn.age_sel = tune9[6]; tn.age_fin = tune9[7]; tn.coop_min = tune9[8];

    const LaneLayout LL = lane_layout(dm.N, dm.P);
    unsigned long long ticket = 0;
    if (n_warps < 1) n_warps = 1;
    struct Warp { Lane L[32]; std::vector<uint32_t> mem; std::vector<uint8_t> cards; int aN = 0, aT = 0, aS = 0, aF = 0; bool force = false, done = false; };
    std::vector<Warp> warps((size_t)n_warps);
    for (auto &w : warps) {
        w.mem.assign((size_t)32 * LL.words, 0); w.cards.assign((size_t)32 * LN_LISTS * LN_CAP, 0);
        for (int l = 0; l < 32; l++) {
            Lane &L = w.L[l];
            L = Lane();
            L.dm = &dm; L.ar = &ar; L.mem = w.mem.data() + (size_t)l * LL.words; L.blob = im.blob.data();
            L.oARM = LL.ARM; L.oOWN = LL.OWN; L.oAL = LL.AL; L.oPOT = LL.POT; L.oPC = LL.PC; L.oNC = LL.NC; L.oRING = LL.RING; L.oGAUSS = LL.GAUSS;
            L.cards = w.cards.data() + (size_t)l * LN_LISTS * LN_CAP;
            L.st = LS_NEW;
        }
    }
    auto count = [](Warp &w, int st) { int n = 0; for (int l = 0; l < 32; l++) n += w.L[l].st == st; return n; };
    auto note = [&](int sec, int n) { if (stats) { stats[sec * 34]++; stats[sec * 34 + 1 + n]++; } };
    int live = n_warps;
    while (live > 0) {
        for (auto &w : warps) {   // one trip of the kernel's loop per warp, round robin (the ticket order is the only coupling)
            if (w.done) continue;
            // the counts the kernel takes at the top of a trip (one REDUX + one ballot)
            const int nN0 = count(w, LS_NEW), nT0 = count(w, LS_TURN), nS0 = count(w, LS_SEL), nD0 = count(w, LS_DICE), nF0 = count(w, LS_FIN);
            const int rare0 = count(w, LS_COOP) + count(w, LS_GAUSS);
            bool ran = nD0 > 0 || rare0 > 0;
            int n;
            if (nN0 > 0) {
                if (w.force || nN0 >= tn.thr_new || w.aN >= tn.age_new) {
                    w.aN = 0; ran = true; note(0, nN0);
                    unsigned long long base = ticket; ticket += (unsigned long long)nN0;
                    int rank = 0;
                    for (int l = 0; l < 32; l++) {
                        Lane &L = w.L[l];
                        if (L.st != LS_NEW) continue;
                        if ((L.fl & LF_HASGAME) && games) L.game_result(games[L.gidx]);
                        L.fl = 0u;
                        unsigned long long idx = base + (unsigned long long)rank++;
                        if (idx < ar.total) L.start_game(idx); else L.st = LS_EXIT;
                    }
                } else w.aN++;
            }
            if (nT0 > 0) {
                if (w.force || nT0 >= tn.thr_turn || w.aT >= tn.age_turn) { w.aT = 0; ran = true; note(1, count(w, LS_TURN)); for (int l = 0; l < 32; l++) if (w.L[l].st == LS_TURN) w.L[l].sec_turn(); }
                else w.aT++;
            }
            for (int l = 0; l < 32; l++) if (w.L[l].st >= LS_SEL && w.L[l].st <= LS_FIN) w.L[l].topup();
            if (nS0 > 0) {
                if (w.force || nS0 >= tn.thr_sel || w.aS >= tn.age_sel) { w.aS = 0; ran = true; note(2, count(w, LS_SEL)); for (int l = 0; l < 32; l++) if (w.L[l].st == LS_SEL) w.L[l].sec_sel(); }
                else w.aS++;
            }
            if (count(w, LS_GAUSS) > 0) { ran = true; for (int l = 0; l < 32; l++) if (w.L[l].st == LS_GAUSS) w.L[l].sec_gauss(); }
            if ((n = count(w, LS_DICE)) > 0) { ran = true; note(3, n); for (int l = 0; l < 32; l++) if (w.L[l].st == LS_DICE) w.L[l].sec_dice(); }
            // COOP after DICE: a lane the warp hands back to DICE has restarted its ring and waits for the next trip's top-up
            if ((n = count(w, LS_COOP)) > 0) { ran = true; if (stats) stats[5 * 34 + 1] += (unsigned long long)n; for (int l = 0; l < 32; l++) if (w.L[l].st == LS_COOP) w.L[l].sec_coop_host(); }
            if (nF0 > 0) {
                if (w.force || nF0 >= tn.thr_fin || w.aF >= tn.age_fin) {
                    w.aF = 0; ran = true; note(4, count(w, LS_FIN));
                    for (int l = 0; l < 32; l++) if (w.L[l].st == LS_FIN) w.L[l].sec_fin();
                    for (int l = 0; l < 32; l++) if (w.L[l].st == LS_ELIM) w.L[l].sec_elim();
                } else w.aF++;
            }
            if (count(w, LS_EXIT) == 32) { w.done = true; live--; }
            w.force = !ran;
            if (stats) stats[5 * 34]++;   // loop trips
        }
    }
    return 0;
}

}  // extern "C"
