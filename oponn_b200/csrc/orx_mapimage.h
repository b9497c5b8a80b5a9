// orx_mapimage.h -- host side: validates an OrxMap / OrxRules pair and lays out everything static a rollout kernel
// needs: the DevMap scalars, the byte image of the map that kernels stage into shared memory, and the two small
// host-computed tables.  Plain C++ (no CUDA) so that orx_api.cu and the CPU emulation harness of the lane-per-game
// kernel (tests/tools/lane_emulator, test infrastructure) build the very same image.
//
// Reference: the SimulationBoard constructor's inputs (O/sim/boards/SimulationBoard.java:102-139),
// CardManager.getCardCashValue (O/CardManager.java:283-341), recalculateContinentBonuses (:956-963),
// BattleSim's single-roll rows (O/BattleSim.java:36-44).
#pragma once

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "orx_devmap.h"

namespace orx {

struct MapImage {
    DevMap dm;                          // device pointers left null
    int nwt = 2;
    std::vector<uint8_t> blob;
    std::vector<int32_t> bonus_table;   // [ORX_BONUS_ROUNDS][C] when continent_increase != 0
    std::vector<int32_t> card_table;    // [ORX_CARD_VALUES] for the exponential progressions
};

// CardManager.getCardCashValue families (O/CardManager.java:283-341)
static inline int parse_progression(const char *cp, double *exp_out, int *five)
{
    *exp_out = 1.0; *five = 0;
    if (!cp) return PROG_UNKNOWN;
    if (strchr(cp, '%')) {
        const char *hat = strchr(cp, '^');
        int pct = 0;
        if (hat) {  // Integer.valueOf(cp.substring(cp.indexOf('^')+1, cp.length()-1))
            std::string t(hat + 1);
            if (!t.empty()) t.pop_back();
            pct = atoi(t.c_str());
        }
        *five = cp[0] == '5';
        *exp_out = 1 + (pct / 100.0);
        return PROG_TABLE;
    }
    if (!strcmp(cp, "0")) return PROG_0;
    if (!strcmp(cp, "5, 5, 5...")) return PROG_555;
    if (!strcmp(cp, "4, 4, 6, 6, 6, 8, 8, 8, 8, 10...")) return PROG_4466688;
    if (!strcmp(cp, "4, 5, 6...")) return PROG_456;
    if (!strcmp(cp, "4, 6, 8...")) return PROG_468;
    if (!strcmp(cp, "3, 6, 9...")) return PROG_369;
    if (!strcmp(cp, "4, 6, 8, 10, 15, 20...")) return PROG_46810;
    if (!strcmp(cp, "5, 10, 15...")) return PROG_51015;
    if (!strcmp(cp, "4, 6, 8, 10, 15, 20, 25, 10, 10, 10...")) return PROG_4681015202510;
    return PROG_UNKNOWN;
}

static inline int32_t java_d2i_host(double d)
{
    if (d != d) return 0;
    if (d >= 2147483647.0) return 2147483647;
    if (d <= -2147483648.0) return (int32_t)(-2147483647 - 1);
    return (int32_t)d;
}

// returns nullptr on success, else the reason the map / rules are rejected
static inline const char *build_map_image(const OrxMap *map, const OrxRules *rules, MapImage &im)
{
    const int N = map->n_territories, C = map->n_continents, P = map->n_players;
    if (N < 1 || N > ORX_MAX_TERRITORIES || P < 2 || P > ORX_MAX_PLAYERS || C < 1 || C > ORX_MAX_CONTINENTS)
        return "map dimensions out of range";
    if (!map->adj_offsets || !map->adj || !map->continent || !map->continent_bonus) return "map arrays missing";
    int maxdeg = 0;
    if (map->adj_offsets[0] != 0) return "adj_offsets[0] != 0";
    for (int t = 0; t < N; t++) {
        int dg = map->adj_offsets[t + 1] - map->adj_offsets[t];
        if (dg < 0 || dg > ORX_MAX_DEGREE) return "territory degree out of range";
        if (dg > maxdeg) maxdeg = dg;
        for (int e = map->adj_offsets[t]; e < map->adj_offsets[t + 1]; e++)
            if (map->adj[e] < 0 || map->adj[e] >= N) return "neighbour code out of range";
        if (map->continent[t] < 0 || map->continent[t] >= C) return "continent id out of range";
        if (map->card_symbol && (map->card_symbol[t] < 0 || map->card_symbol[t] > 2)) return "card symbol out of range";
    }

    DevMap &dm = im.dm;
    memset(&dm, 0, sizeof dm);
    im.nwt = N <= 64 ? 2 : 8;
    dm.N = N; dm.C = C; dm.P = P; dm.NW = im.nwt; dm.NP = 32 * im.nwt;
    dm.degp = ((maxdeg > 0 ? maxdeg : 1) + 3) & ~3;
    dm.cap = (N + 2 + 31) & ~31;
    dm.use_cards = rules->use_cards != 0; dm.transfer_cards = rules->transfer_cards != 0;
    dm.cont_inc = rules->continent_increase;
    dm.max_turns = rules->max_player_turns > 0 ? rules->max_player_turns : (1 << 16);
    dm.can_run_out = P * 5 > N + 2;  // O/CardManager.java:65
    double pexp; int five;
    dm.prog_family = parse_progression(rules->card_progression, &pexp, &five);
    {   // calculateStartingArmies (SimulationBoard.java:518-533)
        double f1 = N / 42.0; f1 -= 1.0; f1 /= 2.0; f1 += 1.0;
        dm.start_armies = (int)floor(f1 * (50 - P * 5) + 0.5);
    }
    OrxLeafLayout L = orx_leaf_layout(N, P);
    dm.leaf_stride = L.stride; dm.off_owner = L.off_owner; dm.off_armies = L.off_armies;
    dm.off_ncards = L.off_ncards; dm.off_cards = L.off_cards;

    // static map image
    int off = 0;
    dm.adj_off = off;  off += (N * dm.degp + 3) & ~3;
    dm.deg_off = off;  off += dm.NP;
    dm.sym_off = off;  off += dm.NP;
    dm.nbr_off = off;  off += 4 * dm.NW * dm.NP;
    dm.cont_off = off; off += 4 * C * dm.NW;
    dm.bonus_off = off; off += 4 * C;
    dm.magic_off = off; off += 4 * 256;
    dm.blob_bytes = (off + 15) & ~15;
    im.blob.assign((size_t)dm.blob_bytes, 0);
    std::vector<uint8_t> &blob = im.blob;
    uint32_t *nbr = (uint32_t *)(blob.data() + dm.nbr_off), *cont = (uint32_t *)(blob.data() + dm.cont_off);
    int32_t *bonus = (int32_t *)(blob.data() + dm.bonus_off);
    uint32_t *magic = (uint32_t *)(blob.data() + dm.magic_off);
    memset(blob.data() + dm.adj_off, 0xFF, (size_t)((N * dm.degp + 3) & ~3));   // unused adjacency slots: no territory
    for (int t = 0; t < N; t++) {
        int dg = map->adj_offsets[t + 1] - map->adj_offsets[t];
        blob[dm.deg_off + t] = (uint8_t)dg;
        for (int j = 0; j < dg; j++) {
            int x = map->adj[map->adj_offsets[t] + j];
            blob[dm.adj_off + t * dm.degp + j] = (uint8_t)x;
            nbr[(x >> 5) * dm.NP + t] |= 1u << (x & 31);
        }
        int sym = map->card_symbol ? map->card_symbol[t] : t % 3;  // O/CardManager.java:54-61
        blob[dm.sym_off + t] = (uint8_t)(1u << sym);
        cont[map->continent[t] * dm.NW + (t >> 5)] |= 1u << (t & 31);
    }
    for (int i = 0; i < C; i++) bonus[i] = map->continent_bonus[i];
    for (uint32_t b = 2; b < 256; b++) magic[b] = (uint32_t)((1ull << 32) / b) + 1u;

    // single-roll thresholds: floor(cdf * 2^53) of O/BattleSim.java:36-44 (float sums widened to double)
    {
        const float a1_d1_al0 = 0.4167f, a1_d2_al0 = 0.2546f, a2_d1_al0 = 0.5787f, a2_d2_al0 = 0.2276f, a2_d2_al1 = 0.3241f,
                    a3_d1_al0 = 0.6597f, a3_d2_al0 = 0.3717f, a3_d2_al1 = 0.3358f;
        const double rows[6][2] = { { a1_d1_al0, 1.0 }, { a1_d2_al0, 1.0 }, { a2_d1_al0, 1.0 },
                                    { a2_d2_al0, (float)(a2_d2_al0 + a2_d2_al1) }, { a3_d1_al0, 1.0 },
                                    { a3_d2_al0, (float)(a3_d2_al0 + a3_d2_al1) } };
        for (int r = 0; r < 6; r++)
            for (int k = 0; k < 2; k++) dm.single_thr[r][k] = (uint64_t)(rows[r][k] * 9007199254740992.0);
    }
    if (dm.cont_inc != 0) {  // recalculateContinentBonuses (SimulationBoard.java:956-963), one row per exponent
        im.bonus_table.resize((size_t)ORX_BONUS_ROUNDS * C);
        for (int e = 0; e < ORX_BONUS_ROUNDS; e++) {
            double factor = pow(1 + dm.cont_inc / 100.0, (double)e);
            for (int i = 0; i < C; i++) im.bonus_table[(size_t)e * C + i] = java_d2i_host(floor(map->continent_bonus[i] * factor));
        }
    }
    if (dm.prog_family == PROG_TABLE) {  // exponential progressions (O/CardManager.java:287-305)
        im.card_table.resize(ORX_CARD_VALUES);
        for (int pos = 0; pos < ORX_CARD_VALUES; pos++) {
            int base, offset;
            if (five) { base = 5 * pos; offset = 6; } else { base = pos <= 4 ? 2 * (pos + 1) : 5 * (pos - 2); offset = 8; }
            im.card_table[pos] = java_d2i_host(floor(base <= 25 ? (double)base : base * pow(pexp, (double)(pos - offset))));
        }
    }
    return nullptr;
}

}  // namespace orx
