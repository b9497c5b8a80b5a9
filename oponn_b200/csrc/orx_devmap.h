// orx_devmap.h -- the static description of a map + rule set that every rollout kernel takes by value (constant bank).
// Plain C++: no CUDA types, so the lane-per-game kernel's logic (orx_lanes.cuh) can also be compiled by a host compiler
// for the CPU emulation harness under tests/tools/lane_emulator (test infrastructure, never shipped).
#pragma once

#include <stdint.h>

#include "../../include/orx_state.h"

#define ORX_FULL 0xFFFFFFFFu
#define ORX_TABLE_LIMIT 100
#define ORX_MAGIC_N 4096

namespace orx {

// Card progression families (CardManager.getCardCashValue, O/CardManager.java:283-341)
enum { PROG_UNKNOWN = 0, PROG_0, PROG_555, PROG_4466688, PROG_456, PROG_468, PROG_369, PROG_46810, PROG_51015,
       PROG_4681015202510, PROG_TABLE };

// Everything static a kernel needs, passed by value (constant bank).  `blob` is a byte image of
// the map staged into shared memory once per block; the *_off fields are byte offsets into it.
struct DevMap {
    int32_t N, C, P, NW;        // territories, continents, players, 32-lane words per board scan
    int32_t NP;                 // N rounded up to 32*NW
    int32_t degp;               // bytes per adjacency row (max out-degree rounded up to 4)
    int32_t cap;                // capacity of a card list = N+2 rounded up to 32
    int32_t use_cards, transfer_cards, cont_inc, prog_family, max_turns, can_run_out;
    int32_t blob_bytes;         // multiple of 16
    int32_t adj_off;            // u8  [N][degp]   Country.getAdjoiningList() order
    int32_t deg_off;            // u8  [NP]
    int32_t sym_off;            // u8  [NP]        1 << card symbol of the territory card
    int32_t nbr_off;            // u32 [NW][NP]    neighbour bitmask of territory t, word-major
    int32_t cont_off;           // u32 [C][NW]     member bitmask of continent c
    int32_t bonus_off;          // i32 [C]         initialContinentBonuses
    // per-warp dynamic slice (bytes from the slice base)
    int32_t w_rng, w_armies, w_owner, w_alist, w_cards, w_bytes;
    int32_t leaf_stride, off_owner, off_armies, off_ncards, off_cards;
    int32_t start_armies;       // calculateStartingArmies base (SimulationBoard.java:518-533)
    const uint8_t  *blob;       // global copy of the static map image
    const int32_t  *bonus_table;  // [ORX_BONUS_ROUNDS][C] when cont_inc != 0, else null
    const int32_t  *card_table;   // [ORX_CARD_VALUES] for the exponential progressions, else null
    const uint32_t *bt_meta;    // [100*100] (row offset << 8) | row count   for (a,d)
    const uint64_t *bt_rows;    // (min(floor(cdf * 2^53), 2^53-1) << 11) | invading << 7 | survivors
    uint64_t single_thr[6][2];  // floor(cdf*2^53) of BattleSim.outcomes rows (O/BattleSim.java:36-44)
    // opponent modelling (MODEL instantiations; kept at the end so the all-SimRandom kernels see the layout they were tuned on)
    int32_t w_move;             // per-warp slice: moveable armies (i32 [NP]); only laid out while modelling is on
    int32_t model_limit;        // ModellingBoard.modellingTurnLimit, 0 = modelling off
    uint32_t agents;            // ORX_AGENT_* of player p in bits 3p..3p+2
    int32_t magic_off;          // u32 [256]       floor(2^32 / b) + 1: small-bound nextInt of the lane-per-game kernel
};

}  // namespace orx
