/*
 * orx_demo.c -- the rollout seam from plain C: no Python, no torch, just include/orx.h and liborx.so.
 *
 *   gcc -O2 -I include examples/orx_demo.c -o orx_demo -L oponn_b200 -lorx -Wl,-rpath,$PWD/oponn_b200
 *   ./orx_demo [leaf-file|-] [playouts] [seed]
 *
 * Creates a context for a 42-territory ring map (6 continents of 7, 6 players, cards on), packs a leaf record by
 * hand (or reads a packed 42-territory / 6-player record from a file) and evaluates it the way
 * MCTSTree.simulation + backpropagation do (MCTSTree.java:359-431): wins per player and the mean game length
 * from the exact integer sums of OrxLeafResult.  tests/test_abi_and_host.py compiles and links this file;
 * tests/test_gpu_parity.py runs it and compares its output with the Python binding.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "orx.h"

#define N 42
#define P 6
#define C 6

int main(int argc, char **argv)
{
    int32_t adj_off[N + 1], adj[2 * N], continent[N], bonus[C] = { 5, 2, 5, 3, 7, 2 };
    for (int t = 0; t < N; t++) {
        adj_off[t] = 2 * t;
        adj[2 * t] = (t + N - 1) % N;
        adj[2 * t + 1] = (t + 1) % N;
        continent[t] = t / 7;
    }
    adj_off[N] = 2 * N;
    OrxMap map = { N, C, P, 0, adj_off, adj, continent, bonus, NULL };
    OrxRules rules = { 1, 1, 0, 0, "4, 6, 8, 10, 15, 20...", 0, 0 };

    OrxCtx *ctx = NULL;
    int rc = orx_create(&map, &rules, 0, &ctx);
    if (rc != ORX_OK) { fprintf(stderr, "orx_create: %s\n", orx_last_error()); return 2; }   /* no CPU fallback */

    const int stride = orx_leaf_stride(ctx);
    OrxLeafLayout L = orx_leaf_layout(N, P);
    uint8_t *leaf = calloc(1, (size_t)stride);
    if (argc > 1 && strcmp(argv[1], "-") != 0) {
        FILE *f = fopen(argv[1], "rb");
        if (!f || fread(leaf, 1, (size_t)stride, f) != (size_t)stride) { fprintf(stderr, "cannot read a %d-byte leaf from %s\n", stride, argv[1]); return 2; }
        fclose(f);
    } else {
        /* territories dealt round-robin, 3 armies each, player 0 about to start its CARDS phase, nobody holds cards,
         * the whole deck (42 territory cards + 2 wildcards) unseen */
        OrxLeafHeader *h = (OrxLeafHeader *)leaf;
        h->turn_count = 1; h->attacking_cc = -1; h->defending_cc = -1; h->card_pos = 1;
        h->current_player = 0; h->phase = ORX_PHASE_CARDS; h->attack_state = ORX_ATTACK_START; h->defending_player = -1;
        h->attack_until_dead = 1; h->n_hand = 0; h->n_deck = N + 2;
        int32_t *armies = (int32_t *)(leaf + L.off_armies);
        for (int t = 0; t < N; t++) { leaf[L.off_owner + t] = (uint8_t)(t % P); armies[t] = 3; }
        for (int i = 0; i < N; i++) leaf[L.off_cards + i] = (uint8_t)i;
        leaf[L.off_cards + N] = ORX_CARD_WILD; leaf[L.off_cards + N + 1] = ORX_CARD_WILD;
        /* CardManager.simReset re-seeds its generator from simDeck.hashCode() (CardManager.java:245): a Java host
         * passes that value; without a JVM the deck hash of orx_state.h stands in for it */
        h->deal_mode = ORX_DEAL_LEAF_LCG; h->deal_seed = orx_deck_hash(leaf + L.off_cards, N + 2);
    }
    const int playouts = argc > 2 ? atoi(argv[2]) : 4096;
    const uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 20261019ull;

    OrxLeafResult r;
    rc = orx_rollout_batch(ctx, leaf, 1, playouts, seed, 0, &r, NULL);
    if (rc != ORX_OK) { fprintf(stderr, "orx_rollout_batch: %s\n", orx_last_error()); return 2; }
    printf("playouts %lld flagged %lld mean_rounds %.3f kernel_ms %.2f\n", (long long)r.n, (long long)r.flagged,
           r.n ? (double)r.len_sum / (double)r.n : 0.0, orx_last_kernel_ms(ctx));
    printf("wins");
    for (int p = 0; p < P; p++) printf(" %lld", (long long)r.wins[p]);
    printf("\n");
    free(leaf);
    orx_destroy(ctx);
    return 0;
}
