/*
 * orx.h -- C ABI of the B200 rollout engine (liborx.so).
 *
 * The reference has no FFI: the rollout seam is three plain Java calls made only from
 * MCTSTree.simulation / backpropagation (src/com/mld46/oponn/MCTSTree.java:372-373,409-410)
 * on boards built once per game (src/com/mld46/oponn/Oponn.java:106-110).  Each entry point
 * below names the reference call(s) it replaces; INTEGRATION.md shows the Java (Panama FFM)
 * and Python (ctypes) bindings a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; the caller owns every buffer it passes; the
 * library owns its device scratch.  Every function returns 0 on success or a negative
 * ORX_E_* code (the reference throws unchecked exceptions instead -- SURVEY.md 8b).
 * A context is bound to one CUDA device and one stream; calls on one context must not
 * overlap, distinct contexts are independent (launches of different contexts are serialised
 * for the instant of the launch itself: the kernels' shared-memory size is set per launch).  There is no CPU fallback: without a usable
 * CUDA device orx_create fails with ORX_E_CUDA.
 */
#ifndef ORX_H
#define ORX_H

#include "orx_state.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORX_OK            0
#define ORX_E_INVALID    -1   /* bad argument / map / rules                               */
#define ORX_E_CUDA       -2   /* CUDA runtime error (orx_last_error has the text)         */
#define ORX_E_LEAF       -3   /* a leaf record failed validation (IllegalArgumentException
                                 analogue of SimulationBoard.java:357,430)                */
#define ORX_E_NOMEM      -4

typedef struct OrxCtx OrxCtx;

/* Replaces: new ModellingBoard(boardState, simAgents, cardManager, i, limit) x CORES
 * (Oponn.java:106-110 -> SimulationBoard.java:102-139) and BattleSim's static table build
 * (BattleSim.java:46-59).  Copies map and rules, stages them on `device`, builds the battle
 * outcome tables for every (a, d) in 1..100 x 1..100 on the GPU.  All playout policies are
 * SimRandom (lux_agent_modelling=false) until orx_set_sim_agents says otherwise. */
int orx_create(const OrxMap *map, const OrxRules *rules, int device, OrxCtx **out);

/* Replaces: letting the boards be garbage collected. */
void orx_destroy(OrxCtx *ctx);

/* Replaces: the `SimAgent[] sas` and `modellingTurnLimit` arguments of `new ModellingBoard(bs, sas, cm, cid, limit)`
 * (ModellingBoard.java:13-19; built by Oponn.createSimAgents, Oponn.java:146-162, from the real agents' names when
 * lux_agent_modelling is on).  agent_ids[p] = ORX_AGENT_* for each of the map's players (orx_state.h; use
 * orx_agent_from_name on Board.getAgentName(p)); modelling_turn_limit <= 0 or NULL ids switch modelling off (the
 * state after orx_create).  Takes effect for the rollouts launched after the call; must not be called while a
 * rollout of this context is in flight. */
int orx_set_sim_agents(OrxCtx *ctx, const int32_t *agent_ids, int modelling_turn_limit);

/* Text of the last error on this thread (never NULL). */
const char *orx_last_error(void);

/* Bytes per leaf record for this context (== orx_leaf_layout(N, P).stride). */
int orx_leaf_stride(const OrxCtx *ctx);

/* Replaces, batched and fused: for each leaf, `playouts_per_leaf` times
 *     board.setupState(boardState); board.simulate();            (MCTSTree.java:372-373)
 *     board.getPlayerOutOnTurn(); board.getSimTurnCount();       (MCTSTree.java:409-410)
 * HOST buffers: leaves (n_leaves * stride bytes) are copied to the device, results are
 * copied back before returning.  Playout k of leaf i uses stream (seed, game id
 * first_game + i*playouts_per_leaf + k), so results do not depend on how a batch is split
 * across calls, contexts or GPUs.  `agg` (n_leaves records) receives exact integer sums;
 * `games` (n_leaves*playouts_per_leaf records) may be NULL. */
int orx_rollout_batch(OrxCtx *ctx, const void *leaves, int n_leaves, int playouts_per_leaf,
                      uint64_t seed, uint64_t first_game,
                      OrxLeafResult *agg, OrxGameResult *games);

/* Same, DEVICE buffers, asynchronous on the context's stream (use orx_sync).  `agg_dev`
 * is overwritten (zeroed, then accumulated).  `games_dev` may be NULL. */
int orx_rollout_batch_device(OrxCtx *ctx, const void *leaves_dev, int n_leaves, int playouts_per_leaf,
                             uint64_t seed, uint64_t first_game,
                             OrxLeafResult *agg_dev, OrxGameResult *games_dev);

/* The same call split in two, on one of two independent in-flight slots (slot 0 or 1, each with its own CUDA
 * stream): begin copies the leaves and launches, end waits and copies the results back.  Two batches in flight let
 * the tail of one (a batch ends with its longest game) overlap the bulk of the next: the batched search
 * (orx_tree.h) alternates slots.  `games` may be NULL in both; a slot must be ended before it is begun again. */
int orx_rollout_begin(OrxCtx *ctx, int slot, const void *leaves, int n_leaves, int playouts_per_leaf,
                      uint64_t seed, uint64_t first_game, int want_games);
int orx_rollout_end(OrxCtx *ctx, int slot, OrxLeafResult *agg, OrxGameResult *games);

/* Replay of injected streams (the "identical injected dice/move stream" harness): game g
 * plays leaf g with the explicit words words[g*words_per_game ...].  HOST buffers. */
int orx_rollout_replay(OrxCtx *ctx, const void *leaves, int n_games,
                       const uint32_t *words, uint32_t words_per_game, OrxGameResult *games);

/* Block until everything queued on the context's stream has finished. */
int orx_sync(OrxCtx *ctx);

/* The CUDA stream (cudaStream_t) the context launches on, for event timing by the host. */
void *orx_stream(OrxCtx *ctx);

/* Milliseconds the last rollout kernel took on the device (CUDA events on the context's
 * stream; valid after orx_sync or a host-buffer call), and how many kernels the context has
 * launched since creation. */
float    orx_last_kernel_ms(OrxCtx *ctx);
uint64_t orx_kernel_launches(const OrxCtx *ctx);

/* Which rollout kernel the last launch of this context used: 1 = warp-per-game (orx_game.cuh: any map, replay
 * streams, modelled opponents: the default everywhere), 2 = lane-per-game (orx_lanes.cuh: all-SimRandom counter-stream
 * batches on maps up to 64 territories; opt-in through the environment, ORX_KERNEL=lanes or ORX_LANE_MIN_TOTAL=n, because
 * it is currently the slower of the two on B200), 0 = none yet.  Both produce identical records. */
int orx_last_kernel_kind(const OrxCtx *ctx);

/* Launch shape of the rollout kernel a counter-stream batch of `total_playouts` would run on with the context's current
 * map and line-up: kind as above, resident blocks per SM, threads per block, dynamic shared memory per block (the
 * board-in-shared-memory footprint BASELINE.json config 4 asks about).  Any output pointer may be NULL. */
int orx_kernel_info(const OrxCtx *ctx, uint64_t total_playouts, int32_t *kind, int32_t *blocks_per_sm,
                    int32_t *threads_per_block, int32_t *smem_bytes_per_block);

/* Replaces: BattleSim.getAttackOutcomes(a, d) (BattleSim.java:174-184) for 1 <= a,d <= 100.
 * Rows are in the reference's order (descending probability, stable).  Returns the row count
 * (<= a+d) or a negative error; arrays hold up to `cap` rows. */
int orx_get_attack_outcomes(OrxCtx *ctx, int attackers, int defenders,
                            float *probability, int32_t *attacker_losses, int32_t *defender_losses,
                            int32_t *invading, int cap);

/* Replaces: n calls of BattleSim.simulateBattle(a, d) (BattleSim.java:66-145); battle i draws
 * from stream (seed, first_game + i).  HOST output buffers. */
int orx_simulate_battles(OrxCtx *ctx, int attackers, int defenders, int n,
                         uint64_t seed, uint64_t first_game,
                         int32_t *attacker_losses, int32_t *defender_losses);

/* Replaces: n calls of BattleSim.simulateIndividualBattle(a, d) (BattleSim.java:152-162). */
int orx_simulate_individual(OrxCtx *ctx, int attackers, int defenders, int n,
                            uint64_t seed, uint64_t first_game, int32_t *attacker_losses);

/* Stream self-test: evaluates java.util.Random derivations on the device over stream
 * (seed, game) or over explicit `words` (NULL = counter mode).  ops[i]: >0 nextInt(ops[i]);
 * 0 nextDouble; -1 nextGaussian; -2 raw word.  out[i] receives the value as a double. */
int orx_stream_draws(OrxCtx *ctx, uint64_t seed, uint64_t game, const uint32_t *words, uint32_t n_words,
                     const int32_t *ops, int n_ops, double *out, uint32_t *pos_out, uint32_t *flags_out);

/* Replaces: CardManager.getCardCashValue(pos) (CardManager.java:283-341) as the device sees it. */
int orx_card_cash_value(OrxCtx *ctx, int pos, int32_t *value);

#ifdef __cplusplus
}
#endif
#endif /* ORX_H */
