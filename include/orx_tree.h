/*
 * orx_tree.h -- C ABI of the native search side of the rollout seam: the leaf producer (move application and
 * legal-move generation) and the batched UCT driver that keeps thousands of leaves in flight on the GPU.
 *
 * SURVEY.md section 8(f) "next" rows f2 and f1.  Reference code restated behind these entry points:
 *   leaf producer  src/com/mld46/oponn/sim/boards/ExplorationBoard.java:116-283 (updateState), :289-706
 *                  (getPossibleMoves and its per-phase generators), :755-779 (sparseIndices), :816-835
 *                  (getBoardState), :838-902 (FortificationAnalysis); move vocabulary src/com/mld46/oponn/moves/
 *   search         src/com/mld46/oponn/MCTSTree.java:159-210 (runMCTS), :226-331 (selection), :335-355 (expansion),
 *                  :399-431 (backpropagation), :541-568 (chooseRobustMove), :621-674 (SearchNode)
 *
 * What differs from the reference, deliberately: runMCTS evaluates one leaf at a time (MCTSTree.java:179-206);
 * orx_tree_search selects `leaves_in_flight` leaves per batch under a virtual loss, evaluates the whole batch with
 * one orx_rollout_batch call, then back-propagates.  With leaves_in_flight = 1 it is the reference's loop.
 *
 * Plain pointers and sizes only; the host side is C++ (the reference is compiled code; no JDK in this image).
 */
#ifndef ORX_TREE_H
#define ORX_TREE_H

#include "orx.h"

#ifdef __cplusplus
extern "C" {
#endif

/* move kinds = the classes of src/com/mld46/oponn/moves/ */
enum {
    ORX_MOVE_COUNTRY_SELECTION = 0,  /* a = cc                                               */
    ORX_MOVE_INITIAL_PLACEMENT = 1,  /* a = cc, b = armies                                   */
    ORX_MOVE_CARD_CASH         = 2,  /* a, b, c = card codes (ORX_CARD_WILD = wildcard)      */
    ORX_MOVE_PLACEMENT         = 3,  /* a = cc, b = armies                                   */
    ORX_MOVE_ATTACK            = 4,  /* a = attackingCC, b = defendingCC, c = attackingArmies, d = defendingArmies, flag = attackUntilDead */
    ORX_MOVE_ATTACK_OUTCOME    = 5,  /* a = attackerLosses, b = defenderLosses, probability, flag = invading (a chance outcome) */
    ORX_MOVE_INVASION          = 6,  /* a = attackingCC, b = defendingCC, c = armies         */
    ORX_MOVE_FORTIFICATION     = 7,  /* a = armies, b = sourceCC, c = destinationCC (-1: mark source as done) */
    ORX_MOVE_NEXT_PHASE        = 8   /* a = nextPhase, b = previousPhase, c = player         */
};

typedef struct OrxMove {
    int32_t type;          /* ORX_MOVE_*                                                     */
    int32_t phase;         /* Move.phase: ORX_PHASE_* (NEXT_PHASE for ORX_MOVE_NEXT_PHASE)   */
    int32_t a, b, c, d;
    float   probability;   /* AttackOutcome.probability                                      */
    int32_t flag;
} OrxMove;

/* The seven pruning switches of ExplorationBoard (defaults = Oponn.java:46-52) + search settings
 * (Oponn.java:43-60, MCTSTree.java:23-27). */
typedef struct OrxTreeOptions {
    int32_t placement_partial_order;   /* default 1 */
    int32_t attack_aggressive;         /* default 1 */
    int32_t attack_repeat_moves;       /* default 1 */
    int32_t attack_partial_order;      /* default 0 */
    int32_t attack_until_dead;         /* default 1 */
    int32_t invasion;                  /* default 1 */
    int32_t fortification_outwards;    /* default 1 */
    int32_t agent_id;                  /* Oponn's player id: whose win counts rank the final move            */
    int32_t playouts_per_leaf;         /* "cores" (MCTSTree.java:69,378)                                     */
    int32_t leaves_in_flight;          /* leaves selected per batch under virtual loss; 1 = the reference     */
    int32_t max_child_policy;          /* 0 = ROBUST_CHILD (default, Oponn.java:57), 1 = MAX_CHILD            */
    int32_t reserved;
    uint64_t seed;                     /* expansion choices, tie-break noise, hidden hands, rollout streams  */
} OrxTreeOptions;

void orx_tree_default_options(OrxTreeOptions *o);

typedef struct OrxTree OrxTree;

/* ctx may be NULL: move generation and application work without a GPU; orx_tree_search needs one. */
int  orx_tree_create(const OrxMap *map, const OrxRules *rules, const OrxTreeOptions *opt, OrxCtx *ctx, OrxTree **out);
void orx_tree_destroy(OrxTree *t);

/* Tree-only state that BoardState carries but a rollout never reads (so it is not in the leaf record):
 * `extra` = int32 moveableArmies[N] (Country.getMoveableArmies) followed by uint8 fortifiedCountries[N]
 * (BoardState.java:55).  NULL means: nothing fortified yet, moveable = armies on the current player's territories
 * when the record is in FORTIFICATION (what setupFortificationPhase would write), else 0. */
static inline int orx_tree_extra_bytes(int n_territories) { return 5 * n_territories; }

/* ExplorationBoard.setupState(boardState) + getPossibleMoves().  `leaf` is a packed record (orx_state.h).
 * Returns the move count (moves beyond `cap` are not written). */
int orx_tree_possible_moves(OrxTree *t, const void *leaf, const void *extra, OrxMove *out, int cap);

/* setupState + updateState(move) + getBoardState(): the successor as a packed record (real hand / deck / progression
 * position are carried over unchanged, as CardManager does).  *ongoing = updateState's return value. */
int orx_tree_apply(OrxTree *t, const void *leaf, const void *extra, const OrxMove *move, void *leaf_out,
                   void *extra_out, int32_t *ongoing);

/* Replaces SimulationBoard.setupNewGame (SimulationBoard.java:141-176): the territories are shuffled
 * (Collections.shuffle with java.util.Random(seed)) and dealt round-robin with one army each; the record starts
 * player 0's INITIAL_PLACEMENT with setupInitialPlacementPhase's allowance; nobody holds cards. */
int orx_new_game(const OrxMap *map, uint64_t seed, void *leaf_out);

/* One root child after a search, in getPossibleMoves() order of the root (so that independent trees merge by
 * index: root-parallel mode, one all-reduce of these integers per move). */
typedef struct OrxRootChild {
    OrxMove move;
    int64_t visits;
    int64_t wins[ORX_MAX_PLAYERS];
    int64_t losses[ORX_MAX_PLAYERS];
    int64_t win_len_sum[ORX_MAX_PLAYERS];    /* averageWinLength[p]  = win_len_sum[p]  / wins[p]   */
    int64_t loss_len_sum[ORX_MAX_PLAYERS];   /* averageLossLength[p] = loss_len_sum[p] / losses[p] */
} OrxRootChild;

typedef struct OrxSearchStats {
    int64_t iterations;       /* leaves evaluated                        */
    int64_t playouts;         /* iterations * playouts_per_leaf          */
    int64_t batches;          /* orx_rollout_batch calls                 */
    int64_t nodes;            /* SearchNodes allocated                   */
    int64_t terminal_hits;    /* selections that ended in a finished game */
    int64_t max_depth;
    double  select_ms, rollout_ms, backprop_ms;
} OrxSearchStats;

/* Replaces MCTSTree.runMCTS(boardState) (MCTSTree.java:159-210): builds a fresh tree, evaluates `iterations`
 * leaves, fills `children` (up to cap, returns the true count in *n_children).  No final choice is made here:
 * call orx_tree_choose on the (possibly merged) children. */
int orx_tree_search(OrxTree *t, const void *root_leaf, const void *extra, int iterations,
                    OrxRootChild *children, int cap, int32_t *n_children, OrxSearchStats *stats);

/* chooseRobustMove / chooseMaxMove (MCTSTree.java:541-595) over root children; returns the index chosen. */
int orx_tree_choose(const OrxTree *t, const OrxRootChild *children, int n_children, uint64_t tie_break);

#ifdef __cplusplus
}
#endif
#endif /* ORX_TREE_H */
