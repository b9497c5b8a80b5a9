/*
 * orx_state.h -- flat game-state encoding shared by every side of the rollout seam.
 *
 * One header, plain C, no CUDA types: included by the CUDA engine (oponn_b200/csrc),
 * by the CPU oracle (oracle/), and mirrored byte-for-byte by the Java / Python hosts
 * (INTEGRATION.md).  It replaces the Java object graph that crosses the seam in the
 * reference:
 *
 *   BoardState                        src/com/mld46/oponn/BoardState.java:16-55
 *   the Country[] it borrows          src/com/mld46/oponn/BoardState.java:118-139
 *   CardManager's real hand / deck    src/com/mld46/oponn/CardManager.java:20-31,240-248
 *   the board's result getters        src/com/mld46/oponn/sim/boards/SimulationBoard.java:1031-1044
 *
 * Everything is little-endian, naturally aligned, and position independent.
 */
#ifndef ORX_STATE_H
#define ORX_STATE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef __CUDACC__
#define ORX_HD __host__ __device__
#else
#define ORX_HD
#endif

/* ---- limits ------------------------------------------------------------ */
#define ORX_MAX_TERRITORIES 252   /* card codes and list sizes (N+2) are bytes       */
#define ORX_MAX_PLAYERS     8
#define ORX_MAX_CONTINENTS  32
#define ORX_MAX_DEGREE      32    /* out-neighbours per territory (one per lane)     */
#define ORX_BONUS_ROUNDS    1024  /* rows of the per-round continent bonus table     */
#define ORX_CARD_VALUES     1024  /* precomputed getCardCashValue(pos), pos < this   */
#define ORX_MAX_ARMIES      (1 << 20)  /* per territory / placeable in a leaf record    */

#define ORX_CARD_WILD   0xFFu     /* Card(-1,3): CardManager.java:14                 */
#define ORX_OWNER_NONE  0xFFu     /* Country owner -1 (only before/during the draft) */

/* SimulationBoard.Phase ordinals (SimulationBoard.java:19-27) */
enum {
    ORX_PHASE_COUNTRY_SELECTION = 0,
    ORX_PHASE_INITIAL_PLACEMENT = 1,
    ORX_PHASE_CARDS             = 2,
    ORX_PHASE_PLACEMENT         = 3,
    ORX_PHASE_ATTACK            = 4,
    ORX_PHASE_FORTIFICATION     = 5,
    ORX_PHASE_NEXT_PHASE        = 6
};

/* SimulationBoard.AttackState ordinals (SimulationBoard.java:29-35) */
enum {
    ORX_ATTACK_START   = 0,
    ORX_ATTACK_EXECUTE = 1,
    ORX_ATTACK_INVADE  = 2,
    ORX_ATTACK_CASH    = 3,
    ORX_ATTACK_PLACE   = 4
};

/* Card sets: com.sillysoft.lux.Card is not part of the reference repository (SDK jar, unpinned).
 * DECLARED ASSUMPTION used by every implementation of this header: three cards are a set when one is a
 * wildcard, or all three symbols are equal, or all three differ; Card.getRandomSet(hand) picks uniformly
 * among the valid triples i<j<k of the hand, enumerated lexicographically, with one nextInt(count). */

/* ---- static per-map / per-ruleset inputs ------------------------------- */

/* The Country graph + board constants the SimulationBoard constructor captures
 * (SimulationBoard.java:102-139).  CSR adjacency keeps Country.getAdjoiningList()
 * ORDER: SimRandom picks "the j-th enemy neighbour in list order"
 * (SimRandom.java:95-109), so the order is part of the semantics. */
typedef struct OrxMap {
    int32_t        n_territories;    /* numberOfCountries                        */
    int32_t        n_continents;     /* numberOfContinents                       */
    int32_t        n_players;        /* numberOfPlayers                          */
    int32_t        reserved;
    const int32_t *adj_offsets;      /* [n_territories+1]                        */
    const int32_t *adj;              /* [adj_offsets[n]] neighbour codes         */
    const int32_t *continent;        /* [n_territories] Country.getContinent()   */
    const int32_t *continent_bonus;  /* [n_continents] initialContinentBonuses   */
    const int32_t *card_symbol;      /* [n_territories] 0..2, or NULL = code % 3
                                        (CardManager.java:54-61: counter cycles in
                                        Country[] iteration order)               */
} OrxMap;

/* Game settings (BoardState.java:22-26) + engine limits. */
typedef struct OrxRules {
    int32_t     use_cards;
    int32_t     transfer_cards;
    int32_t     immediate_cash;       /* carried; compiled out in the reference
                                         (SimulationBoard.java:852-855)          */
    int32_t     continent_increase;   /* percent per round                       */
    const char *card_progression;     /* Lux progression string, e.g.
                                         "4, 6, 8, 10, 15, 20..."                */
    int32_t     max_player_turns;     /* step cap; 0 = default (1<<16)           */
    int32_t     reserved;
} OrxRules;

/* ---- opponent modelling ------------------------- */

/* Playout policies = the SimAgent subclasses of src/com/mld46/oponn/sim/agents/ (SimAgent.getSimAgent,
 * SimAgent.java:248-307).  Every player is SimRandom unless orx_set_sim_agents() says otherwise; with modelling on,
 * ModellingBoard (ModellingBoard.java:29-47) swaps everybody back to SimRandom at the round wrap where
 * simTurnCount + 1 == modelling_turn_limit (Oponn.java:55: 2).  Agent names the reference models but this engine does
 * not yet (cluster, shaft, boscoe, communist, evilpixie, killbot, pixie, quo, yakool) map to ORX_AGENT_RANDOM.
 *
 * DECLARED ASSUMPTIONS (Lux SDK not vendored, version unpinned; "parity unpinned" for every modelled policy):
 *   Country.getNumberEnemyNeighbors()   adjoining countries whose owner differs from this country's owner
 *   Country.getNumberPlayerNeighbors(p) adjoining countries owned by p;  getNumberNeighbors() = adjoining count
 *   Country.getWeakestEnemyNeighbor()   the enemy adjoining country with the fewest armies, the first such in
 *                                       getAdjoiningList() order; null when there is none
 *   PlayerIterator(p) / ArmiesIterator(p, n) / ContinentIterator(c): countries in code order that are owned by p /
 *                                       owned by p with armies >= n (n > 0) or <= -n (n < 0) / lie in continent c; the
 *                                       iterator holds ONE element ahead: next() returns the held element and looks
 *                                       for the following one in the board as it is at that moment
 *   BoardHelper.getPlayerArmies(p)      sum of armies on p's countries
 *   the agents' private `rand` objects  draw from the playout's injected stream like every other draw
 * Moveable armies: the leaf record does not carry Country.moveableArmies (orx_tree.h keeps them tree-side); a leaf in
 * FORTIFICATION starts with setupFortificationPhase's values (armies on the current player's countries, else 0). */
enum { ORX_AGENT_RANDOM = 0, ORX_AGENT_ANGRY = 1, ORX_AGENT_STINKY = 2, ORX_AGENT_COUNT = 3 };

/* SimAgent.getSimAgent(name) (SimAgent.java:248-307): case-insensitive name -> policy id */
static inline int orx_agent_from_name(const char *name)
{
    char b[16]; int i = 0;
    if (!name) return ORX_AGENT_RANDOM;
    for (; name[i] && i < 15; i++) b[i] = (char)((name[i] >= 'A' && name[i] <= 'Z') ? name[i] + 32 : name[i]);
    b[i] = 0;
    if (name[i]) return ORX_AGENT_RANDOM;
    if (b[0] == 'a' && b[1] == 'n' && b[2] == 'g' && b[3] == 'r' && b[4] == 'y' && !b[5]) return ORX_AGENT_ANGRY;
    if (b[0] == 's' && b[1] == 't' && b[2] == 'i' && b[3] == 'n' && b[4] == 'k' && b[5] == 'y' && !b[6]) return ORX_AGENT_STINKY;
    return ORX_AGENT_RANDOM;
}

/* engine limit: SimStinky.placeArmies redraws until it hits a border country of its own (SimStinky.java:68-78);
 * after this many consecutive misses the playout is stopped with ORX_FLAG_STEP_LIMIT */
#define ORX_MAX_REDRAWS 65536

/* ---- the leaf record (one MCTS leaf = one rollout start state) --------- */

/* Fixed 32-byte header followed by four byte arrays whose offsets depend only on
 * (n_territories, n_players): see orx_leaf_layout().  Records in a batch are
 * leaf_stride bytes apart (multiple of 16) so a warp loads one with 128-bit
 * coalesced reads. */
typedef struct OrxLeafHeader {
    int32_t turn_count;        /* BoardState.turnCount -> simStartTurn               */
    int32_t placeable;         /* BoardState.numberOfPlaceableArmies                 */
    int16_t attacking_cc;      /* BoardState.attackingCC, -1 = none                  */
    int16_t defending_cc;      /* BoardState.defendingCC, -1 = none                  */
    int16_t card_pos;          /* CardManager.cardProgressionPos (CardManager.java:247) */
    int8_t  current_player;    /* BoardState.currentPlayer                           */
    int8_t  phase;             /* ORX_PHASE_*                                        */
    int8_t  attack_state;      /* ORX_ATTACK_*                                       */
    int8_t  has_invaded;       /* BoardState.hasInvadedACountry                      */
    int8_t  defending_player;  /* BoardState.defendingPlayer, -1 = none              */
    int8_t  attack_until_dead; /* the board's sticky attackUntilDead bit (SURVEY 8.1
                                  quirk 2; SimulationBoard.java:93,754,763)          */
    uint8_t n_hand;            /* |CardManager.inHand|: the real hand, handed to
                                  whoever is current_player (SimulationBoard.java:215-218) */
    uint8_t n_deck;            /* |CardManager.notInHand|: the unseen deck, in list order */
    uint8_t tree_attack_until_dead; /* tree-only: Attack.attackUntilDead of a pending (EXECUTE) attack as the
                                  exploration board holds it (orx_tree.h); rollouts ignore it            */
    uint8_t deal_mode;         /* ORX_DEAL_*: where CardManager.simDeal's nextInt comes from (below)    */
    int32_t deal_seed;         /* what simDeck.hashCode() returned in CardManager.simReset
                                  (CardManager.java:245); only read when deal_mode == ORX_DEAL_LEAF_LCG  */
    uint8_t reserved[4];
} OrxLeafHeader;

/* Hidden-hand determinisation (CardManager.java:240-263).  CardManager.simReset runs on EVERY
 * setupState and re-creates its generator as `new Random(simDeck.hashCode())`; every simDeal of the playout that
 * follows (the hidden hands dealt in setupCardState, SimulationBoard.java:205-227, and the card dealt after an attack
 * phase with a conquest, :884-896) draws `random.nextInt(deck.size())` from THAT java.util.Random -- so all playouts
 * of one leaf see the same hidden hands and the same deal sequence.  Card.hashCode() belongs to the un-vendored SDK
 * (identity hash unless overridden), so the value cannot be recomputed here: the host injects what the call returned.
 *   ORX_DEAL_LEAF_LCG (default, the reference's behaviour): deals come from a java.util.Random LCG (48-bit state,
 *       scrambled seed (deal_seed ^ 0x5DEECE66D) & (2^48-1), nextInt as documented) private to the playout; they do
 *       NOT consume words of the injected stream.
 *   ORX_DEAL_GAME_STREAM (option): deals draw nextInt from the playout's own word stream, i.e. hidden hands are
 *       re-sampled per playout (the round-1 behaviour; a Java harness must then hand CardManager.random the word
 *       source instead of letting simReset re-seed it).
 * Hosts without a JVM (oponn_b200.state, orx_tree) use orx_deck_hash() below as a stand-in for hashCode(): what
 * matters is "same deck, same hidden hands". */
enum { ORX_DEAL_LEAF_LCG = 0, ORX_DEAL_GAME_STREAM = 1 };

static inline ORX_HD int32_t orx_deck_hash(const uint8_t *deck, int n)
{
    uint64_t h = 1469598103934665603ull;             /* FNV-1a over the deck's card codes, folded to 32 bits */
    for (int i = 0; i < n; i++) { h ^= deck[i]; h *= 1099511628211ull; }
    return (int32_t)(uint32_t)(h ^ (h >> 32));
}

/* Leaf validity (engine contract; the reference would throw, loop or index out of bounds on
 * most of these -- SimulationBoard.java:357,430; here the playout is flagged ORX_FLAG_ERROR):
 *   0 <= current_player < P;  phase in COUNTRY_SELECTION..FORTIFICATION;  attack_state in START..PLACE;
 *   -1 <= attacking_cc, defending_cc < N;  -1 <= defending_player < P;  card_pos >= 0;
 *   |placeable| <= ORX_MAX_ARMIES;  n_hand + n_deck <= N + 2;  every card is a territory code or ORX_CARD_WILD;
 *   every owner is < P (or ORX_OWNER_NONE while phase == COUNTRY_SELECTION);  1 <= armies[t] <= ORX_MAX_ARMIES,
 *   except that armies[defending_cc] may be 0 in an ATTACK / INVADE leaf (ExplorationBoard.getBoardState after an
 *   invading AttackOutcome: finishAttack left it at 0 and executeInvasion overwrites it, SimulationBoard.java:814-825);
 *   deal_mode is an ORX_DEAL_* value.
 * A flagged playout (ORX_FLAG_ERROR or ORX_FLAG_STREAM_END) reports only its first fault -- exactly one of the
 * two bits -- and every other field of its OrxGameResult is zero. */
typedef struct OrxLeafLayout {
    int32_t off_owner;    /* uint8  owner[N]    : 0..P-1, ORX_OWNER_NONE            */
    int32_t off_armies;   /* int32  armies[N]                                       */
    int32_t off_ncards;   /* uint8  ncards[P]   : BoardState.playerNumberOfCards    */
    int32_t off_cards;    /* uint8  cards[N+2]  : hand[0..n_hand) then deck[0..n_deck),
                             each in Java List order; territory code or ORX_CARD_WILD */
    int32_t stride;       /* bytes per record, multiple of 16                       */
} OrxLeafLayout;

static inline ORX_HD OrxLeafLayout orx_leaf_layout(int n_territories, int n_players)
{
    OrxLeafLayout l;
    int off = (int)sizeof(OrxLeafHeader);
    l.off_owner = off;   off += (n_territories + 3) & ~3;
    l.off_armies = off;  off += 4 * n_territories;
    l.off_ncards = off;  off += (n_players + 3) & ~3;
    l.off_cards = off;   off += (n_territories + 2 + 3) & ~3;
    l.stride = (off + 15) & ~15;
    return l;
}

/* ---- results ----------------------------------------------------------- */

#define ORX_FLAG_ERROR        1u  /* the Java board would have thrown (illegal state)    */
#define ORX_FLAG_STEP_LIMIT   2u  /* max_player_turns reached before a terminal position */
#define ORX_FLAG_STREAM_END   4u  /* an injected stream ran out of words                 */

/* One playout: what MCTSTree.backpropagation reads from a board
 * (MCTSTree.java:409-410) plus two trajectory pins used by the parity tests. */
typedef struct OrxGameResult {
    int32_t  player_out_on_turn[ORX_MAX_PLAYERS]; /* -1 winner / sim round eliminated / 0 */
    int32_t  sim_turn_count;                      /* getSimTurnCount()                     */
    uint32_t flags;                               /* ORX_FLAG_*                            */
    uint32_t stream_pos;  /* 32-bit words drawn when the game became terminal            */
    uint32_t board_hash;  /* sum over territories t of orx_mix(t, owner, armies) mod 2^32
                             on the final board (order independent => lane parallel)  */
} OrxGameResult;

static inline ORX_HD uint32_t orx_mix(uint32_t t, uint32_t owner, uint32_t armies)
{
    uint32_t x = armies * 2654435761u ^ ((owner + 1u) * 0x85EBCA77u) ^ ((t + 1u) * 0xC2B2AE3Du);
    x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
    return x;
}

/* Exact integer sums over the playouts of one leaf; every running mean that
 * MCTSTree.backpropagation keeps (MCTSTree.java:412-427) is a ratio of these. */
typedef struct OrxLeafResult {
    int64_t n;                               /* playouts                              */
    int64_t len_sum;                         /* sum of sim_turn_count                 */
    int64_t wins[ORX_MAX_PLAYERS];           /* playouts with player_out_on_turn==-1  */
    int64_t win_len_sum[ORX_MAX_PLAYERS];    /* sum of sim_turn_count over those      */
    int64_t flagged;                         /* playouts with flags != 0              */
    int64_t reserved;
} OrxLeafResult;

/* ---- the injected random stream ---------------------------------------- */
/*
 * Every random draw of a playout comes from ONE sequence of 32-bit words w_0, w_1, ...
 * consumed in program order through java.util.Random's documented derivations
 * (next(bits) = w >>> (32-bits); nextInt(bound); nextDouble(); nextGaussian()).
 * A Java harness reproduces a playout by giving BattleSim.random, SimRandom.random,
 * CardManager.random and Card.getRandomSet the same word source.
 *
 * Counter mode: w_i = Philox4x32-10( ctr = {i / 4, 0, game_lo, game_hi}, key = {seed_lo, seed_hi} )[ i % 4 ]
 * i.e. the four outputs of counter j are words 4j .. 4j+3: a warp that evaluates 32 consecutive counters holds 128
 * consecutive words with lane l owning words 4l .. 4l+3 (two nextDouble()s), and a single thread can walk the same
 * stream four words at a time.  (Round 1 consumed the block component-major; changed in round 2, oracle and goldens
 * in lock-step.)
 * Replay mode: w_i = words[i] from a caller-supplied array.
 */

#ifdef __cplusplus
}
#endif
#endif /* ORX_STATE_H */
