package com.mld46.oponn.gpu;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.List;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT_UNALIGNED;
import static java.lang.foreign.ValueLayout.JAVA_SHORT_UNALIGNED;

import com.mld46.oponn.BoardState;
import com.sillysoft.lux.Card;
import com.sillysoft.lux.Country;

/**
 * Packs a BoardState (src/com/mld46/oponn/BoardState.java:16-55) plus CardManager's real hand / unseen deck
 * (src/com/mld46/oponn/CardManager.java:20-31) into the flat leaf record of include/orx_state.h.
 *
 * NOT COMPILED in the build image (no JDK, Lux SDK not vendored): reference-side glue for a maintainer.
 * CardManager needs two trivial accessors for this: getCardsHeld() exists (:224-227); add
 * getCardsNotInHand() { return notInHand; } and getCardProgressionPos() { return cardProgressionPos; }.
 */
public final class LeafPacker {
    public static final int HEADER = 32;
    public static final byte WILD = (byte) 0xFF, NONE = (byte) 0xFF;

    public static int align4(int x) { return (x + 3) & ~3; }

    /** == orx_leaf_layout(N, P).stride */
    public static int stride(int n, int p) {
        int off = HEADER + align4(n) + 4 * n + align4(p) + align4(n + 2);
        return (off + 15) & ~15;
    }

    /** Writes one record at `dst` (stride(n, p) bytes).  attackUntilDead = the simulation boards' sticky bit
     *  (SURVEY.md 8.1 quirk 2): true once any attack() has run on them. */
    public static void pack(MemorySegment dst, BoardState bs, List<Card> hand, List<Card> deck, int cardProgressionPos,
                            boolean attackUntilDead) {
        final int n = bs.numberOfCountries, p = bs.numberOfPlayers;
        final int offOwner = HEADER, offArmies = offOwner + align4(n), offNcards = offArmies + 4 * n, offCards = offNcards + align4(p);
        dst.fill((byte) 0);
        dst.set(JAVA_INT_UNALIGNED, 0, bs.turnCount);
        dst.set(JAVA_INT_UNALIGNED, 4, bs.numberOfPlaceableArmies);
        dst.set(JAVA_SHORT_UNALIGNED, 8, (short) bs.attackingCC);
        dst.set(JAVA_SHORT_UNALIGNED, 10, (short) bs.defendingCC);
        dst.set(JAVA_SHORT_UNALIGNED, 12, (short) cardProgressionPos);
        dst.set(JAVA_BYTE, 14, (byte) bs.currentPlayer);
        dst.set(JAVA_BYTE, 15, (byte) bs.currentPhase.ordinal());          // SimulationBoard.Phase ordinals == ORX_PHASE_*
        dst.set(JAVA_BYTE, 16, (byte) (bs.attackState == null ? 0 : bs.attackState.ordinal()));
        dst.set(JAVA_BYTE, 17, (byte) (bs.hasInvadedACountry ? 1 : 0));
        dst.set(JAVA_BYTE, 18, (byte) bs.defendingPlayer);
        dst.set(JAVA_BYTE, 19, (byte) (attackUntilDead ? 1 : 0));
        dst.set(JAVA_BYTE, 20, (byte) hand.size());
        dst.set(JAVA_BYTE, 21, (byte) deck.size());
        // CardManager.simReset re-creates its generator as `new Random(simDeck.hashCode())` on every setupState
        // (CardManager.java:240-248): simDeck is a copy of notInHand, List.hashCode() depends only on the elements, so
        // this IS the seed the reference would use; the engine deals from the same java.util.Random LCG (orx_state.h).
        dst.set(JAVA_BYTE, 23, (byte) 0);                                   // ORX_DEAL_LEAF_LCG
        dst.set(JAVA_INT_UNALIGNED, 24, new java.util.ArrayList<Card>(deck).hashCode());
        for (int cc = 0; cc < n; cc++) {
            Country c = bs.countries[cc];
            dst.set(JAVA_BYTE, offOwner + cc, c.getOwner() < 0 ? NONE : (byte) c.getOwner());
            // an ATTACK / INVADE state carries the conquered country at 0 armies (ExplorationBoard after an invading
            // AttackOutcome: finishAttack zeroed it, executeInvasion overwrites it): the record allows exactly that 0
            dst.set(JAVA_INT_UNALIGNED, offArmies + 4L * cc, c.getArmies());
        }
        for (int i = 0; i < p; i++) dst.set(JAVA_BYTE, offNcards + i, (byte) bs.playerNumberOfCards[i]);
        int k = 0;
        for (Card c : hand) dst.set(JAVA_BYTE, offCards + k++, c.getCode() < 0 ? WILD : (byte) c.getCode());
        for (Card c : deck) dst.set(JAVA_BYTE, offCards + k++, c.getCode() < 0 ? WILD : (byte) c.getCode());
    }

    /** Tree-only extras of include/orx_tree.h: int32 moveableArmies[N] then uint8 fortifiedCountries[N]. */
    public static void packExtra(MemorySegment dst, BoardState bs) {
        final int n = bs.numberOfCountries;
        for (int cc = 0; cc < n; cc++) {
            dst.set(JAVA_INT_UNALIGNED, 4L * cc, bs.countries[cc].getMoveableArmies());
            dst.set(JAVA_BYTE, 4L * n + cc, (byte) (bs.fortifiedCountries != null && bs.fortifiedCountries[cc] ? 1 : 0));
        }
    }

    public static MemorySegment allocate(Arena arena, int n, int p, int leaves) {
        return arena.allocate((long) stride(n, p) * leaves, 16);
    }
}
