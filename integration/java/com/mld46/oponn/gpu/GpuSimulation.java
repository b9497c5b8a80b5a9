package com.mld46.oponn.gpu;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import com.mld46.oponn.BoardState;
import com.mld46.oponn.CardManager;

/**
 * Drop-in for the body of MCTSTree.simulation(boardState) + the result reads of backpropagation
 * (src/com/mld46/oponn/MCTSTree.java:359-395, 405-427): `cores` playouts of one leaf become one GPU call.
 *
 * In MCTSTree:
 *     private final GpuSimulation gpu = new GpuSimulation(boardState, cardManager, device, cores);
 *     ...
 *     if (simOngoing) { expansion(path); gpu.simulate(explorationBoard.getBoardState()); }
 *     backpropagation(path):  for j < cores: playerOutOnTurn = gpu.playerOutOnTurn(j); finalTurn = gpu.simTurnCount(j);
 *
 * NOT COMPILED in the build image (no JDK): reference-side glue for a maintainer.  Orx.java is next to this file.
 */
public final class GpuSimulation implements AutoCloseable {
    private static final int GAME_RESULT_BYTES = 48;   // OrxGameResult: int32 playerOutOnTurn[8], simTurnCount, flags, streamPos, boardHash
    private final Orx orx;
    private final CardManager cardManager;
    private final Arena arena = Arena.ofShared();
    private final MemorySegment leaf, games;
    private final int cores, players;
    private final long seed = System.nanoTime();
    private long gameCounter = 0;

    public GpuSimulation(BoardState bs, CardManager cardManager, int device, int cores) throws Throwable {
        this.orx = new Orx(bs, device);
        this.cardManager = cardManager;
        this.cores = cores;
        this.players = bs.numberOfPlayers;
        this.leaf = LeafPacker.allocate(arena, bs.numberOfCountries, bs.numberOfPlayers, 1);
        this.games = arena.allocate((long) GAME_RESULT_BYTES * cores, 8);
    }

    /** setupState + simulate on `cores` boards (MCTSTree.java:372-373), as one call */
    public void simulate(BoardState boardState) throws Throwable {
        LeafPacker.pack(leaf, boardState, cardManager.getCardsHeld(), cardManager.getCardsNotInHand(),
                        cardManager.getCardProgressionPos(), true);
        orx.rolloutGames(leaf, 1, cores, seed, gameCounter, games);   // orx_rollout_batch(..., agg = NULL, games)
        gameCounter += cores;
        // the Java board throws on an illegal state (SimulationBoard.java:357,430,...); the engine flags the playout
        // instead (ORX_FLAG_ERROR = 1) and zeroes its record: never back-propagate such a record
        for (int j = 0; j < cores; j++) {
            int flags = games.get(java.lang.foreign.ValueLayout.JAVA_INT, (long) GAME_RESULT_BYTES * j + 36);
            if ((flags & 1) != 0) throw new IllegalStateException("rollout engine rejected the leaf record (ORX_FLAG_ERROR), playout " + j);
        }
    }

    /** board.getPlayerOutOnTurn() of playout j (MCTSTree.java:409) */
    public int[] playerOutOnTurn(int j) {
        int[] out = new int[players];
        for (int p = 0; p < players; p++) out[p] = games.get(java.lang.foreign.ValueLayout.JAVA_INT, (long) GAME_RESULT_BYTES * j + 4L * p);
        return out;
    }

    /** board.getSimTurnCount() of playout j (MCTSTree.java:410) */
    public int simTurnCount(int j) {
        return games.get(java.lang.foreign.ValueLayout.JAVA_INT, (long) GAME_RESULT_BYTES * j + 32);
    }

    @Override public void close() { orx.close(); arena.close(); }
}
