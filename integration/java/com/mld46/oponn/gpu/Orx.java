package com.mld46.oponn.gpu;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import com.mld46.oponn.BoardState;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama (java.lang.foreign, JDK 22+) binding of liborx.so (include/orx.h): the GPU replacement of the
 * CORES ModellingBoards built in Oponn.setup() (Oponn.java:104-110) and of what MCTSTree.simulation /
 * backpropagation do with them (MCTSTree.java:359-431).
 *
 * NOT COMPILED in the build image (no JDK, Lux SDK not vendored): reference-side glue for a maintainer.
 * Struct layouts (x86-64 / aarch64 LP64):
 *   OrxMap   { i32 n_territories, n_continents, n_players, reserved; ptr adj_offsets, adj, continent,
 *              continent_bonus, card_symbol }                                    56 bytes
 *   OrxRules { i32 use_cards, transfer_cards, immediate_cash, continent_increase; ptr card_progression;
 *              i32 max_player_turns, reserved }                                  32 bytes
 *   OrxLeafResult = 20 x i64, OrxGameResult = 12 x i32 (include/orx_state.h)
 */
public final class Orx implements AutoCloseable
{
	private static final Linker L = Linker.nativeLinker();
	private static final SymbolLookup LIB = SymbolLookup.libraryLookup("liborx.so", Arena.global());

	private static MethodHandle h(String name, FunctionDescriptor d)
	{
		return L.downcallHandle(LIB.find(name).orElseThrow(), d);
	}

	// int orx_create(const OrxMap*, const OrxRules*, int device, OrxCtx**)
	private static final MethodHandle CREATE = h("orx_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
	private static final MethodHandle DESTROY = h("orx_destroy", FunctionDescriptor.ofVoid(ADDRESS));
	private static final MethodHandle STRIDE = h("orx_leaf_stride", FunctionDescriptor.of(JAVA_INT, ADDRESS));
	private static final MethodHandle LAST_ERROR = h("orx_last_error", FunctionDescriptor.of(ADDRESS));
	// int orx_set_sim_agents(OrxCtx*, const int32_t* agent_ids, int modelling_turn_limit)
	private static final MethodHandle SET_AGENTS = h("orx_set_sim_agents", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
	// int orx_rollout_batch(ctx, leaves, n_leaves, playouts_per_leaf, seed, first_game, OrxLeafResult*, OrxGameResult*)
	private static final MethodHandle ROLLOUT = h("orx_rollout_batch",
			FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS));

	/** ORX_AGENT_* of include/orx_state.h == SimAgent.getSimAgent(name) (SimAgent.java:248-307) */
	public static int agentId(String luxAgentName)
	{
		String n = luxAgentName == null ? "" : luxAgentName.toLowerCase();
		if(n.equals("angry")) return 1;
		if(n.equals("stinky")) return 2;
		return 0;   // SimRandom; also every agent this engine does not model yet
	}

	private final Arena arena = Arena.ofShared();
	private final MemorySegment ctx;
	public final int stride;

	/** Replaces `new ModellingBoard(boardState, simAgents, cardManager, i, 2)` x CORES (Oponn.java:106-110). */
	public Orx(BoardState bs, int device) throws Throwable
	{
		int n = bs.numberOfCountries;
		int[] off = new int[n + 1];
		java.util.ArrayList<Integer> adj = new java.util.ArrayList<>();
		int[] cont = new int[n];
		for(int cc = 0; cc < n; cc++)
		{
			for(var nb : bs.countries[cc].getAdjoiningList()) adj.add(nb.getCode());   // list ORDER is semantic
			off[cc + 1] = adj.size();
			cont[cc] = bs.countries[cc].getContinent();
		}
		MemorySegment map = arena.allocate(56, 8), rules = arena.allocate(32, 8), out = arena.allocate(ADDRESS);
		map.set(JAVA_INT, 0, n);
		map.set(JAVA_INT, 4, bs.numberOfContinents);
		map.set(JAVA_INT, 8, bs.numberOfPlayers);
		map.set(ADDRESS, 16, arena.allocateFrom(JAVA_INT, off));
		map.set(ADDRESS, 24, arena.allocateFrom(JAVA_INT, adj.stream().mapToInt(i -> i).toArray()));
		map.set(ADDRESS, 32, arena.allocateFrom(JAVA_INT, cont));
		map.set(ADDRESS, 40, arena.allocateFrom(JAVA_INT, bs.initialContinentBonuses));
		map.set(ADDRESS, 48, MemorySegment.NULL);                 // card symbols: code % 3 (CardManager.java:54-61)
		rules.set(JAVA_INT, 0, bs.useCards ? 1 : 0);
		rules.set(JAVA_INT, 4, bs.transferCards ? 1 : 0);
		rules.set(JAVA_INT, 8, bs.immediateCash ? 1 : 0);
		rules.set(JAVA_INT, 12, bs.continentIncrease);
		rules.set(ADDRESS, 16, arena.allocateFrom(bs.cardProgression));
		rules.set(JAVA_INT, 24, 0);                               // max_player_turns: default
		int rc = (int) CREATE.invokeExact(map, rules, device, out);
		if(rc != 0) throw new IllegalStateException(error());     // there is no CPU fallback
		ctx = out.get(ADDRESS, 0);
		stride = (int) STRIDE.invokeExact(ctx);
	}

	/** Oponn.createSimAgents (Oponn.java:146-162): the modelled line-up and MODELLING_TURN_LIMIT (Oponn.java:55). */
	public void setSimAgents(int[] agentIds, int modellingTurnLimit) throws Throwable
	{
		int rc = (int) SET_AGENTS.invokeExact(ctx, arena.allocateFrom(JAVA_INT, agentIds), modellingTurnLimit);
		if(rc != 0) throw new IllegalArgumentException(error());
	}

	/** setupState + simulate + the two getters for nLeaves x playouts playouts; result[i] = the 20 longs of
	 *  OrxLeafResult { n, lenSum, wins[8], winLenSum[8], flagged, reserved }. */
	public long[][] rollout(MemorySegment packedLeaves, int nLeaves, int playouts, long seed, long firstGame) throws Throwable
	{
		MemorySegment agg = arena.allocate(160L * nLeaves, 8);
		int rc = (int) ROLLOUT.invokeExact(ctx, packedLeaves, nLeaves, playouts, seed, firstGame, agg, MemorySegment.NULL);
		if(rc != 0) throw new IllegalArgumentException(error());   // the reference throws on illegal states too
		long[][] r = new long[nLeaves][20];
		for(int i = 0; i < nLeaves; i++)
			for(int j = 0; j < 20; j++) r[i][j] = agg.get(JAVA_LONG, 160L * i + 8L * j);
		return r;
	}

	/** The same call with per-playout records: games = nLeaves * playouts x OrxGameResult (48 bytes each:
	 *  playerOutOnTurn[8], simTurnCount, flags, streamPos, boardHash). */
	public void rolloutGames(MemorySegment packedLeaves, int nLeaves, int playouts, long seed, long firstGame,
			MemorySegment games) throws Throwable
	{
		int rc = (int) ROLLOUT.invokeExact(ctx, packedLeaves, nLeaves, playouts, seed, firstGame, MemorySegment.NULL, games);
		if(rc != 0) throw new IllegalArgumentException(error());
	}

	private static String error() throws Throwable
	{
		return ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(512).getString(0);
	}

	@Override
	public void close()
	{
		try { DESTROY.invokeExact(ctx); } catch(Throwable t) { }
		arena.close();
	}
}
