"""Hand-derived playouts: each scenario is (map, rules, leaf, words, expected result fields).  The expectations
were worked out by hand from the reference's source (citations in the comments), not produced by the oracle;
they pin oracle and CUDA engine alike.  Table rows come from battlesim_ref.numpy_outcome_table (an independent
restatement of BattleSim.generateAttackOutcomes)."""
import numpy as np

from battlesim_ref import numpy_outcome_table
from oponn_b200.state import (AGENT_ANGRY, AGENT_RANDOM, AGENT_STINKY, ATTACK_EXECUTE, CARD_WILD, PHASE_ATTACK, PHASE_CARDS,
                              PHASE_COUNTRY_SELECTION, PHASE_FORTIFICATION, PHASE_PLACEMENT, BoardState, RiskMap, Rules)
from scripted import Script, board_hash


def line4(n_players=2):
    # 0 - 1 - 2 - 3 ; continents {0,1} bonus 2 and {2,3} bonus 3
    return RiskMap(name="line4", n_players=n_players, adjacency=[[1], [0, 2], [1, 3], [2]], continent=[0, 0, 1, 1],
                   continent_bonus=[2, 3])


def first_row(a, d):
    p, al, dl, inv = numpy_outcome_table(a, d)[0]
    return al, dl, inv


def scenario_sweep():
    """PLACEMENT catch-up, two conquests, elimination, drain of the attackers list."""
    m, r = line4(), Rules(use_cards=False)
    leaf = BoardState(owner=[0, 0, 1, 1], armies=[1, 5, 2, 1], currentPlayer=0, currentPhase=PHASE_PLACEMENT,
                      numberOfPlaceableArmies=3, playerNumberOfCards=[0, 0])
    assert first_row(7, 2) == (0, 2, True) and first_row(4, 1) == (0, 1, True)
    s = Script()
    s.int(3, 2).int(1, 0)          # SimRandom.place: k = 1 + 2 = 3 on candidates[0] = territory 1  -> armies[1] = 8
    s.int(1, 0).int(1, 0)          # attackers [1] -> A = 1 ; defenders of 1 (adjacency order [0, 2], enemy only) = [2]
    s.double_k(0)                  # battle 7 v 2: r = 0 -> first row (0, 2): conquered
    s.int(4, 2)                    # moveArmiesIn: avail 7, min 3, 3 + nextInt(4) = 5 -> armies [1, 3, 5, 1]
    s.int(2, 1).int(1, 0)          # attackers [1, 2] -> A = 2 ; defenders [3]
    s.double_k(0)                  # battle 4 v 1 -> (0, 1): conquered; avail = 4, min 3, 3 + nextInt(1)
    s.int(1, 0)                    #   -> 3 move in: armies [1, 3, 2, 3]; player 1 has no territory: defeated
    term = len(s.words)            # playerOutOnTurn[0] = -1 is written here (SimulationBoard.java:862-866)
    s.int(3, 0).int(2, 0).int(1, 0)  # the phase still drains [1, 2, 3]: no defenders anywhere -> three removals
    exp = dict(pot=[-1, 0], turns=0, flags=0, stream_pos=term, hash=board_hash([0, 0, 0, 0], [1, 3, 2, 3]))
    return m, r, leaf, s.array(pad=4), exp


def scenario_catchup_execute_keeps_owner():
    """a leaf in ATTACK/EXECUTE that conquers runs executeInvasion without setupInvasion
    (SimulationBoard.java:402-411): the defender keeps the territory and receives the invading armies."""
    m, r = line4(), Rules(use_cards=False, max_player_turns=1)
    leaf = BoardState(owner=[0, 0, 1, 1], armies=[1, 8, 2, 3], currentPlayer=0, currentPhase=PHASE_ATTACK,
                      attackState=ATTACK_EXECUTE, attackingCC=1, defendingCC=2, defendingPlayer=1, attackUntilDead=True,
                      playerNumberOfCards=[0, 0])
    al, dl, inv = first_row(2, 5)
    assert (al, dl, inv) == (2, 0, False)
    s = Script()
    s.double_k(0)                  # executeAttack 7 v 2 -> (0, 2): armies[2] = 0 -> INVADE
    s.int(4, 2)                    # moveArmiesIn 3 + 2 = 5: armies [1, 3, 5, 3], owner[2] is STILL player 1
    s.int(1, 0).int(1, 0)          # attackPhase: attackers [1] (3 armies, enemy neighbour 2); defenders [2]
    s.double_k(0)                  # battle 2 v 5 -> first row (2, 0): attackers destroyed -> armies[1] = 1, list empty
    exp = dict(pot=[0, 0], turns=0, flags=2, stream_pos=len(s.words), hash=board_hash([0, 0, 1, 1], [1, 1, 5, 3]))
    return m, r, leaf, s.array(pad=4), exp


def scenario_cards_and_income():
    """CARDS catch-up with a set: cash decision draw, random set, +2 on owned card territories, progression value,
    then PLACEMENT income = max(n/3, 3) + owned continents, one attack, card deal, step limit after the turn."""
    m, r = line4(), Rules(use_cards=True, transfer_cards=True, max_player_turns=1)
    # hand: cards of territories 0, 3 and a wildcard (symbols 0, 0, wild) -> exactly one triple, a set
    leaf = BoardState(owner=[0, 0, 1, 1], armies=[1, 1, 30, 9], currentPlayer=0, currentPhase=PHASE_CARDS,
                      playerNumberOfCards=[0, 0], cardProgressionPos=3, hand=[0, 3, CARD_WILD], deck=[1, 2, CARD_WILD])
    al, dl, inv = first_row(13, 30)
    s = Script()
    s.double(0.25)                 # SimRandom.cardsPhase: 3 cards, r < 0.5 -> cash
    s.int(1, 0)                    # Card.getRandomSet: one valid triple
    # cashCards: value(pos 3) = 8 for "4, 6, 8, 10, 15, 20..."; territory 0 is owned -> +2; territory 3 is not
    #   armies [3, 1, 30, 9], placeable 8, deck [1, 2, W, 0, 3, W]
    # PLACEMENT: income = max(2 / 3, 3) + bonus of continent 0 (owned: 2) = 5 -> placeable 13
    s.int(13, 12).int(1, 0)        # place all 13 on the only border territory (1): armies [3, 14, 30, 9]
    s.int(1, 0).int(1, 0)          # attackers [1] (territory 0 has no enemy neighbour); defenders [2]
    s.double_k(0)                  # battle 13 v 30 -> first row: the attackers die
    armies = [3, 14 - al, 30 - dl, 9]
    owner = [0, 0, 1, 1]
    assert not inv or armies[2] == 0
    if inv:                        # conquest: move in 3 + nextInt(avail - 3)
        avail = armies[1] - 1
        s.int(avail - 3, 0)
        armies[2] = 3; armies[1] -= 3; owner[2] = 0
        # attackers list: [1] stays if armies > 1, territory 2 appended (3 > 1): continue until nothing can attack
        raise AssertionError("scenario expects the most likely 13 v 30 row to be a defeat of the attacker")
    # attacker wiped out to 1 army: list empty; no conquest -> no card dealt; FORTIFICATION -> player 1, step limit
    exp = dict(pot=[0, 0], turns=0, flags=2, stream_pos=len(s.words), hash=board_hash(owner, armies))
    return m, r, leaf, s.array(pad=4), exp


def scenario_overflow_crowns_the_new_current_player():
    """SURVEY 8.1 quirks 5 and 6.  A FORTIFICATION leaf above 100 000 armies: tearDownFortificationPhase advances
    currentPlayer FIRST (SimulationBoard.java:972-978), then checkForOverflow (:984-1015) defeats everybody but the
    largest holder "by" that holder, and playerDefeated writes the winner flag to currentPlayer (:862-866) -- the player
    who has just been defeated.  Player 2 owns nothing at the leaf: never alive in the playout, its entry stays 0.
    No random draw at all."""
    m, r = line4(3), Rules(use_cards=False)
    leaf = BoardState(owner=[0, 0, 1, 1], armies=[60000, 1, 50000, 1], currentPlayer=0, currentPhase=PHASE_FORTIFICATION,
                      playerNumberOfCards=[0, 0, 0])
    s = Script()
    # getNextPlayer: 1 (not a wrap: simTurnCount stays 0); totals 60001 / 50001 / 0 -> 110002 > 100000, maxPlayer = 0;
    # playerDefeated(1, 0): remainingPlayers 2 -> 1, playerOutOnTurn[1] = 0, then playerOutOnTurn[currentPlayer = 1] = -1
    exp = dict(pot=[0, -1, 0], turns=0, flags=0, stream_pos=0, hash=board_hash([0, 0, 1, 1], [60000, 1, 50000, 1]))
    return m, r, leaf, s.array(pad=4), exp


def scenario_invade_leaf_counts_the_vanquished_player():
    """An ATTACK / INVADE leaf whose conquered country was the defender's last: the tree has already moved ownership
    (ExplorationBoard.java:174-177), the record carries the country at 0 armies, and setupPhaseGameState bumps
    remainingPlayers (SimulationBoard.java:339-342) so that finishInvasion can defeat the player and crown the attacker."""
    from oponn_b200.state import ATTACK_INVADE
    m, r = line4(), Rules(use_cards=False)
    leaf = BoardState(owner=[0, 0, 0, 0], armies=[1, 5, 0, 1], currentPlayer=0, currentPhase=PHASE_ATTACK, attackState=ATTACK_INVADE,
                      attackingCC=1, defendingCC=2, defendingPlayer=1, playerNumberOfCards=[0, 0])
    s = Script()
    s.int(1, 0)                    # moveArmiesIn: avail 4, min 3, 3 + nextInt(1) = 3 -> armies [1, 2, 3, 1]
    term = len(s.words)            # finishInvasion: playerCountries[1] == 0 -> defeated, playerOutOnTurn[0] = -1
    # attackPhase: nobody has an enemy neighbour -> empty list; no cards; FORTIFICATION is never played
    exp = dict(pot=[-1, 0], turns=0, flags=0, stream_pos=term, hash=board_hash([0, 0, 0, 0], [1, 2, 3, 1]))
    return m, r, leaf, s.array(pad=4), exp


def scenario_single_roll_catchup():
    """a board whose attackUntilDead bit is still false resolves the pending EXECUTE attack with ONE
    roll of min(a, 3) v min(d, 2) dice (SimulationBoard.java:758-780, BattleSim.java:152-162), not with a battle."""
    m, r = line4(), Rules(use_cards=False, max_player_turns=1)
    leaf = BoardState(owner=[0, 0, 1, 1], armies=[1, 2, 2, 3], currentPlayer=0, currentPhase=PHASE_ATTACK, attackState=ATTACK_EXECUTE,
                      attackingCC=1, defendingCC=2, defendingPlayer=1, attackUntilDead=False, playerNumberOfCards=[0, 0])
    assert first_row(1, 1) == (1, 0, False)
    s = Script()
    s.double_k(0)                  # 1 v 2 dice, r = 0 <= 0.2546: the attacker loses 0, the defender min(1, 2) - 0 = 1 -> armies[2] = 1
    s.int(1, 0).int(1, 0)          # attackPhase: attackers [1] (2 armies); defenders [2]
    s.double_k(0)                  # battle 1 v 1 till dead: first row = the attacker dies -> armies[1] = 1, list empty
    # no conquest; FORTIFICATION -> player 1; one player-turn played: step limit
    exp = dict(pot=[0, 0], turns=0, flags=2, stream_pos=len(s.words), hash=board_hash([0, 0, 1, 1], [1, 1, 1, 3]))
    return m, r, leaf, s.array(pad=4), exp


def scenario_gaussian_battle_in_a_game():
    """Both sides above 500: the closed form with one nextGaussian() (BattleSim.java:71-93).  The polar method is fed
    nextDouble() = 0.5 and 0.75: v1 = 0, v2 = 0.5, s = 0.25 -> the first gaussian is exactly 0, variance = (int)(0 * sqrt) = 0.
    600 v 600: matched = (int)(0.922 * 600) = 553 <= 600 -> all 600 defenders die, the attacker loses min(0 + 553, 599)."""
    m, r = line4(), Rules(use_cards=False)
    leaf = BoardState(owner=[0, 0, 1, 1], armies=[1, 601, 600, 1], currentPlayer=0, currentPhase=PHASE_ATTACK, playerNumberOfCards=[0, 0])
    assert first_row(2, 1) == (0, 1, True)
    s = Script()
    s.int(1, 0).int(1, 0)          # attackers [1]; defenders [2]
    s.double(0.5).double(0.75)     # nextGaussian: one accepted pair, value 0
    s.int(44, 0)                   # conquered: armies[1] = 601 - 553 = 48, avail 47, 3 + nextInt(44) = 3 -> armies [1, 45, 3, 1]
    s.int(2, 1).int(1, 0)          # attackers [1, 2] -> A = 2; defenders [3]
    s.double_k(0)                  # 2 v 1 -> (0, 1): conquered; avail 2 = min -> both move: armies [1, 45, 1, 2]; player 1 is out
    term = len(s.words)
    s.int(2, 0).int(1, 0)          # list [1, 3] (2 left with one army, 3 appended): both have no defender -> removed
    exp = dict(pot=[-1, 0], turns=0, flags=0, stream_pos=term, hash=board_hash([0, 0, 0, 0], [1, 45, 1, 2]))
    return m, r, leaf, s.array(pad=4), exp


def scenario_single_rolls_above_100_in_a_game():
    """One side above 100: single rolls until a side is out or both fit the tables (BattleSim.java:99-116).  101 v 1 is
    one roll of 3 v 1 dice; r = 0 <= 0.6597: the defender loses its army, no table draw follows."""
    m, r = line4(), Rules(use_cards=False)
    leaf = BoardState(owner=[0, 0, 1, 1], armies=[1, 102, 1, 1], currentPlayer=0, currentPhase=PHASE_ATTACK, playerNumberOfCards=[0, 0])
    s = Script()
    s.int(1, 0).int(1, 0)          # attackers [1]; defenders [2]
    s.double_k(0)                  # the single roll
    s.int(98, 0)                   # conquered: avail 101, 3 + nextInt(98) = 3 -> armies [1, 99, 3, 1]
    s.int(2, 1).int(1, 0)          # A = 2; defenders [3]
    s.double_k(0)                  # 2 v 1 -> conquered, both armies move: [1, 99, 1, 2]; player 1 is out
    term = len(s.words)
    s.int(2, 0).int(1, 0)
    exp = dict(pot=[-1, 0], turns=0, flags=0, stream_pos=term, hash=board_hash([0, 0, 0, 0], [1, 99, 1, 2]))
    return m, r, leaf, s.array(pad=4), exp


def line24():
    n = 24
    return RiskMap(name="line24", n_players=2, adjacency=[[x for x in (t - 1, t + 1) if 0 <= x < n] for t in range(n)],
                   continent=[0] * n, continent_bonus=[7])


def scenario_big_hand_forced_cashing_and_the_deal():
    """SURVEY 8.1 quirks 8, 13, 15 on a 24-territory line, deals scripted through the game stream (ORX_DEAL_GAME_STREAM).
    An 8-card hand of one symbol (codes 0, 3, ... 21: code % 3 == 0, every triple is a set): SimRandom.cardsPhase always
    cashes with 5 or more cards, tearDownCardsPhase keeps cashing until 4 or fewer are left (SimulationBoard.java:620-638);
    each cashed card whose country the player owns is worth +2 armies there (:605-616); a conquest earns one card at the
    end of the attack phase, drawn from the sim deck the cashed cards went back to (:884-896, CardManager.java:250-276)."""
    m, r = line24(), Rules(use_cards=True, transfer_cards=True, card_progression="0", max_player_turns=1)
    hand = [0, 3, 6, 9, 12, 15, 18, 21]
    deck = [c for c in range(24) if c not in hand] + [CARD_WILD, CARD_WILD]
    leaf = BoardState(owner=[0] * 9 + [1] * 15, armies=[1] * 24, currentPlayer=0, currentPhase=PHASE_CARDS, playerNumberOfCards=[8, 0],
                      cardProgressionPos=1, hand=hand, deck=deck, dealFromStream=True)
    assert first_row(3, 1) == (0, 1, True) and first_row(2, 1) == (0, 1, True) and first_row(1, 1) == (1, 0, False)
    s = Script()
    s.double(0.9)                  # cardsPhase draws r even though 8 cards always cash (SimRandom.java:43-44)
    s.int(56, 0)                   # getRandomSet: C(8,3) = 56 valid triples, the first is cards 0, 3, 6 -> +2 on countries 0, 3, 6 (owned)
    s.int(10, 9)                   # forced: 5 cards left, C(5,3) = 10 triples, the last is cards 15, 18, 21 (not owned: no bonus)
    # hand [9, 12]; progression "0": both sets are worth 0; income = max(9 / 3, 3) = 3, the continent is not owned
    s.int(3, 2).int(1, 0)          # place 1 + 2 = 3 on the only border country, 8: armies[8] = 4
    s.int(1, 0).int(1, 0)          # attackers [8] (0, 3 and 6 hold 3 armies but touch no enemy); defenders of 8: [9]
    s.double_k(0)                  # 3 v 1 -> conquered; avail 3 = min: all three move, armies[8] = 1, armies[9] = 3
    s.int(1, 0).int(1, 0)          # list [9]; defenders [10]
    s.double_k(0)                  # 2 v 1 -> conquered; both move: armies[9] = 1, armies[10] = 2
    s.int(1, 0).int(1, 0)          # list [10]; defenders [11]
    s.double_k(0)                  # 1 v 1: the likelier row is the attacker's loss: armies[10] = 1, list empty
    s.int(24, 23)                  # tearDownAttackPhase: one card from the 18 + 3 + 3 = 24 of the deck: the last one, card 21
    armies = [3, 1, 1, 3, 1, 1, 3] + [1] * 17
    exp = dict(pot=[0, 0], turns=0, flags=2, stream_pos=len(s.words), hash=board_hash([0] * 11 + [1] * 13, armies))
    return m, r, leaf, s.array(pad=4), exp


ALL = [scenario_sweep, scenario_catchup_execute_keeps_owner, scenario_cards_and_income, scenario_overflow_crowns_the_new_current_player,
       scenario_invade_leaf_counts_the_vanquished_player, scenario_single_roll_catchup, scenario_gaussian_battle_in_a_game,
       scenario_single_rolls_above_100_in_a_game, scenario_big_hand_forced_cashing_and_the_deal]


# ---- modelled opponents: (map, rules, leaf, words, expected, agent ids, modelling turn limit) --------

def scenario_angry_sweep():
    """SimAngry: placement on the country with most enemy neighbours (SimAngry.java:126-151), attack passes over an
    ArmiesIterator that holds one element ahead (:157-195), moveArmiesIn by enemy-neighbour counts (:199-213)."""
    m, r = line4(), Rules(use_cards=False)
    leaf = BoardState(owner=[0, 0, 1, 1], armies=[4, 2, 3, 1], currentPlayer=0, currentPhase=PHASE_PLACEMENT,
                      numberOfPlaceableArmies=3, playerNumberOfCards=[0, 0])
    assert first_row(4, 3) == (0, 3, True) and first_row(3, 1) == (0, 1, True)
    s = Script()
    # placeArmies(3): country 0 has 0 enemy neighbours, country 1 has 1 -> all 3 on country 1: armies [4, 5, 3, 1]
    # pass 1: iterator holds 0 then 1 (then nothing: 2 and 3 are not ours when 1 is handed out)
    #   us = 0: no enemy neighbour.  us = 1: weakest enemy neighbour 2 (3 armies), 5 > 3 -> attack till death
    s.double_k(0)                  # battle 4 v 3 -> first row (0, 3): conquered
    #   moveArmiesIn(1, 2): enemies of 1 = 0, enemies of 2 = 1 (country 3) -> not more at home -> move all 4
    #   armies [4, 1, 4, 1], owner [0, 0, 0, 1]
    # pass 2: holds 0, then 2 (1 has a single army).  us = 2: weakest 3 (1 army), 4 > 1 -> attack
    s.double_k(0)                  # battle 3 v 1 -> (0, 1): conquered; moveArmiesIn(2, 3): 0 v 0 enemies -> all 3
    #   armies [4, 1, 1, 3]; player 1 has no country left -> playerOutOnTurn = [-1, 0], stream position pinned here
    term = len(s.words)
    # pass 3: nothing to attack, madeAttack stays false
    exp = dict(pot=[-1, 0], turns=0, flags=0, stream_pos=term, hash=board_hash([0, 0, 0, 0], [4, 1, 1, 3]))
    return m, r, leaf, s.array(pad=4), exp, [AGENT_ANGRY, AGENT_RANDOM], 2


def scenario_angry_stinky_then_random():
    """SimAngry.fortifyPhase (SimAngry.java:215-279) from a FORTIFICATION leaf, two SimStinky turns (SimStinky.java:66-125),
    and ModellingBoard's switch to SimRandom at the second round wrap (ModellingBoard.java:29-47)."""
    m, r = line4(), Rules(use_cards=False, max_player_turns=5)
    leaf = BoardState(owner=[0, 0, 0, 1], armies=[5, 1, 2, 9], currentPlayer=0, currentPhase=PHASE_FORTIFICATION,
                      playerNumberOfCards=[0, 0])
    assert first_row(11, 3) == (0, 3, True) and first_row(16, 9) == (5, 9, True) and first_row(3, 11) == (3, 0, False)
    s = Script()
    # catch-up fortifyPhase of player 0 (Angry), moveable = [5, 1, 2, 0]:
    #   i = 0: its only neighbour 1 has no enemy neighbour, neither has 0 -> random neighbour
    s.int(1, 0)                    #   -> country 1; fortify min(5, moveable 5, armies - 1 = 4) = 4: armies [1, 5, 2, 9]
    #   i = 1 (moveable 1): neighbour 2 has 1 enemy neighbour > 0 of its own -> fortify 1: armies [1, 4, 3, 9]
    #   i = 2 (moveable 2): no better neighbour, has an enemy neighbour itself -> stays
    # player 1 (Stinky): income max(1 / 3, 3) = 3
    s.int(4, 0).int(4, 3)          # placeArmies: country 0 is not ours -> redraw; country 3 borders an enemy: armies [1, 4, 3, 12]
    s.double_k(0)                  # attackPhase(true): us = 3, weak = 2 (3 armies), 12 > 4.5 -> battle 11 v 3 -> (0, 3)
    #   moveArmiesIn(3, 2): country 3 has no enemy neighbour any more -> 1000000 -> clamped to 11: armies [1, 4, 11, 1]
    # wrap to player 0: simTurnCount 0 -> 1, no switch yet (0 + 1 != 2)
    # player 0 (Angry): income 3 + continent 0 (bonus 2) = 5 on country 1 (1 enemy neighbour): armies [1, 9, 11, 1]
    #   attack: us = 1, weakest 2 (11 armies), 9 > 11 is false.  fortify: 0 -> 1 moves min(1, 1, 0) = 0; nothing else
    # player 1 (Stinky): income 3 + continent 1 (bonus 3) = 6
    s.int(4, 2)                    # placeArmies: country 2 borders country 1: armies [1, 9, 17, 1]
    s.double_k(0)                  # us = 2 (holds 3 next), weak = 1 (9 armies), 17 > 13.5 -> battle 16 v 9 -> (5, 9)
    #   conquered with 12 left; moveArmiesIn(2, 1): no enemy neighbour of 2 -> all 11: armies [1, 11, 1, 1], owner [0, 1, 1, 1]
    #   us = 3: no enemy neighbour.  13 armies in total, not > 300: no second pass
    # wrap to player 0: simTurnCount 1, 1 + 1 == 2 -> everybody is SimRandom from here on; simTurnCount = 2
    s.int(3, 2).int(1, 0)          # SimRandom.place: income 3, 1 + 2 = 3 armies on candidates[0] = country 0: armies [4, 11, 1, 1]
    s.int(1, 0).int(1, 0)          # SimRandom.attackPhase: attackers [0], defenders [1]
    s.double_k(0)                  # battle 3 v 11 -> (3, 0): attackers destroyed: armies [1, 11, 1, 1]
    # fortification -> player 1, fifth player-turn: step limit
    exp = dict(pot=[0, 0], turns=2, flags=2, stream_pos=len(s.words), hash=board_hash([0, 1, 1, 1], [1, 11, 1, 1]))
    return m, r, leaf, s.array(pad=4), exp, [AGENT_ANGRY, AGENT_STINKY], 2


def scenario_angry_draft():
    """SimAngry.pickCountry (SimAngry.java:49-105) with MapHelper's smallest positive empty / open continent
    (MapHelper.java:132-191), against a SimRandom drafter; first initial placements, then the step limit."""
    m, r = line4(), Rules(use_cards=False, max_player_turns=2)
    leaf = BoardState(owner=[-1, -1, -1, -1], armies=[1, 1, 1, 1], currentPlayer=0, currentPhase=PHASE_COUNTRY_SELECTION,
                      playerNumberOfCards=[0, 0])
    s = Script()
    # player 0 (Angry): both continents empty, size 2, positive bonus -> continent 0; no neighbour of ours there ->
    #   fewest neighbours: country 0 (1 neighbour)
    s.int(3, 1)                    # player 1 (SimRandom.pickCountry): unowned [1, 2, 3] -> country 2
    # player 0: no continent is empty any more; smallest open one = continent 0; country 1 neighbours our country 0
    s.int(1, 0)                    # player 1: unowned [3]
    # INITIAL_PLACEMENT: 22 starting armies each (N = 4, P = 2), 2 countries taken -> 20 left, 4 per turn
    #   player 0 (Angry): country 1 has the enemy neighbour -> armies [1, 5, 1, 1]
    s.int(4, 3).int(1, 0)          #   player 1 (SimRandom.place): 1 + 3 = 4 armies on candidates[0] = country 2
    exp = dict(pot=[0, 0], turns=0, flags=2, stream_pos=len(s.words), hash=board_hash([0, 0, 1, 1], [1, 5, 5, 1]))
    return m, r, leaf, s.array(pad=4), exp, [AGENT_ANGRY, AGENT_RANDOM], 2


def scenario_stinky_cannot_draft():
    """SimStinky.pickCountry returns -1 (SimStinky.java:51-54); SimulationBoard.selectCountry prints "Err..." and then
    indexes countries[-1] (SimulationBoard.java:485-494): the playout dies with an exception -> ORX_FLAG_ERROR."""
    m, r = line4(), Rules(use_cards=False)
    leaf = BoardState(owner=[-1, -1, -1, -1], armies=[1, 1, 1, 1], currentPlayer=0, currentPhase=PHASE_COUNTRY_SELECTION,
                      playerNumberOfCards=[0, 0])
    exp = dict(pot=[0, 0], turns=0, flags=1, stream_pos=0, hash=0)
    return m, r, leaf, Script().array(pad=4), exp, [AGENT_STINKY, AGENT_RANDOM], 2


MODELLED = [scenario_angry_sweep, scenario_angry_stinky_then_random, scenario_angry_draft, scenario_stinky_cannot_draft]


def check(games_row, exp, n_players):
    assert [int(x) for x in games_row["player_out_on_turn"][:n_players]] == exp["pot"]
    assert int(games_row["sim_turn_count"]) == exp["turns"]
    assert int(games_row["flags"]) == exp["flags"]
    assert int(games_row["stream_pos"]) == exp["stream_pos"]
    assert int(games_row["board_hash"]) == exp["hash"]
