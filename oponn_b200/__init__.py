"""oponn_b200 -- B200-native rollout engine for Oponn's Monte-Carlo playouts.

The product is ``liborx.so`` (hand-written CUDA for sm_100a behind the C ABI of ``include/orx.h``);
this package is the thin host mirror of the reference's rollout interface over that library.
There is no CPU implementation here: importing :mod:`oponn_b200.engine` without the built library,
or creating an engine without a CUDA device, raises.
"""
from .state import (ATTACK_CASH, ATTACK_EXECUTE, ATTACK_INVADE, ATTACK_PLACE, ATTACK_START, CARD_WILD, FLAG_ERROR,
                    FLAG_STEP_LIMIT, FLAG_STREAM_END, GAME_RESULT_DTYPE, LEAF_RESULT_DTYPE, OWNER_NONE,
                    PHASE_ATTACK, PHASE_CARDS, PHASE_COUNTRY_SELECTION, PHASE_FORTIFICATION,
                    PHASE_INITIAL_PLACEMENT, PHASE_PLACEMENT, BoardState, RiskMap, Rules, leaf_layout, pack_leaves)
from .maps import classic42, synth200

__all__ = [n for n in dir() if not n.startswith("_")]
