"""``Oponn.txt``: the four settings the reference reads at the start of a game (Oponn.readSettings,
src/com/mld46/oponn/Oponn.java:464-533), parsed with the reference's own rules so that an existing file keeps working.

    lux_agent_modelling=true|false   opponent modelling (ModellingBoard) on unless the value is exactly "false"
    iterations=<int>                 playouts per decision (default 1000; an unparsable value keeps the default)
    cores=<int>|n                    the reference's playouts per leaf: min(cores, available), at least 1; "n" = all
    debug=true|false                 Debug.DEBUGGING (no effect here)

Lines without "=" are ignored; keys are matched with ``startswith``; spaces inside the value are dropped.
:func:`search_options` maps them onto the engine: ``iterations / cores`` leaves per move with ``cores`` playouts each is
exactly ``MCTSTree``'s ``loops = iterations / cores`` (MCTSTree.java:69).
"""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class OponnSettings:
    iterations: int = 1000
    opponent_modelling: bool = True
    debugging: bool = False
    cores: int = os.cpu_count() or 1


def read_settings(path: str, available_cores: int = 0) -> OponnSettings:
    s = OponnSettings()
    if available_cores > 0:
        s.cores = available_cores
    try:
        lines = open(path, encoding="utf-8", errors="replace").read().splitlines()
    except OSError:
        return s                                   # "Error reading settings file": the defaults stand
    for line in lines:
        if "=" not in line:
            continue
        value = line[line.index("=") + 1:].replace(" ", "")
        if line.startswith("lux_agent_modelling"):
            s.opponent_modelling = value != "false"
        elif line.startswith("iterations"):
            try:
                s.iterations = java_int(value)
            except ValueError:
                pass
        elif line.startswith("cores"):
            if value != "n":
                try:
                    s.cores = max(1, min(java_int(value), s.cores))
                except ValueError:
                    pass
        elif line.startswith("debug"):
            s.debugging = value == "true"
    return s


def java_int(text: str) -> int:
    """Integer.valueOf: optional sign and decimal digits only, within int range"""
    t = text[1:] if text[:1] in "+-" else text
    if not t or not t.isascii() or not t.isdigit():
        raise ValueError(text)
    v = int(text)
    if not -(1 << 31) <= v < (1 << 31):
        raise ValueError(text)
    return v


def search_options(s: OponnSettings) -> dict:
    """the engine-side meaning: leaves per move, playouts per leaf, modelling turn limit (Oponn.java:55)"""
    ppl = max(1, s.cores)
    return {"iterations": s.iterations // ppl, "playouts_per_leaf": ppl, "modelling_turn_limit": 2 if s.opponent_modelling else 0}
