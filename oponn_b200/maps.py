"""Maps used by the tests and the bench.

``classic42`` is the standard 42-territory Risk board, authored here (the Lux map files are not
part of the reference repository); territory codes are this file's order, which only matters for
card symbols (code % 3, CardManager.java:54-61) and scan order.  Adjacency lists are kept in the
hand-written (unsorted) order on purpose: SimRandom chooses "the j-th enemy neighbour in
getAdjoiningList() order" (SimRandom.java:95-109), so list order is part of the semantics.

``synth200`` is BASELINE.json config 4: a seeded planar-ish graph with ~200 territories,
10 contiguous continents and 8 players.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

from .state import RiskMap

_CLASSIC = [
    # (name, continent, neighbours)
    ("Alaska", 0, ["Northwest Territory", "Alberta", "Kamchatka"]),
    ("Northwest Territory", 0, ["Alaska", "Alberta", "Ontario", "Greenland"]),
    ("Greenland", 0, ["Northwest Territory", "Ontario", "Quebec", "Iceland"]),
    ("Alberta", 0, ["Alaska", "Northwest Territory", "Ontario", "Western United States"]),
    ("Ontario", 0, ["Northwest Territory", "Alberta", "Greenland", "Quebec", "Western United States",
                    "Eastern United States"]),
    ("Quebec", 0, ["Ontario", "Greenland", "Eastern United States"]),
    ("Western United States", 0, ["Alberta", "Ontario", "Eastern United States", "Central America"]),
    ("Eastern United States", 0, ["Ontario", "Quebec", "Western United States", "Central America"]),
    ("Central America", 0, ["Western United States", "Eastern United States", "Venezuela"]),
    ("Venezuela", 1, ["Central America", "Peru", "Brazil"]),
    ("Peru", 1, ["Venezuela", "Brazil", "Argentina"]),
    ("Brazil", 1, ["Venezuela", "Peru", "Argentina", "North Africa"]),
    ("Argentina", 1, ["Peru", "Brazil"]),
    ("Iceland", 2, ["Greenland", "Scandinavia", "Great Britain"]),
    ("Scandinavia", 2, ["Iceland", "Great Britain", "Northern Europe", "Ukraine"]),
    ("Great Britain", 2, ["Iceland", "Scandinavia", "Northern Europe", "Western Europe"]),
    ("Northern Europe", 2, ["Scandinavia", "Great Britain", "Western Europe", "Southern Europe", "Ukraine"]),
    ("Western Europe", 2, ["Great Britain", "Northern Europe", "Southern Europe", "North Africa"]),
    ("Southern Europe", 2, ["Northern Europe", "Western Europe", "Ukraine", "North Africa", "Egypt",
                            "Middle East"]),
    ("Ukraine", 2, ["Scandinavia", "Northern Europe", "Southern Europe", "Ural", "Afghanistan", "Middle East"]),
    ("North Africa", 3, ["Brazil", "Western Europe", "Southern Europe", "Egypt", "East Africa", "Congo"]),
    ("Egypt", 3, ["Southern Europe", "North Africa", "East Africa", "Middle East"]),
    ("East Africa", 3, ["North Africa", "Egypt", "Congo", "South Africa", "Madagascar", "Middle East"]),
    ("Congo", 3, ["North Africa", "East Africa", "South Africa"]),
    ("South Africa", 3, ["Congo", "East Africa", "Madagascar"]),
    ("Madagascar", 3, ["East Africa", "South Africa"]),
    ("Ural", 4, ["Ukraine", "Siberia", "Afghanistan", "China"]),
    ("Siberia", 4, ["Ural", "Yakutsk", "Irkutsk", "Mongolia", "China"]),
    ("Yakutsk", 4, ["Siberia", "Kamchatka", "Irkutsk"]),
    ("Kamchatka", 4, ["Alaska", "Yakutsk", "Irkutsk", "Mongolia", "Japan"]),
    ("Irkutsk", 4, ["Siberia", "Yakutsk", "Kamchatka", "Mongolia"]),
    ("Mongolia", 4, ["Siberia", "Kamchatka", "Irkutsk", "Japan", "China"]),
    ("Japan", 4, ["Kamchatka", "Mongolia"]),
    ("Afghanistan", 4, ["Ukraine", "Ural", "China", "Middle East", "India"]),
    ("China", 4, ["Ural", "Siberia", "Mongolia", "Afghanistan", "India", "Siam"]),
    ("Middle East", 4, ["Southern Europe", "Ukraine", "Egypt", "East Africa", "Afghanistan", "India"]),
    ("India", 4, ["Afghanistan", "China", "Middle East", "Siam"]),
    ("Siam", 4, ["China", "India", "Indonesia"]),
    ("Indonesia", 5, ["Siam", "New Guinea", "Western Australia"]),
    ("New Guinea", 5, ["Indonesia", "Western Australia", "Eastern Australia"]),
    ("Western Australia", 5, ["Indonesia", "New Guinea", "Eastern Australia"]),
    ("Eastern Australia", 5, ["New Guinea", "Western Australia"]),
]
_CLASSIC_BONUS = [5, 2, 5, 3, 7, 2]  # N. America, S. America, Europe, Africa, Asia, Australia


def classic42(n_players: int = 6) -> RiskMap:
    code: Dict[str, int] = {name: i for i, (name, _, _) in enumerate(_CLASSIC)}
    adjacency = [[code[x] for x in nb] for (_, _, nb) in _CLASSIC]
    for t, nb in enumerate(adjacency):  # the board is undirected
        for x in nb:
            assert t in adjacency[x], (t, x)
    return RiskMap(name="classic42", n_players=n_players, adjacency=adjacency,
                   continent=[c for (_, c, _) in _CLASSIC], continent_bonus=list(_CLASSIC_BONUS),
                   territory_names=[n for (n, _, _) in _CLASSIC])


def synth200(seed: int = 4, n_players: int = 8, width: int = 20, height: int = 10, n_continents: int = 10) -> RiskMap:
    """Seeded planar-ish map: a width x height lattice with every 4-neighbour border, a random ~47 %
    of the diagonals, continents grown by multi-source BFS from random seeds (contiguous, uneven
    sizes), bonus = ceil(size / 3).  Neighbour lists are shuffled so list order != code order."""
    rng = np.random.default_rng(seed)
    n = width * height
    idx = lambda x, y: y * width + x
    nb: List[set] = [set() for _ in range(n)]
    for y in range(height):
        for x in range(width):
            if x + 1 < width:
                nb[idx(x, y)].add(idx(x + 1, y)); nb[idx(x + 1, y)].add(idx(x, y))
            if y + 1 < height:
                nb[idx(x, y)].add(idx(x, y + 1)); nb[idx(x, y + 1)].add(idx(x, y))
            if x + 1 < width and y + 1 < height and rng.random() < 0.47:
                if rng.random() < 0.5:
                    a, b = idx(x, y), idx(x + 1, y + 1)
                else:
                    a, b = idx(x + 1, y), idx(x, y + 1)
                nb[a].add(b); nb[b].add(a)
    # continents: randomised multi-source BFS
    cont = [-1] * n
    seeds = rng.choice(n, size=n_continents, replace=False)
    frontier = []
    for c, s in enumerate(seeds):
        cont[int(s)] = c
        frontier.append([int(s)])
    remaining = n - n_continents
    while remaining:
        order = rng.permutation(n_continents)
        for c in order:
            f = frontier[c]
            grown = False
            while f and not grown:
                t = f[int(rng.integers(len(f)))]
                free = [x for x in sorted(nb[t]) if cont[x] == -1]
                if not free:
                    f.remove(t)
                    continue
                x = free[int(rng.integers(len(free)))]
                cont[x] = int(c); f.append(x); remaining -= 1; grown = True
    sizes = [cont.count(c) for c in range(n_continents)]
    adjacency = []
    for t in range(n):
        lst = sorted(nb[t])
        rng.shuffle(lst)
        adjacency.append([int(x) for x in lst])
    return RiskMap(name=f"synth{n}_s{seed}", n_players=n_players, adjacency=adjacency, continent=cont,
                   continent_bonus=[int(math.ceil(s / 3)) for s in sizes])
