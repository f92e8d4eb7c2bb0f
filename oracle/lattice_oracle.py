"""Oracle restatement of the reference's lattice / defect construction.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Builds ``networkx`` graphs
that are *identical* to the reference's -- same node order, same adjacency
insertion order (the order matters: ``physics.py:116-117`` sums neighbours in
adjacency order, and floating-point addition is not associative).

Restates:
  * ``tetrakis_sim/lattice.py:15-76``   build_sheet
  * ``tetrakis_sim/defects.py:55-107``  wedge       (drop edge A-B of the centre cell)
  * ``tetrakis_sim/defects.py:110-135`` black hole  (delete nodes with dist <  radius)
  * ``tetrakis_sim/defects.py:194-247`` singularity (tag nodes with dist <= radius,
                                                     optionally strip all their edges)
  * ``tetrakis_sim/batch.py:124-151``   kick-node rule
"""

from __future__ import annotations

import math
from itertools import combinations

import networkx as nx

QUADRANTS = "ABCD"


def _layer_edges(size: int, z: int | None):
    """Edges of one layer in the reference's insertion order (lattice.py:37-44 / 59-66):
    every cell's 4-clique, then every row clique, then every column clique."""

    def node(r, c, q):
        return (r, c, q) if z is None else (r, c, z, q)

    for r in range(size):
        for c in range(size):
            for qa, qb in combinations(QUADRANTS, 2):
                yield node(r, c, qa), node(r, c, qb)
    for r in range(size):
        for q in QUADRANTS:
            for ca, cb in combinations(range(size), 2):
                yield node(r, ca, q), node(r, cb, q)
    for c in range(size):
        for q in QUADRANTS:
            for ra, rb in combinations(range(size), 2):
                yield node(ra, c, q), node(rb, c, q)


def build_sheet(size: int = 9, dim: int = 2, layers: int = 3) -> nx.Graph:
    """lattice.py:15-76.  2D nodes (r, c, q); 3D nodes (r, c, z, q), r-major, then c, z, q."""
    G = nx.Graph()
    if dim == 2:
        G.add_nodes_from((r, c, q) for r in range(size) for c in range(size) for q in QUADRANTS)
        G.add_edges_from(_layer_edges(size, None))
        return G
    if dim == 3:
        G.add_nodes_from(
            (r, c, z, q)
            for r in range(size)
            for c in range(size)
            for z in range(layers)
            for q in QUADRANTS
        )
        for z in range(layers):
            G.add_edges_from(_layer_edges(size, z))
        # vertical links, lattice.py:68-72
        G.add_edges_from(
            ((r, c, z, q), (r, c, z + 1, q))
            for z in range(layers - 1)
            for r in range(size)
            for c in range(size)
            for q in QUADRANTS
        )
        return G
    raise ValueError("dim must be 2 or 3")


def default_center(size: int, dim: int, layers: int) -> tuple[int, ...]:
    """batch.py:89-92."""
    return (size // 2, size // 2) if dim == 2 else (size // 2, size // 2, layers // 2)


def _distance(node, center) -> float:
    """defects.py:120-131 / 214-219: hypot for a 2-tuple centre, sqrt of squares for a 3-tuple."""
    if len(center) == 2:
        return math.hypot(node[0] - center[0], node[1] - center[1])
    return math.sqrt(
        (node[0] - center[0]) ** 2 + (node[1] - center[1]) ** 2 + (node[2] - center[2]) ** 2
    )


def apply_blackhole(G: nx.Graph, center, radius: float) -> list:
    """defects.py:110-135: strict ``<``; a 2-tuple centre on a 3-D lattice removes a cylinder."""
    doomed = [n for n in list(G.nodes) if _distance(n, center) < radius]
    G.remove_nodes_from(doomed)
    return doomed


def apply_singularity(
    G: nx.Graph,
    center,
    *,
    mass: float = 1000.0,
    potential: float = 0.0,
    radius: float = 0.0,
    prune_edges: bool = False,
) -> list:
    """defects.py:194-247: ``<=``; tagged nodes stay in the graph; prune strips every incident edge."""
    tagged = [n for n in list(G.nodes) if _distance(n, center) <= radius]
    for n in tagged:
        G.nodes[n]["mass"] = float(mass)
        G.nodes[n]["potential"] = float(potential)
        G.nodes[n]["singular"] = True
    if prune_edges:
        for n in tagged:
            G.remove_edges_from([(n, nb) for nb in list(G.neighbors(n))])
    return tagged


def apply_wedge(G: nx.Graph, center=None, layer: int | None = None):
    """defects.py:55-107: remove the single edge A-B of the centre cell (on ``layer`` in 3-D)."""
    if not G.nodes:
        return None
    first = next(iter(G.nodes))
    if center is None:
        rows = sorted({n[0] for n in G.nodes})
        cols = sorted({n[1] for n in G.nodes})
        center = (rows[len(rows) // 2], cols[len(cols) // 2])
    if len(first) == 3:
        e = ((center[0], center[1], "A"), (center[0], center[1], "B"))
    else:
        zs = sorted({n[2] for n in G.nodes})
        if layer is None:
            layer = zs[len(zs) // 2]
        e = ((center[0], center[1], layer, "A"), (center[0], center[1], layer, "B"))
    if G.has_edge(*e):
        G.remove_edge(*e)
        return e
    return None


def build_case(
    size: int,
    dim: int,
    layers: int,
    defect: str = "none",
    *,
    radius: float = 2.5,
    center=None,
    mass: float = 1000.0,
    potential: float = 0.0,
    prune_edges: bool = False,
):
    """Build lattice + defect the way ``batch.py:87-114`` does.  Returns (G, removed_nodes, lattice_center);
    ``center`` overrides the defect centre only (the kick rule always uses the lattice centre)."""
    G = build_sheet(size=size, dim=dim, layers=layers)
    lattice_center = default_center(size, dim, layers)
    if center is None:
        center = lattice_center
    removed: list = []
    if defect == "blackhole":
        removed = apply_blackhole(G, center, radius)
    elif defect == "singularity":
        apply_singularity(
            G, center, mass=mass, potential=potential, radius=radius, prune_edges=prune_edges
        )
    elif defect == "wedge":
        if dim == 2:
            apply_wedge(G, center=center[:2])
        else:
            apply_wedge(G, center=center[:2], layer=lattice_center[2])
    elif defect != "none":
        raise ValueError(f"Unsupported defect type '{defect}'")
    return G, removed, lattice_center


def pick_kick_node(G: nx.Graph, center, dim: int):
    """batch.py:124-151 / evals/dataset.py:23-41: closest (in r, c) connected non-singular node on
    the central layer; ``min`` keeps the first in node order on ties."""
    if dim == 3:
        pool = [n for n in G if len(n) > 3 and n[2] == center[2]]
    else:
        pool = list(G.nodes)
    if not pool:
        raise RuntimeError("No nodes remain after defect application")
    good = [n for n in pool if G.degree[n] > 0 and not G.nodes[n].get("singular", False)]
    pool = good if good else pool
    return min(pool, key=lambda n: (n[0] - center[0]) ** 2 + (n[1] - center[1]) ** 2)


def graph_arrays(G: nx.Graph):
    """Flatten a graph the way ``physics.py:56-102`` reads it.

    Returns (nodes, rowptr[int64 N+1], colidx[int32 nnz] in adjacency order, inv_mass[N], potential[N]).
    """
    import numpy as np

    nodes = list(G.nodes)
    index = {n: i for i, n in enumerate(nodes)}
    rowptr = np.zeros(len(nodes) + 1, dtype=np.int64)
    cols: list[int] = []
    inv_mass = np.empty(len(nodes), dtype=np.float64)
    potential = np.empty(len(nodes), dtype=np.float64)
    for i, n in enumerate(nodes):
        nbrs = [index[m] for m in G[n]]
        cols.extend(nbrs)
        rowptr[i + 1] = rowptr[i] + len(nbrs)
        m = float(G.nodes[n].get("mass", 1.0))
        if m <= 0.0:
            m = 1.0
        inv_mass[i] = 1.0 / m
        potential[i] = float(G.nodes[n].get("potential", 0.0))
    return nodes, rowptr, np.asarray(cols, dtype=np.int32), inv_mass, potential
