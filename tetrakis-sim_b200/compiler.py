"""One-time graph-to-device compiler.

Turns what ``run_wave_sim`` receives into a device plan:

* any ``networkx`` graph  -> :func:`compile_graph`: node list, adjacency-ordered CSR, degrees,
  inv_mass / potential exactly as the reference precomputes them
  (``tetrakis_sim/physics.py:56-64, 93-102``), handed to ``tk_graph_plan_create`` (sliced ELL on
  the device; bit-identical results).
* a :class:`LatticeSpec`  -> :func:`compile_lattice`: the sparse defect description handed to
  ``tk_lattice_plan_create`` (structured clique-sum kernel; lattices of any size).
* a ``networkx`` graph that *is* a tetrakis lattice can also be routed to the structured kernel
  (:func:`structured_from_graph`): the (alive, part, mass, potential) model is read off the
  graph and every edge on which graph and model differ becomes an explicit correction, so the
  result is the same operator.  Used for cross-checking the two kernels.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._lib import Plan
from .lattice import FLAG_ALIVE, FLAG_PART, QUADRANTS, LatticeSpec


@dataclass
class CompiledGraph:
    plan: Plan
    nodes: list  # list(G.nodes)
    index: dict  # node -> row
    max_degree: int
    state_index: np.ndarray | None = None  # row -> device state index (lattice route) or None


def graph_arrays(G):
    """(nodes, index, rowptr, colidx, degree, inv_mass, potential) -- physics.py:56-64, 93-102.

    One pass over the adjacency dicts with C-level iterators (``map`` / ``chain``): no Python-level loop per edge."""
    from itertools import chain

    nodes = list(G.nodes)
    n = len(nodes)
    index = dict(zip(nodes, range(n)))
    adj = G._adj if hasattr(G, "_adj") else {v: G.adj[v] for v in nodes}
    if n and next(iter(adj)) is not nodes[0] or len(adj) != n:  # adjacency dict in another order than the node dict
        rows = [adj[v] for v in nodes]
    else:
        rows = adj.values()
    counts = np.fromiter(map(len, rows), dtype=np.int64, count=n)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    colidx = np.fromiter(map(index.__getitem__, chain.from_iterable(rows)), dtype=np.int32, count=int(rowptr[-1]))
    # G.degree: the adjacency length, a self-loop counting twice (networkx reportviews.DegreeView)
    degree = counts.astype(np.float64)
    if G.is_multigraph():
        deg_map = dict(G.degree(nodes))
        degree = np.fromiter((deg_map[v] for v in nodes), dtype=np.float64, count=n)
    else:
        # a self-loop (i, i) shows up as colidx == row
        loops = np.flatnonzero(colidx == np.repeat(np.arange(n, dtype=np.int32), counts))
        if loops.size:
            np.add.at(degree, colidx[loops], 1.0)
    inv_mass = np.ones(n, dtype=np.float64)
    potential = np.zeros(n, dtype=np.float64)
    node_attrs = G._node if hasattr(G, "_node") else dict(G.nodes(data=True))
    for v, attrs in node_attrs.items():
        if attrs:
            if "mass" in attrs:
                m = float(attrs["mass"])
                if m <= 0.0:
                    m = 1.0
                inv_mass[index[v]] = 1.0 / m
            if "potential" in attrs:
                potential[index[v]] = float(attrs["potential"])
    return nodes, index, rowptr, colidx, degree, inv_mass, potential


def ell_padding_waste(rowptr) -> float:
    """Fraction of sliced-ELL entries (32-row slices, each padded to its own width) that are padding."""
    deg = np.diff(rowptr)
    n = len(deg)
    if n == 0 or deg.sum() == 0:
        return 0.0
    pad = np.zeros(((n + 31) // 32) * 32, dtype=np.int64)
    pad[:n] = deg
    stored = int((pad.reshape(-1, 32).max(axis=1) * 32).sum())
    return 1.0 - float(deg.sum()) / stored


def reorder_rows(rowptr, colidx, method: str):
    """Row permutation for the device layout.  ``perm[k]`` = original row stored at device row k.

    * "degree": stable sort by descending degree -- rows of similar length share a slice, which removes
      the sliced-ELL padding on graphs with hubs;
    * "rcm": reverse Cuthill-McKee (scipy) -- neighbours get nearby indices, so the u gathers of a warp
      touch fewer cache lines on large sparse graphs;
    * "auto": "degree" when more than a quarter of the ELL entries would be padding, else identity.
    The neighbour ORDER inside every row is untouched, so results stay bit-identical to the reference."""
    n = len(rowptr) - 1
    if method in (None, "none") or n == 0:
        return None
    if method == "auto":
        method = "degree" if ell_padding_waste(rowptr) > 0.25 else "none"
        if method == "none":
            return None
    if method == "degree":
        return np.argsort(-np.diff(rowptr), kind="stable")
    if method == "rcm":
        import scipy.sparse as sp
        from scipy.sparse.csgraph import reverse_cuthill_mckee

        A = sp.csr_matrix((np.ones(len(colidx), dtype=np.int8), colidx, rowptr), shape=(n, n))
        return np.asarray(reverse_cuthill_mckee(A, symmetric_mode=True), dtype=np.int64)
    raise ValueError("reorder must be 'auto', 'none', 'degree' or 'rcm'")


def compile_graph(G, device: int = 0, reorder: str = "auto") -> CompiledGraph:
    nodes, index, rowptr, colidx, degree, inv_mass, potential = graph_arrays(G)
    max_degree = int(degree.max()) if len(nodes) else 0
    perm = reorder_rows(rowptr, colidx, reorder)
    if perm is None:
        plan = Plan.graph(rowptr, colidx, degree, inv_mass, potential, device=device)
        return CompiledGraph(plan, nodes, index, max_degree)
    inv = np.empty(len(perm), dtype=np.int64)
    inv[perm] = np.arange(len(perm))
    deg_rows = np.diff(rowptr)[perm]
    rowptr_p = np.zeros(len(perm) + 1, dtype=np.int64)
    np.cumsum(deg_rows, out=rowptr_p[1:])
    # gather each permuted row's neighbour list (original order), renumbered
    take = np.concatenate([np.arange(rowptr[r], rowptr[r + 1]) for r in perm]) if len(colidx) else np.zeros(0, np.int64)
    colidx_p = inv[colidx[take]].astype(np.int32)
    plan = Plan.graph(rowptr_p, colidx_p, degree[perm], inv_mass[perm], potential[perm], device=device)
    return CompiledGraph(plan, nodes, index, max_degree, state_index=inv)


def compile_lattice(spec: LatticeSpec, device: int = 0, z_begin: int = 0, z_end: int | None = None,
                    dtype: str = "float64") -> Plan:
    """``dtype="float32"`` selects the optional fp32 state mode (12 B per update, fp64 sums/Laplacian)."""
    if dtype not in ("float64", "float32"):
        raise ValueError("dtype must be 'float64' or 'float32'")
    idx, flags, im, pot = spec.specials()
    return Plan.lattice(
        spec.size, spec.layers, z_begin=z_begin, z_end=z_end, special_index=idx,
        special_flags=flags, special_inv_mass=im, special_potential=pot,
        corr=spec.corrections(), device=device, f32=dtype == "float32",
    )  # fmt: skip


def detect_lattice_shape(G):
    """(dim, size, layers) if every node looks like a tetrakis node tuple, else None."""
    dim = None
    rmax = cmax = zmax = -1
    for n in G.nodes:
        if not isinstance(n, tuple) or len(n) not in (3, 4) or n[-1] not in QUADRANTS:
            return None
        if dim is None:
            dim = 2 if len(n) == 3 else 3
        elif (len(n) == 3) != (dim == 2):
            return None
        if not all(isinstance(x, (int, np.integer)) and x >= 0 for x in n[:-1]):
            return None
        rmax = max(rmax, n[0])
        cmax = max(cmax, n[1])
        if dim == 3:
            zmax = max(zmax, n[2])
    if dim is None:
        return None
    return dim, max(rmax, cmax) + 1, (zmax + 1 if dim == 3 else 1)


def structured_from_graph(G, shape=None, device: int = 0) -> CompiledGraph:
    """Route a tetrakis-shaped ``networkx`` graph to the structured kernel.

    O(E) host work (it enumerates the model's edges to diff them against G), so this is for
    graphs networkx can hold anyway."""
    shape = shape or detect_lattice_shape(G)
    if shape is None:
        raise ValueError("graph nodes are not tetrakis (r, c[, z], q) tuples")
    dim, S, L = shape
    spec = LatticeSpec(size=S, dim=dim, layers=L)
    nodes = list(G.nodes)
    index = {n: i for i, n in enumerate(nodes)}
    gidx = np.fromiter((spec.index(n) for n in nodes), dtype=np.int64, count=len(nodes))
    deg_map = dict(G.degree(nodes))
    alive = np.zeros(spec.num_slots, dtype=bool)
    part = np.zeros(spec.num_slots, dtype=bool)
    inv_mass = np.ones(spec.num_slots)
    potential = np.zeros(spec.num_slots)
    for n, gi in zip(nodes, gidx):
        alive[gi] = True
        part[gi] = deg_map[n] > 0
        a = G.nodes[n]
        if a:
            m = float(a.get("mass", 1.0))
            inv_mass[gi] = 1.0 / (m if m > 0.0 else 1.0)
            potential[gi] = float(a.get("potential", 0.0))
    sp = np.flatnonzero(~(alive & part) | (inv_mass != 1.0) | (potential != 0.0))
    flags = (alive[sp] * FLAG_ALIVE + part[sp] * FLAG_PART).astype(np.uint8)
    # model edges vs graph edges
    P4 = part.reshape(L, S, S, 4)
    model = set()
    plane = S * S * 4
    for z, r, c, q in zip(*np.nonzero(P4)):
        a = ((z * S + r) * S + c) * 4 + q
        for q2 in range(q + 1, 4):
            if P4[z, r, c, q2]:
                model.add((a, a - q + q2))
        for c2 in np.flatnonzero(P4[z, r, c + 1 :, q]):
            model.add((a, a + (int(c2) + 1) * 4))
        for r2 in np.flatnonzero(P4[z, r + 1 :, c, q]):
            model.add((a, a + (int(r2) + 1) * S * 4))
        if z + 1 < L and P4[z + 1, r, c, q]:
            model.add((a, a + plane))
    graph = set()
    for a, b in G.edges:
        ia, ib = int(gidx[index[a]]), int(gidx[index[b]])
        graph.add((min(ia, ib), max(ia, ib)))
    corr = []
    for a, b in sorted(model - graph):
        corr += [(int(a), int(b), -1.0), (int(b), int(a), -1.0)]
    for a, b in sorted(graph - model):
        corr += [(a, b, 1.0), (b, a, 1.0)]
    plan = Plan.lattice(S, L, special_index=sp, special_flags=flags,
                        special_inv_mass=inv_mass[sp], special_potential=potential[sp], corr=corr,
                        device=device)  # fmt: skip
    max_degree = max(deg_map.values(), default=0)
    return CompiledGraph(plan, nodes, index, int(max_degree), state_index=gidx)
