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


def lattice_tables_from_graph(G, shape=None, max_corrections: int | None = None) -> dict:
    """Read a tetrakis-shaped ``networkx`` graph as "clique model + sparse exceptions" -- host only, NumPy only.

    The structured kernel evaluates the operator of ``tetrakis_sim/lattice.py:30-73`` (cell, row and column cliques,
    vertical links) from per-node bits: ``alive`` (the node exists) and ``part`` (it has edges).  Everything else the
    graph says becomes explicit data: nodes that are missing (black hole, ``defects.py:110-135``), nodes without
    edges (``prune_edges``, ``defects.py:231-238``), ``mass`` / ``potential`` attributes (``defects.py:221-229``), and
    one directed correction ``(a, b, +-1)`` for every edge on which graph and model differ (the wedge cut,
    ``defects.py:80-84``; any edge a caller added or removed by hand).

    Cost: one C-level pass over the adjacency (``graph_arrays``) + O(E) array arithmetic; model edges the graph lacks
    are enumerated only around the nodes whose model degree is not met.  Raises ``ValueError`` when the graph is not
    a tetrakis lattice in this sense (foreign node labels, self-loops, multigraph / directed, or more than
    ``max_corrections`` exceptions) -- ``path="auto"`` then stays on the generic kernel."""
    shape = shape or detect_lattice_shape(G)
    if shape is None:
        raise ValueError("graph nodes are not tetrakis (r, c[, z], q) tuples")
    if G.is_multigraph() or G.is_directed():
        raise ValueError("multigraphs / directed graphs take the generic kernel")
    dim, S, L = shape
    nodes, index, rowptr, colidx, degree, inv_mass_g, potential_g = graph_arrays(G)
    n = len(nodes)
    nslots = L * S * S * 4
    qmap = {q: i for i, q in enumerate(QUADRANTS)}
    if dim == 2:
        coords = np.array([(0, v[0], v[1], qmap[v[2]]) for v in nodes], dtype=np.int64).reshape(-1, 4)
    else:
        coords = np.array([(v[2], v[0], v[1], qmap[v[3]]) for v in nodes], dtype=np.int64).reshape(-1, 4)
    gidx = ((coords[:, 0] * S + coords[:, 1]) * S + coords[:, 2]) * 4 + coords[:, 3]
    counts = np.diff(rowptr)
    rows = np.repeat(np.arange(n, dtype=np.int64), counts)
    if np.any(colidx == rows):
        raise ValueError("self-loops take the generic kernel")
    alive = np.zeros(nslots, dtype=bool)
    part = np.zeros(nslots, dtype=bool)
    alive[gidx] = True
    part[gidx] = counts > 0
    inv_mass = np.ones(nslots)
    potential = np.zeros(nslots)
    inv_mass[gidx] = inv_mass_g
    potential[gidx] = potential_g
    sp = np.flatnonzero(~(alive & part) | (inv_mass != 1.0) | (potential != 0.0))
    flags = (alive[sp] * FLAG_ALIVE + part[sp] * FLAG_PART).astype(np.uint8)

    # ---- graph edges the model does not have: every directed CSR entry, classified at once
    src, dst = gidx[rows], gidx[colidx]
    cs, cd = src >> 2, dst >> 2
    same_q = (src & 3) == (dst & 3)
    cols, cold = cs % S, cd % S
    rs_, rd_ = (cs // S) % S, (cd // S) % S
    zs, zd = cs // (S * S), cd // (S * S)
    in_model = (cs == cd) | (same_q & (zs == zd) & ((rs_ == rd_) | (cols == cold))) | \
               (same_q & (rs_ == rd_) & (cols == cold) & (np.abs(zs - zd) == 1))
    extra = np.flatnonzero(~in_model)
    corr_a, corr_b, corr_s = [src[extra]], [dst[extra]], [np.ones(extra.size)]

    # ---- model edges the graph does not have: only around nodes whose model degree is not met
    P4 = part.reshape(L, S, S, 4).astype(np.int64)
    Pz = np.zeros((L + 2, S, S, 4), dtype=np.int64)
    Pz[1:-1] = P4
    model_deg = P4 * ((P4.sum(3, keepdims=True) - 1) + (P4.sum(2, keepdims=True) - 1) +
                      (P4.sum(1, keepdims=True) - 1) + Pz[:-2] + Pz[2:])
    met = np.bincount(src[in_model], minlength=nslots)
    short = np.flatnonzero(model_deg.ravel() != met)
    if max_corrections is not None and short.size > max_corrections:
        raise ValueError(f"{short.size} nodes differ from the clique model (limit {max_corrections})")
    if short.size:
        row_of_slot = np.full(nslots, -1, dtype=np.int64)
        row_of_slot[gidx] = np.arange(n)
        plane = S * S * 4
        for a in short.tolist():
            q, cell = a & 3, a >> 2
            c0, r0, z0 = cell % S, (cell // S) % S, cell // (S * S)
            cand = [(cell << 2) + np.array([x for x in range(4) if x != q], dtype=np.int64),
                    (((z0 * S + r0) * S + np.delete(np.arange(S), c0)) << 2) + q,
                    (((z0 * S + np.delete(np.arange(S), r0)) * S + c0) << 2) + q]
            if z0 > 0:
                cand.append(np.array([a - plane], dtype=np.int64))
            if z0 < L - 1:
                cand.append(np.array([a + plane], dtype=np.int64))
            cand = np.concatenate(cand)
            cand = cand[part[cand]]
            g = row_of_slot[a]
            have = dst[rowptr[g]:rowptr[g + 1]]
            missing = np.setdiff1d(cand, have, assume_unique=True)
            corr_a.append(np.full(missing.size, a, dtype=np.int64))
            corr_b.append(missing)
            corr_s.append(-np.ones(missing.size))
    corr_a, corr_b, corr_s = np.concatenate(corr_a), np.concatenate(corr_b), np.concatenate(corr_s)
    if max_corrections is not None and corr_a.size > max_corrections:
        raise ValueError(f"{corr_a.size} edge corrections against the clique model (limit {max_corrections})")
    order = np.lexsort((corr_b, corr_a))
    corr = [(int(a), int(b), float(s)) for a, b, s in zip(corr_a[order], corr_b[order], corr_s[order])]
    return {"size": S, "layers": L, "dim": dim, "nodes": nodes, "index": index, "state_index": gidx,
            "special_index": sp, "special_flags": flags, "special_inv_mass": inv_mass[sp],
            "special_potential": potential[sp], "corr": corr, "max_degree": int(degree.max()) if n else 0}


def structured_from_graph(G, shape=None, device: int = 0, max_corrections: int | None = None) -> CompiledGraph:
    """Route a tetrakis-shaped ``networkx`` graph to the structured kernel (24 bytes per node update instead of
    44 + 4 * degree): :func:`lattice_tables_from_graph` + ``tk_lattice_plan_create``."""
    t = lattice_tables_from_graph(G, shape, max_corrections=max_corrections)
    plan = Plan.lattice(t["size"], t["layers"], special_index=t["special_index"], special_flags=t["special_flags"],
                        special_inv_mass=t["special_inv_mass"], special_potential=t["special_potential"],
                        corr=t["corr"], device=device)  # fmt: skip
    return CompiledGraph(plan, t["nodes"], t["index"], t["max_degree"], state_index=t["state_index"])
