"""Shared helpers for the parity tests: golden-vector loading and error norms."""

from __future__ import annotations

import glob
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


AUX_GOLDEN = {"cfg5_spotcheck"}  # .npz files that are not single-simulation cases


def golden_names():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(GOLDEN + "/*.npz"))
    return [n for n in names if n not in AUX_GOLDEN]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("steps", "probe_index", "extra_history0_keys"):
        g[k] = int(g[k])
    for k in ("c", "dt", "damping", "effective_dt", "dominant_freq"):
        g[k] = float(g[k])
    for k in ("metadata", "warning", "lattice"):
        g[k] = json.loads(str(g[k]))
    g["name"] = name
    return g


def lattice_names():
    return [n for n in golden_names() if load_golden(n)["lattice"]]


def rel_inf(a, b):
    """max_t,i |a-b| / max|b|  -- the tolerance norm of BASELINE.md section 3 (most nodes are ~0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def graph_from_golden(g):
    """Rebuild a networkx graph with the golden case's node order, adjacency order and attributes
    (so the drop-in can be exercised through the reference's own call signature on the GPU box,
    where /root/reference does not exist)."""
    import networkx as nx

    nodes = g["nodes"]
    if nodes.shape[1] == 1:
        labels = [int(x[0]) for x in nodes]
    else:
        labels = [tuple(int(v) for v in x[:-1]) + ("ABCD"[int(x[-1])],) for x in nodes]
    G = nx.Graph()
    G.add_nodes_from(labels)
    rowptr, col = g["rowptr"], g["colidx"]
    # inserting (i, j) for every adjacency entry in row-major order reproduces each node's
    # adjacency order only if we add directed entries carefully: networkx appends j to adj[i]
    # AND i to adj[j].  Fill the adjacency dicts directly instead.
    for i, a in enumerate(labels):
        for k in range(rowptr[i], rowptr[i + 1]):
            G._adj[a][labels[col[k]]] = {}
    for i, a in enumerate(labels):
        im = g["inv_mass"][i]
        if im != 1.0:
            G.nodes[a]["mass"] = 1.0 / im
        if g["potential"][i] != 0.0:
            G.nodes[a]["potential"] = float(g["potential"][i])
    # share edge-attribute dicts between the two directions, as networkx does
    for a in labels:
        for b in G._adj[a]:
            G._adj[b][a] = G._adj[a][b]
    return G, labels
