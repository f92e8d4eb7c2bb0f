"""GPU: randomised parity -- beyond the three named defects.

* graph kernel vs the C oracle (itself bit-identical to the reference on every golden case): random
  graphs from several generators with random masses / potentials / self-loops / isolated nodes / kicks /
  damping, every reordering: BIT-EXACT.
* lattice kernel vs the C clique-sum oracle: arbitrary scattered removed / pruned / tagged nodes and random
  edge corrections, 2-D and 3-D, ragged sizes: 1e-12.
"""

import warnings

import networkx as nx
import numpy as np
import pytest

import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo
from tests.helpers import rel_inf
from tetrakis_sim_b200._lib import Plan
from tetrakis_sim_b200.compiler import graph_arrays

pytestmark = pytest.mark.gpu


def _random_graph(seed):
    rng = np.random.default_rng(seed)
    kind = seed % 6
    n = int(rng.integers(2, 700))
    if kind == 0:
        G = nx.gnp_random_graph(n, min(1.0, 6.0 / n), seed=seed)
    elif kind == 1:
        G = nx.barabasi_albert_graph(max(n, 5), 3, seed=seed)
    elif kind == 2:
        G = nx.random_regular_graph(4, n + (n % 2), seed=seed)
    elif kind == 3:
        G = nx.grid_2d_graph(int(rng.integers(2, 25)), int(rng.integers(2, 25)))
    elif kind == 4:
        G = nx.star_graph(n)
    else:
        G = nx.disjoint_union(nx.path_graph(int(rng.integers(1, 40))), nx.complete_graph(int(rng.integers(2, 40))))
        G.add_nodes_from(["lonely", ("tuple", 1)])
    nodes = list(G.nodes)
    for v in rng.choice(len(nodes), max(1, len(nodes) // 5), replace=False):
        G.nodes[nodes[v]]["mass"] = float(rng.choice([-1.0, 0.0, 0.3, 7.0, 1000.0]))
    for v in rng.choice(len(nodes), max(1, len(nodes) // 7), replace=False):
        G.nodes[nodes[v]]["potential"] = float(rng.uniform(-0.5, 30.0))
    if seed % 3 == 0:
        G.add_edge(nodes[0], nodes[0])
    return G, rng


@pytest.mark.parametrize("seed", range(24))
def test_graph_kernel_bit_exact_on_random_graphs(seed):
    G, rng = _random_graph(seed)
    nodes, index, rowptr, colidx, degree, inv_mass, potential = graph_arrays(G)
    steps = int(rng.integers(1, 80))
    c, dt, damping = float(rng.uniform(0.2, 2.0)), float(rng.uniform(0.01, 0.6)), float(rng.choice([0.0, 0.03]))
    kicks = {nodes[i]: float(rng.normal()) for i in rng.choice(len(nodes), min(len(nodes), 4), replace=False)}
    if seed % 5 == 0:
        kicks = None  # default kick at nodes[len(nodes)//2]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        h = tk.run_wave_sim(G, steps=steps, initial_data=kicks, c=c, dt=dt, damping=damping,
                            reorder=["auto", "none", "degree", "rcm"][seed % 4])
    u0 = wo.initial_state(len(nodes), None if kicks is None else {index[k]: v for k, v in kicks.items()})
    ref = wo.run_graph(rowptr, colidx, degree, inv_mass, potential, u0, steps, (c * h.dt) ** 2, damping)
    dt_eff, meta, _ = wo.cfl_clamp(int(degree.max()) if len(nodes) else 0, c, dt, steps, damping)
    assert h.dt == dt_eff and h.metadata == meta
    assert np.array_equal(h.array, ref["history"])


@pytest.mark.parametrize("seed", range(16))
def test_lattice_kernel_with_arbitrary_sparse_defects(seed):
    rng = np.random.default_rng(1000 + seed)
    dim3 = seed % 2 == 1
    S = int(rng.choice([3, 7, 31, 33, 64, 70, 97]))
    L = int(rng.integers(1, 6)) if dim3 else 1
    n = L * S * S * 4
    flags = np.full(n, 3, dtype=np.uint8)
    im, pot = np.ones(n), np.zeros(n)
    k = max(1, n // 40)
    flags[rng.choice(n, k, replace=False)] = 0            # removed (scattered single nodes, not whole cells)
    flags[rng.choice(n, k, replace=False)] &= 1           # pruned where still alive
    tag = rng.choice(n, k, replace=False)
    im[tag], pot[tag] = rng.uniform(0.001, 3.0, k), rng.uniform(0.0, 20.0, k)
    part = np.flatnonzero(flags == 3)
    corr = []
    for _ in range(int(rng.integers(0, 6))):                # random extra / missing edges between part nodes
        a, b = (int(x) for x in rng.choice(part, 2, replace=False))
        if dim3 and abs(a // (S * S * 4) - b // (S * S * 4)) > 1:
            continue
        s = float(rng.choice([-1.0, 1.0]))
        corr += [(a, b, s), (b, a, s)]
    special = np.flatnonzero((flags != 3) | (im != 1.0) | (pot != 0.0))
    u0 = np.zeros(n)
    alive = np.flatnonzero(flags & 1)
    kicked = rng.choice(alive, 5, replace=False)
    u0[kicked] = rng.normal(size=5)
    steps, coeff, damping = int(rng.integers(3, 40)), float(rng.uniform(1e-4, 2e-3)), float(rng.choice([0.0, 0.02]))
    plan = Plan.lattice(S, L, special_index=special, special_flags=flags[special],
                        special_inv_mass=im[special], special_potential=pot[special], corr=corr)
    try:
        plan.set_state_sparse(kicked, u0[kicked])
        out = plan.run(steps, coeff, damping, probes=kicked)
        got = plan.get_state()
    finally:
        plan.close()
    ref = wo.run_structured(S, L, flags, im, pot, corr, u0, steps, coeff, damping, probes=kicked)
    assert rel_inf(got, ref["u"]) < 1e-12
    assert rel_inf(out["probes"], ref["probes"]) < 1e-12
    assert not got[(flags & 1) == 0].any()
