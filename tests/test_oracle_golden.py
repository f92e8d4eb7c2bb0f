"""CPU: pin the oracle against outputs of the unmodified reference (tests/golden/*.npz)."""

import networkx as nx
import numpy as np
import pytest

from oracle import lattice_oracle as lo
from oracle import wave_oracle as wo
from tests.helpers import golden_names, lattice_names, load_golden, rel_inf


@pytest.mark.parametrize("name", golden_names())
def test_generic_c_oracle_is_bit_exact_with_reference(name):
    g = load_golden(name)
    coeff = (g["c"] * g["effective_dt"]) ** 2
    out = wo.run_graph(g["rowptr"], g["colidx"], g["deg"], g["inv_mass"], g["potential"], g["u0"],
                       g["steps"], coeff, g["damping"], probes=[g["probe_index"]])
    assert np.array_equal(out["history"][g["snap_steps"]], g["states"])
    assert np.array_equal(out["probes"][:, 0], g["probe_values"])


@pytest.mark.parametrize("name", ["path3_clamped", "two_node_mass_potential", "t2d_s7_wedge",
                                  "random_graph_multi_kick"])
def test_pure_python_oracle_is_bit_exact_with_reference(name):
    g = load_golden(name)
    coeff = (g["c"] * g["effective_dt"]) ** 2
    hist = wo.run_graph_python(g["rowptr"], g["colidx"], g["deg"], g["inv_mass"], g["potential"],
                               g["u0"], g["steps"], coeff, g["damping"])
    assert np.array_equal(hist[g["snap_steps"]], g["states"])


@pytest.mark.parametrize("name", golden_names())
def test_cfl_metadata_matches_reference(name):
    g = load_golden(name)
    md = g["metadata"]
    dt_eff, meta, warning = wo.cfl_clamp(md["max_degree"], g["c"], g["dt"], g["steps"], g["damping"])
    assert dt_eff == g["effective_dt"]
    assert {k: float(v) for k, v in meta.items()} == {k: float(v) for k, v in md.items()}
    assert ([warning] if warning else []) == g["warning"]


@pytest.mark.parametrize("name", golden_names())
def test_fft_restatement_matches_reference(name):
    g = load_golden(name)
    freq, spec, _ = wo.run_fft(g["probe_values"], g["effective_dt"])
    assert np.array_equal(freq, g["fft_freq"])
    assert np.array_equal(spec, g["fft_spectrum"])
    direct = wo.dft_abs_direct(g["probe_values"])
    assert rel_inf(direct, g["fft_spectrum"]) < 1e-12
    assert abs(abs(freq[np.argmax(direct)]) - abs(g["dominant_freq"])) < 1e-12 or \
        wo.folded_bin(direct) == wo.folded_bin(g["fft_spectrum"])


def _oracle_graph(lat):
    kw = {}
    if lat["defect"] == "blackhole":
        kw = dict(radius=lat["radius"], center=tuple(lat["center"]))
    elif lat["defect"] == "singularity":
        kw = dict(radius=lat["radius"], mass=lat["mass"], potential=lat["potential"],
                  prune_edges=lat["prune_edges"])
    return lo.build_case(lat["size"], lat["dim"], lat["layers"], lat["defect"], **kw)


@pytest.mark.parametrize("name", lattice_names())
def test_lattice_oracle_reproduces_reference_graph_exactly(name):
    """Node order, adjacency ORDER, degrees, attributes and the kick rule all match."""
    g = load_golden(name)
    lat = g["lattice"]
    G, removed, center = _oracle_graph(lat)
    nodes, rowptr, colidx, inv_mass, potential = lo.graph_arrays(G)
    enc = np.array([[*n[:-1], "ABCD".index(n[-1])] for n in nodes], dtype=np.int64)
    assert np.array_equal(enc, g["nodes"])
    assert np.array_equal(rowptr, g["rowptr"])
    assert np.array_equal(colidx, g["colidx"])
    assert np.array_equal(inv_mass, g["inv_mass"])
    assert np.array_equal(potential, g["potential"])
    assert len(removed) == lat["removed_node_count"]
    if name.startswith(("t2d_s7", "t3d_s7", "cfg")):
        k = lo.pick_kick_node(G, center, lat["dim"])
        assert [*k[:-1], "ABCD".index(k[-1])] == lat["kick"]


@pytest.mark.parametrize("name", lattice_names())
def test_clique_sum_restatement_matches_reference(name):
    """Both the NumPy and the C/OpenMP clique-sum restatements track the reference to ~1e-13."""
    g = load_golden(name)
    lat = g["lattice"]
    size, layers = lat["size"], lat["layers"]
    G, _, _ = _oracle_graph(lat)
    flags, inv_mass, potential, corr = wo.structured_inputs_from_graph(G, size, layers)
    idx = np.array([wo.lattice_index(size, layers, n) for n in G.nodes])
    u0 = np.zeros(layers * size * size * 4)
    u0[idx] = g["u0"]
    coeff = (g["c"] * g["effective_dt"]) ** 2
    steps = g["steps"]
    tol = 1e-13 if steps <= 100 else 2e-11  # noise grows ~ steps^2
    if steps <= 100:
        hist = wo.run_structured_numpy(size, layers, flags, inv_mass, potential, corr, u0, steps,
                                       coeff, g["damping"])
        assert rel_inf(hist[g["snap_steps"]][:, idx], g["states"]) < tol
        dead = np.setdiff1d(np.arange(u0.size), idx)
        assert not hist[:, dead].any()
    out = wo.run_structured(size, layers, flags, inv_mass, potential, corr, u0, steps, coeff,
                            g["damping"], probes=[idx[g["probe_index"]]])
    assert rel_inf(out["probes"][:, 0], g["probe_values"]) < tol
    assert rel_inf(out["u"][idx], g["states"][-1]) < tol
    if lat["defect"] in ("none", "blackhole"):
        assert corr == []
    if lat["defect"] == "wedge":
        assert len(corr) == 2 and corr[0][2] == -1.0


def test_config1_known_answers():
    """SURVEY.md 8c known answers for BASELINE config 1."""
    g = load_golden("cfg1_s11l7_blackhole_r2p5")
    assert g["lattice"]["kick"] == [3, 3, 3, 0]
    assert g["lattice"]["removed_node_count"] == 324
    assert np.allclose(g["probe_values"][:6],
                       [1, 1.24, 0.568, -0.456832, -1.03939328, -0.79000177], atol=1e-8)
    assert g["fft_spectrum"][0] == pytest.approx(1.923720427205299, rel=1e-12)
    assert abs(g["dominant_freq"]) == pytest.approx(0.784313725490196, rel=1e-12)
    assert wo.folded_bin(g["fft_spectrum"]) == 8


def test_oracle_builder_small_counts():
    """tests/test_lattice.py-style structure pins on the oracle builder."""
    G = lo.build_sheet(size=3, dim=2)
    assert G.number_of_nodes() == 36
    assert max(dict(G.degree()).values()) == 3 + 2 * 2
    G3 = lo.build_sheet(size=3, dim=3, layers=2)
    assert G3.has_edge((1, 1, 0, "A"), (1, 1, 1, "A"))
    assert nx.is_connected(G3)
