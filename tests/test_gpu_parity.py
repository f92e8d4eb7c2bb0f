"""GPU parity tests proper: the CUDA path, called through the drop-in API / C ABI, against the
golden vectors generated from the unmodified reference and against the oracle.

Bars: generic-graph kernel BIT-EXACT on u(t); structured kernel within
1e-10 relative (inf-norm over all nodes and steps, relative to max|u_ref|; we hold 1e-12 at <= 100
steps); dominant-frequency bin identical modulo the k <-> N-k mirror.
"""

import os
import warnings

import numpy as np
import pytest

import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo
from tests.helpers import golden_names, graph_from_golden, lattice_names, load_golden, rel_inf
from tests.test_cpu_host import _spec_from

pytestmark = pytest.mark.gpu


def _initial_data(g, labels):
    nz = np.flatnonzero(g["u0"])
    return {labels[i]: float(g["u0"][i]) for i in nz}


@pytest.mark.parametrize("name", golden_names())
def test_graph_kernel_is_bit_exact_with_the_reference(name):
    g = load_golden(name)
    G, labels = graph_from_golden(g)
    if name == "random_graph_multi_kick":
        G.add_edge(5, 5)  # restore the self-loop's degree contribution (adjacency unchanged)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        hist = tk.run_wave_sim(G, steps=g["steps"], initial_data=_initial_data(g, labels),
                               c=g["c"], dt=g["dt"], damping=g["damping"])
    assert [str(w.message) for w in caught if w.category is RuntimeWarning] == g["warning"]
    assert isinstance(hist, tk.SimulationHistory) and len(hist) == g["steps"] + 1
    assert hist.dt == g["effective_dt"]
    assert {k: float(v) for k, v in hist.metadata.items()} == {k: float(v) for k, v in g["metadata"].items()}
    assert np.array_equal(hist.array[g["snap_steps"]], g["states"])
    probe = labels[g["probe_index"]]
    assert [s[probe] for s in hist] == g["probe_values"].tolist()
    assert list(hist[0].keys()) == labels
    freq, spectrum, values = tk.run_fft(hist, probe)
    assert np.array_equal(values, g["probe_values"]) and np.array_equal(freq, g["fft_freq"])
    assert rel_inf(spectrum, g["fft_spectrum"]) < 1e-12
    assert wo.folded_bin(spectrum) == wo.folded_bin(g["fft_spectrum"])
    assert abs(tk.dominant_frequency(freq, spectrum)) == pytest.approx(abs(g["dominant_freq"]), rel=1e-12)


@pytest.mark.parametrize("name", lattice_names())
@pytest.mark.parametrize("route", ["spec", "graph"])
def test_structured_kernel_matches_the_reference(name, route):
    g = load_golden(name)
    lat = g["lattice"]
    tol = 1e-12 if g["steps"] <= 100 else 1e-10
    G, labels = graph_from_golden(g)
    init = _initial_data(g, labels)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        if route == "graph":
            hist = tk.run_wave_sim(G, steps=g["steps"], initial_data=init, c=g["c"], dt=g["dt"],
                                   damping=g["damping"], path="structured")
            states = hist.array[g["snap_steps"]]
        else:
            spec = _spec_from(lat)
            probe = labels[g["probe_index"]]
            hist = tk.run_wave_sim(spec, steps=g["steps"], initial_data=init, c=g["c"], dt=g["dt"],
                                   damping=g["damping"], record="all", probes=[probe])
            idx = np.array([spec.index(n) for n in labels])
            states = hist.array[g["snap_steps"]][:, idx]
            dead = np.setdiff1d(np.arange(spec.num_slots), idx)
            assert not hist.array[:, dead].any(), "removed nodes must stay at 0"
            assert rel_inf(hist.probe_series[:, 0], g["probe_values"]) < tol
            freq, spectrum, _ = tk.run_fft(hist, probe)
            assert wo.folded_bin(spectrum) == wo.folded_bin(g["fft_spectrum"])
            assert rel_inf(spectrum, g["fft_spectrum"]) < 10 * tol
    assert hist.dt == g["effective_dt"]
    assert hist.metadata["max_degree"] == g["metadata"]["max_degree"]
    assert rel_inf(states, g["states"]) < tol


# (cells/lane, segments/CTA, rows/CTA, min CTAs/SM, L2 prefetch distance in rows)
VARIANTS = [(1, 8, 16, 3, 2), (1, 2, 32, 3, 0), (1, 1, 64, 3, 5), (2, 8, 16, 1, 2), (2, 2, 32, 1, 1),
            (2, 1, 256, 2, 3), (4, 8, 32, 1, 2), (4, 1, 16, 1, 20), (1, 8, 40, 2, 4), (1, 8, 24, 4, 1),
            (1, 1, 19, 4, 2), (1, 1, 33, 2, 7), (2, 8, 50, 2, 2)]


@pytest.mark.parametrize("cpl,wc,rchunk,minb,pf", VARIANTS)
@pytest.mark.parametrize("case", ["2d_hole", "3d_sing", "3d_wedge_odd"])
def test_every_kernel_variant_against_the_oracle(cpl, wc, rchunk, minb, pf, case, monkeypatch):
    """Medium lattices (several segments, ragged tails, many row chunks) vs the C oracle."""
    monkeypatch.setenv("TK_LAT3D", "0")  # the register-path kernel, also for the 3-D cases
    monkeypatch.setenv("TK_LAT_MINB", str(minb))
    monkeypatch.setenv("TK_LAT_PREFETCH", str(pf))
    _variant_case(cpl, wc, rchunk, case, monkeypatch)


@pytest.mark.parametrize("wz,stages,rchunk", [(8, 4, 0), (8, 3, 16), (8, 6, 40), (4, 4, 7), (16, 3, 64),
                                              (16, 4, 5)])
@pytest.mark.parametrize("case", ["3d_sing", "3d_wedge_odd", "3d_hole_tall"])
def test_3d_staged_kernel_variants_against_the_oracle(wz, stages, rchunk, case, monkeypatch):
    """lat_step3d_kernel (cp.async-staged z neighbours): partial z-groups, ragged segments, short
    and long row chunks (fewer rows than pipeline stages included)."""
    monkeypatch.setenv("TK_LAT3D", "1")
    monkeypatch.setenv("TK_LAT3D_WZ", str(wz))
    monkeypatch.setenv("TK_LAT3D_STAGES", str(stages))
    _variant_case(1, 1, rchunk, case, monkeypatch)


@pytest.mark.parametrize("stages,rchunk", [(4, 0), (2, 16), (3, 40), (6, 5), (4, 3), (3, 128)])
@pytest.mark.parametrize("case", ["2d_hole", "3d_sing", "3d_wedge_odd", "3d_hole_tall"])
def test_tma_staged_kernel_variants_against_the_oracle(stages, rchunk, case, monkeypatch):
    """lat_step_bulk_kernel (cp.async.bulk + mbarrier, a private ring per warp): ragged last segments
    (short bulk copies), row chunks shorter than the ring, defect rows, 2-D and 3-D."""
    monkeypatch.setenv("TK_LAT3D", "0")
    monkeypatch.setenv("TK_LAT_BULK", "1")
    monkeypatch.setenv("TK_LAT_BULK_STAGES", str(stages))
    _variant_case(1, 1, rchunk, case, monkeypatch)


def _variant_case(cpl, wc, rchunk, case, monkeypatch):
    monkeypatch.setenv("TK_LAT_CPL", str(cpl))
    monkeypatch.setenv("TK_LAT_WC", str(wc))
    if rchunk:
        monkeypatch.setenv("TK_LAT_RCHUNK", str(rchunk))
    if case == "2d_hole":
        spec = tk.LatticeSpec(size=150, dim=2, defect="blackhole", radius=9.3, center=(70, 66))
        steps, damping = 30, 0.01
    elif case == "3d_sing":
        spec = tk.LatticeSpec(size=70, dim=3, layers=11, defect="singularity", radius=3.0,
                              mass=2000.0, potential=1.5, prune_edges=True)
        steps, damping = 20, 0.0
    elif case == "3d_hole_tall":
        spec = tk.LatticeSpec(size=45, dim=3, layers=21, defect="blackhole", radius=5.2)
        steps, damping = 12, 0.0
    else:
        spec = tk.LatticeSpec(size=37, dim=3, layers=9, defect="wedge")
        steps, damping = 25, 0.02
    kick = spec.kick_node()
    far = spec.node(spec.index(kick) ^ 3)
    flags = np.full(spec.num_slots, 3, dtype=np.uint8)
    im = np.ones(spec.num_slots)
    pot = np.zeros(spec.num_slots)
    si, sf, sm, sv = spec.specials()
    flags[si], im[si], pot[si] = sf, sm, sv
    u0 = np.zeros(spec.num_slots)
    u0[spec.index(kick)] = 1.0
    u0[spec.index(far)] = -0.25
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        hist = tk.run_wave_sim(spec, steps=steps, initial_data={kick: 1.0, far: -0.25}, c=1.0,
                               dt=0.05, damping=damping, probes=[kick, far])
    coeff = (1.0 * hist.dt) ** 2
    ref = wo.run_structured(spec.size, spec.layers, flags, im, pot, spec.corrections(), u0, steps,
                            coeff, damping, probes=[spec.index(kick), spec.index(far)])
    assert rel_inf(hist.final_state, ref["u"]) < 1e-12
    assert rel_inf(hist.probe_series, ref["probes"]) < 1e-12


def test_auto_path_takes_the_structured_kernel_when_opted_in(monkeypatch):
    """with TETRAKIS_B200_AUTO_STRUCTURED=1 (or install(structured=True)) path="auto" recognises
    the lattice the reference's front-ends build and runs it on the structured kernel; a graph that is not a tetrakis
    lattice, and every graph without the opt-in, stays on the bit-exact generic kernel."""
    import networkx as nx

    from oracle import lattice_oracle as lo
    from tetrakis_sim_b200 import physics as ph

    G, _, center = lo.build_case(13, 3, 5, "blackhole", radius=2.5)
    kick = lo.pick_kick_node(G, center, 3)
    seen = []
    real = ph.structured_from_graph
    monkeypatch.setattr(ph, "structured_from_graph", lambda *a, **k: (seen.append(1), real(*a, **k))[1])
    kw = dict(steps=40, initial_data={kick: 1.0}, dt=0.2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        a = tk.run_wave_sim(G, **kw)
        assert not seen
        monkeypatch.setenv("TETRAKIS_B200_AUTO_STRUCTURED", "1")
        b = tk.run_wave_sim(G, **kw)
        assert len(seen) == 1
        assert rel_inf(b.array, a.array) < 1e-12 and a.metadata == b.metadata and list(b[7]) == list(a[7])
        # not a lattice: falls back to the generic kernel without an error
        H = nx.barabasi_albert_graph(200, 3, seed=1)
        h1 = tk.run_wave_sim(H, steps=10, dt=0.05)
        monkeypatch.delenv("TETRAKIS_B200_AUTO_STRUCTURED")
        h0 = tk.run_wave_sim(H, steps=10, dt=0.05)
        assert np.array_equal(h0.array, h1.array)
        # install(structured=True) is the same switch
        ph.set_auto_structured(True)
        try:
            c = tk.run_wave_sim(G, **kw)
        finally:
            ph.set_auto_structured(None)
        assert len(seen) == 3 and np.array_equal(c.array, b.array)


def test_structured_and_graph_kernels_agree_on_a_medium_singularity_lattice():
    """SURVEY.md 8d config 4 cross-check: structured vs sliced-ELL on the same operator."""
    from oracle import lattice_oracle as lo

    G, _, center = lo.build_case(24, 3, 6, "singularity", radius=2.5, mass=2000.0, prune_edges=True)
    kick = lo.pick_kick_node(G, center, 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        a = tk.run_wave_sim(G, steps=60, initial_data={kick: 1.0}, dt=0.2)
        b = tk.run_wave_sim(G, steps=60, initial_data={kick: 1.0}, dt=0.2, path="structured")
        spec = tk.LatticeSpec(size=24, dim=3, layers=6, defect="singularity", radius=2.5,
                              mass=2000.0, prune_edges=True)
        c = tk.run_wave_sim(spec, steps=60, initial_data={kick: 1.0}, dt=0.2, record="all")
    assert spec.kick_node() == kick
    assert rel_inf(b.array, a.array) < 1e-12
    idx = np.array([spec.index(n) for n in a.nodes])
    assert rel_inf(c.array[:, idx], a.array) < 1e-12
    assert a.metadata == b.metadata
    assert c.metadata == a.metadata  # the reference's keys only; launch counts live in hist.launches


@pytest.mark.parametrize("n", [1, 2, 3, 51, 64, 1000, 10001])
def test_device_dft_matches_numpy_fft(n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n) + np.sin(np.arange(n) * 0.37)
    spec, arg = tk.dft_abs(x)
    ref = np.abs(np.fft.fft(x))
    assert rel_inf(spec, ref) < 1e-12
    assert wo.folded_bin(spec) == wo.folded_bin(ref) and spec[arg] == spec.max()
    xb = rng.normal(size=(5, n))
    sb, ab = tk.dft_abs(xb)
    assert rel_inf(sb, np.abs(np.fft.fft(xb, axis=1))) < 1e-12 and ab.shape == (5,)


@pytest.mark.parametrize("n", [2, 3, 51, 64, 1000, 4097, 10001, 65536, 100003])
def test_chirp_z_spectrum_matches_the_direct_sum_and_numpy(n, monkeypatch):
    """The O(N log N) path (Bluestein chirp-z over a power-of-two Stockham FFT, kernels_fft.cuh) against the direct
    O(N^2) kernel -- its checker -- and numpy's FFT, batched, for lengths around powers of two and BASELINE config 2's
    10 001 samples."""
    rng = np.random.default_rng(n)
    xb = rng.normal(size=(3, n)) + np.sin(np.arange(n) * 0.37)[None, :]
    monkeypatch.setenv("TK_DFT", "chirp")
    sc, ac = tk.dft_abs(xb)
    ref = np.abs(np.fft.fft(xb, axis=1))
    assert rel_inf(sc, ref) < 1e-12
    if n <= 10001:
        monkeypatch.setenv("TK_DFT", "direct")
        sd, ad = tk.dft_abs(xb)
        assert rel_inf(sc, sd) < 1e-12
        assert [wo.folded_bin(a) for a in sc] == [wo.folded_bin(a) for a in sd]
    assert all(sc[i][ac[i]] == sc[i].max() for i in range(3))


def test_long_batched_spectra_take_the_fast_path():
    """Whole-layer spectra of a long run (notebooks cells 12-13 at BASELINE config 2's step count): 256 series of
    10 001 samples -- 2.6e10 terms for the direct sum, a few milliseconds for the chirp-z path picked by default."""
    import time

    n = 10001
    t = np.arange(n)
    x = np.sin(0.05 * t)[None, :] * np.linspace(0.5, 2.0, 256)[:, None] + 0.1 * np.cos(0.31 * t)[None, :]
    t0 = time.perf_counter()
    s, a = tk.dft_abs(x)
    wall = time.perf_counter() - t0
    assert rel_inf(s, np.abs(np.fft.fft(x, axis=1))) < 1e-12
    assert set(np.minimum(a, n - a).tolist()) == {int(round(0.05 * n / (2 * np.pi)))}
    assert wall < 5.0


def test_run_fft_reference_contract():
    """tests/test_physics.py:30-38 and tests/test_invariants.py:96-115 of the reference."""
    h = tk.SimulationHistory([{0: 0.0}, {0: 1.0}, {0: 0.0}, {0: -1.0}], dt=0.5, wave_speed=1.0)
    freq, spectrum, values = tk.run_fft(h, node=0)
    assert np.allclose(values, [0.0, 1.0, 0.0, -1.0]) and freq[1] == pytest.approx(0.5)
    assert np.argmax(spectrum) in (1, 3)
    t = np.arange(100) * 0.1
    h = tk.SimulationHistory([{0: float(v)} for v in np.sin(2 * np.pi * 2.0 * t)], dt=0.1, wave_speed=1.0)
    freq, spectrum, _ = tk.run_fft(h, node=0)
    assert abs(abs(float(freq[int(np.argmax(spectrum))])) - 2.0) <= 0.1 + 1e-9
    with pytest.raises(ValueError, match="empty"):
        tk.run_fft([])


def test_reference_behavioural_contract_small_graphs():
    """The reference's own unit tests for this path, run against the drop-in
    (tests/test_physics.py:19-27, test_physics_mass_potential.py, test_invariants.py:66-93)."""
    import math

    import networkx as nx

    with pytest.warns(RuntimeWarning, match="Clamping dt"):
        h = tk.run_wave_sim(nx.path_graph(3), steps=3, c=1.0, dt=5.0)
    assert h.dt == pytest.approx(1 / math.sqrt(2)) and h.metadata["stability_adjusted"] is True
    assert h.metadata["max_degree"] == 2

    def one_step(**attrs):
        G = nx.Graph()
        G.add_edge(0, 1)
        G.nodes[0].update(attrs)
        return tk.run_wave_sim(G, steps=1, initial_data={0: 1.0}, c=1.0, dt=0.1)[-1][0]

    assert one_step(mass=10.0) > one_step() > one_step(potential=2.0)
    from oracle import lattice_oracle as lo

    G = lo.build_sheet(size=3, dim=2)
    with pytest.warns(RuntimeWarning, match=r"Clamping dt"):
        h = tk.run_wave_sim(G, steps=25, c=1.0, dt=10.0, damping=0.0)
    vals = np.asarray([v for s in h for v in s.values()])
    assert np.isfinite(vals).all() and np.max(np.abs(vals)) < 1e3
    assert h.metadata["stability_limit"] == pytest.approx(1 / math.sqrt(h.metadata["max_degree"]))
    # empty graph, zero steps, kick on a node that is not in the graph
    assert len(tk.run_wave_sim(nx.Graph(), steps=3)) == 4
    h = tk.run_wave_sim(nx.path_graph(4), steps=0, initial_data={"ghost": 2.0, 1: 0.5})
    assert len(h) == 1 and h[0]["ghost"] == 2.0 and h[0][1] == 0.5
    assert G.number_of_edges() == lo.build_sheet(size=3, dim=2).number_of_edges()  # not mutated


def test_size_independent_properties_on_a_large_lattice():
    """Properties that hold at any size (used where no oracle can run):
    (i) with unit masses, no potential and no damping, sum_i u_t(i) = sum(u_0) * (1 + t) exactly in
        exact arithmetic (the Laplacian has zero column sums and u_{-1} = 0);
    (ii) linearity: the response to a + b is the sum of the responses;
    (iii) time reversal: swapping (u, u_prev) and stepping again retraces the trajectory."""
    spec = tk.LatticeSpec(size=1024, dim=2)
    k1, k2 = spec.bench_kick_node(), (17, 1000, "C")
    kw = dict(c=1.0, dt=0.01, damping=0.0, probes=[k1, k2])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ha = tk.run_wave_sim(spec, steps=40, initial_data={k1: 1.0}, **kw)
        hb = tk.run_wave_sim(spec, steps=40, initial_data={k2: -2.0}, **kw)
        hab = tk.run_wave_sim(spec, steps=40, initial_data={k1: 1.0, k2: -2.0}, **kw)
    assert ha.final_state.sum() == pytest.approx(41.0, rel=1e-11)
    assert hab.final_state.sum() == pytest.approx(-41.0, rel=1e-11)
    scale = np.max(np.abs(hab.final_state))
    assert np.max(np.abs(hab.final_state - (ha.final_state + hb.final_state))) < 1e-12 * scale
    assert rel_inf(hab.probe_series, ha.probe_series + hb.probe_series) < 1e-12
    # time reversal through the raw plan API
    plan = tk.compile_lattice(spec)
    try:
        coeff = (1.0 * ha.dt) ** 2
        plan.set_state_sparse([spec.index(k1)], [1.0])
        plan.run(25, coeff, 0.0)
        u, up = plan.get_state(want_prev=True)
        plan.set_state(up, u)
        plan.run(25, coeff, 0.0)
        back, back_prev = plan.get_state(want_prev=True)
    finally:
        plan.close()
    expect = np.zeros(spec.num_slots)
    assert np.max(np.abs(back)) < 1e-9  # u_{-1} = 0
    expect[spec.index(k1)] = 1.0
    assert np.max(np.abs(back_prev - expect)) < 1e-9


@pytest.mark.parametrize("name", ["t2d_s7_sing_m50_v25", "t3d_s7l5_sing_prune_v3", "random_graph_multi_kick",
                                  "cfg1_s11l7_blackhole_r2p5", "t3d_s7l5_wedge"])
def test_energy_series_matches_the_oracle_on_both_kernels(name):
    """The energy diagnostic (not in the reference): both kernels against a longdouble evaluation of
    the same formula on the reference's own history."""
    g = load_golden(name)
    G, labels = graph_from_golden(g)
    if name == "random_graph_multi_kick":
        G.add_edge(5, 5)
    init = _initial_data(g, labels)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        h = tk.run_wave_sim(G, steps=g["steps"], initial_data=init, c=g["c"], dt=g["dt"],
                            damping=g["damping"], energy=True)
    coeff = (g["c"] * g["effective_dt"]) ** 2
    ref = wo.staggered_energy(h.array, g["rowptr"], g["colidx"], g["deg"], g["inv_mass"], g["potential"],
                              coeff) * g["c"] ** 2
    assert h.energy.shape == (g["steps"],)
    assert rel_inf(h.energy, ref) < 1e-11
    if g["lattice"]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            hs = tk.run_wave_sim(_spec_from(g["lattice"]), steps=g["steps"], initial_data=init, c=g["c"],
                                 dt=g["dt"], damping=g["damping"], energy=True)
        assert rel_inf(hs.energy, ref) < 1e-11


def test_energy_is_conserved_without_damping_and_decays_with_it():
    spec = tk.LatticeSpec(size=300, dim=3, layers=9, defect="blackhole", radius=4.5)
    k = spec.kick_node()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        h0 = tk.run_wave_sim(spec, steps=200, initial_data={k: 1.0}, dt=0.01, energy=True)
        h1 = tk.run_wave_sim(spec, steps=200, initial_data={k: 1.0}, dt=0.01, damping=0.01, energy=True)
    assert np.max(np.abs(h0.energy - h0.energy[0])) < 1e-10 * abs(h0.energy[0])
    assert h0.energy[0] > 0 and h1.energy[-1] < 0.5 * h1.energy[0] and np.mean(np.diff(h1.energy) < 0) > 0.9


def test_run_fft_many_equals_per_node_run_fft():
    """N2: whole-layer spectra in one batched call (notebooks/tetrakis_blackhole_waves.ipynb cells 12-13)."""
    g = load_golden("t3d_s7l5_blackhole_r2p5")
    G, labels = graph_from_golden(g)
    h = tk.run_wave_sim(G, steps=g["steps"], initial_data=_initial_data(g, labels), c=g["c"], dt=g["dt"],
                        damping=g["damping"])
    layer = [n for n in labels if n[2] == 2] + [(99, 99, 2, "A")]  # one node that is not in the graph
    freq, spectra, values, bins = tk.run_fft_many(h, layer)
    assert spectra.shape == (len(layer), g["steps"] + 1) and not values[-1].any()
    for i in (0, 7, len(layer) - 2):
        f1, s1, v1 = tk.run_fft(h, layer[i])
        assert np.array_equal(f1, freq) and np.array_equal(v1, values[i]) and np.array_equal(s1, spectra[i])
        assert bins[i] == int(np.argmax(s1))
    assert rel_inf(spectra, np.abs(np.fft.fft(values, axis=1))) < 1e-12


def test_medium_2d_lattice_against_the_c_oracle():
    """SURVEY.md 8d config 2 at reduced size: 2-D 2048 x 2048 (16.8 M nodes, 32 kernel segments per row,
    many row chunks) against the CPU clique-sum oracle."""
    spec = tk.LatticeSpec(size=2048, dim=2)
    kick = spec.bench_kick_node()
    steps = 30
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        h = tk.run_wave_sim(spec, steps=steps, initial_data={kick: 1.0}, c=1.0, dt=0.1, probes=[kick, (0, 2047, "D")])
    u0 = np.zeros(spec.num_slots)
    u0[spec.index(kick)] = 1.0
    ref = wo.run_structured(2048, 1, None, None, None, [], u0, steps, (1.0 * h.dt) ** 2, 0.0,
                            probes=[spec.index(kick), spec.index((0, 2047, "D"))])
    assert h.metadata["max_degree"] == 3 + 2 * 2047 and h.metadata["stability_adjusted"] is True
    assert rel_inf(h.final_state, ref["u"]) < 1e-12
    assert rel_inf(h.probe_series, ref["probes"]) < 1e-12


def test_ten_thousand_steps_against_the_bit_exact_oracle():
    """BASELINE north_star: relative 1e-10 on u(t) after N steps, N = 10 000 (config 2's step count).
    Truth = the generic C oracle, which is bit-identical to the reference (tests/test_oracle_golden.py);
    rounding differences between the reference's neighbour-order sums and the clique sums grow ~N^2
    (2.3e-11 at 5 000 steps is the reference's own float64 noise floor)."""
    from oracle import lattice_oracle as lo

    G = lo.build_sheet(size=15, dim=2)
    kick = (7, 7, "A")
    nodes, rowptr, colidx, inv_mass, potential = lo.graph_arrays(G)
    deg = np.array([G.degree[n] for n in nodes], dtype=np.float64)
    u0 = np.zeros(len(nodes))
    u0[nodes.index(kick)] = 1.0
    steps, dt = 10_000, 0.1
    truth = wo.run_graph(rowptr, colidx, deg, inv_mass, potential, u0, steps, dt * dt, 0.0,
                         probes=[nodes.index(kick)], keep_history=False)
    hg = tk.run_wave_sim(G, steps=steps, initial_data={kick: 1.0}, dt=dt, record="probes")
    assert np.array_equal(hg.final_state, truth["u"]), "graph kernel must stay bit-exact at 10k steps"
    assert np.array_equal(hg.probe_series[:, 0], truth["probes"][:, 0])
    spec = tk.LatticeSpec(size=15, dim=2)
    hs = tk.run_wave_sim(spec, steps=steps, initial_data={kick: 1.0}, dt=dt)
    idx = np.array([spec.index(n) for n in nodes])
    err = rel_inf(hs.final_state[idx], truth["u"])
    err_t = np.max(np.abs(hs.probe_series[:, 0] - truth["probes"][:, 0])) / np.max(np.abs(truth["probes"]))
    print(f"10k-step structured-vs-reference error: final {err:.2e}, probe series {err_t:.2e}")
    assert err < 1e-10 and err_t < 1e-10


def test_full_config3_on_one_gpu_uses_64_bit_indexing():
    """BASELINE config 3 whole (1024 x 1024 x 512 = 2^31 node slots, 34 GB of state) on ONE device: the
    kick and a probe sit in the top layer (linear index > INT32_MAX), and the device-side checksum
    sum_i u_t(i) = sum(u_0) (1 + t) covers every node."""
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~40 GB of free device memory")
    spec = tk.LatticeSpec(size=1024, dim=3, layers=512, defect="blackhole", radius=64.0)
    assert spec.num_slots == 2**31
    top, mid = (512, 515, 511, "C"), spec.kick_node()
    assert spec.index(top) > 2**31 - 2**22 > np.iinfo(np.int32).max - 2**22
    plan = tk.compile_lattice(spec)
    try:
        assert plan.max_degree == 3 + 2 * 1023 + 2
        plan.set_state_sparse([spec.index(top), spec.index(mid)], [-0.75, 1.0])
        steps = 6
        out = plan.run(steps, 0.0221**2, 0.0, probes=[spec.index(top), spec.index(mid), spec.index((512, 515, 510, "C"))])
        assert out["probes"][0].tolist() == [-0.75, 1.0, 0.0]
        assert abs(out["probes"][-1, 0]) > 0.1 and out["probes"][1, 2] != 0.0  # the layer below the kick moved

        class Raw:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

        n_all = (spec.layers + 2) * spec.plane
        sums = [float(torch.as_tensor(Raw(p, n_all), device="cuda:0").sum().item()) for p in plan.state_ptrs()]
        assert any(abs(s - 0.25 * (1 + steps)) < 1e-10 for s in sums), sums
    finally:
        plan.close()


@pytest.mark.parametrize("size,dim,layers,defect", [
    (1, 2, 1, "none"), (2, 2, 1, "wedge"), (3, 2, 1, "blackhole"), (1, 3, 1, "none"), (1, 3, 4, "none"),
    (2, 3, 2, "wedge"), (3, 3, 3, "singularity"), (5, 3, 1, "blackhole"), (33, 2, 1, "singularity"),
    (65, 2, 1, "blackhole"), (64, 3, 2, "wedge"),
])
def test_tiny_and_awkward_lattices_spec_route_equals_graph_route(size, dim, layers, defect):
    """Degenerate shapes (single cell, single layer 3-D, sizes around the 32/64-cell segment widths):
    the LatticeSpec route against the bit-exact graph route on LatticeSpec.to_networkx()."""
    kw = dict(size=size, dim=dim, layers=layers, defect=defect)
    if defect == "blackhole":
        kw["radius"] = 1.2
    if defect == "singularity":
        kw.update(radius=1.0, mass=30.0, potential=0.7, prune_edges=size % 2 == 1)
    spec = tk.LatticeSpec(**kw)
    G = spec.to_networkx()
    nodes = list(G.nodes)
    kick = spec.kick_node()
    other = nodes[-1]
    init = {kick: 1.0, other: -0.3} if other != kick else {kick: 1.0}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        a = tk.run_wave_sim(G, steps=30, initial_data=init, dt=0.2, damping=0.01)
        b = tk.run_wave_sim(spec, steps=30, initial_data=init, dt=0.2, damping=0.01, record="all")
    assert a.metadata["max_degree"] == b.metadata["max_degree"] and a.dt == b.dt
    idx = np.array([spec.index(n) for n in nodes])
    assert rel_inf(b.array[:, idx], a.array) < 1e-12
    assert spec.num_nodes == len(nodes)


def test_snapshots_every_k_steps_equal_the_full_history():
    spec = tk.LatticeSpec(size=40, dim=3, layers=5, defect="blackhole", radius=2.5)
    k = spec.kick_node()
    full = tk.run_wave_sim(spec, steps=23, initial_data={k: 1.0}, dt=0.05, record="all")
    snap = tk.run_wave_sim(spec, steps=23, initial_data={k: 1.0}, dt=0.05, record="all", snapshot_every=5)
    assert snap.array.shape == (5, spec.num_slots)
    assert np.array_equal(snap.array, full.array[::5])
    assert np.array_equal(snap.final_state, full.array[-1]) and np.array_equal(snap.probe_series, full.probe_series)


@pytest.mark.parametrize("reorder", ["none", "degree", "rcm", "auto"])
def test_row_reordering_keeps_the_graph_kernel_bit_exact(reorder):
    """Device row order (padding / locality optimisation) must not change a single bit."""
    g = load_golden("random_graph_multi_kick")
    G, labels = graph_from_golden(g)
    G.add_edge(5, 5)
    h = tk.run_wave_sim(G, steps=g["steps"], initial_data=_initial_data(g, labels), c=g["c"], dt=g["dt"],
                        damping=g["damping"], reorder=reorder)
    assert np.array_equal(h.array[g["snap_steps"]], g["states"])
    import networkx as nx

    B = nx.barabasi_albert_graph(4000, 3, seed=5)
    a = tk.run_wave_sim(B, steps=40, initial_data={0: 1.0, 17: -0.5}, dt=0.05, reorder="none")
    b = tk.run_wave_sim(B, steps=40, initial_data={0: 1.0, 17: -0.5}, dt=0.05, reorder=reorder, energy=True)
    assert np.array_equal(a.array, b.array)


@pytest.mark.parametrize("steps", [8, 9, 31])
def test_cuda_graph_replay_equals_direct_launches(steps, monkeypatch):
    """Launch-bound runs are replayed from a captured CUDA graph (two steps per graph, output rows through
    device counters); results must be identical to direct launches -- both kernels, probes, history, energy,
    odd step counts."""
    import networkx as nx

    spec = tk.LatticeSpec(size=50, dim=3, layers=6, defect="singularity", radius=1.5, mass=40.0, potential=2.0)
    k = spec.kick_node()
    G = nx.barabasi_albert_graph(500, 3, seed=1)

    def runs():
        a = tk.run_wave_sim(spec, steps=steps, initial_data={k: 1.0}, dt=0.05, damping=0.01, record="all",
                            probes=[k, (0, 0, 0, "A")], energy=True)
        b = tk.run_wave_sim(spec, steps=steps, initial_data={k: 1.0}, dt=0.05, probes=[k])
        c = tk.run_wave_sim(G, steps=steps, initial_data={3: 1.0}, dt=0.1, energy=True)
        d = tk.run_wave_sim(G, steps=steps, initial_data={3: 1.0}, dt=0.1, record="probes", probes=[3, 77])
        e = tk.run_wave_sim(spec, steps=steps, initial_data={k: 1.0}, dt=0.05, dtype="float32", record="all")
        return a, b, c, d, e

    monkeypatch.setenv("TK_CUDA_GRAPH", "1")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        g = runs()
        monkeypatch.setenv("TK_CUDA_GRAPH", "0")
        h = runs()
    for x, y in zip(g, h):
        for f in ("array", "probe_series", "final_state", "energy"):
            u, v = getattr(x, f), getattr(y, f)
            assert (u is None) == (v is None)
            if u is not None:
                assert np.array_equal(u, v), f
