"""K-step passes on lattices WITH defects and on 3-D lattices (the cross scheme, kernels_cross.cuh) against the oracle.

The rows / columns of the defect (and of the probes) are stepped one step at a time together with every line sum; the
rest of the lattice takes one pass per K steps.  The result must be the reference's u(t) all the same: every case
is compared with the oracle's step-by-step run (reference tetrakis_sim/physics.py:106-129 on lattice.py:30-73 +
defects.py:55-107, 110-135, 194-247)."""

import warnings

import numpy as np
import pytest

import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo
from tests.helpers import rel_inf

pytestmark = pytest.mark.gpu

DEFECTS = {
    "blackhole": dict(defect="blackhole", radius=3.3),
    "sing_tag": dict(defect="singularity", radius=2.5, mass=50.0, potential=2.0),
    "sing_prune": dict(defect="singularity", radius=2.0, mass=2000.0, prune_edges=True),
    "wedge": dict(defect="wedge"),
    "none": dict(),
}


def _oracle(spec, u0, up0, steps, coeff, damping, probes):
    n = spec.num_slots
    flags, im, pot = np.full(n, 3, np.uint8), np.ones(n), np.zeros(n)
    si, sf, sm, sv = spec.specials()
    flags[si], im[si], pot[si] = sf, sm, sv
    return wo.run_structured(spec.size, spec.layers, flags, im, pot, spec.corrections(), u0.copy(), steps, coeff,
                             damping, probes=probes, uprev0=None if up0 is None else up0.copy())


def _run(spec, steps, kicks, damping, probes, dt=0.05):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return tk.run_wave_sim(spec, steps=steps, initial_data=kicks, c=1.0, dt=dt, damping=damping, probes=probes)


def _case(spec, steps, damping, extra_probe=True):
    kick = spec.kick_node()
    far = spec.node((spec.index(kick) + 4 * (spec.size // 3) * spec.size + 4 * (spec.size // 5) + 2) % spec.plane)
    kicks = {kick: 1.0}
    if spec.contains(far) and far != kick:
        kicks[far] = -0.25
    probes = [kick] + ([far] if extra_probe and spec.contains(far) and far != kick else [])
    hist = _run(spec, steps, kicks, damping, probes)
    u0 = np.zeros(spec.num_slots)
    for n_, v in kicks.items():
        u0[spec.index(n_)] = v
    ref = _oracle(spec, u0, None, steps, (1.0 * hist.dt) ** 2, damping, [spec.index(p) for p in probes])
    return hist, ref


@pytest.fixture(autouse=True)
def _small_lattices_allowed(monkeypatch):
    monkeypatch.setenv("TK_FX_MIN_NODES", "0")
    monkeypatch.setenv("TK_FX_MAXFRAC", "95")


@pytest.mark.parametrize("K", [2, 5, 8, 16, 40])
@pytest.mark.parametrize("defect", ["blackhole", "sing_tag", "sing_prune", "wedge"])
@pytest.mark.parametrize("S,steps,damping", [(24, 40, 0.0), (33, 37, 0.01), (65, 48, 0.0), (130, 41, 0.02)])
def test_sheets_with_defects_in_fused_passes(K, defect, S, steps, damping, monkeypatch):
    monkeypatch.setenv("TK_LAT_FUSE", str(K))
    spec = tk.LatticeSpec(size=S, dim=2, **DEFECTS[defect])
    hist, ref = _case(spec, steps, damping)
    assert hist.cross["steps_per_pass"] >= 2 and hist.fused["fused_steps"] >= steps - 1
    assert 0 < hist.cross["cross_rows"] < S and 0 < hist.cross["cross_cols"] < S
    assert rel_inf(hist.final_state, ref["u"]) < 1e-12
    assert rel_inf(hist.probe_series, ref["probes"]) < 1e-12


@pytest.mark.parametrize("K", [8, 16])
@pytest.mark.parametrize("defect", ["none", "blackhole", "sing_tag", "sing_prune", "wedge"])
@pytest.mark.parametrize("S,L,steps,damping", [(20, 5, 35, 0.0), (33, 12, 40, 0.01), (40, 33, 50, 0.0), (70, 3, 27, 0.02),
                                               (16, 1 + 16, 16, 0.0), (64, 2, 24, 0.0)])
def test_three_d_lattices_in_fused_passes(K, defect, S, L, steps, damping, monkeypatch):
    monkeypatch.setenv("TK_LAT_FUSE", str(K))
    spec = tk.LatticeSpec(size=S, dim=3, layers=L, **DEFECTS[defect])
    hist, ref = _case(spec, steps, damping)
    k_expected = K if steps >= K else (8 if steps >= 8 else 0)
    assert hist.cross["steps_per_pass"] == k_expected
    fused = (steps // k_expected) * k_expected
    fused += 8 if steps - fused >= 8 else 0
    assert hist.fused["fused_steps"] == fused
    assert rel_inf(hist.final_state, ref["u"]) < 1e-12
    assert rel_inf(hist.probe_series, ref["probes"]) < 1e-12


@pytest.mark.parametrize("dim", [2, 3])
def test_fused_passes_continue_from_a_dense_state_with_a_previous_level(dim, monkeypatch):
    """u_{t-1} != 0 and a second run that continues the first: the bulk parts start from sums of both time levels."""
    monkeypatch.setenv("TK_LAT_FUSE", "8")
    spec = (tk.LatticeSpec(size=48, dim=2, **DEFECTS["blackhole"]) if dim == 2
            else tk.LatticeSpec(size=30, dim=3, layers=9, **DEFECTS["sing_prune"]))
    rng = np.random.default_rng(5)
    flags = np.full(spec.num_slots, 3, np.uint8)
    si, sf, _, _ = spec.specials()
    flags[si] = sf
    alive = (flags & 1).astype(float)
    u0 = rng.standard_normal(spec.num_slots) * alive
    up0 = (u0 + 0.01 * rng.standard_normal(spec.num_slots)) * alive
    coeff = 0.04 ** 2
    plan = tk.compile_lattice(spec)
    try:
        plan.set_state(u0, up0)
        out = plan.run(24, coeff, 0.01, probes=[3, 1000])
        assert plan.cross["steps_per_pass"] == 8 and plan.fused["fused_steps"] == 24
        got, gotp = plan.get_state(want_prev=True)
        out2 = plan.run(19, coeff, 0.01, probes=[3])
        got2 = plan.get_state()
    finally:
        plan.close()
    ref = _oracle(spec, u0, up0, 24, coeff, 0.01, [3, 1000])
    assert rel_inf(got, ref["u"]) < 1e-12 and rel_inf(gotp, ref["uprev"]) < 1e-12
    assert rel_inf(out["probes"], ref["probes"]) < 1e-12
    ref2 = _oracle(spec, ref["u"], ref["uprev"], 19, coeff, 0.01, [3])
    assert rel_inf(got2, ref2["u"]) < 1e-12
    assert rel_inf(out2["probes"], ref2["probes"]) < 1e-12


@pytest.mark.parametrize("resync", ["0", "1", "2"])
def test_bulk_parts_follow_their_recurrence_or_are_summed_from_the_data(resync, monkeypatch):
    """3-D: B / D are carried from pass to pass by their recurrence; TK_FX_RESYNC re-anchors them on the data."""
    monkeypatch.setenv("TK_LAT_FUSE", "8")
    monkeypatch.setenv("TK_FX_RESYNC", resync)
    spec = tk.LatticeSpec(size=36, dim=3, layers=10, **DEFECTS["blackhole"])
    hist, ref = _case(spec, 80, 0.0)
    assert hist.fused["fused_steps"] == 80
    assert rel_inf(hist.final_state, ref["u"]) < 1e-12
    assert rel_inf(hist.probe_series, ref["probes"]) < 1e-12


def test_fused_passes_equal_the_step_by_step_kernel_on_a_medium_lattice(monkeypatch):
    """3-D 96 x 96 x 40 with the black hole: K = 16 passes against one step per launch, 100 steps."""
    spec = tk.LatticeSpec(size=96, dim=3, layers=40, defect="blackhole", radius=9.0)
    kick = spec.kick_node()
    probes = [kick, spec.node(5), spec.node(spec.num_slots - 2)]
    monkeypatch.setenv("TK_LAT_FUSE", "0")
    a = _run(spec, 100, {kick: 1.0}, 0.0, probes)
    assert a.fused["steps_per_pass"] == 0 and a.cross["steps_per_pass"] == 0
    monkeypatch.setenv("TK_LAT_FUSE", "16")
    b = _run(spec, 100, {kick: 1.0}, 0.0, probes)
    assert b.cross["steps_per_pass"] == 16 and b.fused["fused_steps"] == 96
    assert rel_inf(b.final_state, a.final_state) < 1e-12
    assert rel_inf(b.probe_series, a.probe_series) < 1e-12
    assert np.all(b.final_state[spec.specials()[0]] == 0.0)  # removed nodes stay exactly zero


@pytest.mark.parametrize("name,kw", [
    ("cfg4", dict(size=512, dim=3, layers=512, defect="singularity", radius=2.5, mass=2000.0, prune_edges=True)),
    ("cfg3-slab", dict(size=1024, dim=3, layers=64, defect="blackhole", radius=64.0)),
])
def test_full_size_lattices_in_fused_passes_against_the_step_by_step_kernel(name, kw):
    """BASELINE configs[3] whole (512^3 singularity, pruned edges: 537 M nodes) and the per-GPU slab of configs[2]
    (1024^2 x 64, black hole r = 64): four K = 16 passes of the cross scheme against 64 single steps of lat_step_kernel
    (itself oracle-checked at every size the oracle finishes) -- every final value and two probe series -- plus
    sum_i u_t(i) = 1 + t and removed / pruned nodes exactly at rest."""
    spec = tk.LatticeSpec(**kw)
    steps = 64
    kick = spec.index(spec.kick_node())
    far = spec.index((3, spec.size - 5, spec.layers - 2, "D"))
    plan = tk.compile_lattice(spec)
    try:
        dt = min(0.1, 1.0 / np.sqrt(plan.max_degree))
        coeff = dt * dt
        plan.set_state_sparse([kick], [1.0])
        a = plan.run(steps, coeff, 0.0, probes=[kick, far])
        assert plan.cross["steps_per_pass"] == 16 and plan.fused["fused_steps"] == steps, (plan.cross, plan.fused)
        ua = plan.get_state()
        plan.set_state_sparse([kick], [1.0])
        b = plan.run(steps, coeff, 0.0, probes=[kick, far], fuse=False)
        ub = plan.get_state()
    finally:
        plan.close()
    assert rel_inf(a["probes"], b["probes"]) < 1e-12
    scale = float(np.max(np.abs(ub)))
    ua -= ub
    assert float(np.max(np.abs(ua))) < 1e-12 * scale
    assert abs(float(ub.sum()) - (steps + 1.0)) < 1e-9 * (steps + 1.0)
    assert np.all(ub[spec.specials()[0]] == 0.0)


def test_the_cross_scheme_is_not_used_where_it_does_not_apply(monkeypatch):
    monkeypatch.setenv("TK_LAT_FUSE", "8")
    spec = tk.LatticeSpec(size=24, dim=3, layers=6, **DEFECTS["blackhole"])
    kick = spec.kick_node()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        full = tk.run_wave_sim(spec, steps=16, initial_data={kick: 1.0}, dt=0.05, record="all")
        en = tk.run_wave_sim(spec, steps=16, initial_data={kick: 1.0}, dt=0.05, energy=True)
        many = tk.run_wave_sim(spec, steps=16, initial_data={kick: 1.0}, dt=0.05,
                               probes=[spec.node(i) for i in range(0, 4 * 17, 4)])
        f32 = tk.run_wave_sim(spec, steps=16, initial_data={kick: 1.0}, dt=0.05, dtype="float32")
    for h in (full, en, many, f32):
        assert h.cross["steps_per_pass"] == 0
    monkeypatch.setenv("TK_FX_MAXFRAC", "10")  # the hole's rows and columns cover far more than 10 % of 24 x 24
    h = _run(spec, 16, {kick: 1.0}, 0.0, [kick])
    assert h.cross["steps_per_pass"] == 0
    monkeypatch.setenv("TK_FX_MAXFRAC", "95")
    monkeypatch.setenv("TK_FX_MIN_NODES", str(1 << 21))  # the default: small lattices are bound by their launches
    h = _run(spec, 16, {kick: 1.0}, 0.0, [kick])
    assert h.cross["steps_per_pass"] == 0
