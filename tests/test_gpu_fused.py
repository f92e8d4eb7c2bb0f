"""K-step fused passes on defect-free 2-D sheets (kernels_fused.cuh) against the oracle.

The clique sums of a sheet without defects obey closed recurrences, so the device advances K steps per
pass; the result must be the reference's u(t) all the same (SURVEY.md 8d config 2 is such a sheet)."""

import warnings

import numpy as np
import pytest

import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo
from tests.helpers import rel_inf

pytestmark = pytest.mark.gpu


def _oracle(S, u0, up0, steps, coeff, damping, probes):
    n = S * S * 4
    flags = np.full(n, 3, dtype=np.uint8)
    return wo.run_structured(S, 1, flags, np.ones(n), np.zeros(n), [], u0.copy(), steps, coeff, damping,
                             probes=probes, uprev0=None if up0 is None else up0.copy())


def _default_kmax(S, damping):
    """Mirror of csrc/host_fused.inc:fuse_choice: longer passes on sheets that are launch-bound; the FP64 and HBM
    rates are the ones the library derives from the device's attributes (tk_device_rates)."""
    rates = tk.device_rates(0)
    ncell = float(S) * S
    t_fp = ncell * (4.0 if damping == 0.0 else 8.0) / rates["fp64_per_s"]
    t_pass = ncell * 128.0 / rates["hbm_bytes_per_s"] + 40e-6
    k = int(0.6 * t_pass / t_fp)
    return max(48 if damping == 0.0 else 24, min(k, 1024)) // 8 * 8


def _expected(steps, kmax):
    """(steps per pass, steps covered by passes): passes of nearly equal length, a single left-over step runs
    as an ordinary step."""
    if kmax < 2 or steps < 2:
        return 0, 0
    npass = -(-steps // kmax)
    k = -(-steps // npass)
    if k > 8 and -(-k // 8) * 8 <= kmax:
        k = -(-k // 8) * 8
    k = min(k, steps)
    rem = steps - (steps // k) * k
    return k, steps - 1 if rem == 1 else steps


def _run(spec, steps, kicks, damping, probes, dt=0.05):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return tk.run_wave_sim(spec, steps=steps, initial_data=kicks, c=1.0, dt=dt, damping=damping, probes=probes)


@pytest.mark.parametrize("K", [2, 3, 8, 16, 29, 48, 64])
@pytest.mark.parametrize("S,steps,damping", [(1, 33, 0.0), (7, 40, 0.02), (31, 50, 0.0), (33, 37, 0.01),
                                             (64, 64, 0.0), (150, 45, 0.03), (257, 32, 0.0)])
def test_fused_passes_match_the_oracle(K, S, steps, damping, monkeypatch):
    monkeypatch.setenv("TK_LAT_FUSE", str(K))
    spec = tk.LatticeSpec(size=S, dim=2)
    kick = spec.kick_node()
    far = spec.node((spec.index(kick) + 4 * (S // 3) * S + 4 * (S // 5) + 2) % spec.num_slots)
    last = spec.node(spec.num_slots - 1)
    kicks = {kick: 1.0, far: -0.25} if far != kick else {kick: 1.0}
    probes = [kick, far, last, spec.node(0)]
    hist = _run(spec, steps, kicks, damping, probes)
    assert (hist.fused["steps_per_pass"], hist.fused["fused_steps"]) == _expected(steps, K)
    u0 = np.zeros(spec.num_slots)
    for n_, v in kicks.items():
        u0[spec.index(n_)] = v
    ref = _oracle(S, u0, None, steps, (1.0 * hist.dt) ** 2, damping, [spec.index(p) for p in probes])
    assert rel_inf(hist.final_state, ref["u"]) < 1e-12
    assert rel_inf(hist.probe_series, ref["probes"]) < 1e-12


@pytest.mark.parametrize("graph", ["0", "1"])
def test_fused_passes_equal_the_step_by_step_kernel(graph, monkeypatch):
    """Same sheet, fused vs one step per launch (direct launches and CUDA-graph replay of the pass)."""
    monkeypatch.setenv("TK_CUDA_GRAPH", graph)
    spec = tk.LatticeSpec(size=200, dim=2)
    kick = spec.kick_node()
    probes = [kick, spec.node(5), spec.node(spec.num_slots - 2)]
    monkeypatch.setenv("TK_LAT_FUSE", "0")
    a = _run(spec, 203, {kick: 1.0}, 0.005, probes)
    assert a.fused["steps_per_pass"] == 0
    monkeypatch.setenv("TK_LAT_FUSE", "8")
    b = _run(spec, 203, {kick: 1.0}, 0.005, probes)
    assert (b.fused["steps_per_pass"], b.fused["fused_steps"]) == _expected(203, 8) == (8, 203)
    assert rel_inf(b.final_state, a.final_state) < 1e-12
    assert rel_inf(b.probe_series, a.probe_series) < 1e-12
    assert b.launches < a.launches


def test_fused_passes_continue_from_a_dense_state_with_a_previous_level():
    """u_{t-1} != 0: the recurrences start from the sums of both time levels."""
    S, steps = 48, 41
    spec = tk.LatticeSpec(size=S, dim=2)
    rng = np.random.default_rng(5)
    u0 = rng.standard_normal(spec.num_slots)
    up0 = u0 + 0.01 * rng.standard_normal(spec.num_slots)
    coeff = 0.04 ** 2
    plan = tk.compile_lattice(spec)
    try:
        plan.set_state(u0, up0)
        out = plan.run(steps, coeff, 0.01, probes=[3, 1000])
        assert (plan.fused["steps_per_pass"], plan.fused["fused_steps"]) == _expected(41, _default_kmax(S, 0.01)) == (41, 41)
        got = plan.get_state()
        out2 = plan.run(16, coeff, 0.01, probes=[3])  # a second call continues the same simulation
        got2 = plan.get_state()
    finally:
        plan.close()
    ref = _oracle(S, u0, up0, steps, coeff, 0.01, [3, 1000])
    assert rel_inf(got, ref["u"]) < 1e-12
    assert rel_inf(out["probes"], ref["probes"]) < 1e-12
    ref2 = _oracle(S, u0, up0, steps + 16, coeff, 0.01, [3])
    assert rel_inf(got2, ref2["u"]) < 1e-12
    assert rel_inf(out2["probes"], ref2["probes"][steps:]) < 1e-12


def test_fusion_is_not_used_where_the_recurrences_do_not_close(monkeypatch):
    """Small lattices with defects / in 3-D (the cross scheme of kernels_cross.cuh starts at 2^21 nodes:
    tests/test_gpu_cross.py), full history, energy series: step by step."""
    monkeypatch.setenv("TK_LAT_FUSE", "8")
    hole = tk.LatticeSpec(size=40, dim=2, defect="blackhole", radius=3.0)
    assert _run(hole, 32, None, 0.0, None).fused["steps_per_pass"] == 0
    wedge = tk.LatticeSpec(size=40, dim=2, defect="wedge")
    assert _run(wedge, 32, None, 0.0, None).fused["steps_per_pass"] == 0
    cube = tk.LatticeSpec(size=16, dim=3, layers=4)
    assert _run(cube, 32, None, 0.0, None).fused["steps_per_pass"] == 0
    sheet = tk.LatticeSpec(size=40, dim=2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        full = tk.run_wave_sim(sheet, steps=32, dt=0.05, record="all")  # every state: nothing to fuse
        en = tk.run_wave_sim(tk.LatticeSpec(size=256, dim=2), steps=32, dt=0.05, energy=True)
    assert full.fused["steps_per_pass"] == 0 and en.fused["steps_per_pass"] == 0
    assert _run(sheet, 32, None, 0.0, None).fused["steps_per_pass"] == 8
    assert _run(sheet, 1, None, 0.0, None).fused["steps_per_pass"] == 0
    many = [sheet.node(i) for i in range(65)]  # more probes than a pass records (64)
    assert _run(sheet, 32, None, 0.0, many).fused["steps_per_pass"] == 0


def test_fused_long_run_stays_within_the_parity_bar():
    """2 000 steps at the clamped (marginally stable) dt on 256^2: 1e-10 relative on u and the probe."""
    S, steps = 256, 2000
    spec = tk.LatticeSpec(size=S, dim=2)
    kick = spec.kick_node()
    hist = _run(spec, steps, {kick: 1.0}, 0.0, [kick], dt=0.1)
    assert hist.metadata["stability_adjusted"] and hist.fused["steps_per_pass"] == _expected(steps, _default_kmax(S, 0.0))[0] == 1000
    u0 = np.zeros(spec.num_slots)
    u0[spec.index(kick)] = 1.0
    ref = _oracle(S, u0, None, steps, (1.0 * hist.dt) ** 2, 0.0, [spec.index(kick)])
    assert rel_inf(hist.final_state, ref["u"]) < 1e-10
    assert rel_inf(hist.probe_series, ref["probes"]) < 1e-10
    freq, spectrum, _ = tk.run_fft(hist, kick)
    assert wo.folded_bin(spectrum) == wo.folded_bin(wo.run_fft(ref["probes"][:, 0], hist.dt)[1])


def test_full_size_config_2_sheet_against_the_oracle():
    """BASELINE configs[1] at its full size (8192^2 = 268 M nodes): two fused passes of 48 steps against the C
    oracle stepping one step at a time, all 268 M final values and the probe series."""
    S, steps = 8192, 96
    spec = tk.LatticeSpec(size=S, dim=2)
    kick = spec.bench_kick_node()
    far = spec.node(spec.index(kick) + 4 * 1000 * S + 4 * 777 + 1)
    hist = _run(spec, steps, {kick: 1.0}, 0.0, [kick, far], dt=0.1)
    assert hist.metadata["stability_adjusted"] and hist.fused == {**hist.fused, "steps_per_pass": 48, "fused_steps": 96}
    n = spec.num_slots
    u0 = np.zeros(n)
    u0[spec.index(kick)] = 1.0
    ref = wo.run_structured(S, 1, None, None, None, [], u0, steps, (1.0 * hist.dt) ** 2, 0.0,
                            probes=[spec.index(kick), spec.index(far)])
    assert rel_inf(hist.probe_series, ref["probes"]) < 1e-12
    scale = np.max(np.abs(ref["u"]))
    assert np.max(np.abs(hist.final_state - ref["u"])) < 1e-12 * scale
    assert abs(hist.final_state.sum() - (1.0 + steps)) < 1e-8  # sum_i u_t(i) = 1 + t


def test_config_2_at_full_size_and_full_step_count_by_properties():
    """BASELINE configs[1] whole: 8192^2 nodes AND 10 000 steps (no CPU oracle finishes that: 2.7e12 node updates), checked
    through properties that hold at any size --
      (i)  sum_i u_t(i) = 1 + t after the 10 000 steps (zero column sums of the Laplacian, u_{-1} = 0);
      (ii) time reversal: with (u, u_prev) swapped, 10 000 more steps retrace the trajectory back to the unit kick
           (the leapfrog map is exactly reversible in exact arithmetic, so what is left is 20 000 steps of rounding);
      (iii) the kick node's series is the same as on the step-by-step kernel for the first 96 steps (the oracle-checked
            range of test_full_size_config_2_sheet_against_the_oracle) -- recorded inside the long run."""
    S, steps = 8192, 10_000
    spec = tk.LatticeSpec(size=S, dim=2)
    kick = spec.index(spec.bench_kick_node())
    dt = min(0.1, 1.0 / np.sqrt(2 * S + 1))  # physics.py:74-80 clamp
    coeff = dt * dt
    plan = tk.compile_lattice(spec)
    try:
        plan.set_state_sparse([kick], [1.0])
        out = plan.run(steps, coeff, 0.0, probes=[kick])
        assert plan.fused["steps_per_pass"] == 48 and plan.fused["fused_steps"] == steps, plan.fused
        u, up = plan.get_state(want_prev=True)
        total = float(u.sum())
        assert abs(total - (steps + 1.0)) < 1e-8 * (steps + 1.0), total
        assert np.isfinite(u).all() and np.max(np.abs(u)) <= 1.0 + 1e-9  # marginally stable dt: no growth
        plan.set_state(up, u)
        del u, up
        plan.run(steps, coeff, 0.0)
        back, back_prev = plan.get_state(want_prev=True)
        assert back_prev[kick] == pytest.approx(1.0, abs=1e-7)
        back_prev[kick] -= 1.0
        err = max(float(np.max(np.abs(back))), float(np.max(np.abs(back_prev))))
        print(f"cfg2 full size, 2 x {steps} steps: sum(u) - (1 + t) = {total - (steps + 1.0):.3e}, time-reversal "
              f"residual {err:.3e}")
        assert err < 1e-7, err
        del back, back_prev
        plan.set_state_sparse([kick], [1.0])  # zeroes both time levels first
        ref = plan.run(96, coeff, 0.0, probes=[kick], fuse=False)
    finally:
        plan.close()
    assert rel_inf(out["probes"][:97], ref["probes"]) < 1e-12


def test_fused_passes_record_eight_probes_sharing_cells_rows_and_warps():
    """Probes in one cell (all four quadrants), the same node twice, neighbours in one warp segment, first and
    last row: the pass records each from the lane that owns its cell."""
    S, steps = 100, 53
    spec = tk.LatticeSpec(size=S, dim=2)
    kick = spec.kick_node()
    ki = spec.index(kick)
    cell = ki - ki % 4
    pidx = [cell, cell + 1, cell + 2, cell + 3, cell + 1, cell + 4, 2, spec.num_slots - 3]
    probes = [spec.node(i) for i in pidx]
    hist = _run(spec, steps, {kick: 1.0, spec.node(cell + 6): 0.5}, 0.01, probes)
    assert (hist.fused["steps_per_pass"], hist.fused["fused_steps"]) == _expected(steps, _default_kmax(S, 0.01)) == (53, 53)
    u0 = np.zeros(spec.num_slots)
    u0[ki], u0[cell + 6] = 1.0, 0.5
    ref = _oracle(S, u0, None, steps, (1.0 * hist.dt) ** 2, 0.01, pidx)
    assert rel_inf(hist.probe_series, ref["probes"]) < 1e-12
    assert rel_inf(hist.final_state, ref["u"]) < 1e-12


def test_fused_passes_record_a_whole_row_of_probes():
    """64 probes (a pass records up to 64; every CTA filters the list down to its own tile): one per 25 nodes along
    a diagonal band of the sheet, so they fall into many different CTAs, rows and warps."""
    S, steps = 200, 45
    spec = tk.LatticeSpec(size=S, dim=2)
    kick = spec.kick_node()
    pidx = [((3 * i) % S * S + (7 * i + 5) % S) * 4 + i % 4 for i in range(64)]
    hist = _run(spec, steps, {kick: 1.0}, 0.0, [spec.node(i) for i in pidx])
    assert hist.fused["fused_steps"] == steps
    u0 = np.zeros(spec.num_slots)
    u0[spec.index(kick)] = 1.0
    ref = _oracle(S, u0, None, steps, (1.0 * hist.dt) ** 2, 0.0, pidx)
    assert rel_inf(hist.probe_series, ref["probes"]) < 1e-12
    assert rel_inf(hist.final_state, ref["u"]) < 1e-12


@pytest.mark.parametrize("stride,steps", [(16, 100), (96, 200), (7, 30), (53, 120)])
def test_snapshots_every_k_steps_ride_on_fused_passes(stride, steps, monkeypatch):
    """record="all" with snapshot_every=s: passes whose length divides s; a prime s larger than K falls back."""
    monkeypatch.setenv("TK_LAT_FUSE", "48")
    S = 90
    spec = tk.LatticeSpec(size=S, dim=2)
    kick = spec.kick_node()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        hist = tk.run_wave_sim(spec, steps=steps, initial_data={kick: 1.0}, c=1.0, dt=0.05, record="all",
                               snapshot_every=stride, probes=[kick])
    want = {16: 16, 96: 48, 7: 7, 53: 0}[stride]
    assert hist.fused["steps_per_pass"] == want
    u0 = np.zeros(spec.num_slots)
    u0[spec.index(kick)] = 1.0
    n = spec.num_slots
    full = wo.run_structured_numpy(S, 1, np.full(n, 3, np.uint8), np.ones(n), np.zeros(n), [], u0, steps,
                                   (1.0 * hist.dt) ** 2, 0.0)
    assert hist.array.shape == (steps // stride + 1, n)
    assert rel_inf(hist.array, full[::stride]) < 1e-12
    assert rel_inf(hist.probe_series[:, 0], full[:, spec.index(kick)]) < 1e-12
    assert rel_inf(hist.final_state, full[-1]) < 1e-12
