"""GPU: the SM-resident ensemble kernel against the golden vectors and the single-run kernels."""

import warnings

import numpy as np
import pytest

import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo
from tests.helpers import load_golden, rel_inf
from tests.test_cpu_host import _spec_from

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cfg1_s11l7_blackhole_r2p5", "cfg5_s13l7_none",
                                  "cfg4small_s13l7_sing_m2000_prune", "t3d_s7l5_wedge",
                                  "t3d_s7l5_blackhole_cyl", "t3d_s7l5_sing_prune_v3",
                                  "t2d_s7_sing_m50_v25", "t2d_s7_blackhole_r2", "t2d_s7_wedge"])
def test_ensemble_member_reproduces_the_reference_run(name):
    g = load_golden(name)
    spec = _spec_from(g["lattice"])
    res = tk.run_ensemble([spec, spec], g["steps"], c=g["c"], dt=g["dt"], damping=[g["damping"], 0.5])
    lat = g["lattice"]
    assert [*res.kick_nodes[0][:-1], "ABCD".index(res.kick_nodes[0][-1])] == lat["kick"]
    assert res.effective_dt[0] == g["effective_dt"]
    assert res.max_degree[0] == g["metadata"]["max_degree"]
    assert bool(res.stability_adjusted[0]) == g["metadata"]["stability_adjusted"]
    assert rel_inf(res.series[0], g["probe_values"]) < 1e-12
    assert rel_inf(res.spectrum[0], g["fft_spectrum"]) < 1e-11
    assert wo.folded_bin(res.spectrum[0]) == wo.folded_bin(g["fft_spectrum"])
    assert abs(res.dominant_freq[0]) == pytest.approx(abs(g["dominant_freq"]), rel=1e-12)
    assert res.dominant_bin[0] == int(np.argmax(res.spectrum[0]))
    assert not np.allclose(res.series[1], res.series[0])  # the heavily damped twin differs


def test_sweep_grid_matches_individual_runs():
    """A small radius x damping x dt grid (BASELINE config 5 shape at reduced counts): every member
    equals the single-run structured kernel, including the per-simulation CFL clamp."""
    radii, dampings, dts = [0.5, 1.7, 2.5, 3.3], [0.0, 0.02, 0.05], [0.02, 0.2]
    specs, damp, dt = tk.sweep_grid(13, 7, radii, dampings, dts)
    res = tk.run_ensemble(specs, 50, c=1.0, dt=dt, damping=damp)
    assert res.series.shape == (24, 51) and res.node_updates > 0
    assert res.stability_adjusted.tolist() == [d > 0.18 for d in dt]
    rng = np.random.default_rng(0)
    for i in rng.choice(24, 6, replace=False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            h = tk.run_wave_sim(specs[i], steps=50, initial_data={res.kick_nodes[i]: 1.0}, c=1.0,
                                dt=float(dt[i]), damping=float(damp[i]))
        assert h.dt == res.effective_dt[i]
        assert rel_inf(res.series[i], h.probe_series[:, 0]) < 1e-12
        f, sp, _ = tk.run_fft(h, res.kick_nodes[i])
        assert rel_inf(res.spectrum[i], sp) < 1e-11
        assert wo.folded_bin(res.spectrum[i]) == wo.folded_bin(sp)


def test_ensemble_argument_errors():
    with pytest.raises(ValueError):
        tk.run_ensemble([tk.LatticeSpec(size=5, dim=2), tk.LatticeSpec(size=6, dim=2)], 5)
    with pytest.raises(ValueError, match="not in the graph"):
        tk.run_ensemble([tk.LatticeSpec(size=7, dim=2, defect="blackhole", radius=2.0)], 5,
                        kicks=[(3, 3, "A")])
    with pytest.raises(ValueError, match="shared memory"):
        tk.run_ensemble([tk.LatticeSpec(size=64, dim=3, layers=8)], 5)


def test_config5_grid_spot_check_against_the_reference():
    """64 random points (seed 0) of BASELINE config 5's 64 x 32 x 32 grid, each run
    through the unmodified reference (tests/golden/cfg5_spotcheck.npz) vs the ensemble kernel running
    the WHOLE 65 536-member grid in one launch."""
    import os

    from tests.helpers import GOLDEN

    z = np.load(os.path.join(GOLDEN, "cfg5_spotcheck.npz"))
    radii, dampings, dts = np.linspace(0.5, 4.0, 64), np.linspace(0.0, 0.05, 32), np.linspace(0.02, 0.2, 32)
    specs, damp, dt = tk.sweep_grid(13, 7, radii, dampings, dts)
    res = tk.run_ensemble(specs, 50, c=1.0, dt=dt, damping=damp)
    assert res.series.shape == (65536, 51)
    picks = z["picks"]
    assert np.array_equal(res.effective_dt[picks], z["effective_dt"])
    for i, g in enumerate(picks):
        k = res.kick_nodes[g]
        assert [*k[:-1], "ABCD".index(k[-1])] == z["kicks"][i].tolist()
        assert rel_inf(res.series[g], z["series"][i]) < 1e-12
        assert rel_inf(res.spectrum[g], z["spectra"][i]) < 1e-11
        assert wo.folded_bin(res.spectrum[g]) == wo.folded_bin(z["spectra"][i])


@pytest.mark.parametrize("steps", [40, 51, 7])
def test_device_features_match_the_host_restatement(steps):
    """The ensemble kernel's epilogue computes the spectral / time-series features of tetrakis-eval generate
    (reference evals/features.py:11-81) on the device; tetrakis_sim_b200.features is the NumPy restatement of the same
    definitions.  Odd and even series lengths (the non-negative half has (n+1)//2 bins), damped and undamped members."""
    from tetrakis_sim_b200.features import spectral_features, timeseries_features

    radii, dampings, dts = [0.5, 1.7, 2.5], [0.0, 0.03], [0.05, 0.2]
    specs, damp, dt = tk.sweep_grid(11, 5, radii, dampings, dts)
    res = tk.run_ensemble(specs, steps, c=1.0, dt=dt, damping=damp, features=True)
    only = tk.run_ensemble(specs, steps, c=1.0, dt=dt, damping=damp, features=True, series=False, spectrum=False)
    assert only.series is None and only.spectrum is None and np.array_equal(only.features, res.features)
    for i in range(len(specs)):
        freq = np.fft.fftfreq(steps + 1, d=float(res.effective_dt[i]))
        want = spectral_features(freq, res.spectrum[i])
        want.update(timeseries_features(res.series[i]))
        got = res.feature_dict(i)
        assert set(got) == set(want)
        for k, v in want.items():
            if k.startswith("peak_i_"):
                assert got[k] == v, (i, k)
            else:
                assert got[k] == pytest.approx(v, rel=1e-11, abs=1e-13), (i, k)
        assert abs(got["dominant_freq"]) == pytest.approx(abs(res.dominant_freq[i]), rel=1e-12)
