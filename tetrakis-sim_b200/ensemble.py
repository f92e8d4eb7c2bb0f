"""Batched ensembles: many independent small-lattice simulations in one launch.

The reference runs parameter sweeps as one Python process per point
(``examples/naked_singularity_sweep.py:211-212``) and builds its eval dataset in a serial loop
(``tetrakis_sim/evals/dataset.py:81-163``).  Here every simulation is one CTA of
``ensemble_kernel`` (state resident in shared memory, probe DFT fused into the epilogue) behind
``tk_ensemble_run`` in include/tetrakis_b200.h.  Each simulation is exactly what
``run_wave_sim(G, steps, {kick: 1.0}, c, dt, damping)`` + ``run_fft(history, kick)`` compute for the
graph ``apply_defect(build_sheet(...))``, including the per-simulation CFL clamp
(``physics.py:74-91``; no warning is emitted per simulation -- ``stability_adjusted`` is returned).
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from .lattice import LatticeSpec


@dataclass
class EnsembleResult:
    series: np.ndarray  # (n_sims, steps+1) probe (= kick node) time series
    spectrum: np.ndarray  # (n_sims, steps+1) |DFT|
    dominant_bin: np.ndarray  # (n_sims,) np.argmax(spectrum)
    dominant_freq: np.ndarray  # (n_sims,) freq[argmax] with freq = fftfreq(steps+1, d=dt_eff)
    effective_dt: np.ndarray  # (n_sims,)
    stability_adjusted: np.ndarray  # (n_sims,) bool
    max_degree: np.ndarray  # (n_sims,)
    kick_nodes: list  # reference node tuples
    elapsed_ms: float  # device time of the ensemble kernel
    node_updates: int  # surviving nodes x steps summed over simulations
    features: np.ndarray | None = None  # (n_sims, 18) device-computed features (tk_ensemble_run_ex), or None

    FEATURE_NAMES = ("dominant_freq", "spectral_centroid", "spectral_bandwidth", "peak_amplitude", "low_high_ratio",
                     "peak_i_0", "peak_i_1", "peak_i_2", "peak_i_3", "peak_i_4",
                     "peak_mag_0", "peak_mag_1", "peak_mag_2", "peak_mag_3", "peak_mag_4", "ts_rms", "ts_peak")

    def feature_dict(self, i: int) -> dict:
        """Record ``i`` with the reference's feature names and units (evals/features.py:11-81): the kernel reports
        frequencies in bins, f = bin / ((steps+1) * dt_eff)."""
        f = self.features[i]
        nt = self.series.shape[1] if self.series is not None else self.nt
        per_bin = 1.0 / (nt * float(self.effective_dt[i]))
        out = {k: float(f[j]) for j, k in enumerate(self.FEATURE_NAMES)}
        if f[17] > 0.0:
            for k in ("dominant_freq", "spectral_centroid", "spectral_bandwidth"):
                out[k] *= per_bin
        return out


def _spec_key(s: LatticeSpec):
    return (s.size, s.dim, s.layers, s.defect, s.center, float(s.radius), float(s.mass),
            float(s.potential), bool(s.prune_edges), s.wedge_layer)  # fmt: skip


def mask_arrays(spec: LatticeSpec, want_degrees: bool = False):
    """(cell_bits[L*S*S] uint8, exceptions (node, inv_mass, potential), corrections, max_degree,
    surviving node count) of one lattice+defect -- the per-mask inputs of ``tk_ensemble_run``."""
    ck = ("mask_arrays", bool(want_degrees))
    if ck in spec._cache:  # a sweep is usually run more than once over the same specs
        return spec._cache[ck]
    S, L = spec.size, spec.layers
    alive = np.ones(spec.num_slots, dtype=bool)
    part = np.ones(spec.num_slots, dtype=bool)
    idx, flags, im, pot = spec.specials()
    alive[idx] = (flags & 1) != 0
    part[idx] = (flags & 2) != 0
    part &= alive
    A, P = alive.reshape(L, S, S, 4), part.reshape(L, S, S, 4)
    w = np.array([1, 2, 4, 8], dtype=np.uint16)
    bits = ((P * w).sum(-1) | ((A * w).sum(-1) << 4)).astype(np.uint8).reshape(-1)
    Pi = P.astype(np.int64)
    Pz = np.zeros((L + 2, S, S, 4), dtype=np.int64)
    Pz[1:-1] = Pi
    deg = Pi * ((Pi.sum(3, keepdims=True) - 1) + (Pi.sum(2, keepdims=True) - 1)
                + (Pi.sum(1, keepdims=True) - 1) + Pz[:-2] + Pz[2:])  # fmt: skip
    deg = deg.reshape(-1)
    corr = spec.corrections()
    for a, _, s in corr:
        deg[a] += int(s)
    max_degree = int(deg[alive].max()) if alive.any() else 0
    keep = alive[idx] & ((im != 1.0) | (pot != 0.0))
    exc = (idx[keep].astype(np.int32), im[keep], pot[keep])
    out = (bits, exc, corr, max_degree, int(alive.sum()), deg, alive) if want_degrees else \
          (bits, exc, corr, max_degree, int(alive.sum()))
    spec._cache[ck] = out
    return out


def run_ensemble(specs, steps: int, c=1.0, dt=0.2, damping=0.0, kicks=None, device: int = 0,
                 spectrum: bool = True, features: bool = False, series: bool = True) -> EnsembleResult:  # fmt: skip
    """Run ``len(specs)`` independent simulations.  ``c``, ``dt``, ``damping`` are scalars or
    per-simulation arrays; ``kicks`` optional list of node tuples (default: the batch.py kick rule).
    All specs must share (size, dim, layers); identical defects share one device mask.
    ``features=True``: the spectral / time-series features of evals/features.py computed on the device
    (``result.features``, ``result.feature_dict(i)``); with ``series=False, spectrum=False`` only those 18 numbers per
    simulation come back to the host."""
    specs = list(specs)
    n = len(specs)
    if n == 0:
        raise ValueError("no simulations")
    # sweeps hold a few distinct LatticeSpec objects referenced many times (sweep_grid: 64 specs for 65 536
    # simulations): everything per-spec is done once per OBJECT, the per-simulation part is array indexing
    sid = list(map(id, specs))
    distinct: dict = {}
    for k, s in zip(sid, specs):
        if k not in distinct:
            distinct[k] = s
    S, L, dim = specs[0].size, specs[0].layers, specs[0].dim
    if any((s.size, s.layers, s.dim) != (S, L, dim) for s in distinct.values()):
        raise ValueError("all simulations of an ensemble must share the lattice shape")
    c = np.broadcast_to(np.asarray(c, dtype=np.float64), (n,))
    dt = np.broadcast_to(np.asarray(dt, dtype=np.float64), (n,))
    damping = np.ascontiguousarray(np.broadcast_to(np.asarray(damping, dtype=np.float64), (n,)))

    masks: dict = {}
    mask_data = []
    mask_kick = []  # per mask: (kick node by the batch.py rule, its index)
    obj_mask = {}
    for k, s in distinct.items():
        key = _spec_key(s)
        if key not in masks:
            masks[key] = len(mask_data)
            mask_data.append(mask_arrays(s))
            node = s.kick_node() if kicks is None else None
            if node is not None and not s.contains(node):
                raise ValueError(f"kick node {node} is not in the graph after defect application")
            mask_kick.append((node, s.index(node) if node is not None else -1))
        obj_mask[k] = masks[key]
    sim_mask = np.fromiter(map(obj_mask.__getitem__, sid), dtype=np.int32, count=n)
    if kicks is None:
        sim_kick = np.array([i for _, i in mask_kick], dtype=np.int32)[sim_mask]
        kick_nodes = list(map([nd for nd, _ in mask_kick].__getitem__, sim_mask.tolist()))
    else:
        sim_kick = np.empty(n, dtype=np.int32)
        kick_nodes = []
        for i, s in enumerate(specs):
            node = kicks[i]
            if not s.contains(node):
                raise ValueError(f"kick node {node} is not in the graph after defect application")
            kick_nodes.append(node)
            sim_kick[i] = s.index(node)
    maxdeg = np.array([m[3] for m in mask_data], dtype=np.int64)[sim_mask]
    # physics.py:74-91 per simulation
    eff = dt.astype(np.float64).copy()
    limit = np.where(maxdeg > 0, 1.0 / np.sqrt(np.maximum(maxdeg, 1)), np.inf)
    clamp = (maxdeg > 0) & (c > 0.0) & (c * eff > limit)
    eff[clamp] = limit[clamp] / c[clamp]
    coeff = np.ascontiguousarray((c * eff) ** 2)

    n_masks = len(mask_data)
    bits = np.ascontiguousarray(np.stack([m[0] for m in mask_data]))
    exc_off = np.zeros(n_masks + 1, dtype=np.int32)
    corr_off = np.zeros(n_masks + 1, dtype=np.int32)
    for j, m in enumerate(mask_data):
        exc_off[j + 1] = exc_off[j] + len(m[1][0])
        corr_off[j + 1] = corr_off[j] + len(m[2])
    exc_node = np.concatenate([m[1][0] for m in mask_data]).astype(np.int32)
    exc_im = np.concatenate([m[1][1] for m in mask_data]).astype(np.float64)
    exc_pot = np.concatenate([m[1][2] for m in mask_data]).astype(np.float64)
    corr_a = np.array([a for m in mask_data for a, _, _ in m[2]], dtype=np.int32)
    corr_b = np.array([b for m in mask_data for _, b, _ in m[2]], dtype=np.int32)
    corr_s = np.array([s for m in mask_data for _, _, s in m[2]], dtype=np.float64)

    nt = int(steps) + 1
    want_series = series
    series = np.empty((n, nt), dtype=np.float64) if want_series else None
    spec_out = np.empty((n, nt), dtype=np.float64) if spectrum else None
    arg = np.empty(n, dtype=np.int64) if spectrum else None
    feat = np.empty((n, 18), dtype=np.float64) if features else None
    ms = ctypes.c_double(0.0)
    p = _lib._p
    d = _lib.EnsembleDesc(S, L, n_masks, p(bits), p(exc_off), p(exc_node), p(exc_im), p(exc_pot),
                          p(corr_off), p(corr_a), p(corr_b), p(corr_s), n, p(sim_mask), p(sim_kick),
                          None, p(coeff), p(damping), int(steps))  # fmt: skip
    _lib.check(_lib.lib().tk_ensemble_run_ex(device, ctypes.byref(d), p(series), p(spec_out), p(arg), p(feat),
                                             ctypes.byref(ms)))  # fmt: skip
    if spectrum:
        freq = np.fft.fftfreq(nt)[arg] / eff  # fftfreq(nt, d=dt)[k] == fftfreq(nt)[k] / dt
    else:
        freq = None
    updates = int(np.array([m[4] for m in mask_data], dtype=np.int64)[sim_mask].sum()) * int(steps)
    res = EnsembleResult(series, spec_out, arg, freq, eff, clamp, maxdeg, kick_nodes, ms.value, updates, feat)
    res.nt = nt
    return res


def run_ensemble_sharded(specs, steps: int, c=1.0, dt=0.2, damping=0.0, kicks=None, *, dist=None,
                         device: int | None = None, gather: bool = True):
    """Ensemble spread over the ranks of a ``torch.distributed`` group (one process per GPU): the
    simulations are independent, so rank r simply runs the contiguous block ``[lo_r, hi_r)`` on its
    own device -- no data-path communication.  With ``gather=True`` every rank
    returns the complete :class:`EnsembleResult` (small per-simulation outputs, one all_gather_object);
    otherwise only its own block plus ``(lo, hi)``."""
    specs = list(specs)
    n = len(specs)
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)

    def part(a):
        a = np.asarray(a, dtype=np.float64)
        return a if a.ndim == 0 else a[lo:hi]

    dev = rank if device is None else device
    mine = run_ensemble(specs[lo:hi], steps, c=part(c), dt=part(dt), damping=part(damping),
                        kicks=None if kicks is None else kicks[lo:hi], device=dev) if hi > lo else None
    if not gather or world == 1:
        return mine, (lo, hi)
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    parts = [p for p in parts if p is not None]
    cat = lambda f: np.concatenate([getattr(p, f) for p in parts])  # noqa: E731
    return EnsembleResult(cat("series"), cat("spectrum"), cat("dominant_bin"), cat("dominant_freq"),
                          cat("effective_dt"), cat("stability_adjusted"), cat("max_degree"),
                          [k for p in parts for k in p.kick_nodes], max(p.elapsed_ms for p in parts),
                          sum(p.node_updates for p in parts)), (lo, hi)


def sweep_grid(size: int, layers: int, radii, dampings, dts, *, dim: int = 3, defect="blackhole",
               **defect_kw):
    """Specs + per-simulation (damping, dt) arrays for a radius x damping x dt grid (BASELINE config 5).
    Ordering: radius-major, then damping, then dt."""
    radii, dampings, dts = map(lambda a: np.asarray(a, dtype=np.float64), (radii, dampings, dts))
    base = [LatticeSpec(size=size, dim=dim, layers=layers, defect=defect, radius=float(r), **defect_kw)
            for r in radii]  # fmt: skip
    specs = [base[i] for i in range(len(radii)) for _ in range(len(dampings) * len(dts))]
    damp = np.tile(np.repeat(dampings, len(dts)), len(radii))
    dt = np.tile(dts, len(radii) * len(dampings))
    return specs, damp, dt
