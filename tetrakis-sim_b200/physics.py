"""Drop-in ``run_wave_sim`` / ``run_fft`` / ``SimulationHistory`` on the B200.

Mirrors ``tetrakis_sim/physics.py`` of the reference: same names, argument meaning, defaults,
metadata keys, warning category and text, exceptions.  The stepping itself
(``physics.py:104-129``) and the spectrum (``physics.py:162-163``) run on the device through the
C ABI in ``include/tetrakis_b200.h``; there is no CPU fallback (missing library or device =>
exception).

Extra keyword-only arguments (all optional; the positional signature is the reference's):
  device   CUDA device ordinal (default 0)
  path     "auto" | "graph" | "structured".  "graph" is the sliced-ELL kernel that reproduces the
           reference bit for bit on any networkx graph; "structured" the clique-sum kernel
           (always used for a LatticeSpec).  "auto" = graph for networkx input, unless
           TETRAKIS_B200_AUTO_STRUCTURED=1 / install(structured=True) opts tetrakis-shaped graphs into the
           structured kernel (compiler.lattice_tables_from_graph; <= 1e-12 instead of bit-identical).
  reorder  graph path only: device row order, "auto" (default) | "none" | "degree" | "rcm"; never
           changes results (the neighbour order inside a row is kept)
  record   "all" (default for graphs: every state, as the reference returns) or "probes"
           (default for a LatticeSpec: only the probe nodes' time series and the final state).
  probes   nodes to record when record="probes" (default: the kicked nodes)
  dtype    "float64" (default; the parity mode) or "float32" for a LatticeSpec: fp32 state, fp64
           row/column sums and Laplacian; relative error vs fp64 about 1e-6 after 100 steps and
           growing ~linearly (tests/test_gpu_fp32.py); not for parity work
  snapshot_every  (LatticeSpec, record="all") keep only states 0, k, 2k, ... in ``history.array`` --
           whole-lattice snapshots for runs whose full history would not fit (the probe series is
           still recorded at every step)
  final_state  (LatticeSpec, record="probes") False: do not copy the whole final state to the host
           (``history.final_state`` stays None) -- 2.1 GB and ~40 ms over PCIe at BASELINE config 2, when all the caller
           wants is the probes' time series
  energy   also return ``history.energy``: the staggered discrete energy between consecutive states,
           ``steps`` values, conserved to rounding when damping == 0 (a new diagnostic; the reference
           computes no energy -- SURVEY.md section 2)
"""

from __future__ import annotations

import warnings
from collections.abc import Hashable, Iterable, Mapping

import numpy as np

from . import _lib
from .compiler import CompiledGraph, compile_graph, compile_lattice, structured_from_graph
from .lattice import LatticeSpec

# (steps+1)*N entries above this -> read-only row views over the history array instead of eager dicts (building
# 51 dicts of 3 064 entries costs 20 ms -- forty times the simulation; every consumer in the reference only reads:
# state.get(node), iteration, .values(): physics.py:162, plot.py:153,274, tests/test_invariants.py:88-89).
# TETRAKIS_B200_EAGER_STATES=1 forces plain dicts.
_EAGER_DICT_LIMIT = 60_000


class SimulationHistory(list):
    """List-like container returned by :func:`run_wave_sim` (reference physics.py:13-27)."""

    def __init__(
        self,
        states: Iterable[Mapping[Hashable, float]],
        *,
        dt: float,
        wave_speed: float,
        metadata: dict | None = None,
    ) -> None:
        super().__init__(states)
        self.dt = dt
        self.wave_speed = wave_speed
        self.metadata = metadata or {}
        # device-run extras (absent on histories built by hand)
        self.array: np.ndarray | None = None  # (steps+1, N) in node order, when record="all"
        self.nodes: list | None = None
        self.probe_nodes: list | None = None
        self.probe_series: np.ndarray | None = None  # (steps+1, n_probes)
        self.final_state: np.ndarray | None = None
        self.elapsed_ms: float | None = None
        self.energy: np.ndarray | None = None  # (steps,) staggered energy, when energy=True
        self._index: dict | None = None  # node -> column of ``array``


class _RowView(Mapping):
    """One recorded state as a read-only mapping node -> value over a row of the history array."""

    __slots__ = ("_row", "_index", "_nodes", "_extra")

    def __init__(self, row, index, nodes, extra=None):
        self._row, self._index, self._nodes, self._extra = row, index, nodes, extra

    def __getitem__(self, node):
        i = self._index.get(node)
        if i is None:
            if self._extra and node in self._extra:
                return self._extra[node]
            raise KeyError(node)
        return float(self._row[i])

    def __iter__(self):
        yield from self._nodes
        if self._extra:
            yield from self._extra

    def __len__(self):
        return len(self._nodes) + (len(self._extra) if self._extra else 0)

    def values(self):
        vals = self._row.tolist()
        if self._extra:
            vals += list(self._extra.values())
        return vals

    def copy(self):
        return dict(self.items())


def _cfl(max_degree: int, c: float, dt: float, steps, damping: float):
    """reference physics.py:64-91: metadata + CFL clamp.  Returns (effective_dt, metadata).
    ``steps`` is recorded as the caller passed it (the reference stores the requested value, negative or not)."""
    effective_dt = float(dt)
    metadata: dict = {
        "steps": steps,
        "requested_dt": dt,
        "wave_speed": c,
        "max_degree": max_degree,
        "damping": damping,
    }
    if max_degree > 0 and c > 0.0:
        cfl_limit = float(1.0 / np.sqrt(max_degree))
        metadata["stability_limit"] = cfl_limit
        if c * effective_dt > cfl_limit:
            effective_dt = cfl_limit / c
            metadata["stability_adjusted"] = True
            metadata["effective_dt"] = float(effective_dt)
            warnings.warn(
                (
                    "Requested time-step is unstable for the current lattice "
                    f"(c*dt={c * dt:.3f} > {cfl_limit:.3f}). Clamping dt to "
                    f"{effective_dt:.3f}"
                ),
                RuntimeWarning,
                stacklevel=4,
            )
    metadata.setdefault("stability_adjusted", False)
    metadata.setdefault("effective_dt", float(effective_dt))
    return effective_dt, metadata


_AUTO_STRUCTURED = None  # None: follow the environment; True / False: set by install(structured=...)


def auto_structured_enabled() -> bool:
    """Does ``path="auto"`` route tetrakis-shaped ``networkx`` graphs to the structured kernel?  Off by default
    (the generic kernel is bit-identical to the reference); on with ``TETRAKIS_B200_AUTO_STRUCTURED=1`` or
    ``install(structured=True)``."""
    import os

    if _AUTO_STRUCTURED is not None:
        return bool(_AUTO_STRUCTURED)
    return os.environ.get("TETRAKIS_B200_AUTO_STRUCTURED", "0") == "1"


def set_auto_structured(flag) -> None:
    global _AUTO_STRUCTURED
    _AUTO_STRUCTURED = flag


def _run_graph(G, steps, initial_data, c, dt, damping, device, path, record, probes, energy, reorder="auto",
               steps_requested=None):
    comp: CompiledGraph | None = None
    if path == "structured":
        comp = structured_from_graph(G, device=device)
    elif path == "auto" and auto_structured_enabled():
        # opt-in (TETRAKIS_B200_AUTO_STRUCTURED=1 / install(structured=True)): a graph that IS a tetrakis lattice, up to
        # a sparse set of exceptions, takes the structured kernel (24 B per update instead of 44 + 4 * degree; results
        # within 1e-12 of the reference instead of bit-identical); anything else stays on the generic kernel
        try:
            comp = structured_from_graph(G, device=device, max_corrections=max(64, G.number_of_edges() // 50))
        except ValueError:
            comp = None
    if comp is None:
        comp = compile_graph(G, device=device, reorder=reorder)
    try:
        nodes, index = comp.nodes, comp.index
        effective_dt, metadata = _cfl(comp.max_degree, c, dt, steps if steps_requested is None else steps_requested,
                                      damping)
        # physics.py:30-38: zero state, then the user's kicks, else a unit kick at nodes[len//2]
        extra0 = None
        if initial_data:
            kicks = {n: float(v) for n, v in initial_data.items() if n in index}
            extra0 = {n: v for n, v in initial_data.items() if n not in index} or None
        elif nodes:
            kicks = {nodes[len(nodes) // 2]: 1.0}
        else:
            kicks = {}
        to_state = (lambda i: i) if comp.state_index is None else (lambda i: int(comp.state_index[i]))
        comp.plan.set_state_sparse([to_state(index[n]) for n in kicks], list(kicks.values()))
        coeff = (c * effective_dt) ** 2  # physics.py:105
        n = len(nodes)
        if record == "all":
            out = comp.plan.run(steps, coeff, damping, history=True, timed=True, energy=energy)
            arr = out["history"]
            if comp.state_index is not None:
                arr = np.ascontiguousarray(arr[:, comp.state_index])
            import os

            if (steps + 1) * n <= _EAGER_DICT_LIMIT or os.environ.get("TETRAKIS_B200_EAGER_STATES") == "1":
                states = [dict(zip(nodes, row)) for row in arr.tolist()]
                if extra0:
                    states[0].update(extra0)
            else:
                states = [_RowView(arr[t], index, nodes, extra0 if t == 0 else None)
                          for t in range(steps + 1)]  # fmt: skip
            hist = SimulationHistory(states, dt=effective_dt, wave_speed=c, metadata=metadata)
            hist.array, hist.nodes, hist._index = arr, nodes, index
            hist.final_state = arr[-1]
        else:
            pnodes = list(probes) if probes is not None else list(kicks)
            pidx = [to_state(index[p]) for p in pnodes]
            out = comp.plan.run(steps, coeff, damping, probes=pidx, timed=True, energy=energy)
            series = out["probes"] if pidx else np.empty((steps + 1, 0))
            states = [dict(zip(pnodes, row)) for row in series.tolist()]
            hist = SimulationHistory(states, dt=effective_dt, wave_speed=c, metadata=metadata)
            hist.probe_nodes, hist.probe_series, hist.nodes = pnodes, series, nodes
            fin = comp.plan.get_state()
            hist.final_state = fin if comp.state_index is None else fin[comp.state_index]
        hist.elapsed_ms = out["elapsed_ms"]
        hist.energy = None if out["energy"] is None else out["energy"] * (c * c)
        return hist
    finally:
        comp.plan.close()


def _run_lattice(spec: LatticeSpec, steps, initial_data, c, dt, damping, device, record, probes, energy,
                 dtype="float64", snapshot_every=1, steps_requested=None, final_state=True):
    plan = compile_lattice(spec, device=device, dtype=dtype)
    try:
        effective_dt, metadata = _cfl(plan.max_degree, c, dt, steps if steps_requested is None else steps_requested,
                                      damping)
        if initial_data:
            kicks = {n: float(v) for n, v in initial_data.items() if spec.contains(n)}
        else:
            kicks = {spec.node(spec.default_node_index()): 1.0}
        plan.set_state_sparse([spec.index(n) for n in kicks], list(kicks.values()))
        coeff = (c * effective_dt) ** 2
        pnodes = list(probes) if probes is not None else list(kicks)
        pidx = [spec.index(p) for p in pnodes]
        if record == "all":
            out = plan.run(steps, coeff, damping, probes=pidx, history=True, timed=True, energy=energy,
                           history_stride=snapshot_every)
        else:
            out = plan.run(steps, coeff, damping, probes=pidx, timed=True, energy=energy)
        series = out["probes"] if pidx else np.empty((steps + 1, 0))
        states = [dict(zip(pnodes, row)) for row in series.tolist()]
        hist = SimulationHistory(states, dt=effective_dt, wave_speed=c, metadata=metadata)
        hist.probe_nodes, hist.probe_series = pnodes, series
        hist.array = out["history"]  # (steps // snapshot_every + 1, num_slots) in device layout when record="all"
        hist.snapshot_every = snapshot_every
        if hist.array is not None and steps % max(snapshot_every, 1) == 0:
            hist.final_state = hist.array[-1]
        elif final_state:
            hist.final_state = plan.get_state()
        hist.elapsed_ms = out["elapsed_ms"]
        hist.energy = None if out["energy"] is None else out["energy"] * (c * c)
        hist.launches = plan.launches  # an attribute, not a metadata key: metadata holds the reference's keys only
        hist.fused = plan.fused  # K steps per pass where the lattice allows (tk_plan_fused_info)
        hist.cross = plan.cross  # lattices with defects / 3-D: the cross scheme (tk_plan_cross_info)
        return hist
    finally:
        plan.close()


def run_wave_sim(
    G,
    steps: int = 100,
    initial_data: dict | None = None,
    c: float = 1.0,
    dt: float = 0.2,
    damping: float = 0.0,
    *,
    device: int = 0,
    path: str = "auto",
    record: str | None = None,
    probes: list | None = None,
    energy: bool = False,
    dtype: str = "float64",
    snapshot_every: int = 1,
    reorder: str = "auto",
    final_state: bool = True,
) -> SimulationHistory:
    """Simulate the discrete wave equation on ``G`` (reference physics.py:41-131).

    ``G`` is a ``networkx`` graph (not mutated) or a :class:`LatticeSpec`.  The CFL check clamps an
    unstable ``dt`` and emits a :class:`RuntimeWarning`, exactly as the reference does.
    """
    steps_requested = steps  # metadata["steps"] is the requested value (physics.py:67)
    steps = max(int(steps), 0)  # range(steps) in the reference simply does not iterate
    if path not in ("auto", "graph", "structured"):
        raise ValueError("path must be 'auto', 'graph' or 'structured'")
    if isinstance(G, LatticeSpec):
        return _run_lattice(G, steps, initial_data, c, dt, damping, device, record or "probes", probes, energy,
                            dtype, max(int(snapshot_every), 1), steps_requested, final_state)
    if dtype != "float64":
        raise ValueError("dtype='float32' is available for LatticeSpec input only")
    return _run_graph(G, steps, initial_data, c, dt, damping, device, path, record or "all", probes, energy,
                      reorder, steps_requested)


def run_fft(
    history: Iterable[Mapping[Hashable, float]],
    node: Hashable | None = None,
    *,
    dt: float | None = None,
    device: int = 0,
) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Spectrum of one node across the history (reference physics.py:140-165).

    Returns ``(freq, |X|, values)``; ``|X|`` is computed by the device DFT kernel, ``freq`` is
    ``np.fft.fftfreq(len(values), d=dt)`` (pure index arithmetic)."""
    history_list = list(history)
    if not history_list:
        raise ValueError("Simulation history is empty")
    sample_state = history_list[0]
    if node is None:
        node = next(iter(sample_state.keys()))
    inferred_dt = dt
    if inferred_dt is None and hasattr(history, "dt"):
        inferred_dt = history.dt
    if inferred_dt is None:
        inferred_dt = 1.0
    values = None
    arr, index = getattr(history, "array", None), getattr(history, "_index", None)
    if arr is not None and index is not None and len(arr) == len(history_list):
        col = index.get(node)  # a column of the recorded array instead of steps+1 dict lookups
        if col is not None:
            values = np.array(arr[:, col], dtype=np.float64)
    if values is None:
        values = np.array([state.get(node, 0.0) for state in history_list], dtype=np.float64)
    spectrum, _ = _lib.dft_abs(values, device=device)
    freq = np.fft.fftfreq(len(values), d=inferred_dt)
    return freq, spectrum, values


def run_fft_many(history, nodes, *, dt: float | None = None, device: int = 0):
    """Spectra of MANY nodes of one history in a single batched device call (the notebooks of the
    reference call ``run_fft`` once per node over the full history to build per-layer spectral maps,
    notebooks/tetrakis_blackhole_waves.ipynb cells 12-13).

    Returns ``(freq, spectra[len(nodes), T], values[len(nodes), T], dominant_bin[len(nodes)])`` with the
    per-node semantics of :func:`run_fft` (missing nodes read as 0.0, physics.py:162)."""
    history_list = list(history)
    if not history_list:
        raise ValueError("Simulation history is empty")
    nodes = list(nodes)
    inferred_dt = dt if dt is not None else getattr(history, "dt", None)
    if inferred_dt is None:
        inferred_dt = 1.0
    arr, index = getattr(history, "array", None), getattr(history, "_index", None)
    T = len(history_list)
    values = np.zeros((len(nodes), T), dtype=np.float64)
    if arr is not None and index is not None and len(arr) == T:
        cols = [index.get(n) for n in nodes]
        have = [i for i, c in enumerate(cols) if c is not None]
        if have:
            values[have] = np.asarray(arr[:, [cols[i] for i in have]], dtype=np.float64).T
        for i, c in enumerate(cols):
            if c is None:
                values[i] = [state.get(nodes[i], 0.0) for state in history_list]
    else:
        for i, n in enumerate(nodes):
            values[i] = [state.get(n, 0.0) for state in history_list]
    if len(nodes) == 0:
        return np.fft.fftfreq(T, d=inferred_dt), values.copy(), values, np.zeros(0, dtype=np.int64)
    spectra = np.empty_like(values)
    bins = np.empty(len(nodes), dtype=np.int64)
    for lo in range(0, len(nodes), 32768):  # the DFT grid's batch dimension is bounded
        sp, ab = _lib.dft_abs(values[lo : lo + 32768], device=device)
        spectra[lo : lo + 32768], bins[lo : lo + 32768] = sp, ab
    return np.fft.fftfreq(T, d=inferred_dt), spectra, values, bins


def dominant_frequency(freq, spectrum) -> float:
    """``float(freq[np.argmax(spectrum)])`` (reference batch.py:205) with the k <-> N-k tie of a real signal's
    two-sided spectrum (equal up to FFT rounding) broken towards the lower bin, so that the
    sign does not depend on the FFT implementation."""
    spectrum = np.asarray(spectrum)
    k = int(np.argmax(spectrum))
    mirror = (len(spectrum) - k) % len(spectrum)
    if mirror < k and abs(spectrum[mirror] - spectrum[k]) <= 1e-9 * abs(spectrum[k]):
        k = mirror
    return float(np.asarray(freq)[k])
