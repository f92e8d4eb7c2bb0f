"""Oracle restatement of ``run_wave_sim`` / ``run_fft`` -- TEST INFRASTRUCTURE ONLY.

Restates ``/root/reference/tetrakis_sim/physics.py``:
  * ``:60-91``   CFL clamp + metadata            -> :func:`cfl_clamp`
  * ``:30-38``   default kick                    -> :func:`initial_state`
  * ``:104-129`` leapfrog loop                   -> :func:`run_graph_python` (pure Python, the
                 reference's literal operation order), :func:`run_graph` (C, bit-identical),
                 :func:`run_structured_numpy` / :func:`run_structured` (clique-sum form)
  * ``:140-165`` probe FFT                       -> :func:`run_fft`

Pinned against outputs of the unmodified reference (``tests/golden``); see ``oracle/__init__.py``.
"""

from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libwave_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile ``wave_oracle.c`` (gcc, no FMA contraction).  Called by ``__graft_entry__.build()``."""
    src = os.path.join(_HERE, "wave_oracle.c")
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    ):
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", _LIB_PATH]
        )
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        P = ctypes.c_void_p
        lib.tko_graph_run.restype = ctypes.c_int
        lib.tko_graph_run.argtypes = [
            ctypes.c_int64, P, P, P, P, P, P, P, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
            P, P, ctypes.c_int64, P,
        ]  # fmt: skip
        lib.tko_structured_run.restype = ctypes.c_int
        lib.tko_structured_run.argtypes = [
            ctypes.c_int32, ctypes.c_int32, P, P, P, ctypes.c_int64, P, P, P, P, P,
            ctypes.c_int64, ctypes.c_double, ctypes.c_double, P, ctypes.c_int64, P,
        ]  # fmt: skip
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------
# physics.py:60-91
# --------------------------------------------------------------------------------------
def cfl_clamp(max_degree: int, c: float, dt: float, steps: int, damping: float):
    """Return (effective_dt, metadata, warning_text_or_None) exactly as physics.py:64-91 does."""
    effective_dt = float(dt)
    metadata = {
        "steps": steps,
        "requested_dt": dt,
        "wave_speed": c,
        "max_degree": max_degree,
        "damping": damping,
    }
    warning = None
    if max_degree > 0 and c > 0.0:
        cfl_limit = float(1.0 / np.sqrt(max_degree))
        metadata["stability_limit"] = cfl_limit
        if c * effective_dt > cfl_limit:
            effective_dt = cfl_limit / c
            metadata["stability_adjusted"] = True
            metadata["effective_dt"] = float(effective_dt)
            warning = (
                "Requested time-step is unstable for the current lattice "
                f"(c*dt={c * dt:.3f} > {cfl_limit:.3f}). Clamping dt to "
                f"{effective_dt:.3f}"
            )
    metadata.setdefault("stability_adjusted", False)
    metadata.setdefault("effective_dt", float(effective_dt))
    return effective_dt, metadata, warning


def initial_state(n: int, kicks: dict[int, float] | None) -> np.ndarray:
    """physics.py:30-38 on node indices: falsy ``kicks`` => 1.0 at index n//2."""
    u = np.zeros(n, dtype=np.float64)
    if kicks:
        for i, v in kicks.items():
            u[i] = v
    elif n:
        u[n // 2] = 1.0
    return u


# --------------------------------------------------------------------------------------
# physics.py:104-129, generic graph
# --------------------------------------------------------------------------------------
def run_graph_python(rowptr, colidx, deg, inv_mass, potential, u0, steps, coeff, damping):
    """Literal pure-Python restatement (Python floats), for small cases.  Returns (steps+1, N) array."""
    n = len(u0)
    nbrs = [[int(j) for j in colidx[rowptr[i] : rowptr[i + 1]]] for i in range(n)]
    degs = [int(d) for d in deg]
    im = [float(x) for x in inv_mass]
    pot = [float(x) for x in potential]
    u = [float(x) for x in u0]
    up = [0.0] * n
    hist = [list(u)]
    for _ in range(steps):
        unew = [0.0] * n
        for i in range(n):
            u_n = u[i]
            up_n = up[i]
            neighbor_sum = 0.0
            for j in nbrs[i]:
                neighbor_sum += u[j]
            lap = neighbor_sum - degs[i] * u_n
            unew[i] = (
                2 * u_n
                - up_n
                + (coeff * im[i]) * lap
                - (coeff * pot[i]) * u_n
                - damping * (u_n - up_n)
            )
        up, u = u, unew
        hist.append(list(u))
    return np.asarray(hist, dtype=np.float64)


def run_graph(
    rowptr, colidx, deg, inv_mass, potential, u0, steps, coeff, damping, *, probes=None,
    keep_history=True,
):  # fmt: skip
    """C restatement, bit-identical to :func:`run_graph_python`.

    Returns dict(history|None, probes|None, u, uprev).
    """
    lib = _load()
    n = len(u0)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    deg = np.ascontiguousarray(deg, dtype=np.float64)
    inv_mass = np.ascontiguousarray(inv_mass, dtype=np.float64)
    potential = np.ascontiguousarray(potential, dtype=np.float64)
    u = np.array(u0, dtype=np.float64)
    up = np.zeros(n, dtype=np.float64)
    hist = np.empty((steps + 1, n), dtype=np.float64) if keep_history else None
    pidx = None if probes is None else np.ascontiguousarray(probes, dtype=np.int64)
    pout = None if probes is None else np.empty((steps + 1, len(pidx)), dtype=np.float64)
    rc = lib.tko_graph_run(
        n, _ptr(rowptr), _ptr(colidx), _ptr(deg), _ptr(inv_mass), _ptr(potential), _ptr(u),
        _ptr(up), steps, float(coeff), float(damping), _ptr(hist), _ptr(pidx),
        0 if pidx is None else len(pidx), _ptr(pout),
    )  # fmt: skip
    if rc != 0:
        raise MemoryError("tko_graph_run failed")
    return {"history": hist, "probes": pout, "u": u, "uprev": up}


# --------------------------------------------------------------------------------------
# clique-sum restatement on the [z][r][c][q] layout
# --------------------------------------------------------------------------------------
def lattice_index(size: int, layers: int, node) -> int:
    """Reference node tuple -> linear index in the device layout [z][r][c][q]."""
    if len(node) == 3:
        r, c, q = node
        z = 0
    else:
        r, c, z, q = node
    return ((z * size + r) * size + c) * 4 + "ABCD".index(q)


def structured_inputs_from_graph(G, size: int, layers: int):
    """Derive (flags, inv_mass, potential, corrections) of the clique-sum model from a networkx
    graph built by lattice.py + defects.py.  flags: bit0 alive, bit1 part (alive and degree > 0 or
    never pruned).  Corrections are the exact edge difference between the model and G."""
    n = layers * size * size * 4
    flags = np.zeros(n, dtype=np.uint8)
    inv_mass = np.ones(n, dtype=np.float64)
    potential = np.zeros(n, dtype=np.float64)
    idx = {}
    for node in G.nodes:
        i = lattice_index(size, layers, node)
        idx[node] = i
        flags[i] = 1 | (2 if G.degree[node] > 0 else 0)
        m = float(G.nodes[node].get("mass", 1.0))
        if m <= 0.0:
            m = 1.0
        inv_mass[i] = 1.0 / m
        potential[i] = float(G.nodes[node].get("potential", 0.0))
    # a 1x1x1 lattice node with no neighbours is "part" with degree 0 in the model too; harmless.
    part = ((flags >> 1) & 1).reshape(layers, size, size, 4).astype(bool)
    model_edges = set()
    P = part
    for z in range(layers):
        for r in range(size):
            for c in range(size):
                for q in range(4):
                    if not P[z, r, c, q]:
                        continue
                    a = ((z * size + r) * size + c) * 4 + q
                    for q2 in range(q + 1, 4):
                        if P[z, r, c, q2]:
                            model_edges.add((a, a - q + q2))
                    for c2 in range(c + 1, size):
                        if P[z, r, c2, q]:
                            model_edges.add((a, ((z * size + r) * size + c2) * 4 + q))
                    for r2 in range(r + 1, size):
                        if P[z, r2, c, q]:
                            model_edges.add((a, ((z * size + r2) * size + c) * 4 + q))
                    if z + 1 < layers and P[z + 1, r, c, q]:
                        model_edges.add((a, a + size * size * 4))
    graph_edges = set()
    for a, b in G.edges:
        ia, ib = idx[a], idx[b]
        graph_edges.add((min(ia, ib), max(ia, ib)))
    corr = []
    for a, b in sorted(model_edges - graph_edges):
        corr += [(a, b, -1.0), (b, a, -1.0)]
    for a, b in sorted(graph_edges - model_edges):
        corr += [(a, b, 1.0), (b, a, 1.0)]
    return flags, inv_mass, potential, corr


def run_structured_numpy(
    size, layers, flags, inv_mass, potential, corr, u0, steps, coeff, damping, *, dtype=np.float64
):
    """NumPy clique-sum restatement; ``dtype=np.longdouble`` gives the noise-floor reference.
    Returns the full history (steps+1, N)."""
    shp = (layers, size, size, 4)
    n = int(np.prod(shp))
    A = np.ones(shp, dtype) if flags is None else (flags & 1).reshape(shp).astype(dtype)
    P = np.ones(shp, dtype) if flags is None else ((flags >> 1) & 1).reshape(shp).astype(dtype)
    IM = np.ones(shp, dtype) if inv_mass is None else np.asarray(inv_mass, dtype).reshape(shp)
    V = np.zeros(shp, dtype) if potential is None else np.asarray(potential, dtype).reshape(shp)
    coeff = dtype(coeff)
    damping = dtype(damping)
    Pz = np.zeros((layers + 2,) + shp[1:], dtype)
    Pz[1:-1] = P
    deg = P * (
        (P.sum(3, keepdims=True) - 1) + (P.sum(2, keepdims=True) - 1) + (P.sum(1, keepdims=True) - 1)
        + Pz[:-2] + Pz[2:]
    )  # fmt: skip
    ca = np.array([c[0] for c in corr], dtype=np.int64)
    cb = np.array([c[1] for c in corr], dtype=np.int64)
    cs = np.array([c[2] for c in corr], dtype=dtype)
    U = np.asarray(u0, dtype).reshape(shp).copy()
    Up = np.zeros(shp, dtype)
    hist = np.empty((steps + 1, n), dtype)
    hist[0] = U.ravel()
    for s in range(steps):
        pu = P * U
        pz = np.zeros((layers + 2,) + shp[1:], dtype)
        pz[1:-1] = pu
        nsum = P * (
            (pu.sum(3, keepdims=True) - pu) + (pu.sum(2, keepdims=True) - pu)
            + (pu.sum(1, keepdims=True) - pu) + pz[:-2] + pz[2:]
        )  # fmt: skip
        lap = nsum - deg * U
        if len(ca):
            lf = lap.ravel()
            uf = U.ravel()
            np.add.at(lf, ca, cs * (uf[cb] - uf[ca]))
        Un = 2 * U - Up + (coeff * IM) * lap - (coeff * V) * U - damping * (U - Up)
        Un *= A
        Up, U = U, Un
        hist[s + 1] = U.ravel()
    return hist


def run_structured(
    size, layers, flags, inv_mass, potential, corr, u0, steps, coeff, damping, *, probes=None,
    uprev0=None,
):  # fmt: skip
    """C/OpenMP clique-sum restatement (the CPU baseline at sizes networkx cannot hold).
    Returns dict(u, uprev, probes)."""
    lib = _load()
    n = layers * size * size * 4
    u = np.array(u0, dtype=np.float64).reshape(-1)
    assert u.size == n
    up = np.zeros(n, dtype=np.float64) if uprev0 is None else np.array(uprev0, dtype=np.float64)
    ca = np.ascontiguousarray([c[0] for c in corr], dtype=np.int64)
    cb = np.ascontiguousarray([c[1] for c in corr], dtype=np.int64)
    cs = np.ascontiguousarray([c[2] for c in corr], dtype=np.float64)
    fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
    im = None if inv_mass is None else np.ascontiguousarray(inv_mass, dtype=np.float64)
    pv = None if potential is None else np.ascontiguousarray(potential, dtype=np.float64)
    pidx = None if probes is None else np.ascontiguousarray(probes, dtype=np.int64)
    pout = None if probes is None else np.empty((steps + 1, len(pidx)), dtype=np.float64)
    rc = lib.tko_structured_run(
        layers, size, _ptr(fl), _ptr(im), _ptr(pv), len(ca), _ptr(ca), _ptr(cb), _ptr(cs),
        _ptr(u), _ptr(up), steps, float(coeff), float(damping), _ptr(pidx),
        0 if pidx is None else len(pidx), _ptr(pout),
    )  # fmt: skip
    if rc < 0:
        raise MemoryError("tko_structured_run failed")
    if rc == 1:
        u, up = up, u
    return {"u": u, "uprev": up, "probes": pout}


def staggered_energy(history, rowptr, colidx, deg, inv_mass, potential, coeff):
    """Discrete energy between consecutive states of a (steps+1, N) history (NOT in the reference --
    a new diagnostic): entry s =
        1/2 sum_i [ m_i (u_i^{s+1} - u_i^s)^2 / coeff + u_i^{s+1} (V_i u_i^s - lap_i(u^s)) ],
    lap = A u - deg * u as in physics.py:115-119.  Exactly conserved by the undamped leapfrog."""
    import scipy.sparse as sp

    h = np.asarray(history, dtype=np.longdouble)
    n = h.shape[1]
    A = sp.csr_matrix((np.ones(len(colidx)), np.asarray(colidx), np.asarray(rowptr)), shape=(n, n))
    out = np.empty(h.shape[0] - 1, dtype=np.float64)
    m = 1.0 / np.asarray(inv_mass, dtype=np.longdouble)
    for s in range(h.shape[0] - 1):
        u, un = h[s], h[s + 1]
        lap = (A @ np.asarray(u, dtype=np.float64)).astype(np.longdouble) - np.asarray(deg) * u
        out[s] = float(0.5 * np.sum(m * (un - u) ** 2 / coeff + un * (np.asarray(potential) * u - lap)))
    return out


# --------------------------------------------------------------------------------------
# physics.py:140-165
# --------------------------------------------------------------------------------------
def run_fft(values, dt: float):
    """physics.py:162-165: (fftfreq, |fft|, values) with numpy's own FFT, as the reference calls it."""
    values = np.asarray(values, dtype=np.float64)
    return np.fft.fftfreq(len(values), d=dt), np.abs(np.fft.fft(values)), values


def dft_abs_direct(values):
    """O(N^2) DFT magnitude in extended precision with exact-integer twiddle reduction: an
    FFT-library-independent check for the device DFT."""
    x = np.asarray(values, dtype=np.longdouble)
    n = len(x)
    k = np.arange(n, dtype=np.int64)
    ph = (np.outer(k, k) % n).astype(np.longdouble) * (2 * np.longdouble(math.pi) / n)
    re = (np.cos(ph) * x).sum(1)
    im = -(np.sin(ph) * x).sum(1)
    return np.sqrt(re * re + im * im).astype(np.float64)


def folded_bin(spectrum) -> int:
    """Dominant bin modulo the k <-> N-k mirror of a real signal."""
    k = int(np.argmax(spectrum))
    return min(k, len(spectrum) - k) if k else 0
