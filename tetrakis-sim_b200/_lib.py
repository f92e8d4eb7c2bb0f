"""ctypes binding of ``include/tetrakis_b200.h`` -- the only way Python reaches the CUDA code.

There is deliberately no fallback: if ``lib/libtetrakis_b200.so`` is missing or fails to load,
:func:`lib` raises :class:`BackendUnavailable`; if no CUDA device is present every compute entry
point returns ``TK_ERR_NO_DEVICE`` and :func:`check` raises :class:`NoDeviceError`.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_void_p

import numpy as np

from ._build import LIB_PATH

TK_OK, TK_ERR_INVALID, TK_ERR_CUDA, TK_ERR_NOMEM, TK_ERR_NO_DEVICE = 0, 1, 2, 3, 4

# every symbol declared in include/tetrakis_b200.h (tests check the .so exports all of them)
ABI_SYMBOLS = (
    "tk_abi_version", "tk_last_error", "tk_device_count", "tk_device_info", "tk_device_rates", "tk_device_trim",
    "tk_graph_plan_create", "tk_lattice_plan_create", "tk_lattice_plan_create_ex",
    "tk_plan_set_stream", "tk_plan_num_nodes", "tk_plan_max_degree",
    "tk_plan_set_state_sparse", "tk_plan_set_state", "tk_plan_get_state",
    "tk_plan_run", "tk_plan_set_probes", "tk_plan_read_probes", "tk_plan_launch_count", "tk_plan_fused_info", "tk_plan_cross_info", "tk_plan_destroy",
    "tk_sheet_band_plan_create", "tk_sheet_exchange_doubles", "tk_sheet_pass_local", "tk_sheet_pass_commit", "tk_sheet_kernel_ms",
    "tk_lattice_halo_ptrs", "tk_lattice_step_begin", "tk_lattice_step_layers",
    "tk_lattice_step_layers_on", "tk_lattice_step_end", "tk_lattice_tiling",
    "tk_lattice_ipc_export", "tk_ipc_open", "tk_ipc_close", "tk_lattice_peer_attach",
    "tk_lattice_set_push_mode", "tk_lattice_push_halos", "tk_lattice_signal", "tk_lattice_wait",
    "tk_lattice_arm_signal", "tk_lattice_signal_pending", "tk_lattice_state_ptrs", "tk_dft_abs", "tk_ensemble_run", "tk_ensemble_run_ex",
)  # fmt: skip


class BackendUnavailable(RuntimeError):
    """The CUDA library is not built / not loadable.  There is no CPU path."""


class NoDeviceError(RuntimeError):
    """No CUDA device.  There is no CPU path."""


class TkError(RuntimeError):
    pass


class LatticeDesc(ctypes.Structure):
    _fields_ = [
        ("size", c_int32), ("layers", c_int32), ("z_begin", c_int32), ("z_end", c_int32),
        ("n_special", c_int64), ("special_index", c_void_p), ("special_flags", c_void_p),
        ("special_inv_mass", c_void_p), ("special_potential", c_void_p),
        ("n_corr", c_int64), ("corr_a", c_void_p), ("corr_b", c_void_p), ("corr_sign", c_void_p),
    ]  # fmt: skip


class RunArgs(ctypes.Structure):
    _fields_ = [
        ("steps", c_int64), ("coeff", c_double), ("damping", c_double),
        ("n_probes", c_int64), ("probe_index", c_void_p), ("probe_out", c_void_p),
        ("history_out", c_void_p), ("energy_out", c_void_p), ("elapsed_ms", c_void_p),
        ("step_kernel_ms", c_void_p), ("history_stride", c_int64), ("flags", ctypes.c_uint64),
    ]  # fmt: skip


class EnsembleDesc(ctypes.Structure):
    _fields_ = [
        ("size", c_int32), ("layers", c_int32), ("n_masks", c_int64), ("cell_bits", c_void_p),
        ("exc_offset", c_void_p), ("exc_node", c_void_p), ("exc_inv_mass", c_void_p),
        ("exc_potential", c_void_p), ("corr_offset", c_void_p), ("corr_a", c_void_p),
        ("corr_b", c_void_p), ("corr_sign", c_void_p), ("n_sims", c_int64), ("sim_mask", c_void_p),
        ("sim_kick", c_void_p), ("sim_kick_value", c_void_p), ("sim_coeff", c_void_p),
        ("sim_damping", c_void_p), ("steps", c_int64),
    ]  # fmt: skip


_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise BackendUnavailable(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  tetrakis-b200 has no CPU fallback."
        )
    try:
        L = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the host
        raise BackendUnavailable(f"cannot load {LIB_PATH}: {exc}") from exc
    P = c_void_p
    L.tk_abi_version.restype = c_int
    L.tk_last_error.restype = c_char_p
    L.tk_device_count.argtypes = [POINTER(c_int)]
    L.tk_device_info.argtypes = [c_int, c_char_p, c_int, POINTER(c_int), POINTER(c_int),
                                 POINTER(c_int), POINTER(c_int64)]  # fmt: skip
    L.tk_device_trim.argtypes = [c_int]
    L.tk_device_rates.argtypes = [c_int, c_void_p, c_void_p]
    L.tk_graph_plan_create.argtypes = [c_int, c_int64, P, P, P, P, P, POINTER(P)]
    L.tk_lattice_plan_create.argtypes = [c_int, POINTER(LatticeDesc), POINTER(P)]
    L.tk_lattice_plan_create_ex.argtypes = [c_int, POINTER(LatticeDesc), ctypes.c_uint32, POINTER(P)]
    L.tk_plan_set_stream.argtypes = [P, P]
    L.tk_plan_num_nodes.argtypes = [P, POINTER(c_int64)]
    L.tk_plan_max_degree.argtypes = [P, POINTER(c_int64)]
    L.tk_plan_set_state_sparse.argtypes = [P, c_int64, P, P]
    L.tk_plan_set_state.argtypes = [P, P, P]
    L.tk_plan_get_state.argtypes = [P, P, P]
    L.tk_plan_run.argtypes = [P, POINTER(RunArgs)]
    L.tk_plan_set_probes.argtypes = [P, c_int64, P, c_int64]
    L.tk_plan_read_probes.argtypes = [P, P, POINTER(c_int64)]
    L.tk_plan_launch_count.argtypes = [P, POINTER(c_int64)]
    L.tk_plan_fused_info.argtypes = [P, POINTER(c_int32), POINTER(c_int64), POINTER(c_int32), POINTER(c_int32)]
    L.tk_plan_cross_info.argtypes = [P, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_double),
                                     POINTER(c_int64)]
    L.tk_sheet_band_plan_create.argtypes = [c_int, c_int32, c_int32, c_int32, POINTER(P)]
    L.tk_sheet_exchange_doubles.argtypes = [P, POINTER(c_int64)]
    L.tk_sheet_pass_local.argtypes = [P, c_int32, c_double, c_double, c_void_p]
    L.tk_sheet_pass_commit.argtypes = [P, c_void_p]
    L.tk_sheet_kernel_ms.argtypes = [P, POINTER(c_double), POINTER(c_int64)]
    L.tk_plan_destroy.argtypes = [P]
    L.tk_lattice_halo_ptrs.argtypes = [P, c_int32, POINTER(P), POINTER(P), POINTER(P), POINTER(P),
                                       POINTER(c_int64)]  # fmt: skip
    L.tk_lattice_step_begin.argtypes = [P, c_double, c_double]
    L.tk_lattice_step_layers.argtypes = [P, c_int32, c_int32]
    L.tk_lattice_step_layers_on.argtypes = [P, c_int32, c_int32, P]
    L.tk_lattice_step_end.argtypes = [P]
    L.tk_lattice_ipc_export.argtypes = [P, P]
    L.tk_ipc_open.argtypes = [c_int, P, POINTER(P)]
    L.tk_ipc_close.argtypes = [c_int, P]
    L.tk_lattice_peer_attach.argtypes = [P, c_int32, P, P, c_int32]
    L.tk_lattice_set_push_mode.argtypes = [P, c_int32]
    L.tk_lattice_push_halos.argtypes = [P, P]
    L.tk_lattice_signal.argtypes = [P, ctypes.c_uint64, P]
    L.tk_lattice_wait.argtypes = [P, ctypes.c_uint64, P]
    L.tk_lattice_arm_signal.argtypes = [P, ctypes.c_uint64]
    L.tk_lattice_signal_pending.argtypes = [P, POINTER(c_int32)]
    L.tk_lattice_state_ptrs.argtypes = [P, POINTER(P), POINTER(P)]
    L.tk_lattice_tiling.argtypes = [P, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32),
                                    POINTER(c_int32)]  # fmt: skip
    L.tk_dft_abs.argtypes = [c_int, P, c_int64, c_int64, P, P]
    L.tk_ensemble_run.argtypes = [c_int, POINTER(EnsembleDesc), P, P, P, POINTER(c_double)]
    L.tk_ensemble_run_ex.argtypes = [c_int, POINTER(EnsembleDesc), P, P, P, P, POINTER(c_double)]
    for name in ABI_SYMBOLS:
        fn = getattr(L, name)
        if name not in ("tk_last_error",):
            fn.restype = c_int
    if L.tk_abi_version() != 1:
        raise BackendUnavailable(f"ABI version mismatch: library reports {L.tk_abi_version()}")
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc == TK_OK:
        return
    msg = lib().tk_last_error().decode("utf-8", "replace")
    if rc == TK_ERR_NO_DEVICE:
        raise NoDeviceError(msg)
    if rc == TK_ERR_INVALID:
        raise ValueError(msg)
    if rc == TK_ERR_NOMEM:
        raise MemoryError(msg)
    raise TkError(msg)


def device_count() -> int:
    n = c_int(0)
    check(lib().tk_device_count(ctypes.byref(n)))
    return n.value


def device_info(device: int = 0) -> dict:
    name = ctypes.create_string_buffer(256)
    sm, maj, mnr, mem = c_int(0), c_int(0), c_int(0), c_int64(0)
    check(lib().tk_device_info(device, name, 256, ctypes.byref(sm), ctypes.byref(maj),
                               ctypes.byref(mnr), ctypes.byref(mem)))  # fmt: skip
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value),
            "total_mem": mem.value}  # fmt: skip


def device_rates(device: int = 0) -> dict:
    """FP64 instructions / s and HBM bytes / s the pass-length heuristics assume for this device (tk_device_rates)."""
    f, b = ctypes.c_double(0.0), ctypes.c_double(0.0)
    check(lib().tk_device_rates(device, ctypes.byref(f), ctypes.byref(b)))
    return {"fp64_per_s": f.value, "hbm_bytes_per_s": b.value}


def device_trim(device: int = 0) -> None:
    """Return the library's cache of freed device blocks to the driver (tk_device_trim)."""
    check(lib().tk_device_trim(device))


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _arr(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class Plan:
    """Owner of one ``tk_plan`` (device buffers freed on :meth:`close` / garbage collection)."""

    def __init__(self, handle, kind: str):
        self._h = handle
        self.kind = kind
        self.f32 = False  # optional fp32 state (lattice plans only)

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def graph(cls, rowptr, colidx, degree, inv_mass, potential, device: int = 0) -> "Plan":
        rowptr = _arr(rowptr, np.int64)
        colidx = _arr(colidx, np.int32)
        degree = _arr(degree, np.float64)
        inv_mass = _arr(inv_mass, np.float64)
        potential = _arr(potential, np.float64)
        n = len(rowptr) - 1
        if not (len(degree) == len(inv_mass) == len(potential) == n):
            raise ValueError("graph arrays disagree on the node count")
        h = c_void_p()
        check(lib().tk_graph_plan_create(device, n, _p(rowptr), _p(colidx), _p(degree),
                                         _p(inv_mass), _p(potential), ctypes.byref(h)))  # fmt: skip
        return cls(h, "graph")

    @classmethod
    def lattice(cls, size, layers, *, z_begin=0, z_end=None, special_index=None,
                special_flags=None, special_inv_mass=None, special_potential=None, corr=None,
                device: int = 0, f32: bool = False) -> "Plan":  # fmt: skip
        si = _arr(special_index if special_index is not None else [], np.int64)
        sf = _arr(special_flags if special_flags is not None else [], np.uint8)
        sm = _arr(special_inv_mass, np.float64)
        sv = _arr(special_potential, np.float64)
        if len(si) != len(sf):
            raise ValueError("special_index / special_flags length mismatch")
        corr = corr or []
        ca = _arr([c[0] for c in corr], np.int64)
        cb = _arr([c[1] for c in corr], np.int64)
        cs = _arr([c[2] for c in corr], np.float64)
        d = LatticeDesc(int(size), int(layers), int(z_begin),
                        int(layers if z_end is None else z_end), len(si), _p(si), _p(sf), _p(sm),
                        _p(sv), len(ca), _p(ca), _p(cb), _p(cs))  # fmt: skip
        h = c_void_p()
        check(lib().tk_lattice_plan_create_ex(device, ctypes.byref(d), 1 if f32 else 0, ctypes.byref(h)))
        plan = cls(h, "lattice")
        plan.f32 = bool(f32)
        return plan

    @classmethod
    def sheet_band(cls, size, row_begin, row_end, device: int = 0) -> "Plan":
        """Rows [row_begin, row_end) of ONE defect-free sheet sharded over GPUs (tk_sheet_band_plan_create)."""
        h = c_void_p()
        check(lib().tk_sheet_band_plan_create(device, int(size), int(row_begin), int(row_end), ctypes.byref(h)))
        return cls(h, "sheet_band")

    @property
    def exchange_doubles(self) -> int:
        n = c_int64(0)
        check(lib().tk_sheet_exchange_doubles(self._h, ctypes.byref(n)))
        return n.value

    def pass_local(self, k: int, coeff: float, damping: float, exchange_ptr: int):
        """k steps on this band (k = 0: only the sums of both time levels); leaves the band's column sums and
        quadrant totals in the device buffer at ``exchange_ptr`` for the caller's all-reduce."""
        check(lib().tk_sheet_pass_local(self._h, int(k), float(coeff), float(damping), c_void_p(exchange_ptr)))

    def pass_commit(self, exchange_ptr: int):
        check(lib().tk_sheet_pass_commit(self._h, c_void_p(exchange_ptr)))

    def sheet_kernel_ms(self):
        """(summed device time of the pass kernels since the last call, number of passes)."""
        ms, n = c_double(0.0), c_int64(0)
        check(lib().tk_sheet_kernel_ms(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if self._h is not None and self._h.value:
            lib().tk_plan_destroy(self._h)
        self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- queries ------------------------------------------------------------------------------
    @property
    def num_nodes(self) -> int:
        n = c_int64(0)
        check(lib().tk_plan_num_nodes(self._h, ctypes.byref(n)))
        return n.value

    @property
    def max_degree(self) -> int:
        n = c_int64(0)
        check(lib().tk_plan_max_degree(self._h, ctypes.byref(n)))
        return n.value

    @property
    def launches(self) -> int:
        n = c_int64(0)
        check(lib().tk_plan_launch_count(self._h, ctypes.byref(n)))
        return n.value

    @property
    def fused(self) -> dict:
        """How the last run() went: steps per fused pass (0 = step by step), steps covered by passes."""
        k, n, r, o = c_int32(0), c_int64(0), c_int32(0), c_int32(0)
        check(lib().tk_plan_fused_info(self._h, ctypes.byref(k), ctypes.byref(n), ctypes.byref(r), ctypes.byref(o)))
        return {"steps_per_pass": k.value, "fused_steps": n.value, "rows_per_cta": r.value, "ctas_per_sm": o.value}

    @property
    def cross(self) -> dict:
        """The cross scheme of the last run() (lattices with defects / 3-D): steps per pass (0 = not used), rows and
        columns of the cross, the fraction of the lattice it covers."""
        k, r, c, f, n = c_int32(0), c_int32(0), c_int32(0), c_double(0.0), c_int64(0)
        check(lib().tk_plan_cross_info(self._h, ctypes.byref(k), ctypes.byref(r), ctypes.byref(c), ctypes.byref(f),
                                       ctypes.byref(n)))
        return {"steps_per_pass": k.value, "cross_rows": r.value, "cross_cols": c.value, "cross_fraction": f.value,
                "passes_since_anchor": n.value}

    @property
    def tiling(self) -> dict:
        a, b, c, d = c_int32(0), c_int32(0), c_int32(0), c_int32(0)
        check(lib().tk_lattice_tiling(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c),
                                      ctypes.byref(d)))  # fmt: skip
        return {"cells_per_lane": a.value, "segs_per_cta": b.value, "rows_per_cta": c.value,
                "ctas_per_sm": d.value}  # fmt: skip

    def set_stream(self, cuda_stream: int):
        check(lib().tk_plan_set_stream(self._h, c_void_p(cuda_stream)))

    # -- state --------------------------------------------------------------------------------
    def set_state_sparse(self, index, value):
        index = _arr(index, np.int64)
        value = _arr(value, np.float64)
        check(lib().tk_plan_set_state_sparse(self._h, len(index), _p(index), _p(value)))

    def set_state(self, u, u_prev=None):
        u = _arr(u, np.float64).reshape(-1)
        up = None if u_prev is None else _arr(u_prev, np.float64).reshape(-1)
        n = self.num_nodes
        if u.size != n or (up is not None and up.size != n):
            raise ValueError(f"state arrays must have {n} entries")
        check(lib().tk_plan_set_state(self._h, _p(u), _p(up)))

    def get_state(self, want_prev: bool = False):
        n = self.num_nodes
        u = np.empty(n, dtype=np.float64)
        up = np.empty(n, dtype=np.float64) if want_prev else None
        check(lib().tk_plan_get_state(self._h, _p(u), _p(up)))
        return (u, up) if want_prev else u

    # -- stepping -----------------------------------------------------------------------------
    def run(self, steps, coeff, damping, *, probes=None, history=False, timed=False,
            probe_out=None, history_out=None, energy=False, history_stride=1,
            kernel_timing=False, fuse=True):  # fmt: skip
        """``steps`` leapfrog updates.  Returns dict(probes, history, energy, elapsed_ms, ...).

        ``timed``: CUDA-event time of the whole stepping loop.  ``kernel_timing``: additionally the summed
        event time of the step kernels alone (bench.py's roofline); it forces direct launches, whereas
        launch-bound runs are otherwise replayed from a captured CUDA graph.  ``fuse=False``: one step per launch
        even where K-step passes would apply (TK_RUN_NO_FUSE)."""
        steps = int(steps)
        pidx = _arr(probes if probes is not None else [], np.int64)
        pout = None
        if len(pidx):
            pout = probe_out if probe_out is not None else np.empty((steps + 1, len(pidx)))
        hout = None
        if history:
            hout = (history_out if history_out is not None
                    else np.empty((steps // max(int(history_stride), 1) + 1, self.num_nodes),
                                  dtype=np.float64))  # fmt: skip
        ms, kms = c_double(0.0), c_double(0.0)
        eout = np.empty(steps, dtype=np.float64) if energy else None
        a = RunArgs(steps, float(coeff), float(damping), len(pidx), _p(pidx), _p(pout), _p(hout),
                    _p(eout), ctypes.cast(ctypes.byref(ms), c_void_p) if timed else None,
                    ctypes.cast(ctypes.byref(kms), c_void_p) if kernel_timing else None,
                    max(int(history_stride), 1), 0 if fuse else 1)  # fmt: skip
        check(lib().tk_plan_run(self._h, ctypes.byref(a)))
        return {"probes": pout, "history": hout, "energy": eout,
                "elapsed_ms": ms.value if timed else None,
                "step_kernel_ms": kms.value if kernel_timing else None}

    def set_probes(self, probes, rows: int):
        """Record these nodes at the next ``rows`` finalise calls of a phase-driven run."""
        pidx = _arr(probes, np.int64)
        self._n_probes = len(pidx)
        check(lib().tk_plan_set_probes(self._h, len(pidx), _p(pidx), int(rows)))

    def read_probes(self, max_rows: int):
        out = np.empty((max_rows, max(getattr(self, "_n_probes", 0), 1)), dtype=np.float64)
        n = c_int64(0)
        check(lib().tk_plan_read_probes(self._h, _p(out), ctypes.byref(n)))
        return out[: n.value, : getattr(self, "_n_probes", 0)]

    # -- slab phases (multi-GPU) --------------------------------------------------------------
    def halo_ptrs(self, next_state: bool = False):
        sl, sh, rl, rh = c_void_p(), c_void_p(), c_void_p(), c_void_p()
        nb = c_int64(0)
        check(lib().tk_lattice_halo_ptrs(self._h, int(next_state), ctypes.byref(sl), ctypes.byref(sh),
                                         ctypes.byref(rl), ctypes.byref(rh), ctypes.byref(nb)))  # fmt: skip
        return {"send_lo": sl.value, "send_hi": sh.value, "recv_lo": rl.value,
                "recv_hi": rh.value, "plane_bytes": nb.value}  # fmt: skip

    def step_begin(self, coeff, damping):
        check(lib().tk_lattice_step_begin(self._h, float(coeff), float(damping)))

    def step_layers(self, zl_begin, zl_end):
        check(lib().tk_lattice_step_layers(self._h, int(zl_begin), int(zl_end)))

    def step_layers_on(self, zl_begin, zl_end, cuda_stream: int):
        check(lib().tk_lattice_step_layers_on(self._h, int(zl_begin), int(zl_end), c_void_p(cuda_stream)))

    # -- fused halo push over peer memory ------------------------------------------------------
    def ipc_export(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        check(lib().tk_lattice_ipc_export(self._h, buf))
        return buf.raw

    def state_ptrs(self):
        a, b = c_void_p(), c_void_p()
        check(lib().tk_lattice_state_ptrs(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def peer_attach(self, side: int, buf0, buf1, peer_layers: int):
        check(lib().tk_lattice_peer_attach(self._h, int(side), c_void_p(buf0), c_void_p(buf1),
                                           int(peer_layers)))

    def set_push_mode(self, in_kernel: bool):
        check(lib().tk_lattice_set_push_mode(self._h, 1 if in_kernel else 0))

    def push_halos(self, cuda_stream: int = 0):
        check(lib().tk_lattice_push_halos(self._h, c_void_p(cuda_stream)))

    def signal(self, value: int, cuda_stream: int = 0):
        check(lib().tk_lattice_signal(self._h, int(value), c_void_p(cuda_stream)))

    def arm_signal(self, value: int):
        """The next step launch over all layers signals ``value`` from inside the kernel (tk_lattice_arm_signal)."""
        check(lib().tk_lattice_arm_signal(self._h, int(value)))

    def signal_pending(self) -> bool:
        v = c_int32(0)
        check(lib().tk_lattice_signal_pending(self._h, ctypes.byref(v)))
        return bool(v.value)

    def wait(self, value: int, cuda_stream: int = 0):
        check(lib().tk_lattice_wait(self._h, int(value), c_void_p(cuda_stream)))

    def step_end(self):
        check(lib().tk_lattice_step_end(self._h))


def ipc_open(handle64: bytes, device: int = 0) -> int:
    p = c_void_p()
    check(lib().tk_ipc_open(device, ctypes.create_string_buffer(handle64, 64), ctypes.byref(p)))
    return p.value


def ipc_close(ptr: int, device: int = 0) -> None:
    check(lib().tk_ipc_close(device, c_void_p(ptr)))


def dft_abs(series, device: int = 0):
    """|DFT| and argmax of one (n,) or many (batch, n) real series on the device."""
    x = np.ascontiguousarray(series, dtype=np.float64)
    single = x.ndim == 1
    x2 = x.reshape(1, -1) if single else x
    spec = np.empty_like(x2)
    arg = np.empty(x2.shape[0], dtype=np.int64)
    check(lib().tk_dft_abs(device, _p(x2), x2.shape[0], x2.shape[1], _p(spec), _p(arg)))
    return (spec[0], int(arg[0])) if single else (spec, arg)


__all__ = ["Plan", "lib", "check", "device_count", "device_info", "device_rates", "dft_abs", "ABI_SYMBOLS",
           "BackendUnavailable", "NoDeviceError", "TkError", "c_uint8"]  # fmt: skip
