"""NumPy/oracle stand-in for ``tk_ensemble_run_ex`` (test infrastructure): reads the ``tk_ensemble_desc`` the host code
hands to the C ABI and computes every member with the oracle's clique-sum restatement, so that the HOST side of the
ensembles -- mask arrays, kick rule, per-simulation CFL clamp, block distribution over ranks, gather -- can be checked
on the CPU against the literal restatement on reference-built graphs (tests/test_ensemble_gloo.py).  On the GPU the
same descriptor drives ``ensemble_kernel`` (tests/test_gpu_ensemble.py)."""

import ctypes

import numpy as np

from oracle import wave_oracle as wo


def _view(ptr, ctype, count):
    if not ptr or count == 0:
        return np.zeros(0, dtype=np.dtype(ctype))
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), (count,))


class FakeEnsembleLib:
    """Just enough of the library for ``run_ensemble``."""

    def __init__(self):
        self.calls = []

    def tk_ensemble_run_ex(self, device, desc_ref, series, spectrum, argmax, features, elapsed_ms):
        d = desc_ref._obj
        S, L, n_masks, n, steps = d.size, d.layers, d.n_masks, d.n_sims, d.steps
        ncell, N, nt = L * S * S, L * S * S * 4, steps + 1
        bits = _view(d.cell_bits, ctypes.c_uint8, n_masks * ncell).reshape(n_masks, ncell)
        exc_off = _view(d.exc_offset, ctypes.c_int32, n_masks + 1)
        exc_node = _view(d.exc_node, ctypes.c_int32, int(exc_off[-1]))
        exc_im = _view(d.exc_inv_mass, ctypes.c_double, int(exc_off[-1]))
        exc_v = _view(d.exc_potential, ctypes.c_double, int(exc_off[-1]))
        corr_off = _view(d.corr_offset, ctypes.c_int32, n_masks + 1)
        corr_a = _view(d.corr_a, ctypes.c_int32, int(corr_off[-1]))
        corr_b = _view(d.corr_b, ctypes.c_int32, int(corr_off[-1]))
        corr_s = _view(d.corr_sign, ctypes.c_double, int(corr_off[-1]))
        sim_mask = _view(d.sim_mask, ctypes.c_int32, n)
        sim_kick = _view(d.sim_kick, ctypes.c_int32, n)
        coeff = _view(d.sim_coeff, ctypes.c_double, n)
        damping = _view(d.sim_damping, ctypes.c_double, n)
        out_series = _view(series, ctypes.c_double, n * nt).reshape(n, nt)
        out_spec = _view(spectrum, ctypes.c_double, n * nt).reshape(n, nt) if spectrum else None
        out_arg = _view(argmax, ctypes.c_int64, n) if argmax else None
        self.calls.append((device, int(n)))
        for i in range(n):
            m = int(sim_mask[i])
            part = (bits[m][:, None] >> np.arange(4)) & 1
            alive = (bits[m][:, None] >> (4 + np.arange(4))) & 1
            flags = (alive + 2 * part).astype(np.uint8).reshape(-1)
            im, pot = np.ones(N), np.zeros(N)
            sl = slice(int(exc_off[m]), int(exc_off[m + 1]))
            im[exc_node[sl]], pot[exc_node[sl]] = exc_im[sl], exc_v[sl]
            cs = slice(int(corr_off[m]), int(corr_off[m + 1]))
            corr = list(zip(corr_a[cs].tolist(), corr_b[cs].tolist(), corr_s[cs].tolist()))
            u0 = np.zeros(N)
            u0[sim_kick[i]] = 1.0
            r = wo.run_structured(S, L, flags, im, pot, corr, u0, steps, float(coeff[i]), float(damping[i]),
                                  probes=[int(sim_kick[i])])
            out_series[i] = r["probes"][:, 0]
            if out_spec is not None:
                out_spec[i] = np.abs(np.fft.fft(out_series[i]))
                if out_arg is not None:
                    out_arg[i] = int(np.argmax(out_spec[i]))
        if elapsed_ms is not None:
            elapsed_ms._obj.value = 1.0
        return 0
