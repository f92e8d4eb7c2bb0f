"""NumPy stand-in for the CUDA row-band plan (tetrakis_sim_b200.sheet.CudaSheetBackend): the same
pass_local / all-reduce / pass_commit protocol, arithmetic restated from kernels_fused.cuh.  Lets the CPU suite
drive SheetRunner with world_size > 1 over gloo (tests/test_sheet_gloo.py)."""

import numpy as np
import torch


class NumpySheetBackend:
    def __init__(self, size, row_begin, row_end):
        self.S, self.R = size, row_end - row_begin
        self.u = np.zeros((self.R, size, 4))
        self.up = np.zeros_like(self.u)
        self.exchange = torch.zeros(8 * size + 8, dtype=torch.float64)
        self.cs = np.zeros((size, 4))   # global column sums (installed by pass_commit)
        self.csp = np.zeros((size, 4))
        self.tot = np.zeros(4)
        self.totp = np.zeros(4)
        self.probe_idx, self.probe_rows = [], []

    def set_state_sparse(self, index, value):
        self.u[:] = 0.0
        self.up[:] = 0.0
        self.u.reshape(-1)[np.asarray(index, dtype=np.int64)] = value

    def set_probes(self, probes, rows):
        self.probe_idx, self.probe_rows = list(probes), []

    def read_probes(self, rows):
        return np.array(self.probe_rows).reshape(len(self.probe_rows), len(self.probe_idx))

    def get_state(self):
        return self.u.reshape(-1).copy()

    def _responses(self, X, Xp, k, coeff, g):
        S = self.S
        ca, cb = (2 - g) - coeff * (2 * S + 4), -(1 - g)
        T, Tp = self.tot.copy(), self.totp.copy()
        y, yp, out = np.zeros_like(X), np.zeros_like(X), []
        for _ in range(k):
            yn = ca * y + cb * yp + coeff * y.sum(1, keepdims=True) + coeff * X
            yp, y = y, yn
            out.append(y)
            Xn = 2 * X - Xp + coeff * (X.sum(1, keepdims=True) + T[None, :] - (S + 4) * X) - g * (X - Xp)
            Tn = 2 * T - Tp + coeff * (T.sum() - 4 * T) - g * (T - Tp)
            Xp, X, Tp, T = X, Xn, T, Tn
        return out

    def pass_local(self, k, coeff, g):
        S = self.S
        if k == 0:
            if self.probe_idx:
                self.probe_rows.append(self.u.reshape(-1)[self.probe_idx].copy())
        else:
            ca, cb = (2 - g) - coeff * (2 * S + 4), -(1 - g)
            yR = self._responses(self.u.sum(1), self.up.sum(1), k, coeff, g)   # row sums are band-local
            yC = self._responses(self.cs.copy(), self.csp.copy(), k, coeff, g)  # column sums are global
            u, up = self.u, self.up
            for j in range(k):
                u, up = ca * u + cb * up + coeff * u.sum(2, keepdims=True), u
                if self.probe_idx:
                    full = u + yR[j][:, None, :] + yC[j][None, :, :]
                    self.probe_rows.append(full.reshape(-1)[self.probe_idx].copy())
            zero = np.zeros_like(yR[0])
            zc = np.zeros_like(yC[0])
            self.u = u + yR[k - 1][:, None, :] + yC[k - 1][None, :, :]
            self.up = up + (yR[k - 2] if k >= 2 else zero)[:, None, :] + (yC[k - 2] if k >= 2 else zc)[None, :, :]
        ex = np.concatenate([self.u.sum(0).ravel(), self.up.sum(0).ravel(), self.u.sum((0, 1)), self.up.sum((0, 1))])
        self.exchange.copy_(torch.from_numpy(ex))

    def pass_commit(self):
        ex = self.exchange.numpy()
        S = self.S
        self.cs = ex[: 4 * S].reshape(S, 4).copy()
        self.csp = ex[4 * S : 8 * S].reshape(S, 4).copy()
        self.tot, self.totp = ex[8 * S : 8 * S + 4].copy(), ex[8 * S + 4 :].copy()
