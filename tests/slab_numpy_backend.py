"""TEST INFRASTRUCTURE: a NumPy stand-in for one slab's CUDA plan (same phase API), built on the
oracle's clique-sum formula, so the multi-rank orchestration in tetrakis-sim_b200/slab.py can be
driven on CPU over gloo (there is no GPU in the build container).  Never used by the product."""

from __future__ import annotations

import numpy as np
import torch


class NumpySlabBackend:
    def __init__(self, spec, z_begin, z_end):
        S, L = spec.size, spec.layers
        self.spec, self.zb, self.ze, self.S = spec, z_begin, z_end, S
        self.Lloc = z_end - z_begin
        flags = np.full(spec.num_slots, 3, dtype=np.uint8)
        im = np.ones(spec.num_slots)
        pot = np.zeros(spec.num_slots)
        si, sf, sm, sv = spec.specials()
        flags[si], im[si], pot[si] = sf, sm, sv

        def padded(a, fill):
            out = np.full((self.Lloc + 2, S, S, 4), fill, dtype=a.dtype)
            g = a.reshape(L, S, S, 4)
            lo, hi = max(z_begin - 1, 0), min(z_end + 1, L)
            out[lo - (z_begin - 1) : hi - (z_begin - 1)] = g[lo:hi]
            return out

        self.A = padded((flags & 1).astype(np.float64), 0.0)
        self.P = padded(((flags >> 1) & 1).astype(np.float64), 0.0)
        self.IM = padded(im, 1.0)
        self.V = padded(pot, 0.0)
        self.corr = [(a, b, s) for a, b, s in spec.corrections()
                     if z_begin <= a // spec.plane < z_end]
        self.buf = [np.zeros((self.Lloc + 2, S, S, 4)), np.zeros((self.Lloc + 2, S, S, 4))]
        self.cur = 0
        self.coeff = self.damping = 0.0

    # ---- the plan's phase API --------------------------------------------------------------
    @property
    def num_nodes(self):
        return self.Lloc * self.spec.plane

    def set_state_sparse(self, idx, val):
        for b in self.buf:
            b[:] = 0.0
        self.cur = 0
        flat = self.buf[0][1:-1].reshape(-1)
        for i, v in zip(idx, val):
            flat[i] = v

    def step_begin(self, coeff, damping):
        self.coeff, self.damping = coeff, damping

    def step_layers(self, a, b):
        U, W, P = self.buf[self.cur], self.buf[self.cur ^ 1], self.P
        sl = slice(a + 1, b + 1)
        pu = P * U
        me, p = pu[sl], P[sl]
        nsum = p * ((me.sum(3, keepdims=True) - me) + (me.sum(2, keepdims=True) - me)
                    + (me.sum(1, keepdims=True) - me) + pu[a : b] + pu[a + 2 : b + 2])
        deg = p * ((p.sum(3, keepdims=True) - 1) + (p.sum(2, keepdims=True) - 1)
                   + (p.sum(1, keepdims=True) - 1) + P[a : b] + P[a + 2 : b + 2])
        u = U[sl]
        lap = nsum - deg * u
        off = (self.zb - 1) * self.spec.plane  # padded flat index = global - off
        uf = U.reshape(-1)
        for ca, cb, s in self.corr:
            zl = ca // self.spec.plane - self.zb
            if a <= zl < b:
                lap.reshape(-1)[ca - off - (a + 1) * self.spec.plane] += s * (uf[cb - off] - uf[ca - off])
        W[sl] = (2 * u - W[sl] + (self.coeff * self.IM[sl]) * lap - (self.coeff * self.V[sl]) * u
                 - self.damping * (u - W[sl])) * self.A[sl]

    def step_end(self):
        self.cur ^= 1

    def planes(self, next_state):
        b = self.buf[self.cur ^ 1 if next_state else self.cur]
        t = lambda k: torch.from_numpy(b[k].reshape(-1))  # noqa: E731  (views: recv writes in place)
        return {"recv_lo": t(0), "send_lo": t(1), "send_hi": t(self.Lloc), "recv_hi": t(self.Lloc + 1)}

    def get_state(self):
        return self.buf[self.cur][1:-1].reshape(-1).copy()
