#!/usr/bin/env python3
"""Where does a run_wave_sim(G) call on a reference-sized graph spend its wall time?  (host-side profiling aid)
Wraps every C-ABI entry point with a timer and prints the per-call breakdown of the first calls."""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetrakis_sim_b200 as tk
from tetrakis_sim_b200 import _lib

size, dim, layers = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 5
spec = tk.LatticeSpec(size=size, dim=dim, layers=layers)
G = spec.to_networkx()
kick = spec.bench_kick_node()
warnings.simplefilter("ignore", RuntimeWarning)
L = _lib.lib()
acc = {}


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t0 = time.perf_counter()
        r = self.fn(*a)
        acc[self.name] = acc.get(self.name, 0.0) + time.perf_counter() - t0
        return r


class Proxy:
    def __getattr__(self, name):
        return Timed(name, getattr(L, name))


_lib._LIB = Proxy()
for i in range(6):
    acc.clear()
    t0 = time.perf_counter()
    h = tk.run_wave_sim(G, steps=50, initial_data={kick: 1.0}, dt=0.1)
    wall = time.perf_counter() - t0
    parts = ", ".join(f"{k[3:]} {1e3 * v:.2f}" for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:6])
    print(f"call {i}: {1e3 * wall:.2f} ms wall, device {h.elapsed_ms:.3f} ms | C ABI: {parts} | python {1e3 * (wall - sum(acc.values())):.2f}")
