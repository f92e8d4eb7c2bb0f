"""Small driver for profiling the ensemble kernel: a 4 736-simulation slice of BASELINE config 5."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetrakis_sim_b200 as tk

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
specs, damp, dt = tk.sweep_grid(13, 7, np.linspace(0.5, 4.0, 64), np.linspace(0.0, 0.05, 32), np.linspace(0.02, 0.2, 32))
sel = np.linspace(0, len(specs) - 1, n).astype(int)
for rep in range(2):
    t0 = time.perf_counter()
    res = tk.run_ensemble([specs[i] for i in sel], 50, c=1.0, dt=dt[sel], damping=damp[sel])
    print(n, "sims kernel ms", res.elapsed_ms, "GNUPS", res.node_updates / res.elapsed_ms / 1e6, "wall", time.perf_counter() - t0)
