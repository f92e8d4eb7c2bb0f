#!/usr/bin/env python3
"""Short run of one 3-D / defect workload through the C ABI (profiling aid: few launches, no CPU leg).

    python scripts/fx_run.py cfg4|cfg3|sheet [steps] [size] [layers]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import tetrakis_sim_b200 as tk

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
if name == "cfg4":
    S, L = (int(sys.argv[3]) if len(sys.argv) > 3 else 512), (int(sys.argv[4]) if len(sys.argv) > 4 else 512)
    spec = tk.LatticeSpec(size=S, dim=3, layers=L, defect="singularity", radius=2.5, mass=2000.0, prune_edges=True)
elif name == "cfg3":
    S, L = (int(sys.argv[3]) if len(sys.argv) > 3 else 1024), (int(sys.argv[4]) if len(sys.argv) > 4 else 64)
    spec = tk.LatticeSpec(size=S, dim=3, layers=L, defect="blackhole", radius=64.0)
else:
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
    spec = tk.LatticeSpec(size=S, dim=2, defect="blackhole", radius=8.0)
plan = tk.compile_lattice(spec)
coeff = min(0.1, 1.0 / np.sqrt(plan.max_degree)) ** 2
k = spec.index(spec.kick_node())
plan.set_state_sparse([k], [1.0])
plan.run(16, coeff, 0.0, probes=[k])
t0 = time.perf_counter()
out = plan.run(steps, coeff, 0.0, probes=[k], timed=True, kernel_timing=True)
wall = time.perf_counter() - t0
print(f"{name} {S}x{S}x{spec.layers}: {steps} steps {out['elapsed_ms']:.3f} ms device, pass kernels {out['step_kernel_ms']:.3f} ms, "
      f"{spec.num_nodes * steps / out['elapsed_ms'] / 1e6:.1f} GNUPS, fused {plan.fused}, cross {plan.cross}, wall {wall:.3f} s")
plan.close()
