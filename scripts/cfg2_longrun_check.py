"""Config 2's step count (10 000, CFL-clamped dt) on a 2048^2 defect-free sheet: the K-step fused passes on the GPU
against the C oracle stepping one step at a time.  Prints one JSON line (kept in profiles/).

    python scripts/cfg2_longrun_check.py [size] [steps]
"""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetrakis_sim_b200 as tk  # noqa: E402
from oracle import wave_oracle as wo  # noqa: E402  (the checker)

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
spec = tk.LatticeSpec(size=size, dim=2)
kick = spec.bench_kick_node()
t0 = time.perf_counter()
with warnings.catch_warnings():
    warnings.simplefilter("ignore", RuntimeWarning)
    hist = tk.run_wave_sim(spec, steps=steps, initial_data={kick: 1.0}, c=1.0, dt=0.1, probes=[kick])
t_gpu = time.perf_counter() - t0
u0 = np.zeros(spec.num_slots)
u0[spec.index(kick)] = 1.0
t0 = time.perf_counter()
ref = wo.run_structured(size, 1, None, None, None, [], u0, steps, (1.0 * hist.dt) ** 2, 0.0, probes=[spec.index(kick)])
t_cpu = time.perf_counter() - t0
scale = float(np.max(np.abs(ref["u"])))
freq, spectrum, _ = tk.run_fft(hist, kick)
ref_spec = wo.run_fft(ref["probes"][:, 0], hist.dt)[1]
print(json.dumps({
    "lattice": f"2-D {size}x{size}, no defect", "steps": steps, "effective_dt": hist.dt, "fused": hist.fused,
    "rel_inf_error_final_state": float(np.max(np.abs(hist.final_state - ref["u"])) / scale),
    "rel_inf_error_probe_series": float(np.max(np.abs(hist.probe_series[:, 0] - ref["probes"][:, 0])) /
                                        np.max(np.abs(ref["probes"]))),
    "sum_u": float(hist.final_state.sum()), "sum_u_expected": 1.0 + steps,
    "dominant_bin_gpu": int(wo.folded_bin(spectrum)), "dominant_bin_oracle": int(wo.folded_bin(ref_spec)),
    "spectrum_rel_error": float(np.max(np.abs(spectrum - ref_spec)) / np.max(ref_spec)),
    "gpu_seconds_wall": t_gpu, "cpu_oracle_seconds": t_cpu, "cpu_threads": os.cpu_count(),
}))
