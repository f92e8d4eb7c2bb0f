"""Microseconds per step on small (L2-resident, launch-bound) defect-free sheets through the public API:
fused passes (default) against one step per launch (TK_LAT_FUSE=0).  Prints a markdown table."""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetrakis_sim_b200 as tk  # noqa: E402

steps = 960
print("| sheet | one step per launch: us/step | fused passes: us/step | GNUPS fused |")
print("|---|---|---|---|")
for size in (11, 41, 128, 512, 1024, 2048, 4096):
    spec = tk.LatticeSpec(size=size, dim=2)
    kick = spec.bench_kick_node()
    row = []
    for fuse in ("0", None):
        if fuse is None:
            os.environ.pop("TK_LAT_FUSE", None)
        else:
            os.environ["TK_LAT_FUSE"] = fuse
        best = None
        for _ in range(3):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                h = tk.run_wave_sim(spec, steps=steps, initial_data={kick: 1.0}, c=1.0, dt=0.1, probes=[kick])
            best = h.elapsed_ms if best is None else min(best, h.elapsed_ms)
        row.append(best)
    print(f"| 2-D {size}^2 | {1e3 * row[0] / steps:.2f} | {1e3 * row[1] / steps:.3f} | "
          f"{spec.num_nodes * steps / (row[1] * 1e-3) / 1e9:.1f} |")
