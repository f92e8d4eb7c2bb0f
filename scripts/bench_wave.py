#!/usr/bin/env python3
"""``scripts/bench_wave.py`` of the reference, on the B200 path.

Same flags and the same CSV (header and column meaning, reference scripts/bench_wave.py:128-146):
``dim,layers,size,steps,repeats,nodes,edges,max_degree,build_s,sim_s_median,sim_s_min,sim_s_max,
effective_dt,stability_adjusted``.  ``build_s`` is the graph-to-device compile, ``sim_s_*`` the wall
time of ``run_wave_sim`` (kick ``(size//2, size//2[, layers//2], "A")``, bench_wave.py:43-46).
``--path graph`` builds the networkx graph (``LatticeSpec.to_networkx``) and runs the bit-exact sliced-ELL
kernel (sizes networkx can hold), ``--path graph-structured`` the same graph recognised as a tetrakis lattice and run
on the structured kernel; the default ``--path spec`` uses a LatticeSpec and scales to any size.
``--gnups`` appends node-updates/s columns.
"""

from __future__ import annotations

import argparse
import os
import statistics
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetrakis_sim_b200 as tk  # noqa: E402

HEADER = ("dim,layers,size,steps,repeats,nodes,edges,max_degree,build_s,sim_s_median,sim_s_min,"
          "sim_s_max,effective_dt,stability_adjusted")


def edge_count(size: int, dim: int, layers: int) -> int:
    """Edges of build_sheet (lattice.py:37-44, 59-72): per layer 6 per cell + row and column cliques."""
    per_layer = 6 * size * size + 2 * 4 * size * (size * (size - 1) // 2)
    L = layers if dim == 3 else 1
    return L * per_layer + (L - 1) * size * size * 4


def bench_one(size, dim, layers, steps, c, dt, damping, repeats, path):
    spec = tk.LatticeSpec(size=size, dim=dim, layers=layers)
    kick = spec.bench_kick_node()
    t0 = time.perf_counter()
    if path in ("graph", "graph-structured"):
        target = spec.to_networkx()
    else:
        target = spec
    build_s = time.perf_counter() - t0
    times, dts, adjusted, hist = [], [], False, None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for _ in range(repeats):
            t1 = time.perf_counter()
            hist = tk.run_wave_sim(target, steps=steps, initial_data={kick: 1.0}, c=c, dt=dt, damping=damping,
                                   final_state=path != "spec")  # the script only reads hist.dt / metadata
            times.append(time.perf_counter() - t1)
            dts.append(float(hist.dt))
            adjusted = adjusted or bool(hist.metadata.get("stability_adjusted", False))
    return dict(dim=dim, layers=layers if dim == 3 else 1, size=size, steps=steps, repeats=repeats,
                nodes=spec.num_nodes, edges=edge_count(size, dim, layers),
                max_degree=int(hist.metadata["max_degree"]), build_s=build_s,
                sim_s_median=statistics.median(times), sim_s_min=min(times), sim_s_max=max(times),
                effective_dt=statistics.median(dts), stability_adjusted=adjusted,
                device_ms=hist.elapsed_ms)


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description="Benchmark run_wave_sim scaling on the GPU.")
    p.add_argument("--sizes", nargs="+", default=["11,21,31"])
    p.add_argument("--dim", type=int, choices=[2, 3], default=2)
    p.add_argument("--layers", type=int, default=5)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--repeats", type=int, default=3)
    p.add_argument("--c", type=float, default=1.0)
    p.add_argument("--dt", type=float, default=0.1)
    p.add_argument("--damping", type=float, default=0.0)
    p.add_argument("--path", choices=["spec", "graph", "graph-structured"], default="spec",
                   help="graph-structured: the networkx graph, recognised as a tetrakis lattice and run on the "
                        "structured kernel (what TETRAKIS_B200_AUTO_STRUCTURED=1 / install(structured=True) does)")
    p.add_argument("--gnups", action="store_true", help="append node-updates/s columns")
    a = p.parse_args(argv)
    if a.path == "graph-structured":
        os.environ["TETRAKIS_B200_AUTO_STRUCTURED"] = "1"
    sizes = [int(t) for item in a.sizes for t in item.split(",") if t.strip()]
    if not sizes:
        raise SystemExit("No sizes provided.")
    print(HEADER + (",gnups_wall,gnups_device" if a.gnups else ""))
    # untimed warm-up: CUDA context creation and lazy kernel loading (~1 s) belong to no particular size
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        wspec = tk.LatticeSpec(size=3, dim=a.dim, layers=2)
        tk.run_wave_sim(wspec.to_networkx() if a.path != "spec" else wspec, steps=9, dt=a.dt)
    for s in sizes:
        r = bench_one(s, a.dim, a.layers, a.steps, a.c, a.dt, a.damping, a.repeats, a.path)
        line = (f"{r['dim']},{r['layers']},{r['size']},{r['steps']},{r['repeats']},{r['nodes']},{r['edges']},"
                f"{r['max_degree']},{r['build_s']:.6f},{r['sim_s_median']:.6f},{r['sim_s_min']:.6f},"
                f"{r['sim_s_max']:.6f},{r['effective_dt']:.6f},{int(r['stability_adjusted'])}")
        if a.gnups:
            nu = r["nodes"] * r["steps"]
            line += f",{nu / r['sim_s_median'] / 1e9:.4f},{nu / (r['device_ms'] * 1e-3) / 1e9:.4f}"
        print(line)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
