#!/usr/bin/env python3
"""Benchmark of the wave-propagation hot path (BASELINE.json metric: node-updates/s and % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one leapfrog update of every node of the lattice.
Workloads (config.workload):
  cfg2-2d-8192      BASELINE configs[1]: 2-D tetrakis 8192x8192, no defect, fp64   (default at N=1).  The sheet
                    has no defect, so tk_plan_run advances it in K-step passes (kernels_fused.cuh, 32/K bytes
                    per update); TK_LAT_FUSE=0 measures the step-by-step kernel (24 bytes per update) instead.
                    At N>1 with this name: N independent replicas.
  cfg2-sharded      the same sheet grown with N (size 8192*sqrt(N)) and held as N row bands, one all-reduce of
                    the column sums per pass; --size fixes the sheet (strong scaling)    (default at N>1, with
                    the cfg3 slab run of the same launch reported under "sharded_3d")
  cfg3-slab-1024x64 BASELINE configs[2] shape: 3-D 1024x1024, 64 layers PER GPU, black hole r=64 at the
                    global centre, layer slabs + halo push over NVLink peer memory (or NCCL)
  cfg4-3d-512       BASELINE configs[3]: 3-D 512^3 singularity, mass 2000, pruned edges (K-step passes with the
                    cross scheme of kernels_cross.cuh; TK_LAT_FUSE=0: one step per launch)
  cfg5-ensemble     BASELINE configs[4]: 65 536 independent 13x13x7 simulations
  cfg1-batch        BASELINE configs[0]: tetrakis-batch --dim 3 --size 11 --layers 7 --radius 2.5 --steps 50 through
                    the drop-in run_wave_sim(G) on the networkx graph, next to the literal CPU restatement (1 core)
The default N=1 line (cfg2) also carries "sharded_3d" (the cfg3 slab through the multi-GPU slab code with no
neighbours: the N=1 point of the 3-D weak-scaling series) and "other_configs" (cfg4, cfg3 fused, cfg5, cfg1).
  ell-tetrakis-64x16 / ell-grid-256   the generic sliced-ELL graph kernel on an explicit CSR (44 + 4w bytes/update)
Prints ONE JSON line (rank 0).  `value` is device-timed with the state resident in HBM; `e2e` is the same
metric through the C ABI with host buffers (kicks H2D, probe series + final state D2H inside the timed
region); `roofline` is the dominant kernel's algorithmic bytes / its CUDA-event time against the measured HBM
peak; `cpu_baseline` is the oracle (oracle/wave_oracle.c, OpenMP) on this box's host cores.
`--impl reference` times that CPU restatement alone (the reference itself is pure Python over networkx and
cannot even build these lattices: 2.2e12 edges for cfg2 -- BASELINE.md section 1).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "node-updates/sec"
UNIT = "GNUPS"
BYTES_PER_UPDATE = 24.0  # read u_t, read u_{t-1}, write u_{t+1}; fp64


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, k=0):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None.
    A capture of a K-step pass kernel only counts for a run with the same K."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t[workload]
        if k and int(e.get("steps_per_pass", k)) != int(k):
            e = t.get(f"{workload}-k{int(k)}")  # a capture at another K, if one was taken
            if e is None:
                return None
        return float(e["bytes"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    nvidia-smi takes a few hundred ms to start, so the sampler is started early (``start()``), the
    caller brackets the timed region with ``mark_begin()`` / ``mark_end()``, and only rows that
    arrived inside the bracket count (falling back to every row under load if the region was shorter
    than one sampling period)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            deadline = time.perf_counter() + 5.0
            while not self.rows and time.perf_counter() < deadline:
                time.sleep(0.01)
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __enter__(self):
        if self.proc is None:
            self.start()
        self.mark_begin()
        return self

    def __exit__(self, *exc):
        self.mark_end()
        time.sleep(0.06)  # let the last sample of the region arrive
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.06]
        scope = "timed region"
        if not inside:
            inside, scope = [r for _, r in self.rows], "warm-up + timed region (region shorter than a sample)"
        sm, mx, reasons, pw = [], [], set(), []
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw), "scope": scope}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C/OpenMP clique-sum restatement on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_run(workload: str, steps: int, warmup: int, budget_s: float = 0.0):
    """The oracle's C/OpenMP clique-sum restatement on a FIXED bounded sample of the workload: 2-D 4096^2 for the
    sheet workloads, 3-D 512 x 512 x 16 for the 3-D ones (no defect), all host threads."""
    from oracle import wave_oracle as wo

    cores = os.cpu_count() or 1
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    three_d = workload.startswith("cfg3") or workload.startswith("cfg4")
    S, L = (512, 16) if three_d else (4096, 1)

    def one(nsteps):
        u0 = np.zeros(L * S * S * 4)
        u0[((L // 2 * S + S // 2) * S + S // 2) * 4] = 1.0
        maxdeg = 3 + 2 * (S - 1) + (2 if L > 1 else 0)
        coeff = 1.0 / maxdeg  # dt at the CFL limit, c = 1
        t0 = time.perf_counter()
        wo.run_structured(S, L, None, None, None, [], u0, nsteps, coeff, 0.0)
        return time.perf_counter() - t0, L * S * S * 4

    if warmup:
        one(warmup)
    t, n = one(steps)
    sample = (f"{'3-D' if three_d else '2-D'} tetrakis {S}x{S}" + (f"x{L}" if three_d else "") +
              f" no defect (fixed sample), {steps} steps, oracle/wave_oracle.c clique-sum restatement, OpenMP {cores} threads")
    return {"value": n * steps / t / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "seconds": t, "ms_per_step": 1e3 * t / max(steps, 1)}


def cpu_literal_cfg1(steps: int = 50):
    """BASELINE configs[0] on one host core with the LITERAL restatement of physics.py:106-129 (per-node neighbour
    loops over the adjacency lists, oracle/wave_oracle.c:run_graph) -- what the reference's arithmetic costs once
    the Python interpreter is taken out; the reference itself (Python over networkx) measured 0.59 s for the same
    run on this image's host."""
    from oracle import lattice_oracle as lo
    from oracle import wave_oracle as wo

    G, removed, center = lo.build_case(11, 3, 7, "blackhole", radius=2.5)
    kick = lo.pick_kick_node(G, center, 3)
    nodes, rowptr, colidx, inv_mass, potential = lo.graph_arrays(G)
    deg = np.array([G.degree[n] for n in nodes], dtype=np.float64)
    u0 = np.zeros(len(nodes))
    u0[nodes.index(kick)] = 1.0
    dt_eff, _, _ = wo.cfl_clamp(int(deg.max()), 1.0, 0.2, steps, 0.0)
    wo.run_graph(rowptr, colidx, deg, inv_mass, potential, u0, 2, dt_eff ** 2, 0.0)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        wo.run_graph(rowptr, colidx, deg, inv_mass, potential, u0, steps, dt_eff ** 2, 0.0)
        best = min(best, time.perf_counter() - t0)
    return {"value": len(nodes) * steps / best / 1e9, "unit": UNIT, "cores": 1, "kind": "port", "seconds": best,
            "sample": f"cfg1 whole (3 064 nodes, {int(rowptr[-1]) // 2} edges, {steps} steps), literal per-edge loop in C, 1 core"}


def reference_arm(args, rank):
    if rank != 0:
        return 0
    workload = args.workload or ("cfg2-2d-8192" if args.gpus == 1 else "cfg2-sharded")
    res = cpu_run(workload, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload, "sample": res["sample"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def pinned_empty(n):
    import torch

    t = torch.empty(n, dtype=torch.float64, pin_memory=True)
    return t, t.numpy()


WORKLOADS = {
    "cfg2-2d-8192": dict(dim=2, size=8192, layers=1, kw={}, dt=0.1,
                         desc="2-D tetrakis {S}x{S}, no defect, fp64, kick (S/2,S/2,'A'), dt 0.1 -> CFL-clamped"),
    "cfg3-slab-1024x64": dict(dim=3, size=1024, layers=64, kw=dict(defect="blackhole", radius=64.0), dt=0.1,
                              desc="3-D tetrakis {S}x{S}x{L}, black hole r=64 at the centre, fp64"),
    "cfg4-3d-512": dict(dim=3, size=512, layers=512,
                        kw=dict(defect="singularity", radius=2.5, mass=2000.0, prune_edges=True), dt=0.1,
                        desc="3-D tetrakis {S}x{S}x{L} singularity mass 2000 + prune_edges, fp64"),
}


def bench_lattice(workload, args, local=0, dist=None, steps=None, warmup=None, size=None, layers=None, fuse=True,
                  clk=None, want_e2e=True):
    """One lattice on one GPU through the C ABI (tk_plan_run).  The timed region -- `steps` steps, synchronised on
    both sides -- is repeated until >= 50 ms and >= 5 repeats have been seen (the state stays resident and the
    simulation simply continues); the reported time is the MEDIAN.  Returns the fields of a bench line."""
    import torch

    import tetrakis_sim_b200 as tk
    from tetrakis_sim_b200 import _lib

    w = WORKLOADS[workload]
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    world = int(os.environ.get("WORLD_SIZE", "1")) if dist else 1
    dev = torch.device(f"cuda:{local}")

    def fence():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    spec = tk.LatticeSpec(size=size or w["size"], dim=w["dim"], layers=layers or w["layers"], **w["kw"])
    desc = w["desc"].format(S=spec.size, L=spec.layers)
    kick = spec.kick_node() if spec.defect != "none" else spec.bench_kick_node()
    f32 = args.dtype == "f32"
    bpu = BYTES_PER_UPDATE / 2 if f32 else BYTES_PER_UPDATE
    t_build = time.perf_counter()
    plan = tk.compile_lattice(spec, device=local, dtype="float32" if f32 else "float64")
    maxdeg = plan.max_degree
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build  # graph-to-device compile (excluded from every rate below)
    dt_eff = min(w["dt"], 1.0 / np.sqrt(maxdeg))
    coeff = dt_eff ** 2
    n_updates = spec.num_nodes  # surviving nodes
    kidx = [spec.index(kick)]

    plan.set_state_sparse(kidx, [1.0])
    plan.run(warmup, coeff, 0.0, probes=kidx, fuse=fuse)
    fence()
    reps = []
    if clk is not None:
        clk.mark_begin()
    while len(reps) < 5 or (sum(r[0] for r in reps) < 50.0 and len(reps) < 40):
        l0 = plan.launches
        fence()
        out = plan.run(steps, coeff, 0.0, probes=kidx, timed=True, kernel_timing=True, fuse=fuse)
        fence()
        reps.append((max_over_ranks(out["elapsed_ms"]), out["step_kernel_ms"], plan.launches - l0))
    if clk is not None:
        clk.mark_end()
    order = sorted(range(len(reps)), key=lambda i: reps[i][0])
    ms, kms, launches = reps[order[len(reps) // 2]]  # the median repeat (its own kernel time and launch count)
    steps_done = warmup + steps * len(reps)
    value = world * n_updates * steps / (ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()

    # ---- the dominant kernel and its algorithmic bytes
    fused, cross = plan.fused, plan.cross
    fk = fused["steps_per_pass"]
    nsingle = steps - fused["fused_steps"]
    slots = spec.num_slots
    if cross["steps_per_pass"]:
        # cross scheme (kernels_cross.cuh): a bulk pass reads u_t, u_{t-1} and writes u_{t+K}, u_{t+K-1} of the BULK nodes
        bulk = slots * (1.0 - cross["cross_fraction"])
        npass = fused["fused_steps"] // fk + (1 if fused["fused_steps"] % fk else 0)
        pass_bytes = 4 * (bpu / 3) * bulk
        alg_bytes = npass * pass_bytes + nsingle * bpu * n_updates
        kernel = (f"fx_bulk3d_kernel<{fk}>" if spec.dim == 3 else f"lat_fused2d_kernel<CROSS> (K={fk})") + \
                 " (the bulk pass of the cross scheme)"
        cross_bytes = fused["fused_steps"] * bpu * slots * cross["cross_fraction"]
        extra = {"cross_stage": {"rows": cross["cross_rows"], "cols": cross["cross_cols"],
                                 "fraction_of_lattice": cross["cross_fraction"],
                                 "algorithmic_bytes": cross_bytes,
                                 "note": "the cross cells and every line sum are stepped one step at a time "
                                         "(fx_cross_step_kernel + fx_lines_kernel, 2 launches per step); their time is "
                                         "the part of the step outside kernel_share_of_step"},
                 "bytes_per_update_all_kernels": (alg_bytes + cross_bytes) / (n_updates * steps)}
        traffic_key = workload + "-fused"
    elif fk:
        npass = -(-fused["fused_steps"] // fk)
        pass_bytes = 4 * (bpu / 3) * n_updates
        alg_bytes = npass * pass_bytes + nsingle * bpu * n_updates
        kernel = f"lat_fused2d_kernel{'<float>' if f32 else ''} (K={fk})"
        extra, traffic_key = {}, workload + "-fused"
    else:
        npass, pass_bytes = 0, bpu * n_updates
        alg_bytes = nsingle * bpu * n_updates
        kernel = "lat_step_f32_kernel" if f32 else "lat_step_kernel"
        extra, traffic_key = {}, workload
    kernel_launches = npass + nsingle
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    traffic = None if f32 else ncu_traffic(traffic_key, fk)
    checksum = float(plan.get_state().sum()) if slots <= 600_000_000 else None

    e2e = None
    if want_e2e:
        # e2e through the C ABI with host buffers, device-synchronised wall clock: kick H2D, `steps` steps, the probe
        # series D2H and its |DFT| + dominant bin (what tetrakis-batch writes out); "with_final_state" adds the D2H
        # of the whole final state into pinned host memory (what run_wave_sim's history.final_state holds)
        hold, final = pinned_empty(plan.num_nodes)
        fence()
        t0 = time.perf_counter()
        plan.set_state_sparse(kidx, [1.0])
        o2 = plan.run(steps, coeff, 0.0, probes=kidx, fuse=fuse)
        spectrum, dom = tk.dft_abs(o2["probes"][:, 0], device=local)
        fence()
        t1 = time.perf_counter()
        _lib.check(_lib.lib().tk_plan_get_state(plan._h, _lib._p(final), None))
        fence()
        t2 = time.perf_counter()
        e_probe, e_full = max_over_ranks(t1 - t0), max_over_ranks(t2 - t0)
        e2e = {"value": world * n_updates * steps / e_probe / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": world * (16.0 + o2["probes"].nbytes) / steps,
               "d2h_bytes_per_step": world * (o2["probes"].nbytes + spectrum.nbytes + 8) / steps,
               "seconds": e_probe,
               "includes": "set_state_sparse (kick H2D) + run + probe series D2H + probe series H2D, |DFT| and "
                           "dominant bin D2H (tk_dft_abs)",
               "with_final_state": {"value": world * n_updates * steps / e_full / 1e9, "seconds": e_full,
                                    "d2h_bytes_per_step": world * (final.nbytes + o2["probes"].nbytes) / steps,
                                    "includes": "the above + the whole final state D2H into pinned host memory"}}
    tiling = plan.tiling
    if fk:
        tiling = {"steps_per_pass": fk, "single_steps": nsingle, "passes": npass,
                  "temporal_blocking": ("every node update is computed, the intermediate states are not stored; " +
                                        ("cross scheme: the rows / columns of the defect and of the probe are stepped "
                                         "explicitly with every line sum, the bulk takes one pass per K steps"
                                         if cross["steps_per_pass"] else
                                         "the row/column sums of the intermediate steps follow from closed recurrences "
                                         "because the sheet has no defect") +
                                        ".  TK_LAT_FUSE=0 runs one step per launch (24 B per update).")}
        if not cross["steps_per_pass"]:
            tiling.update(rows_per_cta=fused["rows_per_cta"], ctas_per_sm=fused["ctas_per_sm"])
    plan.close()
    return {
        "value": value, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
        "timing": {"repeats": len(reps), "median_ms": ms, "min_ms": reps[order[0]][0], "max_ms": reps[order[-1]][0],
                   "rule": "the steps-step region repeated until >= 50 ms and >= 5 repeats; value from the median"},
        "config": {"workload": workload, "lattice": desc.replace("fp64", "fp32 state (fp64 sums)") if f32 else desc,
                   "nodes": n_updates, "max_degree": maxdeg, "effective_dt": dt_eff, "kick": list(map(str, kick)),
                   "l2": f"state {2 * slots * (4 if f32 else 8) / 1e9:.2f} GB >> 126 MB L2 (no flush needed)",
                   "tiling": tiling, "plan_build_s": t_build},
        "roofline": dict({"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                          "traffic": traffic, "peak_source": peak_src,
                          "algorithmic_bytes_per_launch": alg_bytes / max(kernel_launches, 1), "kernel": kernel,
                          "bytes_per_update": alg_bytes / (n_updates * steps),
                          "kernel_ms_per_launch": kms / max(kernel_launches, 1),
                          "kernel_share_of_step": kms / ms, "frac_of_8TBs_nominal": achieved / 8000.0}, **extra),
        "e2e": e2e, "gpu_launches": launches,
        # sum_i u_t(i) = 1 + t for unit masses (zero column sums of the Laplacian): a whole-lattice checksum
        "checksum_sum_u": checksum,
        # (tagged nodes with prune_edges have no edges and stay at rest, so the law also holds for cfg4; tagged nodes
        # that keep their edges conserve sum_i m_i u_i instead)
        "checksum_expected": float(steps_done + 1) if (spec.defect != "singularity" or
                                                       (spec.prune_edges and spec.potential == 0.0)) else None,
    }


def gpu_single(args, dist=None, emit=True):
    """One lattice per GPU.  N = 1: the bench line.  N > 1 (``dist`` given): N independent replicas of the same
    sheet, one per rank -- a 2-D sheet does not shard ("replicas only"), so there is no data-path
    collective; the timed region is bracketed by barriers and the slowest rank's device time counts."""
    import torch

    rank = int(os.environ.get("RANK", "0")) if dist else 0
    world = int(os.environ.get("WORLD_SIZE", "1")) if dist else 1
    local = int(os.environ.get("LOCAL_RANK", str(rank))) if dist else 0
    torch.cuda.init()
    torch.cuda.set_device(local)
    workload = args.workload or "cfg2-2d-8192"
    if workload not in WORKLOADS:
        raise SystemExit(f"unknown workload {workload}")
    clk = ClockSampler(local).start()
    res = bench_lattice(workload, args, local=local, dist=dist, size=args.size, layers=args.layers, clk=clk)
    clk.__exit__()
    res["config"]["parallelism"] = ("single GPU" if world == 1 else
                                    f"{world} independent replicas of the sheet, one per GPU, no data-path collective")
    cpu = None
    if not args.no_cpu and world == 1:
        cpu = cpu_run(workload, steps=100 if workload.startswith("cfg2") else 20, warmup=2)  # ~10 s of CPU work
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": res["config"],
        "roofline": res["roofline"],
        "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": clk.summary(), "timing": res["timing"],
        "checksum_sum_u": res["checksum_sum_u"], "checksum_expected": res["checksum_expected"],
    }
    if world == 1 and args.workload is None and not args.no_extras:
        line.update(default_extras(args, local))
    if emit and rank == 0:
        print(json.dumps(line))
    return line


def default_extras(args, local):
    """What the default N=1 line carries besides cfg2: the N=1 point of the 3-D weak-scaling series (the cfg3 slab
    through the multi-GPU slab code, no neighbours), and short lines of the other BASELINE configs."""
    import argparse as _ap

    from tetrakis_sim_b200.slab import bench_slabs

    out = {}
    a3 = _ap.Namespace(**vars(args))
    a3.workload, a3.size, a3.layers = "cfg3-slab-1024x64", None, None
    slab = bench_slabs(a3, emit=False)
    out["sharded_3d"] = {k: slab[k] for k in ("value", "unit", "ms_per_step", "config", "roofline", "e2e", "gpu_launches",
                                              "timing", "checksum_sum_u", "checksum_expected")}
    out["sharded_3d"]["note"] = ("N=1 point of the weak-scaling series of bench.py --gpus N: the same SlabRunner code, one "
                                 "step per launch, no neighbours.  The K-step passes of the same lattice on one GPU "
                                 "are under other_configs.")
    other = {}
    a = _ap.Namespace(**vars(args))
    a.dtype = "f64"
    for wl, st in (("cfg4-3d-512", 64), ("cfg3-slab-1024x64", 64)):
        r = bench_lattice(wl, a, local=local, steps=st, warmup=16, want_e2e=True)
        other[wl] = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "timing", "roofline", "e2e",
                                       "gpu_launches", "checksum_sum_u", "checksum_expected")}
        other[wl]["config"] = {k: r["config"][k] for k in ("workload", "lattice", "nodes", "tiling")}
        r1 = bench_lattice(wl, a, local=local, steps=20, warmup=5, fuse=False, want_e2e=False)
        other[wl]["one_step_per_launch"] = {"value": r1["value"], "ms_per_step": r1["ms_per_step"],
                                            "roofline_frac": r1["roofline"]["frac"]}
    other["cfg5-ensemble"] = ensemble_short(local)
    other["cfg1-batch"] = cfg1_short(local, cpu=not args.no_cpu)
    if not args.no_cpu:
        other["cpu_reference_tables"] = cpu_reference_tables(local)
    out["other_configs"] = other
    return out


def ensemble_short(local):
    """BASELINE configs[4] on this GPU: 65 536 independent 13x13x7 simulations in one launch (3 sweeps, the last 2 timed)."""
    import tetrakis_sim_b200 as tk

    specs, damp, dt = tk.sweep_grid(13, 7, np.linspace(0.5, 4.0, 64), np.linspace(0.0, 0.05, 32),
                                    np.linspace(0.02, 0.2, 32))
    kms, walls, res = [], [], None
    for it in range(3):
        t0 = time.perf_counter()
        res, _ = tk.run_ensemble_sharded(specs, 50, c=1.0, dt=dt, damping=damp, dist=None, device=local, gather=False)
        walls.append(time.perf_counter() - t0)
        kms.append(res.elapsed_ms)
    k, wl = float(np.mean(kms[1:])), float(np.mean(walls[1:]))
    return {"value": res.node_updates / (k * 1e-3) / 1e9, "unit": UNIT, "ms_per_sweep": k,
            "e2e": {"value": res.node_updates / wl / 1e9, "seconds": wl,
                    "includes": "mask/parameter upload, kernel, series + spectra + dominant bins D2H"},
            "config": {"workload": "cfg5-ensemble", "simulations": 65536, "lattice": "3-D tetrakis 13x13x7, 50 steps each",
                       "note": "state resident in shared memory: the HBM roofline does not apply"},
            "gpu_launches": 1}


def ensemble_sharded_short(dist, local):
    """BASELINE configs[4] over the ranks of this launch: the 65 536 simulations block-distributed, no data-path
    communication; 3 sweeps, the last 2 timed; device time = the slowest rank's kernel (strong scaling: the sweep
    is fixed)."""
    import torch

    import tetrakis_sim_b200 as tk

    specs, damp, dt = tk.sweep_grid(13, 7, np.linspace(0.5, 4.0, 64), np.linspace(0.0, 0.05, 32),
                                    np.linspace(0.02, 0.2, 32))
    kms, walls, res = [], [], None
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, _ = tk.run_ensemble_sharded(specs, 50, c=1.0, dt=dt, damping=damp, dist=dist, device=local, gather=False)
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
        kms.append(res.elapsed_ms)
    t = torch.tensor([float(res.node_updates), float(np.mean(kms[1:])), float(np.mean(walls[1:]))],
                     dtype=torch.float64, device=f"cuda:{local}")
    upd = t[:1].clone()
    dist.all_reduce(upd)
    mx = t[1:].clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    updates, k, wl = float(upd.item()), float(mx[0].item()), float(mx[1].item())
    return {"value": updates / (k * 1e-3) / 1e9, "unit": UNIT, "ms_per_sweep": k, "scaling": "strong",
            "n_gpus": dist.get_world_size(),
            "e2e": {"value": updates / wl / 1e9, "seconds": wl,
                    "includes": "mask/parameter upload, kernel, series + spectra + dominant bins D2H (slowest rank)"},
            "config": {"workload": "cfg5-ensemble", "simulations": 65536, "lattice": "3-D tetrakis 13x13x7, 50 steps each",
                       "parallelism": f"simulations block-distributed over {dist.get_world_size()} GPUs, no communication"}}


def cfg1_short(local, cpu=True):
    """BASELINE configs[0] through the drop-in: run_wave_sim on the networkx graph (generic sliced-ELL kernel, results
    bit-identical to the reference) and on the LatticeSpec (structured kernel), wall clock of the call; beside it the
    literal CPU restatement on one core."""
    import warnings

    import tetrakis_sim_b200 as tk

    spec = tk.LatticeSpec(size=11, dim=3, layers=7, defect="blackhole", radius=2.5)
    kick = spec.kick_node()
    G = spec.to_networkx()
    out = {"config": {"workload": "cfg1-batch", "lattice": "3-D tetrakis 11x11x7, black hole r=2.5, 50 steps, dt 0.2",
                      "nodes": spec.num_nodes}}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for name, obj in (("run_wave_sim(G)", G), ("run_wave_sim(LatticeSpec)", spec)):
            best = 1e9
            for _ in range(5):
                t0 = time.perf_counter()
                h = tk.run_wave_sim(obj, steps=50, initial_data={kick: 1.0}, c=1.0, dt=0.2, device=local)
                best = min(best, time.perf_counter() - t0)
            out[name] = {"wall_s": best, "value": spec.num_nodes * 50 / best / 1e9, "unit": UNIT,
                         "device_ms": h.elapsed_ms}
    if cpu:
        out["cpu_literal_1_core"] = cpu_literal_cfg1()
    out["reference_python_networkx_s"] = 0.593  # SURVEY.md section 3a, measured on this image's host
    return out


def _literal_inputs(spec):
    """CSR + attributes of the reference-built graph of ``spec`` (oracle/lattice_oracle.py), the kick by the batch.py rule
    (bench_wave.py's centre node when there is no defect)."""
    from oracle import lattice_oracle as lo

    kw = {} if spec.defect == "none" else dict(radius=spec.radius)
    G, _, center = lo.build_case(spec.size, spec.dim, spec.layers, spec.defect, **kw)
    nodes, rowptr, colidx, inv_mass, potential = lo.graph_arrays(G)
    deg = np.diff(rowptr).astype(np.float64)
    kick = spec.bench_kick_node() if spec.defect == "none" else lo.pick_kick_node(G, center, spec.dim)
    u0 = np.zeros(len(nodes))
    u0[nodes.index(kick)] = 1.0
    return G, kick, rowptr, colidx, deg, inv_mass, potential, u0


def _cfg5_point(job):
    """One cfg5 grid point through the LITERAL pure-Python restatement of physics.py:106-129 (oracle.run_graph_python):
    what one `run_wave_sim` call of the reference computes, minus networkx's dict lookups.  Runs in a worker process."""
    radius, damping, dt = job
    import tetrakis_sim_b200 as tk
    from oracle import wave_oracle as wo

    spec = tk.LatticeSpec(size=13, dim=3, layers=7, defect="blackhole", radius=radius)
    _, _, rowptr, colidx, deg, inv_mass, potential, u0 = _literal_inputs(spec)
    dt_eff, _, _ = wo.cfl_clamp(int(deg.max()), 1.0, dt, 50, damping)
    t0 = time.perf_counter()
    wo.run_graph_python(rowptr, colidx, deg, inv_mass, potential, u0, 50, dt_eff ** 2, damping)
    return time.perf_counter() - t0, len(u0)


def cpu_reference_tables(local, python_budget_visits=3e7):
    """SURVEY.md 8(d)(i), as far as it can travel to the GPU box (the reference itself lives in /root/reference and does
    not): the reference's seven published `scripts/bench_wave.py` sizes (benchmarks/bench_2d_after.csv, bench_3d_after.csv)
    through (a) the drop-in run_wave_sim(G) on the networkx graph, wall clock, median of 3; (b) the literal per-edge
    restatement in C on one core; (c) the literal restatement in pure Python on one core where it finishes in seconds --
    and 2 x cores grid points of cfg5 (seed 0) through (c) on all host cores, one process each as
    examples/naked_singularity_sweep.py does, extrapolated to the 65 536-point grid."""
    import multiprocessing as mp
    import statistics
    import warnings
    from concurrent.futures import ProcessPoolExecutor

    import tetrakis_sim_b200 as tk
    from oracle import wave_oracle as wo

    rows = []
    for dim, layers, size, steps in ((2, 1, 11, 50), (2, 1, 21, 50), (2, 1, 31, 50), (2, 1, 41, 50),
                                     (3, 5, 7, 30), (3, 5, 11, 30), (3, 5, 15, 30)):
        spec = tk.LatticeSpec(size=size, dim=dim, layers=layers)
        G, kick, rowptr, colidx, deg, inv_mass, potential, u0 = _literal_inputs(spec)
        dt_eff, _, _ = wo.cfl_clamp(int(deg.max()), 1.0, 0.1, steps, 0.0)
        walls = []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            for _ in range(4):
                t0 = time.perf_counter()
                tk.run_wave_sim(G, steps=steps, initial_data={kick: 1.0}, c=1.0, dt=0.1, device=local)
                walls.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        wo.run_graph(rowptr, colidx, deg, inv_mass, potential, u0, steps, dt_eff ** 2, 0.0)
        t_c = time.perf_counter() - t0
        row = {"dim": dim, "layers": layers, "size": size, "steps": steps, "nodes": len(u0), "edges": int(rowptr[-1]) // 2,
               "dropin_wall_s_median": statistics.median(walls[1:]), "literal_c_1_core_s": t_c}
        if rowptr[-1] * steps <= python_budget_visits:
            t0 = time.perf_counter()
            wo.run_graph_python(rowptr, colidx, deg, inv_mass, potential, u0, steps, dt_eff ** 2, 0.0)
            row["literal_python_1_core_s"] = time.perf_counter() - t0
        rows.append(row)
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    radii, damps, dts = np.linspace(0.5, 4.0, 64), np.linspace(0.0, 0.05, 32), np.linspace(0.02, 0.2, 32)
    n_pts = min(64, 2 * cores)
    jobs = [(float(radii[rng.integers(64)]), float(damps[rng.integers(32)]), float(dts[rng.integers(32)]))
            for _ in range(n_pts)]
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(_cfg5_point, jobs))
    wall = time.perf_counter() - t0
    per_sim = statistics.mean(r[0] for r in res)
    return {"bench_wave_sizes": rows,
            "cfg5_points": {"points": n_pts, "seed": 0, "cores": cores, "wall_s": wall, "mean_s_per_simulation": per_sim,
                            "node_updates": int(sum(r[1] for r in res)) * 50,
                            "extrapolated_hours_for_65536": 65536 * per_sim / cores / 3600.0,
                            "kind": "port (oracle.run_graph_python: the literal loop of physics.py:106-129 in pure Python, "
                                    "one process per point)"},
            "note": "the unmodified reference measured 0.0855 / 0.358 / 1.206 / 3.417 s (2-D) and 0.0697 / 0.2495 / 0.6938 s "
                    "(3-D) for these sizes and 1.336 s per cfg5 simulation on this image's host (BASELINE.md section 2); it "
                    "cannot travel to the GPU box"}


def _csr_tetrakis(S, L):
    """Adjacency-ordered CSR of a defect-free 3-D tetrakis lattice built with NumPy (no networkx): node
    (z, r, c, q) -> ((z*S + r)*S + c)*4 + q; neighbours = cell mates, row mates, column mates, z-1, z+1."""
    n = L * S * S * 4
    idx = np.arange(n, dtype=np.int64)
    q, cell = idx & 3, idx >> 2
    c, r, z = cell % S, (cell // S) % S, cell // (S * S)
    cols = []
    for dq in (1, 2, 3):
        cols.append((cell << 2) + ((q + dq) & 3))
    for d in range(1, S):
        cols.append(((((z * S + r) * S + (c + d) % S)) << 2) + q)
    for d in range(1, S):
        cols.append(((((z * S + (r + d) % S) * S + c)) << 2) + q)
    nb = np.stack(cols, axis=1)
    width = nb.shape[1]
    zm = np.where(z > 0, idx - S * S * 4, -1)
    zp = np.where(z < L - 1, idx + S * S * 4, -1)
    nb = np.concatenate([nb, zm[:, None], zp[:, None]], axis=1)
    keep = nb >= 0
    deg = keep.sum(1)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    return rowptr, nb[keep].astype(np.int32), deg.astype(np.float64), width + 2


def _csr_grid(m):
    """7-point grid graph on m^3 nodes (non-periodic)."""
    n = m ** 3
    idx = np.arange(n, dtype=np.int64)
    x, y, z = idx % m, (idx // m) % m, idx // (m * m)
    nb = np.stack([np.where(x > 0, idx - 1, -1), np.where(x < m - 1, idx + 1, -1),
                   np.where(y > 0, idx - m, -1), np.where(y < m - 1, idx + m, -1),
                   np.where(z > 0, idx - m * m, -1), np.where(z < m - 1, idx + m * m, -1)], axis=1)
    keep = nb >= 0
    deg = keep.sum(1)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    return rowptr, nb[keep].astype(np.int32), deg.astype(np.float64), 6


def gpu_ell(args):
    """--workload ell-tetrakis-64x16 | ell-grid-256: the generic sliced-ELL graph kernel (what any networkx
    graph runs on) on a synthetic CSR.  Algorithmic bytes per update = 44 + 4*w: u, u_prev, u_next,
    degree, inv_mass, potential + one 4-byte index per padded neighbour slot; the gathered u is cache traffic."""
    import torch

    import tetrakis_sim_b200 as tk

    torch.cuda.init()
    if args.workload == "ell-grid-256":
        m = args.size or 256
        rowptr, colidx, deg, w = _csr_grid(m)
        desc = f"7-point grid graph {m}^3, sliced ELL width 6"
    else:
        S, L = args.size or 64, args.layers or 16
        rowptr, colidx, deg, w = _csr_tetrakis(S, L)
        desc = f"3-D tetrakis {S}x{S}x{L} as an explicit graph (degree {int(deg.max())}), sliced ELL"
    n = len(deg)
    t0 = time.perf_counter()
    plan = tk.Plan.graph(rowptr, colidx, deg, np.ones(n), np.zeros(n), device=0)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    coeff = 1.0 / float(deg.max())
    kidx = [n // 2]
    clk = ClockSampler(0).start()
    plan.set_state_sparse(kidx, [1.0])
    plan.run(args.warmup, coeff, 0.0, probes=kidx)
    torch.cuda.synchronize()
    l0 = plan.launches
    with clk:
        out = plan.run(args.steps, coeff, 0.0, probes=kidx, timed=True, kernel_timing=True)
        torch.cuda.synchronize()
    launches = plan.launches - l0
    ms, kms = out["elapsed_ms"], out["step_kernel_ms"]
    bpu = 44.0 + 4.0 * w
    peak, peak_src = measured_peak_gbs()
    achieved = n * bpu * args.steps / (kms * 1e-3) / 1e9
    final_t, final = pinned_empty(n)
    from tetrakis_sim_b200 import _lib

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan.set_state_sparse(kidx, [1.0])
    o2 = plan.run(args.steps, coeff, 0.0, probes=kidx)
    _lib.check(_lib.lib().tk_plan_get_state(plan._h, _lib._p(final), None))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    checksum = float(final.sum())
    plan.close()
    line = {
        "metric": METRIC, "value": n * args.steps / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "graph": desc, "nodes": n, "edges_directed": int(rowptr[-1]),
                   "plan_build_s": t_build, "l2": f"index array {4 * w * n / 1e9:.2f} GB"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "kernel": "graph_step_kernel",
                     "bytes_per_update": bpu, "algorithmic_bytes_per_launch": n * bpu,
                     "kernel_ms_per_launch": kms / args.steps, "kernel_share_of_step": kms / ms},
        "cpu_baseline": None,
        "e2e": {"value": n * args.steps / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": 16.0 / args.steps,
                "d2h_bytes_per_step": (final.nbytes + o2["probes"].nbytes) / args.steps, "seconds": e2e_s},
        "gpu_launches": launches, "clocks": clk.summary(),
        "checksum_sum_u": checksum, "checksum_expected": float(args.steps + 1),
    }
    print(json.dumps(line))
    return 0


def gpu_ensemble(args):
    """--workload cfg5-ensemble: BASELINE configs[4], 65 536 independent 13x13x7 simulations over a
    radius x damping x dt grid, 50 steps each; simulations are block-distributed over the ranks, no
    communication on the data path.  One "step" = one complete sweep."""
    import torch

    import tetrakis_sim_b200 as tk

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    specs, damp, dt = tk.sweep_grid(13, 7, np.linspace(0.5, 4.0, 64), np.linspace(0.0, 0.05, 32),
                                    np.linspace(0.02, 0.2, 32))
    sim_steps = 50
    kms, walls, res = [], [], None
    clk = ClockSampler(local).start()
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            clk.mark_begin()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, (lo, hi) = tk.run_ensemble_sharded(specs, sim_steps, c=1.0, dt=dt, damping=damp, dist=dist,
                                                device=local, gather=False)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if it >= args.warmup:
            kms.append(res.elapsed_ms)
            walls.append(wall)
    clk.__exit__()
    tot = torch.tensor([float(res.node_updates), max(kms), max(walls), float(np.mean(kms)), float(np.mean(walls))],
                       dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        upd = tot[:1].clone()
        dist.all_reduce(upd)
        mx = tot[1:].clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        tot = torch.cat([upd, mx])
    updates, _, _, k_mean, w_mean = tot.tolist()
    if rank == 0:
        line = {
            "metric": METRIC, "value": updates / (k_mean * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": k_mean, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg5-ensemble",
                       "lattice": "65 536 simulations of 3-D tetrakis 13x13x7 (4 732 nodes), black hole radius in "
                                  "64 values [0.5, 4] x damping in 32 values [0, 0.05] x dt in 32 values [0.02, 0.2], "
                                  "50 steps each, probe series + |DFT| + dominant bin per simulation",
                       "node_updates_per_sweep": updates, "parallelism": f"simulations block-distributed over {world} GPU(s)",
                       "l2": "state resident in shared memory (no HBM state traffic): the HBM roofline does not apply"},
            "roofline": {"bound": "hbm", "achieved": updates * BYTES_PER_UPDATE / (k_mean * 1e-3) / 1e9,
                         "peak": measured_peak_gbs()[0], "unit": "GB/s",
                         "frac": updates * BYTES_PER_UPDATE / (k_mean * 1e-3) / 1e9 / measured_peak_gbs()[0],
                         "traffic": None, "kernel": "ensemble_kernel",
                         "note": "nominal 24 B/update for comparability only; the state never leaves the SM"},
            "cpu_baseline": None,
            "e2e": {"value": updates / w_mean / 1e9, "unit": UNIT, "seconds": w_mean,
                    "h2d_bytes_per_step": 65536 * 24.0 + 64 * 1183, "d2h_bytes_per_step": 65536 * (2 * 51 + 1) * 8.0,
                    "includes": "mask/parameter upload, kernel, series + spectra + dominant bins D2H"},
            "gpu_launches": args.steps, "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true",
                    help="N=1 default: only the cfg2 line, without sharded_3d / other_configs")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="f32 = optional fp32 state mode (12 B/update; NOT the parity mode)")
    args = ap.parse_args()
    if os.environ.get("TK_WATCHDOG"):  # diagnostics: dump every thread's stack and exit if the run stalls
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["TK_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    # stdout carries exactly one JSON line: NCCL's own banner / debug output (it goes to stdout when NCCL_DEBUG is
    # set in the environment) is sent to stderr instead
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.impl == "reference":
        return reference_arm(args, rank)
    if args.workload == "cfg5-ensemble":
        return gpu_ensemble(args)
    if args.workload in ("ell-tetrakis-64x16", "ell-grid-256"):
        return gpu_ell(args)
    if args.workload == "cfg2-sharded":  # ONE sheet over N GPUs by row bands, one all-reduce per pass
        from tetrakis_sim_b200.sheet import bench_sheet

        return bench_sheet(args)
    if args.gpus == 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        gpu_single(args)
        return 0
    from tetrakis_sim_b200.slab import bench_slabs  # multi-GPU: layer slabs + halo exchange

    if args.workload == "cfg2-2d-8192":  # N independent replicas of the 8192^2 sheet, no collective
        import torch
        import torch.distributed as dist

        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        gpu_single(args, dist=dist)
        dist.destroy_process_group()
        return 0
    if args.workload is not None:
        return bench_slabs(args)
    # N > 1, default: the metric's configuration grown with N -- ONE defect-free sheet held as N row bands, every
    # GPU holding as many nodes as the single-GPU 8192^2 sheet, one all-reduce of the column sums per pass (weak
    # scaling) -- followed in the same launch by the layer-slab run of the 3-D lattice with the black hole
    from tetrakis_sim_b200.sheet import bench_sheet

    import torch.distributed as dist

    line = bench_sheet(args, emit=False, keep_pg=True)
    slab = bench_slabs(args, emit=False, keep_pg=True)
    ens = ensemble_sharded_short(dist, int(os.environ.get("LOCAL_RANK", str(rank))))
    dist.destroy_process_group()
    if rank == 0:
        line["sharded_3d"] = {k: slab[k] for k in ("value", "unit", "ms_per_step", "config", "roofline", "e2e",
                                                   "gpu_launches", "single_gpu_same_workload", "timing", "parity",
                                                   "checksum_sum_u", "checksum_expected") if k in slab}
        line["ensemble"] = ens
        print(json.dumps(line), flush=True)
    return 0


def _main_with_clean_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version banner to
    stdout on communicator creation): run with file descriptor 1 pointing at stderr and give the JSON line, which
    the arms print through ``print``, the real stdout at the end."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    captured = []
    import builtins

    orig_print = builtins.print

    def json_print(*a, **k):
        if len(a) == 1 and isinstance(a[0], str) and a[0].startswith("{") and k.get("file") is None:
            captured.append(a[0])
        else:
            orig_print(*a, **k)

    builtins.print = json_print
    try:
        rc = main()
    finally:
        builtins.print = orig_print
        sys.stdout.flush()
        os.dup2(real, 1)
        os.close(real)
    for line in captured:
        print(line, flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(_main_with_clean_stdout())
