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
  cfg4-3d-512       BASELINE configs[3]: 3-D 512^3 singularity, mass 2000, pruned edges
  cfg5-ensemble     BASELINE configs[4]: 65 536 independent 13x13x7 simulations
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


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return float(json.load(f)[workload]["bytes"])
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
def cpu_run(workload: str, steps: int, warmup: int, budget_s: float):
    from oracle import wave_oracle as wo

    cores = os.cpu_count() or 1
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    three_d = workload.startswith("cfg3")

    def shape(size):
        return (size, 16) if three_d else (size, 1)

    def one(size, nsteps):
        S, L = shape(size)
        u0 = np.zeros(L * S * S * 4)
        u0[((L // 2 * S + S // 2) * S + S // 2) * 4] = 1.0
        maxdeg = 3 + 2 * (S - 1) + (2 if L > 1 else 0)
        coeff = 1.0 / maxdeg  # dt at the CFL limit, c = 1
        t0 = time.perf_counter()
        wo.run_structured(S, L, None, None, None, [], u0, nsteps, coeff, 0.0)
        return time.perf_counter() - t0, L * S * S * 4

    t, n = one(256, 4)  # calibrate seconds per node-update
    per_nu = max(t / (4 * n), 1e-12)
    size = 256
    for cand in ((128, 256, 512, 1024) if three_d else (512, 1024, 2048, 4096)):
        S, L = shape(cand)
        if (steps + warmup) * L * cand * cand * 4 * per_nu <= budget_s:
            size = cand
    if warmup:
        one(size, warmup)
    t, n = one(size, steps)
    S, L = shape(size)
    sample = (f"{'3-D' if three_d else '2-D'} tetrakis {S}x{S}" + (f"x{L}" if three_d else "") +
              f" no defect, {steps} steps, oracle/wave_oracle.c clique-sum restatement, OpenMP {cores} threads")
    return {"value": n * steps / t / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "seconds": t, "ms_per_step": 1e3 * t / max(steps, 1)}


def reference_arm(args, rank):
    if rank != 0:
        return 0
    workload = args.workload or ("cfg2-2d-8192" if args.gpus == 1 else "cfg2-sharded")
    res = cpu_run(workload, args.steps, args.warmup, budget_s=150.0)
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


def gpu_single(args, dist=None, emit=True):
    """One lattice per GPU.  N = 1: the bench line.  N > 1 (``dist`` given): N independent replicas of the same
    sheet, one per rank -- a 2-D sheet does not shard ("replicas only"), so there is no data-path
    collective; the timed region is bracketed by barriers and the slowest rank's device time counts."""
    import torch

    import tetrakis_sim_b200 as tk

    rank = int(os.environ.get("RANK", "0")) if dist else 0
    world = int(os.environ.get("WORLD_SIZE", "1")) if dist else 1
    local = int(os.environ.get("LOCAL_RANK", str(rank))) if dist else 0
    torch.cuda.init()
    torch.cuda.set_device(local)
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

    workload = args.workload or "cfg2-2d-8192"
    if workload == "cfg2-2d-8192":
        spec = tk.LatticeSpec(size=args.size or 8192, dim=2)
        dt_req, desc = 0.1, "2-D tetrakis {0}x{0}, no defect, fp64, kick (S/2,S/2,'A'), dt 0.1 -> CFL-clamped"
    elif workload == "cfg3-slab-1024x64":
        spec = tk.LatticeSpec(size=args.size or 1024, dim=3, layers=args.layers or 64, defect="blackhole",
                              radius=64.0)
        dt_req, desc = 0.1, "3-D tetrakis {0}x{0}x" + str(spec.layers) + ", black hole r=64 at the centre, fp64"
    elif workload == "cfg4-3d-512":
        spec = tk.LatticeSpec(size=args.size or 512, dim=3, layers=args.layers or 512, defect="singularity",
                              radius=2.5, mass=2000.0, prune_edges=True)
        dt_req, desc = 0.1, "3-D tetrakis {0}^3 singularity mass 2000 + prune_edges, fp64"
    else:
        raise SystemExit(f"unknown workload {workload}")
    desc = desc.format(spec.size)
    kick = spec.kick_node() if spec.defect != "none" else spec.bench_kick_node()
    f32 = args.dtype == "f32"
    bpu = BYTES_PER_UPDATE / 2 if f32 else BYTES_PER_UPDATE
    t_build = time.perf_counter()
    plan = tk.compile_lattice(spec, device=local, dtype="float32" if f32 else "float64")
    maxdeg = plan.max_degree
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build  # graph-to-device compile (excluded from every rate below)
    limit = 1.0 / np.sqrt(maxdeg)
    dt_eff = min(dt_req, limit)
    coeff = dt_eff ** 2
    n_updates = spec.num_nodes  # surviving nodes
    kidx = [spec.index(kick)]

    clk = ClockSampler(local).start()
    plan.set_state_sparse(kidx, [1.0])
    plan.run(args.warmup, coeff, 0.0, probes=kidx)
    fence()
    l0 = plan.launches
    with clk:
        fence()
        out = plan.run(args.steps, coeff, 0.0, probes=kidx, timed=True, kernel_timing=True)
        fence()
    launches = plan.launches - l0
    ms = max_over_ranks(out["elapsed_ms"])  # device time (CUDA events on the plan's stream), slowest rank
    kms = out["step_kernel_ms"]
    value = world * n_updates * args.steps / (ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()
    # Defect-free 2-D sheets advance K steps per pass (kernels_fused.cuh): a pass reads u_t, u_{t-1} and
    # writes u_{t+K}, u_{t+K-1} = 32 B per node per pass; the remaining steps (steps mod K) move 24 B each.
    fused = plan.fused
    fk = fused["steps_per_pass"]
    npass = -(-fused["fused_steps"] // fk) if fk else 0
    nsingle = args.steps - fused["fused_steps"]
    alg_bytes = n_updates * (npass * (4 * bpu / 3) + nsingle * bpu)
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    kernel_launches = npass + nsingle
    checksum = float(plan.get_state().sum()) if spec.num_slots <= 600_000_000 else None

    # e2e: the whole run through the C ABI with host buffers (kicks H2D; probe series and the final
    # state D2H into pinned host memory), device-synchronised wall clock
    hold, final = pinned_empty(plan.num_nodes)
    from tetrakis_sim_b200 import _lib
    fence()
    t0 = time.perf_counter()
    plan.set_state_sparse(kidx, [1.0])
    o2 = plan.run(args.steps, coeff, 0.0, probes=kidx)
    _lib.check(_lib.lib().tk_plan_get_state(plan._h, _lib._p(final), None))
    fence()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * n_updates * args.steps / e2e_s / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": world * 16.0 / args.steps,
           "d2h_bytes_per_step": world * (final.nbytes + o2["probes"].nbytes) / args.steps,
           "seconds": e2e_s, "includes": "set_state_sparse + run + probe series D2H + final state D2H (pinned)"}
    tiling = plan.tiling
    if fk:
        tiling = {"steps_per_pass": fk, "cells_per_lane": 1, "segs_per_cta": tiling["segs_per_cta"],
                  "rows_per_cta": fused["rows_per_cta"], "ctas_per_sm": fused["ctas_per_sm"],
                  "single_steps": nsingle,
                  "temporal_blocking": (f"{npass} passes of <= {fk} steps: every node update is computed (cell-local, in "
                                        "registers), the intermediate states are not stored; the row/column sums of the "
                                        "intermediate steps follow from closed recurrences because the sheet has no "
                                        "defect.  TK_LAT_FUSE=0 runs one step per launch (24 B per update).")}
    plan.close()

    cpu = None
    if not args.no_cpu and world == 1:
        cpu = cpu_run(workload, steps=100, warmup=2, budget_s=40.0)  # ~10 s of CPU work
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload, "lattice": desc.replace("fp64", "fp32 state (fp64 sums)") if f32 else desc, "nodes": n_updates, "max_degree": maxdeg,
                   "effective_dt": dt_eff, "kick": list(map(str, kick)),
                   "l2": f"state {2 * spec.num_slots * 8 / 1e9:.2f} GB >> 126 MB L2 (no flush needed)",
                   "tiling": tiling, "plan_build_s": t_build,
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} independent replicas of the sheet, one per GPU, no data-path collective (a 2-D sheet "
                   "does not shard; the sharded 3-D slab run of the same launch is reported under sharded_3d)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": None if f32 else ncu_traffic(workload + ("-fused" if fk else "")),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes / kernel_launches,
                     "kernel": (f"lat_fused2d_kernel{'<float>' if f32 else ''} (K={fk})" if fk
                                else ("lat_step_f32_kernel" if f32 else "lat_step_kernel")),
                     "bytes_per_update": alg_bytes / (n_updates * args.steps),
                     "kernel_ms_per_launch": kms / kernel_launches,
                     "kernel_share_of_step": kms / ms, "frac_of_8TBs_nominal": achieved / 8000.0},
        "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary(),
        # sum_i u_t(i) = 1 + t for unit masses (zero column sums of the Laplacian): a whole-lattice checksum
        "checksum_sum_u": checksum,
        "checksum_expected": float(args.warmup + args.steps + 1) if spec.defect != "singularity" else None,
    }
    if emit and rank == 0:
        print(json.dumps(line))
    return line


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

    line = bench_sheet(args, emit=False, keep_pg=True)
    slab = bench_slabs(args, emit=False)
    if rank == 0:
        line["sharded_3d"] = {k: slab[k] for k in ("value", "unit", "ms_per_step", "config", "roofline", "e2e",
                                                   "gpu_launches", "single_gpu_same_workload",
                                                   "checksum_sum_u", "checksum_expected")}
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
