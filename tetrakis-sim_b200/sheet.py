"""ONE defect-free 2-D sheet sharded over GPUs by rows (one process per GPU).

A tetrakis sheet couples every node to its whole row and column (reference ``tetrakis_sim/lattice.py:58-66``),
so a row band needs the column sums of the complete sheet: stepping one step at a time that is an all-reduce
of 4*size doubles per step.  With the K-step passes of ``kernels_fused.cuh`` the column
sums of the K-1 intermediate steps follow from closed recurrences, so the exchange shrinks to ONE all-reduce of
(8*size + 8) doubles -- column sums and quadrant totals of the two time levels -- per K steps:

    pass_local(K)  : response tables from the global sums, K steps on the band in registers, local sums
    all_reduce     : band column sums + totals -> global   (torch.distributed, NCCL over NVLink; SUM)
    pass_commit    : install the global sums for the next pass

Row sums never leave the rank.  ``torch.distributed`` is plumbing only; all arithmetic is in the CUDA plan
(``tk_sheet_*`` in include/tetrakis_b200.h).  The orchestration is backend-agnostic (anything with
``pass_local / pass_commit / exchange``), which is how tests/test_sheet_gloo.py drives it on CPU with
world_size 2 over gloo.
"""

from __future__ import annotations

import json
import os
import time

import numpy as np


def band_rows(size: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [begin, end) of rank ``rank``: contiguous, sizes differing by at most one."""
    base, extra = divmod(size, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def pass_lengths(steps: int, kmax: int) -> list[int]:
    """Passes of (nearly) equal length covering ``steps`` (every pass moves the whole state once)."""
    if steps <= 0:
        return []
    npass = -(-steps // max(kmax, 1))
    k = -(-steps // npass)
    if k > 8 and -(-k // 8) * 8 <= kmax:  # the pass kernel's step loop is unrolled by 8 (csrc/host_fused.inc:fuse_choice)
        k = -(-k // 8) * 8
    k = min(k, steps)
    out = [k] * (steps // k)
    if steps % k:
        out.append(steps % k)
    return out


class CudaSheetBackend:
    """The CUDA plan of one row band + the exchange buffer as a torch tensor."""

    def __init__(self, size: int, row_begin: int, row_end: int, device: int):
        import torch

        from ._lib import Plan

        self.torch = torch
        self.device = device
        self.size, self.row_begin, self.row_end = size, row_begin, row_end
        self.plan = Plan.sheet_band(size, row_begin, row_end, device=device)
        self.plan.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.exchange = torch.zeros(self.plan.exchange_doubles, dtype=torch.float64, device=f"cuda:{device}")
        self.sums_valid = False      # the plan holds the global sums of the current state (after a commit)
        self.probes_pending = False  # set_probes was called: the next run must record row 0

    def pass_local(self, k, coeff, damping):
        self.plan.pass_local(k, coeff, damping, self.exchange.data_ptr())
        if k == 0:
            self.probes_pending = False

    def pass_commit(self):
        self.plan.pass_commit(self.exchange.data_ptr())
        self.sums_valid = True

    def set_state_sparse(self, index, value):
        self.plan.set_state_sparse(index, value)
        self.sums_valid = False

    def set_probes(self, probes, rows):
        self.plan.set_probes(probes, rows)
        self.probes_pending = True

    def read_probes(self, rows):
        return self.plan.read_probes(rows)

    def get_state(self):
        return self.plan.get_state()

    def close(self):
        self.plan.close()


class SheetRunner:
    """Steps one sheet held as row bands by ``world`` ranks.  ``dist`` = torch.distributed (or None for one
    rank); ``backend`` = CudaSheetBackend (or the NumPy stand-in of the CPU tests)."""

    def __init__(self, size: int, backend, rank: int, world: int, dist):
        self.size, self.backend, self.rank, self.world, self.dist = size, backend, rank, world, dist
        self.row_begin, self.row_end = band_rows(size, rank, world)
        self.started = False

    # global node index ((r*size + c)*4 + q) <-> band-local index
    def owns(self, gidx: int) -> bool:
        r = gidx // (4 * self.size)
        return self.row_begin <= r < self.row_end

    def local_index(self, gidx: int) -> int:
        return gidx - self.row_begin * self.size * 4

    def _exchange(self):
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(self.backend.exchange)  # SUM
        self.backend.pass_commit()

    def begin(self, coeff: float, damping: float):
        """Sums of the initial state (both time levels) -> global; records the probes' row 0."""
        self.coeff, self.damping = coeff, damping
        b = self.backend
        # a run that continues the previous one (same state, nothing to record at row 0) already has the sums
        if not (getattr(b, "sums_valid", False) and not getattr(b, "probes_pending", True)):
            b.pass_local(0, coeff, damping)
            self._exchange()
        self.started = True

    def advance(self, k: int):
        assert self.started
        self.backend.pass_local(k, self.coeff, self.damping)
        self._exchange()

    def run(self, steps: int, coeff: float, damping: float, kmax: int | None = None):
        if kmax is None:
            kmax = int(os.environ.get("TK_LAT_FUSE", "0") or 0) or (48 if damping == 0.0 else 24)
        self.begin(coeff, damping)
        for k in pass_lengths(steps, kmax):
            self.advance(k)


def run_sheet_sharded(size: int, steps: int, kicks: dict, c: float = 1.0, dt: float = 0.2, damping: float = 0.0,
                      probes=None, dist=None, rank: int = 0, world: int = 1, device: int | None = None,
                      kmax: int | None = None):
    """``run_wave_sim`` for one defect-free sheet held by ``world`` GPUs.  ``kicks`` / ``probes`` use global node
    indices ((r*size + c)*4 + q).  Returns dict(effective_dt, probes (steps+1, n_probes) -- each column is
    filled on the rank owning that node, zeros elsewhere --, runner, backend); the caller closes the backend."""
    import warnings

    maxdeg = 2 * size + 1  # lattice.py:30-73 without defects
    limit = 1.0 / np.sqrt(maxdeg)
    effective_dt = dt
    if c > 0 and c * dt > limit:  # physics.py:66-80
        effective_dt = limit / c
        warnings.warn(
            "Requested time-step is unstable for the current lattice "
            f"(c*dt={c * dt:.3f} > {limit:.3f}). Clamping dt to {effective_dt:.3f}",
            RuntimeWarning, stacklevel=2,
        )  # fmt: skip
    backend = CudaSheetBackend(size, *band_rows(size, rank, world), device=rank if device is None else device)
    runner = SheetRunner(size, backend, rank, world, dist)
    mine = {g: v for g, v in kicks.items() if runner.owns(g)}
    backend.set_state_sparse([runner.local_index(g) for g in mine], list(mine.values()))
    probes = list(probes or [])
    own = [i for i, g in enumerate(probes) if runner.owns(g)]
    if own:
        backend.set_probes([runner.local_index(probes[i]) for i in own], steps + 1)
    runner.run(steps, (c * effective_dt) ** 2, damping, kmax)
    series = np.zeros((steps + 1, len(probes)))
    if own:
        series[:, own] = backend.read_probes(steps + 1)
    return {"effective_dt": effective_dt, "probes": series, "runner": runner, "backend": backend}


def sheet_parity_check(dist, rank: int, world: int, local: int, steps: int = 60, kmax: int = 16):
    """In-run parity of the row-band path: a reduced sheet (512 x 512) stepped by the SAME SheetRunner and NCCL
    all-reduce as the timed run, against ONE plan holding the whole sheet on rank 0 (final state of every node and
    the kick node's time series).  Returns the dict for the bench line (rank 0) or None."""
    import torch

    size = 512
    dev = torch.device(f"cuda:{local}")
    kick = ((size // 2) * size + size // 2) * 4
    far = ((size - 3) * size + 7) * 4 + 2  # a node of the last band
    res = run_sheet_sharded(size, steps, {kick: 1.0}, c=1.0, dt=1.0 / np.sqrt(2 * size + 1), probes=[kick, far],
                            dist=dist, rank=rank, world=world, device=local, kmax=kmax)
    torch.cuda.synchronize()
    mine = torch.from_numpy(res["backend"].get_state()).to(dev)
    sizes = [(band_rows(size, r, world)[1] - band_rows(size, r, world)[0]) * size * 4 for r in range(world)]
    parts = [torch.empty(n, dtype=torch.float64, device=dev) for n in sizes]
    dist.all_gather(parts, mine)
    ser = torch.from_numpy(res["probes"]).to(dev)
    dist.all_reduce(ser)  # each column is filled on its owner, zeros elsewhere
    res["backend"].close()
    out = None
    if rank == 0:
        from .compiler import compile_lattice
        from .lattice import LatticeSpec

        plan = compile_lattice(LatticeSpec(size=size, dim=2), device=local)
        try:
            plan.set_state_sparse([kick], [1.0])
            o = plan.run(steps, res["effective_dt"] ** 2, 0.0, probes=[kick, far], fuse=False)
            ref = plan.get_state()
        finally:
            plan.close()
        whole = torch.cat(parts).cpu().numpy()
        den = float(np.abs(ref).max())
        out = {"lattice": f"ONE 2-D tetrakis sheet {size}x{size} as {world} row bands, {steps} steps in passes of "
                          f"<= {kmax}, NCCL all-reduce per pass",
               "reference": "ONE plan holding the whole sheet on rank 0, one step per launch (tests/ compare that plan "
                            "with the oracle)",
               "max_rel_err": float(np.abs(whole - ref).max() / den),
               "bit_identical": bool(np.array_equal(whole, ref)),
               "probe_series_max_rel_err": float(np.abs(ser.cpu().numpy() - o["probes"]).max() / den),
               "sum_u": float(whole.sum()), "sum_u_expected": float(steps + 1)}
    dist.barrier()
    return out


def bench_sheet(args, emit: bool = True, keep_pg: bool = False):
    """``bench.py --gpus N`` (N > 1) / ``--workload cfg2-sharded``: ONE sheet over N GPUs.  Weak scaling: the sheet grows with
    N so that every GPU holds as many nodes as the single-GPU cfg2 sheet (size = 8192 * sqrt(N), rounded to a
    multiple of 32*N); ``--size`` fixes the sheet instead (strong scaling)."""
    import torch
    import torch.distributed as dist

    from bench import METRIC, UNIT, ClockSampler, measured_peak_gbs, ncu_traffic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    strong = args.size is not None
    if strong:
        size = args.size
    else:
        q = 32 * world
        size = int(round(8192 * np.sqrt(world) / q)) * q
    main = torch.cuda.Stream(dev)
    torch.cuda.set_stream(main)
    backend = CudaSheetBackend(size, *band_rows(size, rank, world), device=local)
    runner = SheetRunner(size, backend, rank, world, dist if world > 1 else None)
    dt_eff = min(0.1, 1.0 / np.sqrt(2 * size + 1))
    coeff = dt_eff**2
    kick = ((size // 2) * size + size // 2) * 4
    kicks = ([runner.local_index(kick)], [1.0]) if runner.owns(kick) else ([], [])
    kmax = int(os.environ.get("TK_LAT_FUSE", "0") or 0) or 48

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    clk = ClockSampler(local).start()
    backend.set_state_sparse(*kicks)
    runner.run(args.warmup, coeff, 0.0, kmax)
    fence()
    # the timed region (args.steps steps, barrier + synchronize on both sides, slowest rank counts) is repeated
    # until >= 50 ms and >= 5 repeats have been seen; the line reports the median
    reps, kms, kpasses, launches = [], 0.0, 0, 0
    with clk:
        while len(reps) < 5 or (sum(reps) < 50.0 and len(reps) < 40):
            l0 = backend.plan.launches
            backend.plan.sheet_kernel_ms()  # reset the pass-kernel timer
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fence()
            e0.record(main)
            runner.run(args.steps, coeff, 0.0, kmax)
            e1.record(main)
            fence()
            launches = backend.plan.launches - l0
            kms, kpasses = backend.plan.sheet_kernel_ms()
            ms_t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            reps.append(float(ms_t.item()))
    ms = float(np.median(reps))
    steps_done = args.warmup + args.steps * len(reps)
    n_updates = size * size * 4
    value = n_updates * args.steps / (ms * 1e-3) / 1e9

    final_t = torch.empty(backend.plan.num_nodes, dtype=torch.float64, pin_memory=True)
    final = final_t.numpy()
    from . import _lib

    _lib.check(_lib.lib().tk_plan_get_state(backend.plan._h, _lib._p(final), None))
    cs = torch.tensor([float(final.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cs)

    # e2e with host buffers, as at N = 1 (bench.py:bench_lattice): kick H2D, the steps, the kick node's time series D2H
    # and its |DFT| + dominant bin on the rank that owns the node; "with_final_state" adds every band's D2H
    fence()
    t0 = time.perf_counter()
    backend.set_state_sparse(*kicks)
    if kicks[0]:
        backend.set_probes(kicks[0], args.steps + 1)
    runner.run(args.steps, coeff, 0.0, kmax)
    probe_bytes = 0
    if kicks[0]:
        ser = backend.read_probes(args.steps + 1)
        spectrum, _ = _lib.dft_abs(np.ascontiguousarray(ser[:, 0]), device=local)
        probe_bytes = ser.nbytes + spectrum.nbytes + 8
    fence()
    t1 = time.perf_counter()
    _lib.check(_lib.lib().tk_plan_get_state(backend.plan._h, _lib._p(final), None))
    fence()
    e2e_t = torch.tensor([t1 - t0, time.perf_counter() - t0, float(probe_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_probe_s, e2e_s, probe_bytes = (float(x) for x in e2e_t.tolist())
    passes = pass_lengths(args.steps, kmax)
    peak, peak_src = measured_peak_gbs()
    line = None
    if rank == 0:
        pass_bytes = 32.0 * backend.plan.num_nodes  # rank 0's band: read u_t, u_{t-1}; write u_{t+K}, u_{t+K-1}
        achieved = pass_bytes * kpasses / (kms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2-sharded",
                       "lattice": f"ONE 2-D tetrakis sheet {size}x{size}, no defect, fp64, {world} row bands",
                       "nodes": n_updates, "max_degree": 2 * size + 1, "effective_dt": dt_eff,
                       "passes": passes,
                       "parallelism": f"row bands x{world}; one NCCL all-reduce of {8 * size + 8} doubles (column "
                                      "sums + totals of two time levels) per pass",
                       "l2": "band state >> 126 MB L2 (no flush needed)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # ncu capture of the same kernel at the same K on the 8192^2 sheet, scaled to the band (or null)
                         "traffic": (lambda t: None if t is None else t * backend.plan.num_nodes / 268435456.0)(
                             ncu_traffic("cfg2-2d-8192-fused", max(passes))),
                         "peak_source": peak_src, "kernel": f"lat_fused2d_kernel (K={max(passes)}), rank 0's band",
                         "algorithmic_bytes_per_launch": pass_bytes, "bytes_per_update": 32.0 * len(passes) / args.steps,
                         "kernel_ms_per_launch": kms / max(kpasses, 1),
                         "kernel_share_of_step": kms / ms, "frac_of_8TBs_nominal": achieved / 8000.0},
            "cpu_baseline": None,
            "e2e": {"value": n_updates * args.steps / e2e_probe_s / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": (16.0 + probe_bytes / 2) / args.steps,
                    "d2h_bytes_per_step": probe_bytes / args.steps, "seconds": e2e_probe_s,
                    "includes": "kick H2D + run (all ranks, all-reduce per pass) + the kick node's series D2H, its "
                                "|DFT| and dominant bin (tk_dft_abs) on the owning rank; slowest rank",
                    "with_final_state": {"value": n_updates * args.steps / e2e_s / 1e9, "seconds": e2e_s,
                                         "d2h_bytes_per_step": (world * final.nbytes + probe_bytes) / args.steps,
                                         "includes": "the above + every band's final state D2H into pinned memory"}},
            "gpu_launches": launches, "clocks": clk.summary(),
            "timing": {"repeats": len(reps), "median_ms": ms, "min_ms": min(reps), "max_ms": max(reps),
                       "rule": "the steps-step region repeated until >= 50 ms and >= 5 repeats; value from the median; "
                               "the roofline kernel time is that of the last repeat"},
            "checksum_sum_u": float(cs.item()), "checksum_expected": float(steps_done + 1),
        }
    backend.close()
    if world > 1:
        dist.barrier()
        if os.environ.get("TK_BENCH_PARITY", "1") != "0":
            par = sheet_parity_check(dist, rank, world, local)
            if line is not None:
                line["parity"] = par
    if line is not None and emit:
        print(json.dumps(line), flush=True)
    if world > 1 and not keep_pg:
        dist.destroy_process_group()
    return 0 if emit else line
