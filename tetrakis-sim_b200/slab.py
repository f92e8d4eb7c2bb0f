"""Layer-slab sharding of 3-D lattices across GPUs (one process per GPU).

Only the z direction of a tetrakis lattice is a true stencil: row / column cliques never leave a
layer (reference ``tetrakis_sim/lattice.py:58-66``) and layers couple only through
``(r,c,z,q)-(r,c,z+1,q)`` (``lattice.py:68-72``).  So rank g owns a contiguous block of layers
plus one halo plane on each side, row/column sums stay rank-local, and a step needs exactly one
exchange: each rank sends its first and last owned plane of u_{t+1} to its z-neighbours.  No
all-reduce is involved (SURVEY.md sections 0.3 and 8e).

Overlap: the two boundary layers are updated first, their planes go out on a side stream (NCCL
send/recv over NVLink) while the interior layers are updated on the main stream, and the next
step waits for the receive.  ``torch.distributed`` is plumbing only; all arithmetic is in the CUDA
plan (``tk_lattice_step_*`` in include/tetrakis_b200.h).

The orchestration is backend-agnostic (anything with the plan's phase API and four plane tensors),
which is how tests/test_slab_gloo.py drives it on CPU with world_size 2 over gloo.
"""

from __future__ import annotations

import json
import os
import time

import numpy as np

from .lattice import LatticeSpec


class _RawCuda:
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3,
        }  # fmt: skip


class CudaSlabBackend:
    """The CUDA plan of one slab + torch views of its halo planes."""

    def __init__(self, spec: LatticeSpec, z_begin: int, z_end: int, device: int):
        import torch

        from .compiler import compile_lattice

        self.torch = torch
        self.device = device
        self.plan = compile_lattice(spec, device=device, z_begin=z_begin, z_end=z_end)
        self.plan.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self._views: dict = {}

    def planes(self, next_state: bool):
        p = self.plan.halo_ptrs(next_state)
        esz, typestr = (4, "<f4") if self.plan.f32 else (8, "<f8")  # fp32 plans exchange fp32 planes
        n = p["plane_bytes"] // esz
        out = {}
        for k in ("send_lo", "send_hi", "recv_lo", "recv_hi"):
            key = p[k]
            if key not in self._views:
                self._views[key] = self.torch.as_tensor(_RawCuda(key, n, typestr), device=f"cuda:{self.device}")
            out[k] = self._views[key]
        return out

    def attach_peers(self, dist, rank: int, world: int, spec: LatticeSpec):
        """Map the z-neighbours' state buffers (cudaIpc) so the step kernel can push halos itself."""
        from . import _lib

        handles = [None] * world
        dist.all_gather_object(handles, self.plan.ipc_export())
        self._ipc = []
        for side, nb in ((0, rank - 1), (1, rank + 1)):
            if 0 <= nb < world:
                b0 = _lib.ipc_open(handles[nb][:64], self.device)
                b1 = _lib.ipc_open(handles[nb][64:], self.device)
                self._ipc += [b0, b1]
                zb, ze = spec.slab(nb, world)
                self.plan.peer_attach(side, b0, b1, ze - zb)

    def detach_peers(self):
        from . import _lib

        for side in (0, 1):
            self.plan.peer_attach(side, None, None, 0)
        for ptr in getattr(self, "_ipc", []):
            _lib.ipc_close(ptr, self.device)
        self._ipc = []

    def __getattr__(self, name):  # step_begin / step_layers / step_end / set_state_sparse / ...
        return getattr(self.plan, name)


def p2p_possible(dist, rank: int, world: int, device: int) -> bool:
    """The peer-memory transports need every z-neighbour on the same host and peer-accessible; all ranks must take
    the same decision.  False => the NCCL transport (multi-node runs, topologies without P2P)."""
    import socket

    import torch

    where = [None] * world
    dist.all_gather_object(where, (socket.gethostname(), int(device)))
    ok = True
    for nb in (rank - 1, rank + 1):
        if 0 <= nb < world:
            host, dev = where[nb]
            if host != where[rank][0]:
                ok = False
            elif dev != device and not torch.cuda.can_device_access_peer(device, dev):
                ok = False
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=f"cuda:{device}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


class SlabRunner:
    """Steps one rank's slab and exchanges halos with the z-neighbours."""

    def __init__(self, spec: LatticeSpec, backend, rank: int, world: int, dist=None,
                 comm_stream=None, main_stream=None, transport: str = "nccl"):  # fmt: skip
        if spec.dim != 3:
            raise ValueError("slab sharding is for 3-D lattices (2-D sheets run as replicas)")
        if spec.layers < world:
            raise ValueError("fewer layers than ranks")
        self.spec, self.backend, self.rank, self.world, self.dist = spec, backend, rank, world, dist
        self.z_begin, self.z_end = spec.slab(rank, world)
        self.n_local_layers = self.z_end - self.z_begin
        self.lo = rank - 1 if rank > 0 else None
        self.hi = rank + 1 if rank < world - 1 else None
        self.comm_stream, self.main_stream = comm_stream, main_stream
        self._pending = []
        # "nccl": send/recv of the boundary planes; "p2p": the step kernel stores them into the
        # neighbours' halos through peer-mapped pointers, synchronised by stream-ordered flags
        self.transport = transport if (dist is not None and world > 1 and comm_stream is not None) else "nccl"
        self._push_count = 0
        self.interior_first = os.environ.get("TK_SLAB_ORDER", "boundary_first") == "interior_first"
        self.split = os.environ.get("TK_SLAB_SPLIT", "0") == "1"
        self.copy_push = False
        # in-kernel early arrival signal (tk_lattice_arm_signal): measured SLOWER than the plain stream-ordered flag on
        # 2 x B200 (1.286 vs 1.229 ms per step, profiles/r02_slab_transports.md) -- scheduling the boundary layer groups
        # first costs their z-neighbours' L2 hits -- so it is opt-in
        self.early_signal = os.environ.get("TK_SLAB_EARLY", "0") == "1" and hasattr(backend, "arm_signal")
        if self.transport == "p2p-copy":  # peer memory + flags, planes moved by copy engines after the step
            self.transport, self.copy_push = "p2p", True
        if self.transport == "p2p" and hasattr(backend, "device") and not p2p_possible(dist, rank, world, backend.device):
            import warnings

            warnings.warn("z-neighbours are not peer-accessible GPUs of one host: falling back to the NCCL transport",
                          RuntimeWarning, stacklevel=2)
            self.transport, self.copy_push = "nccl", False
        if self.transport == "p2p":
            backend.attach_peers(dist, rank, world, spec)
            backend.set_push_mode(not self.copy_push)
        self._ev_copy = None

    # -- index helpers --------------------------------------------------------------------------
    def owns(self, global_index: int) -> bool:
        z = global_index // self.spec.plane
        return self.z_begin <= z < self.z_end

    def local_index(self, global_index: int) -> int:
        return global_index - self.z_begin * self.spec.plane

    # -- halo exchange --------------------------------------------------------------------------
    def _post(self, next_state: bool):
        if self.world == 1 or self.dist is None:
            return []
        d = self.dist
        pl = self.backend.planes(next_state)
        ops = []
        # fixed global order (lower neighbour first) keeps paired sends/recvs matched
        if self.lo is not None:
            ops.append(d.P2POp(d.isend, pl["send_lo"], self.lo))
            ops.append(d.P2POp(d.irecv, pl["recv_lo"], self.lo))
        if self.hi is not None:
            ops.append(d.P2POp(d.isend, pl["send_hi"], self.hi))
            ops.append(d.P2POp(d.irecv, pl["recv_hi"], self.hi))
        if not ops:
            return []
        if self.comm_stream is not None:
            import torch

            ev = torch.cuda.Event()
            ev.record(self.main_stream)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                return d.batch_isend_irecv(ops)
        return d.batch_isend_irecv(ops)

    def _wait(self):
        for r in self._pending:
            r.wait()  # NCCL: orders the current (main) stream after the transfer; gloo: blocks
        self._pending = []

    # -- simulation -----------------------------------------------------------------------------
    def begin(self, coeff: float, damping: float):
        """Row/column sums of the initial state + first halo fill."""
        self.backend.step_begin(coeff, damping)
        if self.world == 1:
            return
        if self.transport == "p2p":
            import torch

            self.dist.barrier()  # nobody still reads the halos of a previous run
            st = self.comm_stream if self.split else self.main_stream
            if self.split:
                ev = torch.cuda.Event()
                ev.record(self.main_stream)
                st.wait_event(ev)
            self.backend.push_halos(st.cuda_stream)
            self._push_count += 1
            self.backend.signal(self._push_count, st.cuda_stream)
            self._ev_end = torch.cuda.Event()
            self._ev_end.record(self.main_stream)
            return
        self._pending = self._post(next_state=False)
        if self.comm_stream is not None:
            import torch

            self._ev_end = torch.cuda.Event()
            self._ev_end.record(self.main_stream)
        else:
            self._wait()

    def step(self):
        if self.world == 1:  # no neighbours: the same slab code, one launch over all layers + finalise
            self.backend.step_layers(0, self.n_local_layers)
            self.backend.step_end()
            return
        if self.transport == "p2p":
            return self._step_p2p()
        if self.comm_stream is not None:
            return self._step_cuda()
        L = self.n_local_layers
        b = self.backend
        self._wait()  # halos of u_t are in place
        b.step_layers(0, 1)
        if L > 1:
            b.step_layers(L - 1, L)
        self._pending = self._post(next_state=True)  # boundary planes of u_{t+1} leave now ...
        if L > 2:
            b.step_layers(1, L - 1)  # ... while the interior is updated
        b.step_end()

    def _step_cuda(self):
        """Boundary layers + halo exchange on the (high-priority) side stream, concurrently with the
        interior layers on the main stream; the finalise/swap joins both.

        The two boundary launches are small (one layer each: a few warps per SM, latency-bound, ~0.1
        ms) -- run back to back with the interior they cost 0.4 ms per step (measured at N=2); run
        beside it they are free and the NCCL transfer starts ~0.8 ms before it is needed."""
        import torch

        L, b = self.n_local_layers, self.backend
        side, main = self.comm_stream, self.main_stream
        side.wait_event(self._ev_end)  # row/column sums of u_t (previous finalise) are ready
        with torch.cuda.stream(side):
            self._wait()  # halos of u_t have arrived (orders the side stream after NCCL)
            b.step_layers_on(0, 1, side.cuda_stream)
            if L > 1:
                b.step_layers_on(L - 1, L, side.cuda_stream)
            ev_bnd = torch.cuda.Event()
            ev_bnd.record(side)
            self._pending = self._post_current_stream(next_state=True)
        if L > 2:
            b.step_layers(1, L - 1)
        main.wait_event(ev_bnd)
        b.step_end()
        self._ev_end = torch.cuda.Event()
        self._ev_end.record(main)

    def _step_p2p(self):
        """Fused boundary update + halo push: no NCCL in the loop.  Flag value k = "k pushes done".

        Default: ONE launch over all owned layers -- the CTAs that update the first / last layer store
        their results into the neighbours' halos as they go -- then the flag, then the finalise.  All
        ranks run the same kernel on the same amount of data, so the wait for the neighbours' flag at
        the top of the next step is a few microseconds; nothing is split, nothing needs overlapping.
        TK_SLAB_SPLIT=1 keeps the boundary-first / interior-concurrent schedule (needed for NCCL)."""
        if self.copy_push:
            return self._step_p2p_copy()
        if not self.split:
            b = self.backend
            st = self.main_stream.cuda_stream
            b.wait(self._push_count, st)
            self._push_count += 1
            if self.early_signal:  # the kernel itself raises the neighbours' flags once its boundary warps are done
                b.arm_signal(self._push_count)
            b.step_layers(0, self.n_local_layers)
            # ... and the stream-ordered write of the same value follows the kernel in any case: the in-kernel store is
            # an optimisation (the neighbour may start its next step while this rank's interior is still being
            # updated), the stream write is the guarantee
            b.signal(self._push_count, st)
            b.step_end()
            return
        import torch

        L, b = self.n_local_layers, self.backend
        side, main = self.comm_stream, self.main_stream
        interior_first = self.interior_first
        if interior_first and L > 2:
            b.step_layers(1, L - 1)
        side.wait_event(self._ev_end)
        b.wait(self._push_count, side.cuda_stream)  # the neighbours' planes of u_t sit in my halos
        b.step_layers_on(0, 1, side.cuda_stream)  # ... and these launches push u_{t+1} into theirs
        if L > 1:
            b.step_layers_on(L - 1, L, side.cuda_stream)
        ev_bnd = torch.cuda.Event()
        ev_bnd.record(side)
        self._push_count += 1
        b.signal(self._push_count, side.cuda_stream)
        if not interior_first and L > 2:
            b.step_layers(1, L - 1)
        main.wait_event(ev_bnd)
        b.step_end()
        self._ev_end = torch.cuda.Event()
        self._ev_end.record(main)

    def _step_p2p_copy(self):
        """Peer memory + flags, but the planes travel by cudaMemcpyAsync on a side stream (copy engines)
        while the finalise runs: the SMs issue no remote stores."""
        import torch

        b, main, side = self.backend, self.main_stream, self.comm_stream
        b.wait(self._push_count, main.cuda_stream)  # the neighbours' planes of u_t have landed
        if self._ev_copy is not None:
            main.wait_event(self._ev_copy)  # my own outgoing copies of two steps ago no longer read this buffer
        b.step_layers(0, self.n_local_layers)
        ev_k = torch.cuda.Event()
        ev_k.record(main)
        b.step_end()  # finalise on the main stream; the new state is now "current"
        side.wait_event(ev_k)
        b.push_halos(side.cuda_stream)
        self._push_count += 1
        b.signal(self._push_count, side.cuda_stream)
        self._ev_copy = torch.cuda.Event()
        self._ev_copy.record(side)

    def _post_current_stream(self, next_state: bool):
        if self.world == 1 or self.dist is None:
            return []
        d = self.dist
        pl = self.backend.planes(next_state)
        ops = []
        if self.lo is not None:
            ops += [d.P2POp(d.isend, pl["send_lo"], self.lo), d.P2POp(d.irecv, pl["recv_lo"], self.lo)]
        if self.hi is not None:
            ops += [d.P2POp(d.isend, pl["send_hi"], self.hi), d.P2POp(d.irecv, pl["recv_hi"], self.hi)]
        return d.batch_isend_irecv(ops) if ops else []

    def finish(self):
        self._wait()
        if self.comm_stream is not None:
            self.main_stream.wait_stream(self.comm_stream)


def global_max_degree(plan_max: int, dist, device=None) -> int:
    """CFL input: max degree over all slabs (reference physics.py:64) -- a one-off scalar MAX."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return plan_max
    import torch

    t = torch.tensor([plan_max], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


class _Solo:
    """Stand-in for torch.distributed on one rank (bench.py --gpus 1 runs the same slab code with no neighbours)."""

    class ReduceOp:
        MAX = SUM = None

    @staticmethod
    def barrier():
        pass

    @staticmethod
    def all_reduce(t, op=None):
        return t

    @staticmethod
    def is_initialized():
        return False


def slab_parity_check(dist, rank: int, world: int, local: int, main, comm, transport: str, steps: int = 40):
    """In-run parity of the multi-GPU slab path: a reduced lattice (96 x 96 x 6 layers per rank, black hole r = 9)
    stepped by the SAME SlabRunner / transport as the timed run, against ONE plan holding the whole lattice on
    rank 0 -- final state (every node) and the kick node's time series.  Returns the dict for the bench line."""
    import torch

    dev = torch.device(f"cuda:{local}")
    spec = LatticeSpec(size=96, dim=3, layers=6 * world, defect="blackhole", radius=9.0)
    backend = CudaSlabBackend(spec, *spec.slab(rank, world), device=local)
    runner = SlabRunner(spec, backend, rank, world, dist, comm_stream=comm, main_stream=main, transport=transport)
    maxdeg = global_max_degree(backend.max_degree, dist, dev)
    coeff = min(0.1, 1.0 / np.sqrt(maxdeg)) ** 2
    kick = spec.index(spec.kick_node())
    far = spec.index((7, 11, spec.layers - 1, "C"))  # a node of the last slab
    probes = [g for g in (kick, far) if runner.owns(g)]
    backend.set_state_sparse(*(([runner.local_index(kick)], [1.0]) if runner.owns(kick) else ([], [])))
    if probes:
        backend.set_probes([runner.local_index(g) for g in probes], steps + 1)
    runner.begin(coeff, 0.0)
    for _ in range(steps):
        runner.step()
    runner.finish()
    torch.cuda.synchronize()
    series = {g: backend.read_probes(steps + 1)[:, i].copy() for i, g in enumerate(probes)}
    mine = torch.from_numpy(backend.plan.get_state()).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    all_series = [None] * world
    dist.all_gather_object(all_series, series)
    out = None
    if rank == 0:
        from .compiler import compile_lattice

        whole = torch.cat(parts).cpu().numpy()
        got = {}
        for d in all_series:
            got.update(d)
        plan = compile_lattice(spec, device=local)
        try:
            plan.set_state_sparse([kick], [1.0])
            o = plan.run(steps, coeff, 0.0, probes=[kick, far], fuse=False)
            ref = plan.get_state()
            os.environ["TK_FX_MIN_NODES"] = "0"  # the same lattice through the K-step passes (cross scheme) as well
            os.environ["TK_FX_MAXFRAC"] = "95"
            plan.set_state_sparse([kick], [1.0])
            plan.run(steps, coeff, 0.0, probes=[kick, far])
            fused, fused_k = plan.get_state(), plan.cross["steps_per_pass"]
            del os.environ["TK_FX_MIN_NODES"], os.environ["TK_FX_MAXFRAC"]
        finally:
            plan.close()
        den = float(np.abs(ref).max())
        ser = np.stack([got[kick], got[far]], axis=1)
        out = {"lattice": f"3-D tetrakis 96x96x{spec.layers} ({6} layers per rank), black hole r=9, {steps} steps, "
                          f"transport={runner.transport}",
               "reference": "ONE plan holding the whole lattice on rank 0, one step per launch (tests/ compare that "
                            "plan with the oracle)",
               "max_rel_err": float(np.abs(whole - ref).max() / den),
               "bit_identical": bool(np.array_equal(whole, ref)),
               "probe_series_max_rel_err": float(np.abs(ser - o["probes"]).max() / den),
               "probe_series_bit_identical": bool(np.array_equal(ser, o["probes"])),
               "sum_u": float(whole.sum()), "sum_u_expected": float(steps + 1),
               "fused_passes_vs_stepwise": {"steps_per_pass": fused_k,
                                            "max_rel_err": float(np.abs(fused - ref).max() / den)}}
    dist.barrier()
    if runner.transport == "p2p":
        backend.detach_peers()
        backend.set_push_mode(True)
    backend.plan.close()
    dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): weak scaling, 64 layers of 1024x1024 per GPU (BASELINE configs[2])
# ------------------------------------------------------------------------------------------------
def bench_slabs(args, emit: bool = True, keep_pg: bool = False):
    import torch
    import torch.distributed as dist

    from . import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    if world == 1:
        dist = _Solo()  # noqa: F811 -- the same code path with no neighbours
    elif not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    workload = args.workload or "cfg3-slab-1024x64"
    size = args.size or 1024
    per_gpu = args.layers or 64
    spec = LatticeSpec(size=size, dim=3, layers=per_gpu * world, defect="blackhole", radius=64.0)

    # a dedicated non-blocking stream: the legacy default stream synchronises implicitly with every
    # blocking stream of the process, which we do not want around stream-ordered flag waits
    main = torch.cuda.Stream(dev)
    torch.cuda.set_stream(main)
    comm = torch.cuda.Stream(dev, priority=int(os.environ.get("TK_SLAB_PRIO", "-1")))  # split mode: boundary layers (+ NCCL)
    if not os.environ.get("TK_WATCHDOG"):  # a stalled exchange must fail loudly, not hang the box
        import faulthandler

        faulthandler.dump_traceback_later(600, exit=True)
    backend = CudaSlabBackend(spec, *spec.slab(rank, world), device=local)
    mode = os.environ.get("TK_SLAB_MODE", "full")  # diagnostics: "nocomm" skips the exchange, "serial"
    transport = os.environ.get("TK_SLAB_TRANSPORT", "p2p")
    runner = SlabRunner(spec, backend, rank, world, None if mode == "nocomm" else dist,
                        comm_stream=None if mode == "serial" else comm, main_stream=main,
                        transport=transport)
    maxdeg = global_max_degree(backend.max_degree, dist, dev)
    dt_eff = min(0.1, 1.0 / np.sqrt(maxdeg))
    coeff = dt_eff**2
    kick = spec.index(spec.kick_node())
    kicks = ([runner.local_index(kick)], [1.0]) if runner.owns(kick) else ([], [])

    def run(steps):
        runner.begin(coeff, 0.0)
        for _ in range(steps):
            runner.step()
        runner.finish()

    from bench import ClockSampler  # same sampler as the single-GPU arm

    clk = ClockSampler(local).start()
    backend.set_state_sparse(*kicks)
    run(args.warmup)
    torch.cuda.synchronize()
    dist.barrier()
    # the timed region (args.steps steps, barrier + synchronize on both sides, slowest rank counts) is repeated
    # until >= 50 ms and >= 5 repeats have been seen; the line reports the median
    reps, host_enqueue_s, launches = [], 0.0, 0
    with clk:
        while len(reps) < 5 or (sum(reps) < 50.0 and len(reps) < 40):
            l0 = backend.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            dist.barrier()
            e0.record(main)
            h0 = time.perf_counter()
            runner.begin(coeff, 0.0)
            for _ in range(args.steps):
                runner.step()
            runner.finish()
            e1.record(main)
            host_enqueue_s = time.perf_counter() - h0
            torch.cuda.synchronize()
            dist.barrier()
            launches = backend.launches - l0
            ms_t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            reps.append(float(ms_t.item()))
    ms = float(np.median(reps))
    steps_done = args.warmup + args.steps * len(reps)
    n_updates = spec.num_nodes
    value = n_updates * args.steps / (ms * 1e-3) / 1e9

    # whole-lattice checksum: sum_i u_t(i) = 1 + t (unit masses; removed nodes stay 0)
    final_t, final = None, None
    final_t = torch.empty(backend.num_nodes, dtype=torch.float64, pin_memory=True)
    final = final_t.numpy()
    _lib.check(_lib.lib().tk_plan_get_state(backend.plan._h, _lib._p(final), None))
    cs = torch.tensor([float(final.sum())], dtype=torch.float64, device=dev)
    dist.all_reduce(cs)

    # e2e: kicks H2D, K steps with halo exchange, final state D2H to pinned host memory, per rank
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    backend.set_state_sparse(*kicks)
    if kicks[0]:
        backend.set_probes(kicks[0], args.steps + 1)
    run(args.steps)
    probe_bytes = 0
    if kicks[0]:  # the kick node's time series D2H, its |DFT| and dominant bin on the rank that owns the node
        ser = backend.read_probes(args.steps + 1)
        spectrum, _ = _lib.dft_abs(np.ascontiguousarray(ser[:, 0]), device=local)
        probe_bytes = ser.nbytes + spectrum.nbytes + 8
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.perf_counter()
    _lib.check(_lib.lib().tk_plan_get_state(backend.plan._h, _lib._p(final), None))
    torch.cuda.synchronize()
    dist.barrier()
    e2e_t = torch.tensor([t1 - t0, time.perf_counter() - t0, float(probe_bytes)], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_probe_s, e2e_s, probe_bytes = (float(x) for x in e2e_t.tolist())
    backend.set_probes([], 0)

    # roofline of the step kernel on this rank's slab, no communication (rank 0 reports)
    backend.set_state_sparse(*kicks)
    o = backend.plan.run(min(args.steps, 50), coeff, 0.0, timed=True, kernel_timing=True, fuse=False)
    kms = o["step_kernel_ms"] / min(args.steps, 50)
    alone_ms = o["elapsed_ms"] / min(args.steps, 50)  # this rank's slab stepped alone: step + finalise
    local_nodes = backend.num_nodes  # slots; the black hole removes < 1 %
    from bench import BYTES_PER_UPDATE, METRIC, UNIT, measured_peak_gbs, ncu_traffic

    peak, peak_src = measured_peak_gbs()
    achieved = local_nodes * BYTES_PER_UPDATE / (kms * 1e-3) / 1e9
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload,
                       "lattice": f"3-D tetrakis {size}x{size}x{spec.layers} ({per_gpu} layers per GPU), "
                                  f"black hole r=64 at the global centre, fp64",
                       "nodes": n_updates, "max_degree": maxdeg, "effective_dt": dt_eff,
                       "parallelism": (f"z-slabs x{world}, 1-plane halos; boundary layers on a side stream "
                                       "concurrent with the interior; transport=" + runner.transport +
                                       (" (step kernel stores the planes into the neighbours' halos over "
                                        "NVLink peer pointers, stream-ordered flags)" if runner.transport == "p2p"
                                        else " (NCCL send/recv)")),
                       "halo_bytes_per_direction": size * size * 32,
                       "l2": "slab state >> 126 MB L2 (no flush needed)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic("cfg3-slab-1024x64"),
                         "algorithmic_bytes_per_launch": local_nodes * BYTES_PER_UPDATE,
                         "peak_source": peak_src,
                         "kernel": "lat_step_kernel (one slab, no communication)",
                         "bytes_per_update": BYTES_PER_UPDATE, "kernel_ms_per_launch": kms},
            "cpu_baseline": None,
            "e2e": {"value": n_updates * args.steps / e2e_probe_s / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": (16.0 + probe_bytes / 2) / args.steps,
                    "d2h_bytes_per_step": probe_bytes / args.steps, "seconds": e2e_probe_s,
                    "includes": "kick H2D + run (all ranks, halo exchange per step) + the kick node's series D2H, its "
                                "|DFT| and dominant bin (tk_dft_abs) on the owning rank; slowest rank",
                    "with_final_state": {"value": n_updates * args.steps / e2e_s / 1e9, "seconds": e2e_s,
                                         "d2h_bytes_per_step": (world * final.nbytes + probe_bytes) / args.steps,
                                         "includes": "the above + every slab's final state D2H into pinned memory"}},
            "gpu_launches": launches, "clocks": clk.summary(),
            "host_enqueue_ms_per_step": 1e3 * host_enqueue_s / args.steps, "slab_mode": mode,
            # the same 64-layer slab on one GPU with no neighbours (rank 0, measured in this run): the
            # denominator for weak scaling of THIS workload (bench.py --gpus 1 defaults to the 2-D cfg2)
            "single_gpu_same_workload": {"value": local_nodes / (alone_ms * 1e-3) / 1e9, "unit": UNIT,
                                         "ms_per_step": alone_ms},
            "timing": {"repeats": len(reps), "median_ms": ms, "min_ms": min(reps), "max_ms": max(reps),
                       "rule": "the steps-step region repeated until >= 50 ms and >= 5 repeats; value from the median"},
            "checksum_sum_u": float(cs.item()), "checksum_expected": float(steps_done + 1),
        }
    dist.barrier()
    if runner.transport == "p2p":
        backend.detach_peers()
        backend.set_push_mode(True)
    backend.plan.close()
    dist.barrier()
    if world > 1 and os.environ.get("TK_BENCH_PARITY", "1") != "0":
        par = slab_parity_check(dist, rank, world, local, main, comm, transport)
        if line is not None:
            line["parity"] = par
    if line is not None and emit:
        print(json.dumps(line), flush=True)
    if world > 1 and not keep_pg:
        dist.destroy_process_group()
    return 0 if emit else line
