#!/usr/bin/env python3
"""Full-size cross-check of BASELINE config 3 (3-D 1024 x 1024 x 512, black hole r = 64; 2^31 node slots):
the layer-slab multi-GPU run against the SAME lattice stepped on one GPU.

    torchrun --nproc-per-node 2 scripts/cfg3_crosscheck.py [--steps 12] [--transport p2p|nccl]

Every rank owns 512/N layers.  Probes sit at the kick node, directly across each slab boundary and in the
top layer (linear indices beyond 2^31 - 2^22, i.e. past INT32_MAX in the global numbering).  Rank 0 then frees
its slab, builds the whole lattice on its own GPU (34 GB of state) and compares the probe series: the two runs
use the same kernel on the same layers, so they must agree to the last bit.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import tetrakis_sim_b200 as tk  # noqa: E402
from tetrakis_sim_b200.slab import CudaSlabBackend, SlabRunner, global_max_degree  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=512)
    ap.add_argument("--transport", default="p2p")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    spec = tk.LatticeSpec(size=a.size, dim=3, layers=a.layers, defect="blackhole", radius=64.0)
    S, L = spec.size, spec.layers
    kick = spec.kick_node()
    top = (S // 2, S // 2 + 3, L - 1, "C")
    kicks = {kick: 1.0, top: -0.75}
    probes = [kick, top, (S // 2, S // 2 + 3, L - 2, "C"), (0, 0, 0, "A"), (S - 1, S - 1, L - 1, "D")]
    for r in range(1, world):  # both sides of every slab boundary, above the hole's column
        zb = spec.slab(r, world)[0]
        probes += [(S // 2, S // 2 + 70, zb - 1, "B"), (S // 2, S // 2 + 70, zb, "B")]
    main_s = torch.cuda.Stream(dev)
    torch.cuda.set_stream(main_s)
    backend = CudaSlabBackend(spec, *spec.slab(rank, world), device=local)
    runner = SlabRunner(spec, backend, rank, world, dist, comm_stream=torch.cuda.Stream(dev, priority=-1),
                        main_stream=main_s, transport=a.transport)
    maxdeg = global_max_degree(backend.max_degree, dist, dev)
    coeff = min(0.1, 1.0 / np.sqrt(maxdeg)) ** 2
    mine = {runner.local_index(spec.index(n)): v for n, v in kicks.items() if runner.owns(spec.index(n))}
    backend.set_state_sparse(list(mine), list(mine.values()))
    owned = [(i, runner.local_index(spec.index(n))) for i, n in enumerate(probes) if runner.owns(spec.index(n))]
    backend.plan.set_probes([li for _, li in owned], a.steps + 1)
    t0 = time.perf_counter()
    runner.begin(coeff, 0.0)
    for _ in range(a.steps):
        runner.step()
    runner.finish()
    torch.cuda.synchronize()
    t_slab = time.perf_counter() - t0
    series = backend.plan.read_probes(a.steps + 1)
    gathered = [None] * world
    dist.all_gather_object(gathered, ([i for i, _ in owned], series))
    dist.barrier()
    if a.transport == "p2p":
        backend.detach_peers()
    backend.plan.close()
    dist.barrier()
    ok = True
    if rank == 0:
        multi = np.zeros((a.steps + 1, len(probes)))
        for ids, ser in gathered:
            for j, i in enumerate(ids):
                multi[:, i] = ser[:, j]
        t0 = time.perf_counter()
        plan = tk.compile_lattice(spec, device=local)
        t_build = time.perf_counter() - t0
        assert plan.max_degree == maxdeg
        plan.set_state_sparse([spec.index(n) for n in kicks], list(kicks.values()))
        out = plan.run(a.steps, coeff, 0.0, probes=[spec.index(n) for n in probes], timed=True)
        single = out["probes"]
        # whole-lattice checksum on the device: sum_i u_t(i) = sum(u_0) * (1 + t)
        b0, b1 = plan.state_ptrs()
        n_all = (L + 2) * spec.plane

        class Raw:
            def __init__(self, ptr):
                self.__cuda_array_interface__ = {"shape": (n_all,), "typestr": "<f8", "data": (ptr, False), "version": 3}

        sums = [float(torch.as_tensor(Raw(p), device=dev).sum().item()) for p in (b0, b1)]
        expect = 0.25 * (1 + a.steps)
        plan.close()
        diff = float(np.max(np.abs(multi - single)))
        # identical tiling (TK_LAT_RCHUNK set) => identical summation order => bit-identical; otherwise the
        # column partials are associated differently and the last bits may differ
        ok = diff <= 1e-13 * float(np.max(np.abs(single))) and any(abs(s - expect) < 1e-9 * abs(expect) for s in sums)
        print(json.dumps({
            "check": "cfg3 full lattice: slabs over N GPUs vs one GPU", "n_gpus": world, "transport": a.transport,
            "lattice": f"{S}x{S}x{L}", "node_slots": spec.num_slots, "max_linear_index_probed": spec.index(probes[4]),
            "steps": a.steps, "probes": len(probes), "max_abs_diff": diff, "bit_identical": diff == 0.0,
            "max_abs_probe": float(np.max(np.abs(single))), "sum_u_single_gpu": sums, "sum_u_expected": expect,
            "single_gpu_ms_per_step": out["elapsed_ms"] / a.steps, "single_gpu_gnups": spec.num_nodes * a.steps / out["elapsed_ms"] / 1e6,
            "slab_wall_s": t_slab, "single_gpu_plan_build_s": t_build, "ok": ok}))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
