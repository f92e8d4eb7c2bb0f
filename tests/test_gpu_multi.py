"""GPU (needs >= 2 devices; skipped otherwise): the real multi-process slab path -- one process per
GPU, NCCL halo exchange on a side stream, CUDA plan per slab -- against the single-domain oracle."""

import os
import socket

import numpy as np
import pytest

import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo
from tests.helpers import rel_inf

pytestmark = pytest.mark.gpu

SPEC = dict(size=48, dim=3, layers=13, defect="blackhole", radius=4.2)
STEPS, COEFF, DAMPING = 25, 0.003, 0.005


def _worker(rank, world, port, out, transport, jitter=False, steps=STEPS):
    import time

    import torch
    import torch.distributed as dist

    from tetrakis_sim_b200.slab import CudaSlabBackend, SlabRunner, global_max_degree

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if transport == "p2p-split":
        os.environ["TK_SLAB_SPLIT"] = "1"
        transport = "p2p"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        spec = tk.LatticeSpec(**SPEC)
        dev = torch.device(f"cuda:{rank}")
        if jitter or os.environ.get("TK_TEST_OWN_STREAM") == "1":
            torch.cuda.set_stream(torch.cuda.Stream(dev))  # a stream of its own, as bench.py uses (not the legacy one)
        backend = CudaSlabBackend(spec, *spec.slab(rank, world), device=rank)
        runner = SlabRunner(spec, backend, rank, world, dist,
                            comm_stream=torch.cuda.Stream(dev, priority=-1),
                            main_stream=torch.cuda.current_stream(dev), transport=transport)
        md = global_max_degree(backend.max_degree, dist, dev)
        kicks = {spec.index(spec.kick_node()): 1.0, spec.index((0, 0, 0, "B")): -0.5,
                 spec.index((spec.size - 1, 1, spec.layers - 1, "D")): 0.25}
        mine = {runner.local_index(g): v for g, v in kicks.items() if runner.owns(g)}
        backend.set_state_sparse(list(mine), list(mine.values()))
        if jitter:  # load the delay kernel now: a first-use module load in the middle of a run of stream-ordered
            torch.cuda._sleep(1000)  # waits can need the context to drain
            torch.cuda.synchronize()
            dist.barrier()
        runner.begin(COEFF, DAMPING)
        rng = np.random.default_rng(1234 + 17 * rank)  # every rank draws its own delays
        for _ in range(steps):
            if jitter:
                # a rank that runs ahead or falls behind by random amounts: up to ~2 ms of device stall on its
                # main stream and up to 1 ms of host stall, so flags and halos are produced and consumed out of step
                if rng.random() < 0.5 and os.environ.get("TK_STRESS_DEVICE_DELAY", "1") == "1":
                    torch.cuda._sleep(int(rng.integers(1, 4_000_000)))
                if rng.random() < 0.3:
                    time.sleep(float(rng.random()) * 1e-3)
            runner.step()
        runner.finish()
        torch.cuda.synchronize()
        np.save(os.path.join(out, f"slab{rank}.npy"), backend.get_state())
        if rank == 0:
            np.save(os.path.join(out, "maxdeg.npy"), np.array([md]))
        dist.barrier()
        if transport.startswith("p2p"):
            backend.detach_peers()
        backend.plan.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["nccl", "p2p", "p2p-split", "p2p-copy"])
@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_slab_run_matches_the_oracle(world, transport, tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), transport), nprocs=world, join=True)
    spec = tk.LatticeSpec(**SPEC)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)])
    flags = np.full(spec.num_slots, 3, dtype=np.uint8)
    si, sf, _, _ = spec.specials()
    flags[si] = sf
    u0 = np.zeros(spec.num_slots)
    u0[spec.index(spec.kick_node())] = 1.0
    u0[spec.index((0, 0, 0, "B"))] = -0.5
    u0[spec.index((spec.size - 1, 1, spec.layers - 1, "D"))] = 0.25
    ref = wo.run_structured(spec.size, spec.layers, flags, None, None, [], u0, STEPS, COEFF, DAMPING)["u"]
    assert rel_inf(got, ref) < 1e-12
    assert int(np.load(tmp_path / "maxdeg.npy")[0]) == 3 + 2 * (spec.size - 1) + 2


@pytest.mark.parametrize("transport", ["p2p", "p2p-split", "p2p-copy", "nccl"])
def test_halo_flag_protocol_survives_random_delays(transport, tmp_path):
    """Stress of the stream-ordered flag protocol (cuStreamWriteValue64 / cuStreamWaitValue64 on peer memory): every
    rank stalls its stream and its host thread by random, rank-specific amounts between steps, so neighbours run up to
    a step ahead of or behind each other for 60 steps; the result must still be the oracle's (a halo plane that was
    overwritten too early, or read too early, would show up here)."""
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    steps = 60
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), transport, True, steps), nprocs=world, join=True)
    spec = tk.LatticeSpec(**SPEC)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)])
    flags = np.full(spec.num_slots, 3, dtype=np.uint8)
    si, sf, _, _ = spec.specials()
    flags[si] = sf
    u0 = np.zeros(spec.num_slots)
    u0[spec.index(spec.kick_node())] = 1.0
    u0[spec.index((0, 0, 0, "B"))] = -0.5
    u0[spec.index((spec.size - 1, 1, spec.layers - 1, "D"))] = 0.25
    ref = wo.run_structured(spec.size, spec.layers, flags, None, None, [], u0, steps, COEFF, DAMPING)["u"]
    assert rel_inf(got, ref) < 1e-12
