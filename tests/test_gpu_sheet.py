"""ONE defect-free sheet as row bands (tk_sheet_* in include/tetrakis_b200.h, tetrakis_sim_b200/sheet.py):
several band plans on one GPU stepped in lock-step with the all-reduce done by hand, and the real one-process-per-GPU
run over NCCL (skipped with fewer devices), against the oracle."""

import os
import socket

import numpy as np
import pytest

import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo
from tests.helpers import rel_inf
from tetrakis_sim_b200.sheet import CudaSheetBackend, SheetRunner, band_rows, pass_lengths

pytestmark = pytest.mark.gpu


def _oracle(S, kicks, steps, coeff, damping, probes):
    n = S * S * 4
    u0 = np.zeros(n)
    for g, v in kicks.items():
        u0[g] = v
    return wo.run_structured(S, 1, np.full(n, 3, np.uint8), np.ones(n), np.zeros(n), [], u0, steps, coeff, damping,
                             probes=probes)


def _kicks_probes(S):
    kicks = {((S // 2) * S + S // 2) * 4: 1.0, (S - 1) * 4 + 3: -0.5, ((S - 1) * S + 2 % S) * 4 + 1: 0.25}
    probes = [((S // 2) * S + S // 2) * 4, ((S - 1) * S + 2 % S) * 4 + 1, ((S // 3) * S + (2 * S) // 3) * 4 + 2]
    return kicks, probes


@pytest.mark.parametrize("S,nbands,steps,kmax,damping", [(70, 2, 37, 8, 0.01), (257, 3, 50, 48, 0.0),
                                                         (33, 4, 21, 5, 0.02), (64, 1, 25, 24, 0.0),
                                                         (9, 3, 13, 1, 0.0), (130, 2, 97, 48, 0.0)])
def test_lock_stepped_bands_on_one_gpu_match_the_oracle(S, nbands, steps, kmax, damping):
    import torch

    coeff = 0.9 / (2 * S + 1)
    kicks, probes = _kicks_probes(S)
    bands = [CudaSheetBackend(S, *band_rows(S, r, nbands), device=0) for r in range(nbands)]
    runners = [SheetRunner(S, b, r, nbands, None) for r, b in enumerate(bands)]
    try:
        for rn, b in zip(runners, bands):
            mine = {rn.local_index(g): v for g, v in kicks.items() if rn.owns(g)}
            b.set_state_sparse(list(mine), list(mine.values()))
            own = [i for i, g in enumerate(probes) if rn.owns(g)]
            b.set_probes([rn.local_index(probes[i]) for i in own], steps + 1)
            b.own = own
        for k in [0] + pass_lengths(steps, kmax):
            for b in bands:
                b.pass_local(k, coeff, damping)
            total = torch.stack([b.exchange for b in bands]).sum(0)  # what the all-reduce does
            for b in bands:
                b.exchange.copy_(total)
                b.pass_commit()
        got = np.concatenate([b.get_state() for b in bands])
        series = np.zeros((steps + 1, len(probes)))
        for b in bands:
            if b.own:
                series[:, b.own] = b.read_probes(steps + 1)
    finally:
        for b in bands:
            b.close()
    ref = _oracle(S, kicks, steps, coeff, damping, probes)
    assert rel_inf(got, ref["u"]) < 1e-12
    assert rel_inf(series, ref["probes"]) < 1e-12


def test_one_band_covering_the_sheet_equals_the_ordinary_plan():
    S, steps = 96, 64
    spec = tk.LatticeSpec(size=S, dim=2)
    kick = spec.index(spec.kick_node())
    coeff = 0.8 / (2 * S + 1)
    plan = tk.compile_lattice(spec)
    try:
        plan.set_state_sparse([kick], [1.0])
        a = plan.run(steps, coeff, 0.0, probes=[kick])
        ua = plan.get_state()
    finally:
        plan.close()
    out = tk.sheet.run_sheet_sharded(S, steps, {kick: 1.0}, c=1.0, dt=np.sqrt(coeff), probes=[kick], device=0)
    try:
        ub = out["backend"].get_state()
    finally:
        out["backend"].close()
    assert rel_inf(ub, ua) < 1e-13 and rel_inf(out["probes"], a["probes"]) < 1e-13


def test_band_plans_refuse_the_single_plan_run_call():
    b = CudaSheetBackend(32, 0, 16, device=0)
    try:
        with pytest.raises(ValueError, match="row-band"):
            b.plan.run(4, 0.01, 0.0)
    finally:
        b.close()
    with pytest.raises(ValueError):
        CudaSheetBackend(32, 16, 16, device=0)


SIZE, STEPS, DAMPING, KMAX = 200, 61, 0.004, 16


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        kicks, probes = _kicks_probes(SIZE)
        res = tk.sheet.run_sheet_sharded(SIZE, STEPS, kicks, c=1.0, dt=0.04, damping=DAMPING, probes=probes,
                                         dist=dist, rank=rank, world=world, device=rank, kmax=KMAX)
        torch.cuda.synchronize()
        np.save(os.path.join(out, f"band{rank}.npy"), res["backend"].get_state())
        np.save(os.path.join(out, f"probes{rank}.npy"), res["probes"])
        dist.barrier()
        res["backend"].close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_sheet_run_matches_the_oracle(world, tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"band{r}.npy") for r in range(world)])
    series = sum(np.load(tmp_path / f"probes{r}.npy") for r in range(world))
    kicks, probes = _kicks_probes(SIZE)
    ref = _oracle(SIZE, kicks, STEPS, 0.04 ** 2, DAMPING, probes)
    assert rel_inf(got, ref["u"]) < 1e-12
    assert rel_inf(series, ref["probes"]) < 1e-12
