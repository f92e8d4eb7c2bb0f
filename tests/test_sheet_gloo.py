"""CPU, world_size 2 and 3 over gloo: ONE sheet held as row bands (tetrakis_sim_b200.sheet.SheetRunner) -- band
ownership, pass lengths, one all-reduce of column sums + totals per pass -- reproduces the single-domain oracle.
The arithmetic here is NumPy (tests/sheet_numpy_backend.py); on the GPU the same runner drives the CUDA plan
(tests/test_gpu_sheet.py)."""

import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wave_oracle as wo
from tests.sheet_numpy_backend import NumpySheetBackend
from tetrakis_sim_b200.sheet import SheetRunner, band_rows, pass_lengths

SIZE, STEPS, COEFF, DAMPING, KMAX = 11, 23, 0.03, 0.01, 6
KICKS = {((5 * SIZE + 5) * 4 + 0): 1.0, ((0 * SIZE + 10) * 4 + 3): -0.5, ((10 * SIZE + 2) * 4 + 1): 0.25}
PROBES = [(5 * SIZE + 5) * 4 + 0, (10 * SIZE + 2) * 4 + 1, (3 * SIZE + 7) * 4 + 2]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        backend = NumpySheetBackend(SIZE, *band_rows(SIZE, rank, world))
        runner = SheetRunner(SIZE, backend, rank, world, dist)
        mine = {runner.local_index(g): v for g, v in KICKS.items() if runner.owns(g)}
        backend.set_state_sparse(list(mine), list(mine.values()))
        own = [i for i, g in enumerate(PROBES) if runner.owns(g)]
        backend.set_probes([runner.local_index(PROBES[i]) for i in own], STEPS + 1)
        runner.run(STEPS, COEFF, DAMPING, KMAX)
        series = np.zeros((STEPS + 1, len(PROBES)))
        if own:
            series[:, own] = backend.read_probes(STEPS + 1)
        np.save(os.path.join(out, f"band{rank}.npy"), backend.get_state())
        np.save(os.path.join(out, f"probes{rank}.npy"), series)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sheet_runner_matches_single_domain_oracle(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"band{r}.npy") for r in range(world)])
    probes = sum(np.load(tmp_path / f"probes{r}.npy") for r in range(world))
    n = SIZE * SIZE * 4
    u0 = np.zeros(n)
    for g, v in KICKS.items():
        u0[g] = v
    ref = wo.run_structured(SIZE, 1, np.full(n, 3, np.uint8), np.ones(n), np.zeros(n), [], u0, STEPS, COEFF,
                            DAMPING, probes=PROBES)
    assert np.max(np.abs(got - ref["u"])) <= 1e-13 * np.max(np.abs(ref["u"]))
    assert np.max(np.abs(probes - ref["probes"])) <= 1e-13 * np.max(np.abs(ref["probes"]))


def test_band_ownership_and_pass_lengths():
    for size in (1, 7, 64, 8192):
        for world in (1, 2, 3, 8):
            if world > size:
                continue
            bands = [band_rows(size, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == size
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert max(e - b for b, e in bands) - min(e - b for b, e in bands) <= 1
    assert pass_lengths(200, 48) == [40] * 5
    assert pass_lengths(203, 8) == [8] * 25 + [3]
    assert pass_lengths(33, 16) == [16, 16, 1]
    assert pass_lengths(100, 48) == [40, 40, 20]
    assert pass_lengths(1, 48) == [1] and pass_lengths(0, 48) == []
    assert all(sum(pass_lengths(s, k)) == s for s in range(1, 60) for k in (1, 2, 7, 48))
