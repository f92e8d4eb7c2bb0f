"""CPU, world_size 2 over gloo: the layer-slab orchestration (slab ownership, which planes go to
which neighbour, ordering of boundary / exchange / interior / swap) reproduces the single-domain
oracle.  The arithmetic here is NumPy (tests/slab_numpy_backend.py); on the GPU the same SlabRunner
drives the CUDA plan (tests/test_gpu_slab.py)."""

import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo
from tests.slab_numpy_backend import NumpySlabBackend
from tetrakis_sim_b200.slab import SlabRunner, global_max_degree

CASES = {
    "sing_prune_across_boundary": dict(size=6, dim=3, layers=5, defect="singularity", radius=1.5,
                                       mass=50.0, potential=2.0, prune_edges=True),
    "blackhole": dict(size=7, dim=3, layers=6, defect="blackhole", radius=2.5),
    "wedge": dict(size=5, dim=3, layers=4, defect="wedge"),
}
STEPS, COEFF, DAMPING = 12, 0.02, 0.01


def _worker(rank, world, port, case, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = tk.LatticeSpec(**CASES[case])
        zb, ze = spec.slab(rank, world)
        backend = NumpySlabBackend(spec, zb, ze)
        runner = SlabRunner(spec, backend, rank, world, dist)
        assert global_max_degree(10 + rank, dist) == 10 + world - 1
        kicks = {spec.index(spec.kick_node()): 1.0, spec.index((0, 0, 0, "B")): -0.5,
                 spec.index((spec.size - 1, 1, spec.layers - 1, "D")): 0.25}
        mine = {runner.local_index(g): v for g, v in kicks.items() if runner.owns(g)}
        backend.set_state_sparse(list(mine), list(mine.values()))
        runner.begin(COEFF, DAMPING)
        for _ in range(STEPS):
            runner.step()
        runner.finish()
        np.save(os.path.join(out, f"slab{rank}.npy"), backend.get_state())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("world", [2, 3])
def test_slab_runner_matches_single_domain_oracle(case, world, tmp_path):
    spec = tk.LatticeSpec(**CASES[case])
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)])
    flags = np.full(spec.num_slots, 3, dtype=np.uint8)
    im, pot = np.ones(spec.num_slots), np.zeros(spec.num_slots)
    si, sf, sm, sv = spec.specials()
    flags[si], im[si], pot[si] = sf, sm, sv
    u0 = np.zeros(spec.num_slots)
    u0[spec.index(spec.kick_node())] = 1.0
    u0[spec.index((0, 0, 0, "B"))] = -0.5
    u0[spec.index((spec.size - 1, 1, spec.layers - 1, "D"))] = 0.25
    ref = wo.run_structured(spec.size, spec.layers, flags, im, pot, spec.corrections(), u0, STEPS,
                            COEFF, DAMPING)["u"]
    assert np.max(np.abs(ref)) > 0
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))


def test_slab_ownership_partitions_the_layers():
    spec = tk.LatticeSpec(size=4, dim=3, layers=11)
    for world in (1, 2, 3, 4, 8):
        slabs = [spec.slab(r, world) for r in range(world)]
        assert slabs[0][0] == 0 and slabs[-1][1] == 11
        assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
        assert max(e - b for b, e in slabs) - min(e - b for b, e in slabs) <= 1
    with pytest.raises(ValueError):
        SlabRunner(tk.LatticeSpec(size=4, dim=2), None, 0, 2)
