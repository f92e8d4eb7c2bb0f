"""GPU: the CUDA plan's slab mode (z_begin/z_end, halo planes, phase API) on ONE device -- several
slab plans stepped in lockstep in one process with device-to-device halo copies (no kernels wait on
one another), compared with the single-domain CUDA plan and the oracle."""

import warnings

import numpy as np
import pytest

import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo
from tests.helpers import rel_inf

pytestmark = pytest.mark.gpu

CASES = {
    "blackhole_across_slabs": dict(size=40, dim=3, layers=9, defect="blackhole", radius=3.7),
    "sing_prune_across_slabs": dict(size=33, dim=3, layers=7, defect="singularity", radius=1.5,
                                    mass=2000.0, potential=0.5, prune_edges=True),
    "wedge": dict(size=17, dim=3, layers=5, defect="wedge"),
}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("world", [2, 3])
def test_fused_halo_push_in_lockstep_matches_the_oracle(case, world):
    """The peer-store path on one device: neighbouring plans attach each other's buffers (raw pointers
    instead of cudaIpc), the step kernel pushes the boundary planes, flags order the streams."""
    import torch

    from tetrakis_sim_b200.slab import CudaSlabBackend

    spec = tk.LatticeSpec(**CASES[case])
    steps, coeff, damping = 15, 0.004, 0.01
    kicks = {spec.index(spec.kick_node()): 1.0, spec.index((0, 0, 0, "B")): -0.5,
             spec.index((spec.size - 1, 1, spec.layers - 1, "D")): 0.25}
    slabs = [spec.slab(r, world) for r in range(world)]
    backs = [CudaSlabBackend(spec, zb, ze, device=0) for zb, ze in slabs]
    streams = [torch.cuda.Stream() for _ in backs]
    try:
        for i, b in enumerate(backs):
            b.plan.set_stream(streams[i].cuda_stream)
            if i > 0:
                b.plan.peer_attach(0, *backs[i - 1].plan.state_ptrs(), slabs[i - 1][1] - slabs[i - 1][0])
            if i + 1 < world:
                b.plan.peer_attach(1, *backs[i + 1].plan.state_ptrs(), slabs[i + 1][1] - slabs[i + 1][0])
        for (zb, ze), b in zip(slabs, backs):
            mine = {g - zb * spec.plane: v for g, v in kicks.items() if zb <= g // spec.plane < ze}
            b.set_state_sparse(list(mine), list(mine.values()))
            b.step_begin(coeff, damping)
        torch.cuda.synchronize()
        count = 1
        for b in backs:
            b.push_halos()
            b.signal(count)
        for _ in range(steps):
            for (zb, ze), b in zip(slabs, backs):
                b.wait(count)
                b.step_layers(0, ze - zb)  # one launch per slab: it pushes both boundary planes
            count += 1
            for b in backs:
                b.signal(count)
                b.step_end()
        torch.cuda.synchronize()
        got = np.concatenate([b.get_state() for b in backs])
    finally:
        for b in backs:
            b.plan.close()
    flags = np.full(spec.num_slots, 3, dtype=np.uint8)
    im, pot = np.ones(spec.num_slots), np.zeros(spec.num_slots)
    si, sf, sm, sv = spec.specials()
    flags[si], im[si], pot[si] = sf, sm, sv
    u0 = np.zeros(spec.num_slots)
    for g, v in kicks.items():
        u0[g] = v
    ref = wo.run_structured(spec.size, spec.layers, flags, im, pot, spec.corrections(), u0, steps,
                            coeff, damping)["u"]
    assert rel_inf(got, ref) < 1e-12


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("world", [2, 3])
def test_slab_plans_in_lockstep_match_single_domain(case, world):
    import torch

    from tetrakis_sim_b200.slab import CudaSlabBackend

    spec = tk.LatticeSpec(**CASES[case])
    steps, coeff, damping = 15, 0.004, 0.01
    kicks = {spec.index(spec.kick_node()): 1.0, spec.index((0, 0, 0, "B")): -0.5,
             spec.index((spec.size - 1, 1, spec.layers - 1, "D")): 0.25}
    slabs = [spec.slab(r, world) for r in range(world)]
    backs = [CudaSlabBackend(spec, zb, ze, device=0) for zb, ze in slabs]
    try:
        assert max(b.max_degree for b in backs) == tk.compile_lattice(spec).max_degree
        for (zb, ze), b in zip(slabs, backs):
            mine = {g - zb * spec.plane: v for g, v in kicks.items() if zb <= g // spec.plane < ze}
            b.set_state_sparse(list(mine), list(mine.values()))
            b.step_begin(coeff, damping)

        def exchange(next_state):
            pl = [b.planes(next_state) for b in backs]
            for lo, hi in zip(pl, pl[1:]):
                hi["recv_lo"].copy_(lo["send_hi"])
                lo["recv_hi"].copy_(hi["send_lo"])

        exchange(False)
        for _ in range(steps):
            for (zb, ze), b in zip(slabs, backs):
                L = ze - zb
                b.step_layers(0, 1)
                if L > 1:
                    b.step_layers(L - 1, L)
            exchange(True)
            for (zb, ze), b in zip(slabs, backs):
                if ze - zb > 2:
                    b.step_layers(1, ze - zb - 1)
                b.step_end()
        torch.cuda.synchronize()
        got = np.concatenate([b.get_state() for b in backs])
    finally:
        for b in backs:
            b.plan.close()
    flags = np.full(spec.num_slots, 3, dtype=np.uint8)
    im, pot = np.ones(spec.num_slots), np.zeros(spec.num_slots)
    si, sf, sm, sv = spec.specials()
    flags[si], im[si], pot[si] = sf, sm, sv
    u0 = np.zeros(spec.num_slots)
    for g, v in kicks.items():
        u0[g] = v
    ref = wo.run_structured(spec.size, spec.layers, flags, im, pot, spec.corrections(), u0, steps,
                            coeff, damping)["u"]
    assert rel_inf(got, ref) < 1e-12
    single = tk.compile_lattice(spec)
    try:
        single.set_state_sparse(list(kicks), list(kicks.values()))
        single.run(steps, coeff, damping)
        assert rel_inf(single.get_state(), ref) < 1e-12
    finally:
        single.close()
