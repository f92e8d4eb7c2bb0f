"""Debug aid: two slab plans on ONE GPU in lockstep with the in-kernel early arrival signal."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tetrakis_sim_b200 as tk
from tetrakis_sim_b200.slab import CudaSlabBackend

def run(spec, world, steps, early):
    slabs = [spec.slab(r, world) for r in range(world)]
    backs = [CudaSlabBackend(spec, zb, ze, device=0) for zb, ze in slabs]
    streams = [torch.cuda.Stream() for _ in backs]
    for i, b in enumerate(backs):
        b.plan.set_stream(streams[i].cuda_stream)
        if i > 0:
            b.plan.peer_attach(0, *backs[i - 1].plan.state_ptrs(), slabs[i - 1][1] - slabs[i - 1][0])
        if i + 1 < world:
            b.plan.peer_attach(1, *backs[i + 1].plan.state_ptrs(), slabs[i + 1][1] - slabs[i + 1][0])
    kick = spec.index(spec.kick_node())
    for (zb, ze), b in zip(slabs, backs):
        mine = [kick - zb * spec.plane] if zb <= kick // spec.plane < ze else []
        b.set_state_sparse(mine, [1.0] * len(mine))
        b.step_begin(0.003, 0.0)
    torch.cuda.synchronize()
    count = 1
    for b in backs:
        b.push_halos()
        b.signal(count)
    for s in range(steps):
        for (zb, ze), b in zip(slabs, backs):
            b.wait(count)
            if early:
                b.arm_signal(count + 1)
            b.step_layers(0, ze - zb)
        count += 1
        for b in backs:
            if not early or b.signal_pending():
                b.signal(count)
            b.step_end()
        torch.cuda.synchronize()
        print("step", s, "done", flush=True)
    got = np.concatenate([b.get_state() for b in backs])
    for b in backs:
        b.plan.close()
    return got

for spec, world in ((tk.LatticeSpec(size=48, dim=3, layers=13, defect="blackhole", radius=4.2), 2),
                    (tk.LatticeSpec(size=200, dim=3, layers=52, defect="blackhole", radius=9.0), 2)):
    a = run(spec, world, 4, False)
    print("stream-signal ok", a.sum(), flush=True)
    b = run(spec, world, 4, True)
    print("early-signal ok", b.sum(), "identical", np.array_equal(a, b), flush=True)
