"""Debug aid: 2 ranks, p2p transport with the in-kernel early signal, stack dump if it stalls."""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch, torch.distributed as dist
import tetrakis_sim_b200 as tk
from tetrakis_sim_b200.slab import CudaSlabBackend, SlabRunner

faulthandler.dump_traceback_later(40, exit=True)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
size, layers = int(sys.argv[1]), int(sys.argv[2])
spec = tk.LatticeSpec(size=size, dim=3, layers=layers, defect="blackhole", radius=4.2 if size < 512 else 64.0)
dev = torch.device(f"cuda:{rank}")
legacy = len(sys.argv) > 3 and sys.argv[3] == "legacy"
sync_every = int(sys.argv[4]) if len(sys.argv) > 4 else 5
main = torch.cuda.current_stream(dev) if legacy else torch.cuda.Stream(dev)
torch.cuda.set_stream(main)
backend = CudaSlabBackend(spec, *spec.slab(rank, world), device=rank)
runner = SlabRunner(spec, backend, rank, world, dist, comm_stream=torch.cuda.Stream(dev, priority=-1), main_stream=main, transport="p2p")
kick = spec.index(spec.kick_node())
backend.set_state_sparse(*(([runner.local_index(kick)], [1.0]) if runner.owns(kick) else ([], [])))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
for rep in range(reps - 1):
    runner.begin(0.003, 0.0)
    for s in range(10):
        runner.step()
    runner.finish()
    torch.cuda.synchronize()
    dist.barrier()
    print(f"rank {rank} rep {rep} done", flush=True)
runner.begin(0.003, 0.0)
nsteps = int(sys.argv[6]) if len(sys.argv) > 6 else 40
jit = len(sys.argv) > 7 and sys.argv[7] == "jitter"
rng = np.random.default_rng(1234 + 17 * rank)
for s in range(nsteps):
    if jit and rng.random() < 0.3:
        time.sleep(float(rng.random()) * 1e-3)
    runner.step()
    if s % sync_every == sync_every - 1:
        torch.cuda.synchronize()
        flags = torch.as_tensor(tk.slab._RawCuda(backend.plan.halo_ptrs(False)["recv_lo"], 1), device=dev)
        print(f"rank {rank} step {s} synced, early={runner.early_signal}", flush=True)
runner.finish()
torch.cuda.synchronize()
print(f"rank {rank} done, sum {backend.get_state().sum():.6f}", flush=True)
dist.barrier()
backend.detach_peers()
backend.plan.close()
dist.destroy_process_group()
