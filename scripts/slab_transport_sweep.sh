#!/bin/bash
# N-GPU (NGPU, default 2) comparison of the slab transports on the cfg3 slab workload; every run under its own timeout
N=${NGPU:-2}
run() {
  TK_WATCHDOG=50 TK_BENCH_PARITY=0 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload cfg3-slab-1024x64 --steps 100 --warmup 10 2>gpurun_out/sweep_err_$1.log | python -c "import sys,json; s=sys.stdin.read(); d=json.loads(s) if s.strip() else None; print('$1', 'FAILED' if d is None else (round(d['value'],1), round(d['ms_per_step'],4), round(d['timing']['min_ms']/100,4), d['single_gpu_same_workload']['ms_per_step'], d['checksum_sum_u'], d['checksum_expected']))"
}
for v in "$@"; do
  case $v in
    early) TK_SLAB_TRANSPORT=p2p TK_SLAB_EARLY=1 run early ;;
    early-noreorder) TK_SLAB_TRANSPORT=p2p TK_SLAB_EARLY=1 TK_SLAB_BND_FIRST=0 run early-noreorder ;;
    stream) TK_SLAB_TRANSPORT=p2p TK_SLAB_EARLY=0 run stream ;;
    copy) TK_SLAB_TRANSPORT=p2p-copy run copy ;;
    nccl) TK_SLAB_TRANSPORT=nccl run nccl ;;
  esac
done
