set -x
timeout 600 python -m pytest tests/test_gpu_sheet.py -q -x -k multi_gpu 2>&1 | tail -5
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29521 bench.py --gpus 2 --workload cfg2-sharded --steps 480 --warmup 48 2>/dev/null | tail -1 > gpurun_out/sheet_n2_weak.json
$T --master-port 29522 bench.py --gpus 2 --workload cfg2-sharded --size 8192 --steps 480 --warmup 48 2>/dev/null | tail -1 > gpurun_out/sheet_n2_strong.json
python bench.py --workload cfg2-sharded --steps 480 --warmup 48 2>/dev/null | tail -1 > gpurun_out/sheet_n1.json
