#!/bin/bash
# Round-2 multi-GPU evidence (gpurun --gpus 2): the real NCCL / peer-memory tests, the default bench at N=2 launched as
# the driver launches it, and the tests / bench lines of the host-side changes made after the one-GPU evidence run.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/r02_gpu_multi.sh'
set -u
O=gpurun_out
mkdir -p $O
N=${N:-2}
nvidia-smi -L > $O/r02_multi_box.txt 2>&1
step() { echo "=== $1 ($(date +%T))"; }
step multi_tests
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sheet.py -m gpu -x -q > $O/r02_multi_tests.log 2>&1
echo "multi tests rc=$?" | tee -a $O/r02_multi_tests.log
tail -4 $O/r02_multi_tests.log
step new_tests
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ensemble.py tests/test_gpu_frontends.py -m gpu -x -q > $O/r02_new_tests.log 2>&1
echo "new tests rc=$?" | tee -a $O/r02_new_tests.log
tail -3 $O/r02_new_tests.log
step bench_n$N
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err
echo "bench n$N rc=$?"
step bench_ref_n$N
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/r02_bench_reference_n$N.json 2> $O/r02_bench_reference_n$N.err
echo "bench ref n$N rc=$?"
step bench_cfg5
timeout 300 python bench.py --workload cfg5-ensemble --steps 5 --warmup 2 --no-cpu > $O/r02_bench_cfg5_b.json 2> $O/r02_bench_cfg5_b.err
step bench_wave_graph_structured
timeout 300 python scripts/bench_wave.py --sizes 11,21,31,41 --dim 2 --steps 50 --repeats 5 --path graph-structured --gnups > $O/r02_bench_wave_graph_structured.csv 2>&1
timeout 300 python scripts/bench_wave.py --sizes 11,21,31,41 --dim 2 --steps 50 --repeats 5 --path graph --gnups > $O/r02_bench_wave_graph.csv 2>&1
du -sh $O
step done
