#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun): GPU test suite, smoke, the bench lines, ncu launch lists and
# `ncu --set full` captures of the dominant kernels.  Everything lands in gpurun_out/ and is summarised into
# profiles/ afterwards (scripts/ncu_summary.py, scripts/launch_table.py).
#   gpurun --timeout 1500 -- 'bash scripts/r02_gpu_evidence.sh'
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02_box.txt 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv >> $O/r02_box.txt 2>&1

step() { echo "=== $1 ($(date +%T))"; }

if [ "${TESTS:-1}" = 1 ]; then
step tests
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gputests.log 2>&1
echo "gpu tests rc=$?" | tee -a $O/r02_gputests.log
tail -3 $O/r02_gputests.log
fi

step smoke
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/r02_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/r02_smoke.log

step bench_default
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2> $O/r02_bench_default.err
echo "bench default rc=$?"
step bench_reference
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
echo "bench reference rc=$?"
step bench_cfg2_200
timeout 300 python bench.py --steps 200 --warmup 10 --no-extras --no-cpu > $O/r02_bench_cfg2_200.json 2> $O/r02_bench_cfg2_200.err
step bench_cfg2_960
timeout 300 python bench.py --steps 960 --warmup 48 --no-extras --no-cpu > $O/r02_bench_cfg2_960.json 2> $O/r02_bench_cfg2_960.err
step bench_cfg4
timeout 300 python bench.py --workload cfg4-3d-512 --steps 64 --warmup 16 --no-cpu > $O/r02_bench_cfg4.json 2> $O/r02_bench_cfg4.err
step bench_cfg3
timeout 300 python bench.py --workload cfg3-slab-1024x64 --steps 64 --warmup 16 --no-cpu > $O/r02_bench_cfg3_fused.json 2> $O/r02_bench_cfg3_fused.err
TK_LAT_FUSE=0 timeout 300 python bench.py --workload cfg3-slab-1024x64 --steps 20 --warmup 5 --no-cpu > $O/r02_bench_cfg3_unfused.json 2> $O/r02_bench_cfg3_unfused.err
step bench_cfg5
timeout 300 python bench.py --workload cfg5-ensemble --steps 5 --warmup 2 --no-cpu > $O/r02_bench_cfg5.json 2> $O/r02_bench_cfg5.err

NCU="ncu --clock-control none"
step launches_default
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches_cfg2.csv \
    python bench.py --steps 20 --warmup 5 --no-extras --no-cpu > $O/r02_launches_cfg2.out 2>&1
step launches_cfg4
timeout 600 $NCU --metrics gpu__time_duration.sum -c 320 --csv --log-file $O/r02_launches_cfg4.csv \
    python bench.py --workload cfg4-3d-512 --steps 64 --warmup 16 --no-cpu > $O/r02_launches_cfg4.out 2>&1

full() {  # name, kernel regex, launches to skip, command...
    local name=$1 k=$2 skip=$3
    shift 3
    step "ncu_$name"
    timeout 600 $NCU --set full --import-source on -k "regex:$k" -s "$skip" -c 1 -f -o $O/r02_ncu_$name "$@" \
        > $O/r02_ncu_$name.out 2>&1
    echo "ncu $name rc=$?"
    # gpurun brings back at most 64 MiB and a report with sources is ~20 MB: keep the summaries, drop the report
    python scripts/ncu_summary.py $O/r02_ncu_$name.ncu-rep > $O/r02_ncu_$name.md 2>> $O/r02_ncu_$name.out
    ncu -i $O/r02_ncu_$name.ncu-rep --page raw --csv > $O/r02_ncu_$name.raw.csv 2>> $O/r02_ncu_$name.out
    if [ "${SRC:-0}" = 1 ]; then
        ncu -i $O/r02_ncu_$name.ncu-rep --page source --csv > $O/r02_ncu_$name.source.csv 2>> $O/r02_ncu_$name.out
    fi
    rm -f $O/r02_ncu_$name.ncu-rep
}
full cfg2_fused_k20 lat_fused2d_kernel 2 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu
SRC=1 full cfg4_bulk3d fx_bulk3d_kernel 1 python bench.py --workload cfg4-3d-512 --steps 64 --warmup 16 --no-cpu
full cfg4_cross_step fx_cross_step_kernel 40 python bench.py --workload cfg4-3d-512 --steps 64 --warmup 16 --no-cpu
full cfg3_bulk3d fx_bulk3d_kernel 1 python bench.py --workload cfg3-slab-1024x64 --steps 64 --warmup 16 --no-cpu
SRC=1 full cfg3_cross_step fx_cross_step_kernel 40 python bench.py --workload cfg3-slab-1024x64 --steps 64 --warmup 16 --no-cpu
SRC=1 TK_LAT_FUSE=0 full cfg3_step lat_step_kernel 6 python bench.py --workload cfg3-slab-1024x64 --steps 20 --warmup 5 --no-cpu
full cfg5_ensemble ensemble_kernel 1 python bench.py --workload cfg5-ensemble --steps 1 --warmup 1 --no-cpu
du -sh $O; ls -la $O
step done
