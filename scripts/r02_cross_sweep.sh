#!/bin/bash
# Cross-stage sweeps on one B200: resident CTAs per SM of fx_cross_step_kernel (TK_FX_MINB), rows per column-band task
# (TK_FX_RCC), pass length; cfg4 (512^3, cross = 2 %) and the cfg3 slab (cross = 23.5 %).  One JSON line per run.
#   gpurun --timeout 600 -- 'bash scripts/r02_cross_sweep.sh'
set -u
O=gpurun_out
mkdir -p $O
OUT=$O/r02_cross_sweep.jsonl
: > $OUT
run() {  # label, workload, env...
    local label=$1 wl=$2
    shift 2
    local line
    line=$(env "$@" timeout 200 python bench.py --workload $wl --steps 64 --warmup 16 --no-cpu 2>> $O/r02_cross_sweep.err)
    echo "{\"label\": \"$label\", \"env\": \"$*\", \"line\": ${line:-null}}" >> $OUT
    echo "$label $wl $* -> $(echo "$line" | python -c 'import json,sys; d=json.load(sys.stdin); print("%.1f GNUPS %.4f ms/step share %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_share_of_step"]))' 2>/dev/null)"
}
for mb in 2 3 4; do
    run minb$mb cfg4-3d-512 TK_FX_MINB=$mb
    run minb$mb cfg3-slab-1024x64 TK_FX_MINB=$mb
done
for rcc in 16 32 128; do
    run rcc$rcc cfg4-3d-512 TK_FX_MINB=${BEST_MINB:-3} TK_FX_RCC=$rcc
done
run rcc16_minb4 cfg4-3d-512 TK_FX_MINB=4 TK_FX_RCC=16
run rcc16_minb2 cfg4-3d-512 TK_FX_MINB=2 TK_FX_RCC=16
step() { echo "=== $1 ($(date +%T))"; }
step tests_minb4
TK_FX_MINB=4 timeout 300 python -m pytest tests/test_gpu_cross.py -m gpu -x -q 2>&1 | tail -2
step tests_minb3
TK_FX_MINB=3 timeout 300 python -m pytest tests/test_gpu_cross.py -m gpu -x -q 2>&1 | tail -2
step done
