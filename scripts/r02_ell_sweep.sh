#!/bin/bash
# Generic graph kernel: index-prefetch pipelining on / off (TK_GRAPH_PIPE), explicit-CSR workloads of bench.py.
set -u
O=gpurun_out; mkdir -p $O; : > $O/r02_ell_sweep.jsonl
for wl in ell-tetrakis-64x16 ell-grid-256; do
  for pipe in 0 1; do
    line=$(TK_GRAPH_PIPE=$pipe timeout 200 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu 2>> $O/r02_ell_sweep.err)
    echo "{\"pipe\": $pipe, \"line\": ${line:-null}}" >> $O/r02_ell_sweep.jsonl
    echo "$wl pipe=$pipe $(echo "$line" | python -c 'import json,sys; d=json.load(sys.stdin); print("%.2f GNUPS frac %.3f" % (d["value"], d["roofline"]["frac"]))' 2>/dev/null)"
  done
done
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
