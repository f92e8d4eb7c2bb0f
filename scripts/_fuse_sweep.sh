set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --no-cpu 2>&1 | tail -1
python bench.py --steps 1000 --warmup 48 --no-cpu 2>&1 | tail -1
TK_LAT_FUSE=32 python bench.py --steps 1024 --warmup 32 --no-cpu 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_cfg2_fused.csv python bench.py --steps 192 --warmup 48 --no-cpu > gpurun_out/ncu_l.log 2>&1
