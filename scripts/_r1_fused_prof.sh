set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 400 --warmup 16 > gpurun_out/bench_cfg2_fused.json 2> gpurun_out/bench_cfg2_fused.err; tail -c 600 gpurun_out/bench_cfg2_fused.json
TK_LAT_FUSE=0 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_cfg2_unfused.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg2_fused.csv python bench.py --steps 64 --warmup 16 --no-cpu > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lat_fused2d -s 2 -c 1 -o gpurun_out/prof_cfg2_fused -f python bench.py --steps 32 --warmup 16 --no-cpu > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/
