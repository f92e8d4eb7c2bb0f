set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r01_bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --steps 960 --warmup 48 --no-cpu > gpurun_out/r01_bench_cfg2_960steps.json 2>/dev/null
TK_LAT_FUSE=0 python bench.py --no-cpu > gpurun_out/r01_bench_cfg2_unfused.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches_cfg2_fused.csv python bench.py --steps 96 --warmup 48 --no-cpu > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lat_fused2d -s 2 -c 1 -o gpurun_out/prof_cfg2_fused48 -f python bench.py --steps 96 --warmup 48 --no-cpu > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
