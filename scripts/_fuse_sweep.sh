set -x
timeout 1500 python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -15
for cfg in "4 3" "8 3" "16 3" "16 2" "16 4" "24 3" "32 3" "32 4" "48 3" "64 3"; do
  set -- $cfg
  echo "== cfg2 fuse K=$1 minb=$2"; TK_LAT_FUSE=$1 TK_LAT_FUSE_MINB=$2 timeout 300 python bench.py --steps 384 --warmup 16 --no-cpu 2>&1 | tail -1
done
echo "== K=16 pf sweep"
for pf in 0 2; do echo "== pf $pf"; TK_LAT_PREFETCH=$pf TK_LAT_FUSE=16 timeout 300 python bench.py --steps 384 --warmup 16 --no-cpu 2>&1 | tail -1; done
