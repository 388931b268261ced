# usage: bash tools/variants_bench.sh  -- quick k_pass roofline of every library under rgbd360_b200/variants
mkdir -p gpurun_out
for f in rgbd360_b200/variants/*.so; do
  R360_LIB=$PWD/$f python bench.py --steps 2 --warmup 3 --pairs ${PAIRS:-256} --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-40s value %.0f pairs/s  k_pass %.0f GB/s frac %.3f avg_launch %.3f ms share %.2f' % ('$f', d['value'], r['achieved'], r['frac'], r['avg_launch_ms'], r['kernel_share_of_step']))"
done | tee gpurun_out/variants.txt
