# usage: bash tools/variants_bench_full.sh -- value / e2e / k_pass frac of the in-tree library and every variant
mkdir -p gpurun_out
for f in rgbd360_b200/librgbd360_b200.so rgbd360_b200/variants/*.so; do
  R360_LIB=$PWD/$f python bench.py --steps 3 --warmup 3 --pairs ${PAIRS:-512} --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-40s value %.0f pairs/s e2e %.0f ms/step %.2f  k_pass frac %.3f share %.2f' % ('$f', d['value'], d['e2e']['value'], d['ms_per_step'], r['frac'], r['kernel_share_of_step']))"
done | tee gpurun_out/variants_full.txt
