# usage: bash tools/variants_pyr.sh -- pyramid-build time and whole-step value of every library under rgbd360_b200/variants
mkdir -p gpurun_out
for f in rgbd360_b200/variants/*.so; do
  R360_LIB=$PWD/$f python bench.py --steps 3 --warmup 3 --pairs ${PAIRS:-512} --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-40s value %.0f pairs/s ms/step %.2f pyramids %.2f ms  k_pass frac %.3f' % ('$f', d['value'], d['ms_per_step'], d['pyramid_ms_per_step'], r['frac']))"
done | tee gpurun_out/variants_pyr.txt
