# usage: bash tools/gpu_ab.sh -- A/B of every library under rgbd360_b200/variants (+ R360_SPEC_MARGIN sweep on $MARGIN_LIB)
mkdir -p gpurun_out
: > gpurun_out/ab.txt
run() {   # label, lib, extra env
  env R360_LIB=$PWD/$2 $3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-34s value %.0f pairs/s ms/step %.2f pyr %.2f  k_pass frac %.4f avg_launch %.3f ms passes %s iters %s ok %d mhz %s %s' % ('$1', d['value'], d['ms_per_step'], d['pyramid_ms_per_step'], r['frac'], r['avg_launch_ms'], [round(x,3) for x in d['config']['mean_passes_per_level']], [round(x,3) for x in d['config']['mean_accepted_iters_per_level']], d['config']['pairs_ok'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" | tee -a gpurun_out/ab.txt
}
for f in rgbd360_b200/variants/*.so; do run $(basename $f .so) $f ""; done
if [ -n "$MARGIN_LIB" ]; then
  for m in $MARGINS; do run "margin=$m" $MARGIN_LIB "R360_SPEC_MARGIN=$m"; done
fi
run "base again" rgbd360_b200/variants/base.so ""
