# usage: bash tools/gpu_ab_mix.sh -- CTA-synchronous build vs barrier-free build at several dynamic shares, alternating, GPU suite first
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
run() {
  env R360_LIB=$PWD/$1 R360_DYN_PERMILLE=$2 timeout 300 python bench.py --steps 8 --warmup 3 --no-extra-configs --no-cpu-baseline --no-copy-ceiling > gpurun_out/ab_mix.json 2> gpurun_out/ab_mix.err
  python - "$1 dyn=$2" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab_mix.json'))
print("%-52s value %8.1f  ms/step %.2f  frac %.4f  pass_ms/launch %.4f  clocks %s verify %s" % (sys.argv[1], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks']['sm_mhz'], d['verify']['ok']))
PY
}
for rep in 1 2; do
  run rgbd360_b200/variants/lib_cta_sync.so 150
  run rgbd360_b200/variants/lib_nobarrier.so 150
  run rgbd360_b200/variants/lib_nobarrier.so 80
  run rgbd360_b200/variants/lib_nobarrier.so 30
  run rgbd360_b200/variants/lib_nobarrier.so 250
done 2>&1 | tee gpurun_out/ab_mix.txt
