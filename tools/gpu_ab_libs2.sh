# usage: bash tools/gpu_ab_libs2.sh lib1 lib2 ... -- A/B of library builds (R360_LIB), alternating, 2 rounds, no test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
for rep in 1 2; do
for lib in "$@"; do
  env R360_LIB=$PWD/$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-extra-configs --no-cpu-baseline --no-copy-ceiling > gpurun_out/ab_lib.json 2> gpurun_out/ab_lib.err
  python - "$lib" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab_lib.json'))
print("%-44s value %8.1f  ms/step %.2f  frac %.4f  pass_ms/launch %.4f  pyr %.2f ms  clocks %s verify %s" % (sys.argv[1], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['pyramid_ms_per_step'], d['clocks']['sm_mhz'], d['verify']['ok']))
PY
done
done 2>&1 | tee gpurun_out/ab_lib2.txt
