# usage: bash tools/gpu_ab_env.sh VAR v1 v2 ... -- A/B of one environment switch of the library, same box, alternating, 2 rounds
VAR=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
for rep in 1 2 3; do
for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-extra-configs --no-cpu-baseline --no-copy-ceiling > gpurun_out/ab_env.json 2> gpurun_out/ab_env.err
  python - "$VAR=$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab_env.json'))
print("%-24s value %8.1f  ms/step %.2f  frac %.4f  pass_ms/launch %.4f  pyr %.2f ms  clocks %s verify %s" % (sys.argv[1], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['pyramid_ms_per_step'], d['clocks']['sm_mhz'], d['verify']['ok']))
PY
done
done 2>&1 | tee gpurun_out/ab_env_$VAR.txt
