# usage: bash tools/gpu_ab_head2.sh libA.so libB.so ... -- A/B of pyramid-head variants (R360_LIB): for every library first the
# plane / parity tests THROUGH that library, then the bench alternating, 3 rounds
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
for lib in "$@"; do
  echo "== parity through $lib"
  env R360_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_random.py -m gpu -q -x 2>&1 | tail -2
done
for rep in 1 2 3; do
for lib in "$@"; do
  env R360_LIB=$PWD/$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-extra-configs --no-cpu-baseline --no-copy-ceiling > gpurun_out/ab_lib.json 2> gpurun_out/ab_lib.err
  python - "$lib" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab_lib.json'))
print("%-44s value %8.1f  ms/step %.2f  frac %.4f  pass_ms/launch %.4f  pyr %.2f ms  clocks %s verify %s" % (sys.argv[1], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['pyramid_ms_per_step'], d['clocks']['sm_mhz'], d['verify']['ok']))
PY
done
done 2>&1 | tee gpurun_out/ab_head2.txt
