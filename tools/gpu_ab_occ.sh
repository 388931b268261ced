# usage: bash tools/gpu_ab_occ.sh libA.so libB.so -- occlusion tests through every library, then bench --occlusion 1 / 2 alternating
mkdir -p gpurun_out
for lib in "$@"; do echo "== occlusion tests through $lib"; env R360_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_occlusion.py -m gpu -q -x 2>&1 | tail -1; done
for rep in 1 2; do for occ in 1 2; do for lib in "$@"; do
  env R360_LIB=$PWD/$lib timeout 300 python bench.py --steps 4 --warmup 2 --no-extra-configs --no-cpu-baseline --no-copy-ceiling --occlusion $occ > gpurun_out/ab_occ.json 2> gpurun_out/ab_occ.err
  python - "$lib" $occ <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab_occ.json'))
print("%-40s occ %s value %8.1f  ms/step %.2f  clocks %s verify %s" % (sys.argv[1], sys.argv[2], d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['verify']['ok']))
PY
done; done; done 2>&1 | tee gpurun_out/ab_occ.txt
