# usage: bash tools/gpu_r2_quick.sh -- GPU suite + one short bench line (no extra configs, no CPU baseline)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 8 --warmup 3 --no-extra-configs --no-cpu-baseline --no-copy-ceiling > gpurun_out/quick.json 2> gpurun_out/quick.err; tail -3 gpurun_out/quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick.json'))
print("value %.1f  ms/step %.2f  pyr %.2f ms  frac %.4f  e2e %.1f  clocks %s verify %s" % (d['value'], d['ms_per_step'], d['pyramid_ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']['sm_mhz'], d['verify']['ok']))
PY
