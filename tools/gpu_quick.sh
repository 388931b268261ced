set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value',d['value'],'e2e',d['e2e'],'roofline',d['roofline']['achieved'],d['roofline']['frac'],'share',d['roofline']['kernel_share_of_step'],'ms/step',d['ms_per_step'],'clocks',d['clocks'])
PY
