# A/B of the speculative error-only passes: tests, then bench both ways
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 0 1; do
  R360_SPECULATE=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_spec$v.json 2> gpurun_out/bench_spec$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_spec$v.json'))
print('SPECULATE=$v value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'share',round(d['roofline']['kernel_share_of_step'],3),'ms/step',round(d['ms_per_step'],2),'launches',d['gpu_launches'],'passes',[round(x,3) for x in d['config']['mean_passes_per_level']],'iters',[round(x,3) for x in d['config']['mean_accepted_iters_per_level']], d['clocks']['sm_mhz'])
PY
done
