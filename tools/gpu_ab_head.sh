# A/B of the fused pyramid head (k_pyr_head) against the separate level-0 kernels: tests, then bench both ways
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 0 1; do
  R360_PYR_HEAD=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_head$v.json 2> gpurun_out/bench_head$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_head$v.json'))
print('HEAD=$v value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'share',round(d['roofline']['kernel_share_of_step'],3),'ms/step',round(d['ms_per_step'],2),'launches',d['gpu_launches'])
PY
done
