# tests + one bench line (value, pyramids, k_pass frac)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_quick.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']
print('value %.0f pairs/s e2e %.0f ms/step %.2f pyramids %.2f ms k_pass frac %.3f share %.3f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['pyramid_ms_per_step'], r['frac'], r['kernel_share_of_step'], d['gpu_launches']))
PY
