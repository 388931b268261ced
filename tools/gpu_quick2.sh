# tests + one bench line (value, e2e, pyramids, k_pass frac)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_quick.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']
print('value %.0f pairs/s e2e %.0f (%.2f ms/step) ms/step %.2f pyramids %.2f ms k_pass frac %.3f share %.3f launches %d clocks %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'], d['pyramid_ms_per_step'], r['frac'], r['kernel_share_of_step'], d['gpu_launches'], d['clocks']))
PY
