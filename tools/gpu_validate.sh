# usage: bash tools/gpu_validate.sh -- full GPU suite + smoke + one bench line + compute-sanitizer over the occlusion kernels and smoke()
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_validate.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_validate.json')); r=d['roofline']
print('value %.0f pairs/s e2e %.0f ms/step %.2f pyramids %.2f ms k_pass frac %.3f share %.3f launches %d clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['pyramid_ms_per_step'], r['frac'], r['kernel_share_of_step'], d['gpu_launches'], d['clocks']))
PY
{
for tool in memcheck racecheck initcheck; do
  echo "== $tool: occlusion kernels (k_occ_scatter, pipelined k_occ_eval)"
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_occlusion.py -m gpu -x -q \
    -k "(occlusion_evaluations and holes) or (occlusion_align and loop)" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" | head -6
done
for tool in memcheck racecheck; do
  echo "== $tool: smoke() (fused + error-only passes)"
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|smoke ok|Error|RACECHECK SUMMARY|hazard" | head -6
done
} 2>&1 | tee gpurun_out/sanitizer_v8.txt
