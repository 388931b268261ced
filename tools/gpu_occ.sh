# usage: bash tools/gpu_occ.sh -- occlusion variants: GPU tests, then A/B of the pipelined k_occ_eval against the simple one (R360_OCC_PIPE=0)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_occlusion.py -m gpu -x -q 2>&1 | tail -3
R360_OCC_PIPE=0 timeout 300 python -m pytest tests/test_occlusion.py -m gpu -x -q 2>&1 | tail -1
: > gpurun_out/occ_ab.txt
for occ in 1 2; do
  for pipe in 0 1; do
    R360_OCC_PIPE=$pipe timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --occlusion $occ 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('occ $occ pipe $pipe value %.0f pairs/s ms/step %.2f pyr %.2f e2e %.0f passes %s iters %s ok %d mhz %s' % (d['value'], d['ms_per_step'], d['pyramid_ms_per_step'], d['e2e']['value'], [round(x,3) for x in d['config']['mean_passes_per_level']], [round(x,3) for x in d['config']['mean_accepted_iters_per_level']], d['config']['pairs_ok'], d['clocks']['sm_mhz']))" | tee -a gpurun_out/occ_ab.txt
  done
done
