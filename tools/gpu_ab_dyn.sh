# usage: bash tools/gpu_ab_dyn.sh -- A/B of the dynamic share of k_pass (R360_DYN_PERMILLE), same box, same command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do
for d in 0 100 150 250 400; do
  R360_DYN_PERMILLE=$d timeout 300 python bench.py --steps 8 --warmup 3 --no-extra-configs --no-cpu-baseline --no-copy-ceiling > gpurun_out/ab_dyn_$d.json 2> gpurun_out/ab_dyn_$d.err
  python - "$d" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab_dyn_%s.json'%sys.argv[1]))
print("dyn %4s  value %8.1f  frac %.4f  pass_ms/launch %.4f  pyr %.2f ms  clocks %s verify %s" % (sys.argv[1], d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['pyramid_ms_per_step'], d['clocks']['sm_mhz'], d['verify']['ok']))
PY
done
done 2>&1 | tee gpurun_out/ab_dyn.txt
# SM activity balance of the level-0 launches, static vs dynamic
for d in 0 150; do
R360_DYN_PERMILLE=$d timeout 300 ncu --metrics sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.min,sm__cycles_elapsed.max,gpu__time_duration.sum --clock-control none -k regex:k_pass -s 81 -c 5 --csv --log-file gpurun_out/balance_dyn_$d.csv python bench.py --one-step > gpurun_out/b_ncu_bal.log 2>&1
done
