# usage: bash tools/gpu_chunk.sh -- occlusion tests + bench (active-only head reset), pyramid chunk-size variants, occlusion launch list
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_occlusion.py -m gpu -x -q 2>&1 | tail -1
: > gpurun_out/occ_reset.txt
for occ in 1 2; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --occlusion $occ 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('occ $occ value %.0f pairs/s ms/step %.2f pyr %.2f e2e %.0f launches %d ok %d mhz %s' % (d['value'], d['ms_per_step'], d['pyramid_ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['config']['pairs_ok'], d['clocks']['sm_mhz']))" | tee -a gpurun_out/occ_reset.txt
done
bash tools/variants_pyr.sh
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_occ1_pipe.csv \
    python bench.py --steps 1 --warmup 0 --pairs 64 --no-cpu-baseline --occlusion 1 > gpurun_out/b_occ_ncu.log 2>&1
