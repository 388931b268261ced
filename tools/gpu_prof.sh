# ncu evidence: launch list (durations) + one full capture of the fused pass kernel at level 0
set -x
mkdir -p gpurun_out
PAIRS=${PAIRS:-64}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --pairs $PAIRS --no-cpu-baseline > gpurun_out/b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pass -s 33 -c 2 -f -o gpurun_out/prof_pass \
    python bench.py --steps 1 --warmup 0 --pairs $PAIRS --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
