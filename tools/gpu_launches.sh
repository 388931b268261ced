set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --pairs ${PAIRS:-64} --no-cpu-baseline > gpurun_out/b_ncu1.log 2>&1
