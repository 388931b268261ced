# usage: bash tools/gpu_all.sh -- tests + smoke + bench + ncu launch list + one full capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tools/index_report.py 2>&1 | tee gpurun_out/index_report.txt | tail -6
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
PAIRS=${PAIRS:-64}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --pairs $PAIRS --no-cpu-baseline > gpurun_out/b_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 33 -c 2 -f -o gpurun_out/prof_pass \
    python bench.py --steps 1 --warmup 0 --pairs $PAIRS --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
