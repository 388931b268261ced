# usage: bash tools/gpu_final.sh -- evidence for the committed build in one bounded call:
# bench line (with the CPU baseline), ncu launch list, ncu full captures of the level-0 passes, then the GPU tests + smoke
mkdir -p gpurun_out
t0=$(date +%s); stamp() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
stamp bench
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
stamp launches
PAIRS=${PAIRS:-64}
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --pairs $PAIRS --no-cpu-baseline > gpurun_out/b_ncu1.log 2>&1
stamp full capture
# 64 pairs: 4 levels x (14 fused + 13 error-only) k_pass launches per step; level 0 of the first step starts at launch 81
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 81 -c 8 -f -o gpurun_out/prof_pass \
    python bench.py --steps 1 --warmup 0 --pairs $PAIRS --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
stamp tests
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
stamp smoke
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
stamp done
ls -la gpurun_out
