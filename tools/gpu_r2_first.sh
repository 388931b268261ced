# usage: bash tools/gpu_r2_first.sh -- first GPU call of round 2 (one B200):
# GPU suite (no -x: every failure is wanted) + smoke, the full bench line (configs 2/4/5, verify, CPU baseline),
# the ncu launch list of exactly ONE 512-pair step.
mkdir -p gpurun_out
t0=$(date +%s); stamp() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
stamp tests
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -40 > gpurun_out/gputests.txt; tail -25 gpurun_out/gputests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
stamp bench
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cut -c1-3000 gpurun_out/bench.json
stamp "launch list of one step (512 pairs)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_one_step.csv python bench.py --one-step > gpurun_out/b_ncu1.log 2>&1
tail -1 gpurun_out/b_ncu1.log
stamp done
