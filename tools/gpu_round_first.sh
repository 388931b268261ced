# usage: bash tools/gpu_round_first.sh -- first GPU call of a round (one B200, about 4 minutes):
# GPU suite + smoke, the bench line with the CPU baseline, the ncu launch list of exactly ONE 512-pair step,
# full captures of the level-0 k_pass launches (both instantiations) and of k_pyr_head.
mkdir -p gpurun_out
t0=$(date +%s); stamp() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
stamp tests
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
stamp "opt-in tests written at the end of round 1 without GPU minutes (random problems replayed through the oracle)"
R360_TEST_RANDOM_GPU=1 timeout 600 python -m pytest tests/test_gpu_random.py -m gpu -q 2>&1 | tail -15
stamp bench
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
stamp "launch list of one step (512 pairs)"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_one_step.csv python bench.py --one-step > gpurun_out/b_ncu1.log 2>&1
tail -1 gpurun_out/b_ncu1.log
stamp "full captures"
# one 512-pair step: level 0 starts after 3 levels x (14 fused + 13 error-only) = 81 k_pass launches
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 81 -c 6 -f -o gpurun_out/prof_pass \
    python bench.py --one-step > gpurun_out/b_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pyr_head -s 1 -c 1 -f -o gpurun_out/prof_pyr_head \
    python bench.py --one-step --pairs 128 > gpurun_out/b_ncu3.log 2>&1
stamp done
ls -la gpurun_out
# then, on two GPUs:  gpurun --gpus 2 -- 'R360_TEST_MULTI_GPU=1 python -m pytest tests/test_multi_gpu.py -m gpu -q'
