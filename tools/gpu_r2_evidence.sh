# usage: bash tools/gpu_r2_evidence.sh -- evidence for the committed build in one bounded call (one B200):
# the full bench line, the ncu launch list of exactly ONE 512-pair step, full captures of the level-0 k_pass launches (both
# instantiations) and of k_pyr_head, compute-sanitizer over the kernels that are new or changed in round 2.
mkdir -p gpurun_out
t0=$(date +%s); stamp() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
stamp bench
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err; cut -c1-600 gpurun_out/bench_full.json
stamp "launch list of one step (512 pairs)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_one_step.csv python bench.py --one-step > gpurun_out/b_ncu1.log 2>&1
tail -1 gpurun_out/b_ncu1.log
stamp "full captures"
# one 512-pair step: level 0 starts after 3 levels x (14 fused + 13 error-only) = 81 k_pass launches
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 81 -c 5 -f -o gpurun_out/prof_pass \
    python bench.py --one-step > gpurun_out/b_ncu2.log 2>&1
# per chunk of 128 frames: k_pyr_head<0> (raw input), then k_pyr_head<2> for levels 1, 2, 3 -- launch 4 is the head of the second chunk, launch 1 the level-1 pass
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pyr_head -s 4 -c 1 -f -o gpurun_out/prof_pyr_head \
    python bench.py --one-step --pairs 128 > gpurun_out/b_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pyr_head -s 1 -c 1 -f -o gpurun_out/prof_pyr_mid \
    python bench.py --one-step --pairs 128 > gpurun_out/b_ncu4.log 2>&1
stamp sanitizer
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool: smoke()"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|smoke ok|Error|RACECHECK SUMMARY|hazard" | head -8
done 2>&1 | tee gpurun_out/sanitizer.txt
echo "== memcheck + racecheck: k_pass (dynamic tail, fixed-point flush), k_pyr_head (column strips, partial tiles, 8x16, f32 depth), rig, stitch, latency mode" | tee -a gpurun_out/sanitizer.txt
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_rig.py tests/test_ingest.py -m gpu -x -q \
    -k "edge_cases or invalid_depth or deterministic or slot_reset or (rig_equals and L3) or (stitch_random and 0) or align_with_guess" 2>&1 \
    | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" | head -8 | tee -a gpurun_out/sanitizer.txt
done
stamp done
ls -la gpurun_out | head -40
