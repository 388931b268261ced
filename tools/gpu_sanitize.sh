# usage: bash tools/gpu_sanitize.sh -- compute-sanitizer (memcheck, racecheck, initcheck) over smoke() and two parity tests
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool: smoke()" 
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|smoke ok|Error|RACECHECK SUMMARY|hazard" | head -8
done 2>&1 | tee gpurun_out/sanitizer.txt
echo "== memcheck: pytest (warp maps, error/hessgrad, align, ingest)" | tee -a gpurun_out/sanitizer.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_ingest.py -m gpu -x -q -k "warp_maps or error_and_hessgrad or align_identity or stitch or edge" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Error" | head -8 | tee -a gpurun_out/sanitizer.txt
echo "== memcheck + racecheck: k_pyr_head (partial tiles, 8x16, f32 depth), occlusion and pinhole kernels" | tee -a gpurun_out/sanitizer.txt
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_occlusion.py tests/test_pinhole.py -m gpu -x -q \
    -k "edge_cases or invalid_depth or (occlusion_evaluations and holes) or (occlusion_align and loop) or (pinhole_align and holes) or pinhole_batch" 2>&1 \
    | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" | head -8 | tee -a gpurun_out/sanitizer.txt
done
