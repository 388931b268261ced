# usage: bash tools/gpu_prof_head.sh -- one ncu --set full capture of k_pyr_head (second launch of a 128-pair step), summary printed
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pyr_head -s 1 -c 1 -f -o gpurun_out/prof_pyr_head \
    python bench.py --one-step --pairs 128 > gpurun_out/b_ncu3.log 2>&1
tail -2 gpurun_out/b_ncu3.log
python tools/ncu_summary.py gpurun_out/prof_pyr_head.ncu-rep 2>&1 | head -60 | tee gpurun_out/prof_pyr_head_summary.txt
