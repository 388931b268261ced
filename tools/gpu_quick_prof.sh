set -x
bash tools/gpu_quick.sh
ncu --set full --clock-control none --import-source on -k regex:k_pass -s 33 -c 1 -f -o gpurun_out/prof_pass \
    python bench.py --steps 1 --warmup 0 --pairs 64 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
