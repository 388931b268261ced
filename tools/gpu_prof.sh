# ncu evidence: one full capture of the fused pass kernel at level 0 (64 pairs)
set -x
mkdir -p gpurun_out
PAIRS=${PAIRS:-64}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 33 -c 1 -f -o gpurun_out/prof_pass \
    python bench.py --steps 1 --warmup 0 --pairs $PAIRS --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
