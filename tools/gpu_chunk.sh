# usage: bash tools/gpu_chunk.sh -- pyramid chunk-size variants, then the occlusion launch list (time + DRAM bytes per kernel)
mkdir -p gpurun_out
bash tools/variants_pyr.sh
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_occ1_pipe.csv \
    python bench.py --steps 1 --warmup 0 --pairs 64 --no-cpu-baseline --occlusion 1 > gpurun_out/b_occ_ncu.log 2>&1
