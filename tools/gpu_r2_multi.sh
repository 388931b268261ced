# usage: gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N' -- the multi-GPU evidence of round 2: the C++ all-gather demo through
# r360_allgather_results (2 host threads / contexts / NCCL ranks), the ingest + full GPU suite on this box, bench at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
t0=$(date +%s); stamp() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -14
stamp "gpu tests (incl. the 2-GPU C ABI all-gather)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/gputests_n$N.txt
stamp "bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -4 gpurun_out/bench_n$N.err; cut -c1-1500 gpurun_out/bench_n$N.json
stamp done
