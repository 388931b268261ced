# usage: gpurun --gpus N -- 'bash tools/gpu_r2_scale.sh N' -- bench.py at N GPUs exactly as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_n%s.json'%sys.argv[1]))
print("N=%s value %.0f  frac %.4f  e2e %.0f  h2d/rank %.1f GB/s  ceiling/rank %s  frac_of_ceiling %s  verify %s" % (sys.argv[1], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_copy_ceiling_gbs_per_rank'], d['e2e']['frac_of_copy_ceiling'], d['verify']))
for k,v in d.get('configs',{}).items(): print("  ", k, "%.0f pairs/s"%v['value'], "ok", v['pairs_ok'], "ids", v['gathered_ids_complete'], "gt", v['ground_truth_within_5mrad_1cm'], "/", v['ground_truth_checked'], "frac %.3f"%v['roofline']['frac'], "gather %.1f ms"%v['allgather_ms'])
PY
