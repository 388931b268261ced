"""Write-only, read-only and copy bandwidth of the box (torch, CUDA events): context for the write-dominated pyramid kernels."""
import os, torch, json
os.makedirs("gpurun_out", exist_ok=True)
n = 1 << 30                                   # 1 Gi bytes x 2 buffers
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
def best(f, byts, reps=10):
    f(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return byts / (min(t) * 1e-3) / 1e9
a32 = a.view(torch.float32)
out = {"fill_gbs": best(lambda: a.zero_(), n), "copy_gbs": best(lambda: b.copy_(a), 2 * n),
       "read_sum_gbs": best(lambda: a32.sum(), n)}
open("gpurun_out/fill_bw.json", "w").write(json.dumps(out)); print(json.dumps(out))
