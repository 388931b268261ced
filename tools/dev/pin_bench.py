"""Pinhole registration throughput (development tool): alignFrames on a batch of synthetic 640x480 pairs, device time."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle"))
import rgbd360_b200 as r360
import orc

rows, cols, L, n = 480, 640, 4, 256
cam = (525.0, 525.0, 319.5, 239.5)
p = r360.pinhole_params(n_levels=L)
ctx = r360.Context(rows, cols, 2 * n, n, p)
ctx.set_camera(*cam)
base = [orc.synth_pinhole_frame(0, k, rows, cols, *cam) for k in range(16)]
rgb = np.stack([base[k % 16][0] for k in range(2 * n)]); dep = np.stack([base[k % 16][1] for k in range(2 * n)])
ctx.set_frames(0, rgb, dep, np.array([r360.ROLE_TARGET, r360.ROLE_SOURCE] * n, np.uint8))
pyr_ms = ctx.last_device_ms()
trg = np.arange(0, 2 * n, 2, dtype=np.int32); src = trg + 1
t = []
for rep in range(4):
    res = ctx.register_pairs(src, trg)
    t.append(ctx.last_device_ms())
ms = min(t[1:])
print(json.dumps({"pairs": n, "size": [rows, cols], "levels": L, "register_ms": ms, "pairs_per_s": n / ms * 1e3,
                  "ok": int((res["status"] == 0).sum()), "iters_mean": res["iters"][:, :L].mean(axis=0).tolist(),
                  "launches": ctx.kernel_launches()}))
