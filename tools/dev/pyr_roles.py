"""Pyramid-build time by role mix (development tool): all targets / all sources / alternating / both roles."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rgbd360_b200 as r360

rows, cols, L, n = 1024, 2048, 4, 512
ctx = r360.Context(rows, cols, n, 4, r360.default_params(n_levels=L))
rgb = torch.empty((n, rows, cols, 3), dtype=torch.uint8, device="cuda")
dep = torch.empty((n, rows, cols), dtype=torch.int16, device="cuda")
ctx.synth_frames_dev(0, 0, n, rgb.data_ptr(), dep.data_ptr())
mixes = {"all_target": [r360.ROLE_TARGET] * n, "all_source": [r360.ROLE_SOURCE] * n,
         "alternating": [r360.ROLE_TARGET, r360.ROLE_SOURCE] * (n // 2), "both": [r360.ROLE_BOTH] * n}
out = {}
for name, roles in mixes.items():
    roles = np.array(roles, np.uint8)
    t = []
    for rep in range(4):
        ctx.set_frames_ptr(0, n, rgb.data_ptr(), dep.data_ptr(), roles, device=True)
        t.append(ctx.last_device_ms())
    out[name] = {"ms_per_512_frames": min(t[1:]), "us_per_frame": 1e3 * min(t[1:]) / n}
print(json.dumps(out, indent=1))
