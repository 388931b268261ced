"""Cross-check of the packed index path against the scalar pinned sequence at BASELINE sizes (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rgbd360_b200 as r360
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import small_pose
rows, cols, L = 1024, 2048, 4
ctx = r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L))
rgb, dep = ctx.synth_frames(0, 0, 2)
ctx.set_frames(0, rgb, dep)
rng = np.random.default_rng(1)
poses = [np.eye(4)] + [small_pose(*(rng.uniform(-1, 1, 3) * 0.5), *(rng.uniform(-1, 1, 3) * 3.0)) for _ in range(15)]
for level in range(L):
    tot = dict(valid=0, scalar=0, mismatch=0)
    for T in poses:
        st = ctx.index_stats(1, 0, level, T)
        for k in tot:
            tot[k] += st[k]
    print("level %d (%dx%d): pixels %d, sent to the scalar pinned code %d (%.2e), packed != scalar on %d"
          % (level, cols >> level, rows >> level, tot["valid"], tot["scalar"], tot["scalar"] / tot["valid"], tot["mismatch"]))
