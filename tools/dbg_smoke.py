import sys; sys.path.insert(0,'.')
import numpy as np
import rgbd360_b200 as r360
from oracle import orc
rows, cols, L = 128, 256, 3
P = orc.default_params(n_levels=L)
rgb_t, d_t = orc.synth_frame(0, 0, rows, cols)
rgb_s, d_s = orc.synth_frame(0, 1, rows, cols)
trg = orc.Frame(rgb_t, d_t, P, True); src = orc.Frame(rgb_s, d_s, P, False)
res_o, tr_o = orc.align(src, trg, None, P, trace=True)
ctx = r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L), device=0)
ctx.set_frames(0, np.stack([rgb_s, rgb_t]), np.stack([d_s, d_t]), [1,2])
res, tr = ctx.register_pairs([0], [1], trace=True)
for a,b in zip(tr_o, tr):
    if a.used or b.used:
        print('O', a.level, a.it, a.accepted, a.used, a.err2, a.n_valid, a.n_visible, '| G', b.level, b.it, b.accepted, b.used, b.err2, b.n_valid, b.n_visible, 'dpose', np.abs(np.array(a.pose)-np.array(b.pose)).max())
print(list(res_o.iters)[:L], res[0]['iters'][:L], res_o.final_n_valid, res[0]['final_n_valid'])
