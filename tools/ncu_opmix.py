"""Executed SASS opcode mix of the kernels in an .ncu-rep, per pixel-pair iteration (n_iter given on the command line)."""
import csv, subprocess, collections, sys
rep = sys.argv[1]; n_iter = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr = None; ops = collections.Counter(); tot = 0
for r in csv.reader(src.splitlines()):
    if len(r) > 5 and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr):
        try: inst = float(r[hdr.index('Instructions Executed')] or 0)
        except ValueError: continue
        op = r[hdr.index('Source')].strip().split()
        if not op: continue
        o = op[1] if op[0].startswith('@') else op[0]
        ops[o.split('.')[0]] += inst; tot += inst
for o, v in ops.most_common(45): print('%-10s %8.2f  %5.1f%%' % (o, v / n_iter, 100 * v / tot))
print('total %.2f' % (tot / n_iter))
