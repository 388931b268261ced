"""Summarise an .ncu-rep: headline metrics, stall reasons, per-source-line instruction/stall shares."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed.sum']
for i, h in enumerate(H):
    if h in want:
        print(f"{h} [{U[i]}] = {[r[i] for r in rows[2:]]}")
st = []
for i, h in enumerate(H):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        st.append((float(rows[2][i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
print('stalls/issue:', ', '.join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = {}
for r in csv.reader(src.splitlines()):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1]; continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and cur:
        try:
            ln = int(r[0]); inst = float(r[hdr.index('Instructions Executed')] or 0); s = float(r[hdr.index('Warp Stall Sampling (All Samples)')] or 0)
        except ValueError:
            continue
        a = agg.setdefault((cur.split('/')[-1], ln, r[1].strip()[:80]), [0, 0]); a[0] += inst; a[1] += s
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-16s %4d inst%%=%5.2f stall%%=%5.2f  %s' % (k[0], k[1], 100 * v[0] / ti, 100 * v[1] / ts, k[2]))
