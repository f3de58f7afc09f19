"""Summarise an .ncu-rep of the fused kernel: key counters + hottest source lines (run here, no GPU)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else None
lines = []
def P(*a):
    s = " ".join(str(x) for x in a); print(s); lines.append(s)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keep = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio']
P("# counters")
for k in keep:
    if k in d: P(f"{k:75s} {d[k][0]:16s} {d[k][1]}")
P("# warp stall samples")
st = [(int(v), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for h, (u, v) in d.items()
      if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
tot = sum(s[0] for s in st)
for n, h in sorted(st, reverse=True)[:10]: P(f"  {h:28s} {n:9d} {100*n/max(tot,1):5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = [r for r in rows if r and r[0] == 'Line No'][0]
ci = hdr.index('Instructions Executed'); cs = hdr.index('Warp Stall Sampling (All Samples)')
cur = None; acc = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 10 or r[0] in ('Line No', ''): continue
    try: acc.append((int(r[ci]), int(r[cs] or 0), cur, r[0], r[1].strip()[:100]))
    except Exception: pass
ti = sum(a[0] for a in acc); ts = sum(a[1] for a in acc)
P(f"# hottest source lines (of {ti} executed warp-instructions, {ts} samples)")
for a in sorted(acc, reverse=True)[:28]:
    P(f"{100*a[0]/ti:5.1f}% inst {100*a[1]/max(ts,1):5.1f}% smp  {a[2]}:{a[3]:>4s}  {a[4]}")
if out: open(out, "w").write("\n".join(lines) + "\n")
