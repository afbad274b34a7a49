"""Summarise an ncu report for profiles/: python scripts/ncu_summary.py report.ncu-rep [kernel-index] > profiles/xxx.txt
(raw metrics the roofline uses + the per-instruction stall hot spots of the source page)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
r = rows[2 + which]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
print(f"# {rep} (ncu --set full --clock-control none; per-launch values, cold-cache replays)")
for w in want:
    if w in H:
        print(f"{w:75s} {r[H.index(w)]}  [{rows[1][H.index(w)]}]")
for h in H:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = r[H.index(h)]
        try:
            if float(v) >= 0.05:
                print(f"{h:75s} {v}")
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
ks = [i for i, rr in enumerate(rows) if rr and rr[0] == "Kernel Name"]
ki = ks[which]
H = rows[ki + 1]
body = rows[ki + 2: ks[which + 1] if which + 1 < len(ks) else len(rows)]
si, ai, ei = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
cols = [c for c in H if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(x[si]) for x in body if len(x) > si and x[si].isdigit())
print(f"\n# source page: {tot} stall samples over {len(body)} SASS instructions; the 16 hottest instructions")
top = sorted([(int(x[si]), i) for i, x in enumerate(body) if len(x) > si and x[si].isdigit()], reverse=True)[:16]
for s_, i in sorted(top, key=lambda t: t[1]):
    x = body[i]
    st = sorted([(int(x[H.index(c)]), c.replace("stall_", "")) for c in cols if x[H.index(c)].isdigit()], reverse=True)[:2]
    print(f"  #{i:5d} {100 * s_ / tot:5.1f}%  executed={x[ei]:>9}  {x[ai].strip()[:72]:72s} {st[0][1]}={st[0][0]}")
