"""Summarise an .ncu-rep (raw page + source page) into text: key metrics, stall mix, instruction mix, hot lines."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_uniform.sum", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
name_col = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[name_col][:100])
    for i, h in enumerate(hdr):
        if h in want:
            print(f"   {h:80s} {r[i]:>16s} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for k, hi in enumerate(his[:1]):
    hdr = rows[hi]
    end = his[k + 1] - 1 if k + 1 < len(his) else len(rows)
    data = rows[hi + 1:end]
    col = {h: i for i, h in enumerate(hdr)}

    def f(r, c):
        try:
            return float(r[col[c]])
        except Exception:
            return 0.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(f(r, "# Samples") for r in data)
    print("\n== source page of", rows[hi - 1][1][:90] if hi > 0 else "?", " samples", tot)
    agg = {s: sum(f(r, s) for r in data) for s in stalls}
    for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
        print(f"   {s:28s} {100 * v / max(tot, 1):5.1f}%")
    ops = collections.Counter()
    for r in data:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]].strip())
        ops[m.group(2).split(".")[0] if m else "?"] += f(r, "Instructions Executed")
    tot_i = sum(ops.values())
    print("   instruction mix:", ", ".join(f"{o} {100 * v / tot_i:.1f}%" for o, v in ops.most_common(14)), f" total {tot_i:.3g}")
    print("   hottest lines:")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top_n]:
        st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
        print(f"   {100 * f(r, '# Samples') / max(tot, 1):5.2f}%  exec={f(r, 'Instructions Executed'):11.0f}  "
              f"{r[col['Source']].strip()[:74]:74s} {st[0][1][6:]}:{st[0][0]:.0f} {st[1][1][6:]}:{st[1][0]:.0f}")
