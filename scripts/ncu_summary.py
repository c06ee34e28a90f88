"""Summarise an .ncu-rep: headline raw metrics + top source lines by instructions / stall samples."""
import collections, csv, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit", "lts__t_bytes.sum", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_per_inst_issued", "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_alu.avg.pct", "l1tex__data_bank_conflicts",
        "smsp__thread_inst_executed_per_inst_executed"]
print("== raw metrics:", rows[2][4] if len(rows[2]) > 4 else "")
for h, u, v in zip(hdr, units, vals):
    if any(k in h for k in keys) and ".max" not in h and ".min" not in h and ".sum.p" not in h:
        try:
            if float(v.replace(",", "")) == 0 and "stalled" in h:
                continue
        except ValueError:
            pass
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur, hdr2, agg = None, None, collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr2 = r
    elif hdr2 and r[0].isdigit():
        d = dict(zip(hdr2, r))
        try:
            agg[(cur, int(r[0]))] = (r[1].strip()[:100], int(d.get("# Samples", 0) or 0), int(d.get("Instructions Executed", 0) or 0))
        except ValueError:
            pass
ts = sum(v[1] for v in agg.values()) or 1
ti = sum(v[2] for v in agg.values()) or 1
print(f"== source lines (total warp-instructions {ti}, stall samples {ts})")
for (f, l), (s, sm, ins) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{f}:{l:4d} inst {100*ins/ti:5.1f}% samples {100*sm/ts:5.1f}%  {s}")
