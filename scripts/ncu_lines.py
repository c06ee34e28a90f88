"""Per-source-line instruction / stall-sample table of an ncu report (kernel compiled with -lineinfo)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]
    elif r[0] == "Line No": hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try: agg[(cur, int(r[0]))] = (int(d.get("Instructions Executed", 0) or 0), int(d.get("# Samples", 0) or 0), r[1].strip()[:90])
        except ValueError: pass
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-instr {tot}, samples {ts}")
byfile = collections.Counter()
for (f, l), (i, s, t) in agg.items(): byfile[f] += i
print({k: f"{100*v/tot:.1f}%" for k, v in byfile.most_common(8)})
for (f, l), (i, s, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{l:4d} inst {100*i/tot:5.1f}% samp {100*s/max(ts,1):5.1f}%  {t}")
