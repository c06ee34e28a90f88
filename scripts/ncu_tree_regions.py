"""Region-level instruction breakdown of a tree_solve_kernel capture by line ranges of match_tree.cuh / blossom_regs.cuh."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]; shots = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]
    elif r[0] == "Line No": hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try: agg[(cur, int(r[0]))] = (int(d.get("Instructions Executed", 0) or 0), int(d.get("# Samples", 0) or 0))
        except ValueError: pass
tot = sum(v[0] for v in agg.values())
MT = [(0, "prologue/other"), (198, "init"), (213, "root loop"), (220, "phase start"), (237, "scan queue"), (249, "SCAN"), (250, "scan queue"), (253, "search"), (269, "sel src/stale check"), (272, "stale recompute"),
      (297, "grow"), (303, "augment(free/bdry target)"), (311, "grow"), (340, "bdry augment"), (345, "shrink handler"), (388, "expand handler"), (436, "phase end"), (460, "after"), (468, "dual_start"), (571, "solve_predict"), (640, "x")]
import os
text = open(os.path.join(os.path.dirname(__file__), "..", "graphqec_b200", "csrc", "blossom_regs.cuh")).read().split("\n")
fn_starts = [(i + 1, m.group(1)) for i, l in enumerate(text) for m in [re.match(r"^(?:template.*>\s*)?__device__.*?\b(\w+)\(", l)] if m]
reg = collections.Counter(); samp = collections.Counter()
for (f, l), (i, s) in agg.items():
    if f == "match_tree.cuh":
        name = [n for a, n in MT if a <= l][-1]
    elif f == "blossom_regs.cuh":
        c = [n for a, n in fn_starts if a <= l]; name = "regs:" + (c[-1] if c else "?")
    else: name = f
    reg[name] += i; samp[name] += s
ts = sum(samp.values())
for n, v in reg.most_common(): print(f"{n:34s} {100*v/tot:5.1f}%  {v/shots:8.1f}/shot  samples {100*samp[n]/ts:5.1f}%")
