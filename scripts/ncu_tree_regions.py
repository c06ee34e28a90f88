"""Region-level instruction breakdown of a tree_solve_kernel capture (ncu --set full --import-source on) by marker
comments of match_tree.cuh / function boundaries of blossom_regs.cuh.  The sources must match the profiled build.
usage: ncu_tree_regions.py <report.ncu-rep> [shots decoded by the captured launch]"""
import collections
import csv
import os
import re
import subprocess
import sys

rep = sys.argv[1]
shots = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = hdr = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            agg[(cur, int(r[0]))] = (int(d.get("Instructions Executed", 0) or 0), int(d.get("# Samples", 0) or 0))
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "graphqec_b200", "csrc")
mt_text = open(os.path.join(csrc, "match_tree.cuh")).read().split("\n")


def at(marker, start=0):
    for i in range(start, len(mt_text)):
        if marker in mt_text[i]:
            return i + 1
    raise KeyError(marker)


s0 = at("__device__ __forceinline__ int solve_tree(")
g0 = at("if (type == 0) {", s0)
MT = [
    (0, "prologue/other"),
    (at("for (int x = lane; x < n; x += GQW_LANES) {", s0) - 1, "init"),
    (at("for (int rk = 0; rk < K; ++rk) {", s0) - 1, "root loop"),
    (at("// ---- phase: one alternating tree", s0), "phase start"),
    (at("// ---- the root's row, scanned once", s0), "root scan + 1-event fast path"),
    (at("// ---- one step further", s0), "2-event fast path"),
    (at("slfrom[k] = ek[k] != EK_NONE ? root : -1;", s0) - 2, "tree setup"),
    (at("// ---- scan the rows of the vertices", s0), "scan queue"),
    (at("GQT_SCAN(sv);", s0), "SCAN"),
    (at("GQT_SCAN(sv);", s0) + 1, "scan queue"),
    (at("// ---- earliest event", s0), "search"),
    (at("int src = -1;", s0), "sel src"),
    (g0, "grow"),
    (at("// the node is free (m == -1", g0) - 1, "augment(free/bdry target)"),
    (at("GQM_STAT(2, 1);", g0), "grow"),
    (at("} else if (type == 3) {", s0), "bdry augment"),
    (at("} else if (type == 1) {", s0), "shrink handler"),
    (at("// an entry whose source now sits", s0), "shrink: drop void entries"),
    (at("GQM_STAT(7, 1);", s0) - 1, "expand handler"),
    (at("// ---- phase end", s0) - 1, "phase end"),
    (at("return 0;", s0), "after"),
    (at("void dual_start("), "dual_start"),
    (at("int solve_predict("), "solve_predict"),
    (at("int decode_shot("), "x"),
]
MT.sort()
text = open(os.path.join(csrc, "blossom_regs.cuh")).read().split("\n")
fn_starts = [(i + 1, m.group(1)) for i, l in enumerate(text) for m in [re.match(r"^(?:template.*>\s*)?__device__.*?\b(\w+)\(", l)] if m]
reg = collections.Counter()
samp = collections.Counter()
for (f, l), (i, s) in agg.items():
    if f == "match_tree.cuh":
        name = [n for a, n in MT if a <= l][-1]
    elif f == "blossom_regs.cuh":
        c = [n for a, n in fn_starts if a <= l]
        name = "regs:" + (c[-1] if c else "?")
    else:
        name = f
    reg[name] += i
    samp[name] += s
ts = max(1, sum(samp.values()))
for n, v in reg.most_common():
    print(f"{n:34s} {100 * v / tot:5.1f}%  {v / shots:8.1f}/shot  samples {100 * samp[n] / ts:5.1f}%")
