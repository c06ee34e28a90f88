"""Region-level instruction breakdown of a match_fast_kernel profile (source files must match the profiled build)."""
import csv, subprocess, collections, sys
rep, regs_path, match_path = sys.argv[1], sys.argv[2], sys.argv[3]
fn = sys.argv[4] if len(sys.argv) > 4 else "int solve_abs("
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
        try: agg[(cur, int(r[0]))] = (int(d.get("Instructions Executed", 0) or 0), int(d.get("# Samples", 0) or 0), r[1].strip()[:70])
        except ValueError: pass
tot = sum(v[0] for v in agg.values())
text = open(regs_path).read().split("\n")
def find(s, start=0):
    for i in range(start, len(text)):
        if s in text[i]: return i + 1
    return None
f0 = find(fn)
marks = [("rebase", find("void rebase(")), ("augment", find("void augment(")), ("augment_half", find("void augment_half(")), ("parent_S", find("int parent_S(")),
         ("shrink", find("int shrink(")), ("expand_relabel", find("void expand_relabel(")), ("other solvers(before)", find("template <int K>", find("void expand_relabel("))),
         ("S:macros", f0), ("S:init", find("for (int x = lane; x < nn; x += 32) { f.base[x]", f0)), ("S:greedy", find("// ---- greedy matching", f0)),
         ("S:phase start", find("// ---- phases:", f0)), ("S:search", find("// ---- earliest event", f0) or find("// ---- minimum-slack event", f0)),
         ("S:validate/recompute", find("int src = -1;", f0)), ("S:advance/dual", find("D = t;", f0) or find("// ---- dual update", f0)),
         ("S:grow/augment", find("if (type == 0) {", f0)), ("S:shrink handler", find("} else if (type == 1) {", f0)), ("S:expand handler", find("const int b = cap + arg;", f0)),
         ("S:scan_where(tail)", find("_SCAN_WHERE(news);", f0)), ("S:phase end", find("// ---- phase end", f0) or find("return 0;", f0)), ("other solvers(after)", find("#undef", f0)), ("end", len(text) + 1)]
marks = [(a, b) for a, b in marks if b]
mt = open(match_path).read().split("\n")
k0 = next(i for i, l in enumerate(mt) if "match_fast_kernel(MatchArgs a)" in l) + 1
def findm(s):
    for i in range(k0, len(mt)):
        if s in mt[i]: return i + 1
mm = [("K:setup", k0), ("K:extract defects", findm("int n0 = 0;")), ("K:overflow", findm("if (n0 > CAP - 1)")), ("K:fill W", findm("const int stride = n + 1;")),
      ("K:solve call", findm("#if defined(GQ_PLAIN_TREE)") or findm("rc = gqr::solve")), ("K:predict", findm("if (rc == 0) {")), ("K:end", findm("// host: graph + tables"))]
reg = collections.OrderedDict()
for (f, l), (i, s, t) in agg.items():
    name = f
    if f == "blossom_regs.cuh":
        name = "pre"
        for (nm, ln), (nm2, ln2) in zip(marks, marks[1:]):
            if ln <= l < ln2: name = nm
    elif f == "matching.cu":
        name = "K:other"
        for (nm, ln), (nm2, ln2) in zip(mm, mm[1:]):
            if ln and ln2 and ln <= l < ln2: name = nm
    a = reg.get(name, [0, 0]); a[0] += i; a[1] += s; reg[name] = a
ts = sum(v[1] for v in reg.values())
print(f"total warp-instr {tot} = {tot/2**20/1000:.1f}k per shot (2^20 shots)")
for k, (i, s) in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} inst {100*i/tot:5.1f}% ({i/2**20/1000:6.2f}k/shot)  stall-samples {100*s/ts:5.1f}%")
