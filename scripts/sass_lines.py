"""Static SASS instruction count per source line of one kernel (object built with -lineinfo).
usage: sass_lines.py <nvdisasm --print-line-info output> <mangled kernel name> [top]"""
import re, sys, collections, glob
txt = open(sys.argv[1]).read().split("\n")
kern = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
start = next(i for i, l in enumerate(txt) if ".text." + kern in l and ".section" in l)
cnt = collections.Counter(); cur = None; n = 0
for l in txt[start + 1:]:
    if l.lstrip().startswith(".section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): cnt[cur] += 1; n += 1
print("total", n)
src = {}
for (f, ln), c in cnt.most_common(top):
    if f not in src:
        g = glob.glob("/root/repo/graphqec_b200/csrc/" + f)
        src[f] = open(g[0]).read().split("\n") if g else None
    print(f"{c:5d} {f}:{ln}  {src[f][ln-1].strip()[:90] if src[f] else ''}")
