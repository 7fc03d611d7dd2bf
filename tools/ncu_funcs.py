"""Attribute an `ncu --page source --csv --print-source cuda,sass` dump to the (innermost inlined) source functions of
box2d_sim_env_b200/csrc.  Usage: ncu_funcs.py dump.csv [top]"""
import csv, collections, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = csv.reader(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cur = None; per = collections.Counter(); smp = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if r[0] and r[0].isdigit():
        k = (cur, int(r[0]))
        try: per[k] += int(r[iI]); smp[k] += int(r[iS])
        except ValueError: pass
ts = sum(smp.values()); ti = sum(per.values())
allf = collections.Counter(); alli = collections.Counter()
for fname in sorted({k[0] for k in per}):
    path = os.path.join(ROOT, "box2d_sim_env_b200", "csrc", fname)
    if not os.path.exists(path): continue
    funcs = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        if l.startswith(("DEV", "__global__", "template")):
            m = re.search(r"(\w+)\(", l)
            if m and m.group(1) not in ("template",): funcs.append((i, m.group(1)))
    for (f, ln), v in smp.items():
        if f != fname: continue
        name = "?"
        for s, n in funcs:
            if s <= ln: name = n
            else: break
        allf[fname + ":" + name] += v; alli[fname + ":" + name] += per[(f, ln)]
print("total warp instructions %d, samples %d" % (ti, ts))
for n, v in allf.most_common(top): print("   %-42s samples %5.2f%%  inst %5.2f%%" % (n, 100 * v / ts, 100 * alli[n] / ti))
