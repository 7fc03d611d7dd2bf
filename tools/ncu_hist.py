"""Histogram of SASS instructions by execution count from `ncu --page source --csv --print-source cuda,sass`:
groups instructions that run the same number of times (= the same loop nest), which separates solver phases.
Usage: ncu_hist.py dump.csv [min_share_pct]"""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1]))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
seen = {}
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "Line No": hdr = r; iA = 2; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr and r[0] == "" and len(r) > iI and r[iA].startswith("0x"):
        try: seen[r[iA]] = (int(r[iI]), int(r[iS]), r[3].strip())
        except ValueError: pass
tot = sum(v[0] for v in seen.values()); ts = sum(v[1] for v in seen.values())
groups = collections.defaultdict(lambda: [0, 0, 0])
for a, (n, s, txt) in seen.items():
    g = groups[n]; g[0] += 1; g[1] += n; g[2] += s
print("distinct SASS instructions %d, executed %d, samples %d" % (len(seen), tot, ts))
print("%12s %8s %8s %8s" % ("exec/inst", "#sass", "inst%", "samples%"))
for n, g in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    if 100.0 * g[1] / tot < minshare: continue
    print("%12d %8d %7.2f%% %7.2f%%" % (n, g[0], 100.0 * g[1] / tot, 100.0 * g[2] / max(ts, 1)))
