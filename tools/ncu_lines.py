"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line: instructions executed and
stall samples per line, per file.  Usage: ncu_lines.py dump.csv [top]"""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
per_line = collections.Counter(); samples = collections.Counter(); src = {}
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if r[0] and r[0].isdigit():
        key = (cur, int(r[0]))
        try:
            per_line[key] += int(r[iI]); samples[key] += int(r[iS])
        except ValueError:
            pass
        src[key] = r[1]
tot = sum(per_line.values()); ts = sum(samples.values())
print("total warp instructions %d, samples %d" % (tot, ts))
byfile = collections.Counter(); sfile = collections.Counter()
for k, v in per_line.items(): byfile[k[0]] += v; sfile[k[0]] += samples[k]
for f, v in byfile.most_common(): print("  %-20s inst %5.1f%%  samples %5.1f%%" % (f, 100.0 * v / tot, 100.0 * sfile[f] / max(ts, 1)))
for k, v in per_line.most_common(top):
    print("%5.2f%% i %5.2f%% s  %s:%d  %s" % (100.0 * v / tot, 100.0 * samples[k] / max(ts, 1), k[0], k[1], src[k].strip()[:110]))
