"""Attribute one stall reason of an `ncu --page source --csv --print-source cuda,sass` dump to the OUTER (non-inlined, DEVN /
__global__) function whose SASS range contains the instruction, using the sass section of the dump.
Usage: ncu_stall_funcs.py dump.csv stall_long_sb [top]   (works on the cuda-source view: groups by file:line ranges)"""
import csv, collections, sys
col = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = csv.reader(open(sys.argv[1]))
cur = None; S = collections.Counter(); src = {}; tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ic = hdr.index(col); continue
    if r[0].isdigit():
        try: v = int(r[ic])
        except ValueError: continue
        S[(cur, int(r[0]))] += v; tot += v; src[(cur, int(r[0]))] = r[1].strip()[:100]
print("total %s samples: %d" % (col, tot))
for k, v in S.most_common(top): print("%5.2f%%  %s:%d  %s" % (100.0 * v / max(tot, 1), k[0], k[1], src[k]))
