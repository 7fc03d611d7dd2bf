"""Phase map of a kernel: walks the SASS of `ncu --page source --csv --print-source sass` in address order, cuts it into
segments of (nearly) constant execution count and prints, per segment, its share of executed warp instructions and of the
stall samples, with the enclosing (non-inlined) function and source-line span from an `nvdisasm -g -c` listing of the same
cubin.  Usage: ncu_phase_map.py sass.csv kernel.sass warp_steps [min_share_pct]"""
import csv, re, sys
rows = csv.reader(open(sys.argv[1]))
listing = sys.argv[2]
warp_steps = float(sys.argv[3])
minshare = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
ins = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "Address": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iL = hdr.index("stall_long_sb"); iB = hdr.index("stall_barrier"); iW = hdr.index("stall_wait"); iT = hdr.index("Thread Instructions Executed"); continue
    if hdr and r[0].startswith("0x"):
        ins.append((int(r[0], 16), int(r[iI]), int(r[iS]), int(r[iL]), int(r[iB]), int(r[iW]), int(r[iT]), r[1].strip()))
base = ins[0][0]
# listing: offset -> (function, file:line)
fn = "main"; line = "?"; where = {}
for l in open(listing):
    m = re.match(r"^(\$?[A-Za-z_][\w\$]*):", l)
    if m and not l.startswith(".L_"):
        name = m.group(1)
        mm = re.search(r"\$_ZN3b2g(\d+)", name)
        if mm:
            i = name.index(mm.group(0)) + len(mm.group(0)); fn = name[i:i + int(mm.group(1))]
        elif "k_step" in name: fn = "k_step"
        else: fn = name.strip("$")[:30]
        continue
    m = re.match(r'^\s*//## File "[^"]*/([^/"]+)", line (\d+)', l)
    if m: line = m.group(1).replace("b2g_", "").replace(".cuh", "") + ":" + m.group(2); continue
    m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/", l)
    if m: where[int(m.group(1), 16)] = (fn, line)
tot = sum(i[1] for i in ins); ts = sum(i[2] for i in ins)
print("executed warp instr %d (%.0f per warp-step), samples %d, avg active threads %.1f" % (tot, tot / warp_steps, ts, sum(i[6] for i in ins) / max(tot, 1)))
print("%-22s %-30s %6s %9s %7s %7s | %6s %6s %6s" % ("function", "lines", "#sass", "exec/wstep", "inst%", "smpl%", "longsb", "barr", "wait"))
seg = []
def flush():
    if not seg: return
    n = sum(i[1] for i in seg); s = sum(i[2] for i in seg)
    if 100.0 * n / tot >= minshare or 100.0 * s / ts >= minshare:
        f0 = where.get(seg[0][0] - base, ("?", "?")); f1 = where.get(seg[-1][0] - base, ("?", "?"))
        lines = set(where.get(i[0] - base, ("?", "?"))[1] for i in seg)
        files = sorted(set(x.split(":")[0] for x in lines))
        span = ",".join("%s:%s-%s" % (f, min(int(x.split(":")[1]) for x in lines if x.startswith(f + ":")), max(int(x.split(":")[1]) for x in lines if x.startswith(f + ":"))) for f in files if f != "?")
        print("%-22s %-30s %6d %9.1f %6.2f%% %6.2f%% | %5.1f%% %5.1f%% %5.1f%%" % (f0[0][:22], span[:30], len(seg), n / len(seg) / warp_steps, 100.0 * n / tot, 100.0 * s / ts,
              100.0 * sum(i[3] for i in seg) / max(s, 1), 100.0 * sum(i[4] for i in seg) / max(s, 1), 100.0 * sum(i[5] for i in seg) / max(s, 1)))
    del seg[:]
prev = None
for i in ins:
    f = where.get(i[0] - base, ("?", "?"))[0]
    if prev is not None and (f != prev[0] or not (0.6 * prev[1] <= i[1] <= 1.6 * prev[1])):
        flush()
    seg.append(i)
    avg = sum(x[1] for x in seg) / len(seg)
    prev = (f, avg)
flush()
