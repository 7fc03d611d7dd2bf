"""Stepping-kernel mode sweep: ms per 100-step rollout launch of one config for several batch sizes and (lanes per world,
worlds per warp) choices.  Usage: sweep_modes.py <config> "<n1,n2,...>" "<lpw:lanes,lpw:lanes,...>"  (lanes 0 = default)"""
import os, sys, subprocess
cfg, ns, modes = sys.argv[1], sys.argv[2].split(","), sys.argv[3].split(",")
here = os.path.dirname(os.path.abspath(__file__))
print("%-8s" % "worlds" + "".join("%12s" % m for m in modes))
for n in ns:
    row = "%-8s" % n
    for m in modes:
        lpw, lanes = m.split(":")
        env = dict(os.environ, B2G_LPW=lpw)
        if lanes != "0":
            env["B2G_LANES"] = lanes
        else:
            env.pop("B2G_LANES", None)
        out = subprocess.run([sys.executable, os.path.join(here, "prof_cfg.py"), cfg, n, "3"], env=env, capture_output=True, text=True).stdout
        try:
            row += "%12.3f" % float(out.split(":")[1].split("ms")[0])
        except Exception:
            row += "%12s" % "fail"
    print(row, flush=True)
