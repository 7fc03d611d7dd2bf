"""Print the interesting fields of a bench.py JSON line (file argument or stdin)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
l = json.loads(txt.strip().splitlines()[-1])
print({k: l.get(k) for k in ("value", "ms_per_step", "gpu_launches", "n_gpus")}, "e2e", l["e2e"]["value"], "cpu", (l.get("cpu_baseline") or {}).get("value"))
print("run", l.get("run"), "roofline frac", l["roofline"]["frac"], "kernel ms", l["roofline"]["kernel_ms_per_launch"])
p = l.get("parity")
if p:
    print("parity pass", p["pass"], p["single_step"])
    print({k: p["rollout"][k] for k in ("touching_contacts", "bit_identical_worlds", "toi_events", "toi_calls", "touching_pair_sets_identical")})
    lm = p["rollout_vs_libm"]
    print("vs libm single", lm["single_step"]["max_abs_dq"], lm["single_step"]["max_abs_dqd"], "rollout max dq", lm["max_abs_dq"][-1], "mean dq", lm["mean_abs_dq"][-1])
if "multi_gpu" in l:
    print("multi", l["multi_gpu"])
if "cfg5" in l:
    c = l["cfg5"]
    print("cfg5", {k: c[k] for k in ("value", "ms_per_step")}, "e2e", c["e2e"]["value"], c.get("multi_gpu", {}).get("gather_ms_max_over_ranks"))
