#!/usr/bin/env python
"""bench.py — world-steps/s of the batched Box2D hot path on N B200s (BASELINE.json metric).

One bench "step" = one planner rollout over a batch of synthetic worlds:
    set_state (fresh caches)  ->  10 x [setTargetVelocity ; stepPhysics(10)]  ->  get_state
The default workload is BASELINE.json configs[1] (test_env.yaml x 4096 with random robot velocity controls, 1 B200);
--config selects the other named configs (cfg3 rigid robot x 65536, cfg4 clutter x 65536, cfg5 1M-world sweep).
With --gpus N every rank owns its own batch (weak scaling; cfg5: a slice of the fixed total, strong scaling) and the
final [n, 2, nq] states are all-gathered over NCCL, which is the only collective the path has.

    value     device-timed throughput with inputs resident in HBM (CUDA events on the batch's stream)
    e2e       same metric through the host-pointer C-ABI call a planner makes (b2g_batch_rollout: pinned host buffers in,
              pinned host buffers out; H2D of the step's inputs and D2H of its results inside the timed region; with
              --gpus N > 1 the device path plus explicit copies, because the NCCL gather needs the states on the device)
    roofline  algorithmic bytes of the rollout kernel / its CUDA-event duration, against the measured HBM peak
    cpu_baseline   the CPU oracle ("port" of the reference's arithmetic; the reference itself cannot be built
              here) on the box's host cores, one world per thread, bounded sample

`--impl reference` times that CPU oracle alone on the same config/metric (the reference arm for this tier).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STEPS_PER_TARGET = 10
N_TARGETS = 10
ROLLOUT_STEPS = STEPS_PER_TARGET * N_TARGETS
BASE_Q = np.array([0, 0, 0, 0, 0, 0, 0, 0.9, 1.0, -1.0, 0.0], np.float32)

# BASELINE.json configs -> concrete workloads.  `bytes` = algorithmic bytes per world-step
# (fixed part, per live contact), DESIGN.md §2.
CONFIGS = {
    "cfg2": dict(env="test_env.yaml", worlds=4096, robot_dofs=7, bytes=(940.0, 64.0), scaling="weak",
                 text="test_env.yaml x %d worlds per GPU x 100 physics steps per step, random robot velocity targets "
                      "resampled every 10 steps (BASELINE.json configs[1])"),
    "cfg3": dict(env="test_env_rigid_robot.yaml", worlds=65536, robot_dofs=3, bytes=(348.0, 64.0), scaling="weak",
                 text="test_env_rigid_robot.yaml x %d worlds per GPU x 100 physics steps per step, rigid robot pushing with "
                      "ground friction joints, random base velocity targets resampled every 10 steps (BASELINE.json configs[2])"),
    "cfg4": dict(env="clutter", worlds=65536, robot_dofs=7, bytes=(4660.0, 64.0), scaling="weak",
                 text="synthetic clutter (robot + 32 convex polygons, seed 42) x %d worlds per GPU x 100 physics steps per step "
                      "(propagate), batched checkCollision timed separately (BASELINE.json configs[3])"),
    "cfg5": dict(env="test_env.yaml", worlds=1048576, robot_dofs=7, bytes=(940.0, 64.0), scaling="strong",
                 text="planner rollout sweep: test_env.yaml x %d worlds in total x 100 physics steps per step, sharded over "
                      "the GPUs, final NCCL all-gather of states (BASELINE.json configs[4])"),
}


def load_model(cfg):
    import box2d_sim_env_b200 as b2g
    if cfg["env"] == "clutter":
        from box2d_sim_env_b200 import envgen
        return b2g.Model(envgen.clutter_env())
    return b2g.Model(b2g.data_path(cfg["env"]))


def oracle_env(cfg):
    from oracle import b2o
    if cfg["env"] == "clutter":
        from box2d_sim_env_b200 import envgen
        return envgen.clutter_env()
    return b2o.load_env_yaml(os.path.join(ROOT, "box2d_sim_env_b200", "data", cfg["env"]))


def synth_inputs(limits, n, seed, cfg=None, lo=0, hi=None):
    """Inputs of one bench step for worlds lo:hi of an n-world batch: jittered initial states and robot
    targets ~ U(velocity limits) resampled every 10 steps.  Counter-based (Philox) and indexed by world, so the CPU
    and GPU arms — and any GPU count — draw identical controls for the same world."""
    cfg = cfg or CONFIGS["cfg2"]
    hi = n if hi is None else hi
    nd = cfg["robot_dofs"]
    rng = np.random.Generator(np.random.Philox(key=seed))
    if cfg["env"] == "test_env.yaml":
        q = np.tile(BASE_Q, (n, 1))
        q[:, 7:9] += rng.uniform(-0.2, 0.2, (n, 2)).astype(np.float32)
        q[:, 9] = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
        qd = np.zeros_like(q)
    elif cfg["env"] == "test_env_rigid_robot.yaml":
        # robot(3) obj1(3) obj2(1, static hinge): obj1 in front of the pusher, jittered
        q = np.tile(np.array([0, 0, 0, 0.9, 1.0, -1.0, 0.0], np.float32), (n, 1))
        q[:, 3:5] += rng.uniform(-0.3, 0.3, (n, 2)).astype(np.float32)
        q[:, 5] = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
        qd = np.zeros_like(q)
    else:
        from box2d_sim_env_b200 import envgen
        q, qd = envgen.clutter_states(n, seed=seed)
    tg = rng.uniform(limits[:nd, 2], limits[:nd, 3], (N_TARGETS, n, nd)).astype(np.float32)
    return q[lo:hi], qd[lo:hi], np.ascontiguousarray(tg[:, lo:hi])


def algorithmic_bytes_per_world_step(mean_contacts, cfg=None):
    # e.g. shape A = 6 dynamic bodies x 64 B + 6 active friction joints x 24 B + 4 active revolute
    # joints x 40 B + 7 fixtures on dynamic bodies x 32 B + 28 B of controls, + 64 B per live contact
    fixed, per_contact = (cfg or CONFIGS["cfg2"])["bytes"]
    return fixed + per_contact * mean_contacts


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows = []
        self.stop = False
        self.device = device
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_oracle_throughput(cfg, limits, n_worlds, repeats, threads, seed0=9000):
    """CPU oracle stepping the same workload, one world per thread-pool slot.  Returns (units per repeat, [seconds])."""
    from oracle import b2o
    batch = b2o.OracleBatch(oracle_env(cfg), n_worlds)
    times = []
    for r in range(repeats):
        q, qd, tg = synth_inputs(limits, n_worlds, seed0 + r, cfg)
        # a fresh batch per repeat would also time construction; the reference's clone() is not part of the metric,
        # so worlds are reused and only set_state + stepping is timed (setWorldState on a live world)
        t0 = time.perf_counter()
        batch.rollout(q, qd, 0, tg, STEPS_PER_TARGET, threads)
        times.append(time.perf_counter() - t0)
    return n_worlds * ROLLOUT_STEPS, times


def parity_gate(cfg, batch, limits, n, seed, threads, sample=1024):
    """Correctness gate that accompanies the measurement: the first `sample` worlds of one bench step's
    inputs go through the CUDA path and through the CPU oracle (here as the checker); compared after ONE step from the
    identical fresh-cache state and after the whole 100-step rollout: DOF positions / velocities (max |diff|, worlds that
    are bit-identical) and the touching-pair sets with their manifold point counts."""
    from oracle import b2o
    m = min(sample, n)
    q, qd, tg = synth_inputs(limits, n, seed, cfg)
    out = {"worlds": m, "physics_steps": ROLLOUT_STEPS}
    env = b2o.OracleEnv(oracle_env(cfg))
    for label, n_tg, spt in (("single_step", 1, 1), ("rollout", N_TARGETS, STEPS_PER_TARGET)):
        ob = b2o.OracleBatch(env, m)   # freshly loaded worlds: the CUDA side starts from fresh clones (reset_caches) too
        gq, gqd = batch.rollout(0, tg[:n_tg], spt, q, qd, reset_caches=True)
        counts, cio, _ = batch.contacts(max_per_world=64)
        oq, oqd = ob.rollout(q[:m], qd[:m], 0, np.ascontiguousarray(tg[:n_tg, :m]), spt, threads)
        dq = np.abs(gq[:m].astype(np.float64) - oq)
        dv = np.abs(gqd[:m].astype(np.float64) - oqd)
        same = np.all(gq[:m] == oq, axis=1) & np.all(gqd[:m] == oqd, axis=1)
        pairs_ok, touching = 0, 0
        for w in range(m):
            oio, _ = ob.world(w).contacts()
            k = int(counts[w])
            g = sorted((int(r[0]), int(r[1]), int(r[4])) for r in cio[w, :k] if r[2])      # fixtureA, fixtureB, pointCount
            o = sorted((int(r[0]), int(r[1]), int(r[4])) for r in oio if r[2])
            pairs_ok += int(g == o)
            touching += len(o)
        out[label] = {"max_abs_dq": float(dq.max()), "max_abs_dqd": float(dv.max()), "mean_abs_dq": float(dq.mean()),
                      "bit_identical_worlds": int(same.sum()), "touching_pair_sets_identical": pairs_ok,
                      "touching_contacts": touching}
    # BASELINE.json: single-step state error <= 1e-5 and bit-exact contact-pair sets / manifold point counts; multi-step
    # trajectories are reported with their divergence statistics (here they are bit-identical as well)
    out["pass"] = bool(out["single_step"]["max_abs_dq"] <= 1e-5 and out["single_step"]["max_abs_dqd"] <= 1e-5 and
                       out["single_step"]["touching_pair_sets_identical"] == m and
                       out["rollout"]["touching_pair_sets_identical"] == m)
    return out


def workload_text(cfg, n):
    return cfg["text"] % n


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference cannot be compiled here
    (Box2D / sim_env / yaml-cpp / Eigen / Boost are absent), so this times the oracle port on all host
    threads.  Each step is a bounded sample of the arm's workload: the same worlds/inputs generator, at most
    `--ref-worlds` worlds per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    limits = load_model(cfg).dof_limits()
    threads = host_threads()
    full = args.worlds or cfg["worlds"]
    n = min(full, args.ref_worlds)
    units, times = cpu_oracle_throughput(cfg, limits, n, args.warmup + args.steps, threads)
    timed = times[args.warmup:]
    total = sum(timed)
    value = units * len(timed) / total
    sample = ("%d of the %d worlds x %d steps x %d timed repeats, one world per thread, %d threads"
              % (n, full, ROLLOUT_STEPS, len(timed), threads))
    line = {
        "impl": "reference", "metric": "world-steps/sec", "value": value, "unit": "world-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(cfg, full), "name": args.config, "worlds_per_step_sampled": n,
                   "physics_steps_per_step": ROLLOUT_STEPS, "dt": 0.01, "velocity_iterations": 15, "position_iterations": 15},
        "cpu_baseline": {"value": value, "unit": "world-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import box2d_sim_env_b200 as b2g
    from box2d_sim_env_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the batched backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = CONFIGS[args.config]
    strong = cfg["scaling"] == "strong"
    n_cfg = args.worlds or cfg["worlds"]
    if strong:   # total work fixed: rank r owns a contiguous slice of the n_cfg worlds
        n_total = n_cfg
        lo, hi = sharding.world_slice(n_total, rank, world)
    else:        # per-GPU work fixed
        n_total = n_cfg * world
        lo, hi = rank * n_cfg, (rank + 1) * n_cfg
    n = hi - lo
    L = b2g.lib()
    model = load_model(cfg)
    batch = b2g.BatchBox2DWorld(n, device=local).loadWorld(model)
    limits = model.dof_limits()
    nq = model.nq
    nd = cfg["robot_dofs"]
    stream = torch.cuda.ExternalStream(batch.stream(), device=dev)

    # device-resident inputs / outputs
    d_q = torch.empty((n, nq), dtype=torch.float32, device=dev)
    d_qd = torch.empty_like(d_q)
    d_tg = torch.empty((N_TARGETS, n, nd), dtype=torch.float32, device=dev)
    d_out = torch.empty((n, 2, nq), dtype=torch.float32, device=dev)     # row w = [q_w | qd_w]
    d_all = torch.empty((n_total, 2, nq), dtype=torch.float32, device=dev) if world > 1 else None
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # pinned host staging for the end-to-end path
    h_q = torch.empty((n, nq), dtype=torch.float32).pin_memory()
    h_qd = torch.empty_like(h_q).pin_memory()
    h_tg = torch.empty((N_TARGETS, n, nd), dtype=torch.float32).pin_memory()
    h_out = torch.empty((n, 2, nq), dtype=torch.float32).pin_memory()
    h_oq = torch.empty((n, nq), dtype=torch.float32).pin_memory()
    h_oqd = torch.empty((n, nq), dtype=torch.float32).pin_memory()
    d_oq = torch.empty((n, nq), dtype=torch.float32, device=dev)
    d_oqd = torch.empty_like(d_oq)

    def load_inputs(step_idx, to_device):
        # weak scaling: every rank draws its own batch; strong scaling: rows lo:hi of the one global batch
        if strong:
            q, qd, tg = synth_inputs(limits, n_total, 1234 + step_idx, cfg, lo, hi)
        else:
            q, qd, tg = synth_inputs(limits, n, 1234 + 1000 * rank + step_idx, cfg)
        h_q.copy_(torch.from_numpy(q))
        h_qd.copy_(torch.from_numpy(qd))
        h_tg.copy_(torch.from_numpy(tg))
        if to_device:
            with torch.cuda.stream(stream):
                d_q.copy_(h_q, non_blocking=True)
                d_qd.copy_(h_qd, non_blocking=True)
                d_tg.copy_(h_tg, non_blocking=True)

    def rollout_device():
        batch.set_state_dev(d_q.data_ptr(), d_qd.data_ptr(), reset_caches=True)
        batch.rollout_dev(0, d_tg.data_ptr(), N_TARGETS, STEPS_PER_TARGET)
        batch.get_state_dev(d_oq.data_ptr(), d_oqd.data_ptr())
        with torch.cuda.stream(stream):
            d_out[:, 0].copy_(d_oq)
            d_out[:, 1].copy_(d_oqd)
            if world > 1:   # the path's only collective: final gather of the rollout states over NCCL
                sharding.gather_states(d_out, n_total, out=d_all)

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_pass(n_steps, e2e, first_idx):
        total_ms = 0.0
        for s in range(n_steps):
            load_inputs(first_idx + s, to_device=not e2e)
            with torch.cuda.stream(stream):
                flush.zero_()  # evict L2 between timed iterations
            barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if e2e and world == 1:
                # the call a planner makes: ONE host-pointer C-ABI call (b2g_batch_rollout) on pinned host buffers --
                # H2D of states and targets, fresh-clone set_state, the rollout kernel, get_state, D2H of the results
                batch.rollout_host(0, h_q.data_ptr(), h_qd.data_ptr(), h_tg.data_ptr(), N_TARGETS, STEPS_PER_TARGET,
                                   h_oq.data_ptr(), h_oqd.data_ptr())
            else:
                if e2e:
                    with torch.cuda.stream(stream):
                        d_q.copy_(h_q, non_blocking=True)
                        d_qd.copy_(h_qd, non_blocking=True)
                        d_tg.copy_(h_tg, non_blocking=True)
                rollout_device()
                if e2e:
                    with torch.cuda.stream(stream):
                        h_out.copy_(d_out, non_blocking=True)
            e1.record(stream)
            barrier()
            total_ms += e0.elapsed_time(e1)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # warm-up (nothing is JIT-compiled: the library is prebuilt), then the timed region
    warm = max(args.warmup, 3)
    timed_pass(warm, False, 0)
    timed_pass(1, True, 50)   # the end-to-end path once, untimed: first use allocates its staging buffers
    batch.take_kernel_ms()
    launches0 = L.b2g_launch_count()
    with ClockSampler(local) as clocks:
        dev_ms = timed_pass(args.steps, False, 100)
        kernel_ms, kernel_launches = batch.take_kernel_ms()
        launches = int(L.b2g_launch_count() - launches0)
        counts, _, _ = batch.contacts(max_per_world=8)
        e2e_ms = timed_pass(args.steps, True, 100)
    status = batch.status()
    extras = {}
    if args.config == "cfg4":   # batched checkCollision on the rolled-out worlds, device-timed
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d_counts = torch.empty(n, dtype=torch.int32, device=dev)
        batch.check_collision_dev(d_counts.data_ptr())
        ev0.record(stream)
        for _ in range(5):
            batch.check_collision_dev(d_counts.data_ptr())
        ev1.record(stream)
        stream.synchronize()
        extras["check_collision_worlds_per_s"] = n * 5 / (ev0.elapsed_time(ev1) * 1e-3)

    units_per_step = n_total * ROLLOUT_STEPS
    value = units_per_step * args.steps / (dev_ms * 1e-3)
    e2e_value = units_per_step * args.steps / (e2e_ms * 1e-3)
    mean_contacts = float(counts.mean())
    peak, peak_kind = measured_peak()
    bytes_per_launch = algorithmic_bytes_per_world_step(mean_contacts, cfg) * n * ROLLOUT_STEPS
    achieved = bytes_per_launch * kernel_launches / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0

    if rank == 0:
        config = {"workload": workload_text(cfg, n_cfg), "name": args.config, "worlds_per_gpu": n, "worlds_total": n_total,
                  "physics_steps_per_step": ROLLOUT_STEPS, "dt": 0.01, "velocity_iterations": 15, "position_iterations": 15,
                  "l2": "flushed between timed iterations (384 MB write)",
                  "status_nonzero_worlds": int((status != 0).sum()), "mean_live_contacts": mean_contacts}
        config.update(extras)
        line = {
            "metric": "world-steps/sec", "value": value, "unit": "world-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "world-steps/s",
                    "h2d_bytes_per_step": int(h_q.numel() * 4 * 2 + h_tg.numel() * 4), "d2h_bytes_per_step": int(h_out.numel() * 4)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": MEASURED_TRAFFIC.get(args.config), "kernel": "k_step", "peak_kind": peak_kind,
                         "sm_issue_active_pct": MEASURED_ISSUE_ACTIVE_PCT.get(args.config),
                         "kernel_ms_per_launch": kernel_ms / max(kernel_launches, 1),
                         "algorithmic_bytes_per_world_step": algorithmic_bytes_per_world_step(mean_contacts, cfg),
                         "note": "FP32 issue / dependent-latency bound, not bandwidth bound: DESIGN.md §3, profiles/"},
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            nref = min(n, args.ref_worlds)
            units, times = cpu_oracle_throughput(cfg, limits, nref, 1 + args.cpu_repeats, threads)
            timed = times[1:]
            line["cpu_baseline"] = {"value": units * len(timed) / sum(timed), "unit": "world-steps/s", "cores": threads,
                                    "kind": "port",
                                    "sample": "%d worlds x %d steps x %d repeats of the same workload, one world per thread"
                                              % (nref, ROLLOUT_STEPS, len(timed))}
            # correctness gate run with the measurement (outside every timed region; the oracle is the checker)
            line["parity"] = parity_gate(cfg, batch, limits, n, 1234 + 100 + args.steps - 1, threads)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of one rollout launch (k_step / k_step_narrow) from `ncu --set full`, bytes per
# launch of the config's default world count (profiles/r1e_<config>_<worlds>_k_step_ncu_full.txt)
# smsp__issue_active.avg.pct_of_peak_sustained_active of the same captures: the path is bound by issue slots / dependent
# latency, not by HBM, so the SM issue utilisation is reported next to the HBM fraction
MEASURED_ISSUE_ACTIVE_PCT = {"cfg2": 21.6, "cfg3": 38.9, "cfg4": 26.7}

MEASURED_TRAFFIC = {
    # 4096 worlds: the 32 MB state stays L2-resident across the 100 steps of a launch, DRAM sees far less than the
    # algorithmic bytes (6.44 MB read + 0.46 MB written; profiles/r1f_cfg2_4096_k_step_ncu_full.txt)
    "cfg2": 6444800 + 456960,
    # 65536 rigid-robot worlds: 31.6 MB read + 24.5 MB written -- the touched part of the state stays in L2 as well
    "cfg3": 31556352 + 24489728,
    # 65536 clutter worlds: the ~9 KB a step touches per world (590 MB for the batch) does not fit the 126 MB L2, every
    # step streams it: 69.1 GB read + 36.1 GB written per 6.55 M world-steps = 16 KB per world-step (algorithmic: 5.0 KB)
    "cfg4": 69142817000 + 36054000000,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json config (default: the one the "
                    "metric is quoted on, test_env.yaml x 4096 worlds)")
    ap.add_argument("--worlds", type=int, default=0, help="override the config's world count (per GPU; total for cfg5)")
    ap.add_argument("--ref-worlds", type=int, default=4096, help="worlds per step of the CPU oracle sample")
    ap.add_argument("--cpu-repeats", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
