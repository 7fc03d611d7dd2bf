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

# BASELINE.json configs -> concrete workloads.  `bytes` = algorithmic bytes per world-step
# (fixed part, per live contact), DESIGN.md §2.
CONFIGS = {
    "cfg2": dict(env="test_env.yaml", worlds=4096, robot_dofs=7, bytes=(940.0, 64.0), scaling="weak", obstacle=(0.0, 1.2, 0.0),
                 text="test_env.yaml x %d worlds per GPU x 100 physics steps per step, random robot velocity targets "
                      "resampled every 10 steps (BASELINE.json configs[1])"),
    # the round-1 inputs of cfg2 (box 1.3 m ahead of the robot, obstacle at its YAML pose: next to no contact within a rollout);
    # rides along with the default run as the `r1_inputs` sub-record so that the two rounds' numbers stay comparable
    "cfg2r1": dict(env="test_env.yaml", worlds=4096, robot_dofs=7, bytes=(940.0, 64.0), scaling="weak", placement="far",
                   text="test_env.yaml x %d worlds per GPU x 100 physics steps per step, random robot velocity targets, round-1 "
                        "inputs (box 1.3 m ahead of the robot: contact-free)"),
    "cfg3": dict(env="test_env_rigid_robot.yaml", worlds=65536, robot_dofs=3, bytes=(348.0, 64.0), scaling="weak", obstacle=(0.0, 1.1, 0.0),
                 text="test_env_rigid_robot.yaml x %d worlds per GPU x 100 physics steps per step, rigid robot pushing with "
                      "ground friction joints, random base velocity targets resampled every 10 steps (BASELINE.json configs[2])"),
    "cfg4": dict(env="clutter", worlds=65536, robot_dofs=7, bytes=(4660.0, 64.0), scaling="weak",
                 text="synthetic clutter (robot + 32 convex polygons, seed 42) x %d worlds per GPU x 100 physics steps per step "
                      "(propagate), batched checkCollision timed separately (BASELINE.json configs[3])"),
    "cfg5": dict(env="test_env.yaml", worlds=1048576, robot_dofs=7, bytes=(940.0, 64.0), scaling="strong", obstacle=(0.0, 1.2, 0.0),
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


# Top edge of the robot base's rectangle in the robot frame (the robot pose is the base link's centre of mass,
# `use_center_of_mass: true`): the 0.8 m x 0.2 m rectangle's upper side, y = 0.2 m - com_y
BASE_TOP_Y = {"test_env.yaml": 0.2 + 0.016667, "test_env_rigid_robot.yaml": 0.2 - 0.046667}
OBJ_HALF = 0.05   # obj1 is a 0.1 m x 0.1 m box whose pose is its lower-left corner


def place_obstacle(target, cfg, n):
    """The static hinged obstacle obj2 goes half a metre in front of the finger tips (the YAML leaves it 3 m away), so that
    the pushed box and the faster robots run into a STATIC body within the rollout: dynamic-vs-static contacts are what
    Box2D's continuous step (TOI sub-stepping) works on.  About one world in twenty has a TOI event per rollout and 0.3
    b2TimeOfImpact evaluations per world; DESIGN.md §3 lists how the throughput moves with the distance (the TOI path is the
    most divergent part of the step on a GPU).  A static object's pose is part of the world state a clone keeps, so it is
    set once per batch.  `target`: a BatchBox2DWorld, a MultiGpuBatch or an OracleBatch."""
    pose = cfg.get("obstacle")
    if pose is None:
        return
    if hasattr(target, "setObjectPose"):
        target.setObjectPose("obj2", np.tile(np.asarray(pose, np.float32), (n, 1)))
    else:   # OracleBatch: creation order robot, obj1, obj2
        for w in range(n):
            target.world(w).set_object_pose(2, *[float(x) for x in pose])


def robot_velocity_limits(env):
    """[robot_dofs, 2] velocity limits of the first robot of an environment description (plain dicts): base x, y, theta,
    then the joints in file order -- the same rows as Model.dof_limits()[:, 2:4], without touching the CUDA library."""
    r = env["robots"][0]
    ba = r["base_actuation"]
    rows = [ba["translational_velocity_limits"], ba["translational_velocity_limits"], ba["rotational_velocity_limits"]]
    rows += [j["velocity_limits"] for j in r["joints"]]
    return np.array(rows, np.float32)


def synth_inputs(vel_limits, n, seed, cfg=None, lo=0, hi=None):
    """Inputs of one bench step for worlds lo:hi of an n-world batch.  Counter-based (Philox) and indexed
    by world, so the CPU and GPU arms -- and any GPU count -- draw identical inputs for the same world.

    cfg2 / cfg3 / cfg5 (robot + free box obj1 + hinged obstacle): the box starts in the robot's palm, between the fingers
    and in front of the base -- a quarter of the worlds with the box resting against the base's front edge (within the
    2 mm polygon skin: touching from the first step), the rest up to 0.3 m ahead with a random heading -- and the robot
    pushes: targets ~ U(velocity limits), resampled every 10 steps, with the base's forward component folded to >= 0.
    The rollout therefore runs the narrow phase, the contact solver and restitution on most worlds (the round-1
    placement left the box 1.3 m away: 0 touching contacts in 1024 worlds)."""
    cfg = cfg or CONFIGS["cfg2"]
    hi = n if hi is None else hi
    nd = cfg["robot_dofs"]
    vel_limits = np.asarray(vel_limits, np.float32)
    if vel_limits.shape[1] == 6:          # Model.dof_limits() rows: position, velocity, acceleration (lo, hi)
        vel_limits = vel_limits[:, 2:4]
    rng = np.random.Generator(np.random.Philox(key=seed))
    if cfg["env"] in BASE_TOP_Y:
        ytop = BASE_TOP_Y[cfg["env"]]
        resting = rng.uniform(0.0, 1.0, n) < 0.25
        cx = rng.uniform(-0.17, 0.17, n)
        cy = rng.uniform(ytop + 0.08, ytop + 0.30, n)
        th = rng.uniform(-np.pi, np.pi, n)
        gap = rng.uniform(0.0002, 0.0015, n)
        cy = np.where(resting, ytop + gap + OBJ_HALF, cy)
        th = np.where(resting, 0.0, th)
        if cfg.get("placement") == "far":
            cy = np.full(n, ytop + 1.3)
        c, s = np.cos(th), np.sin(th)
        first = 7 if nd == 7 else 3      # robot DOFs come first, then obj1 (x, y, theta), then obj2's hinge
        q = np.zeros((n, first + 4), np.float32)
        q[:, first] = cx - (c * OBJ_HALF - s * OBJ_HALF)
        q[:, first + 1] = cy - (s * OBJ_HALF + c * OBJ_HALF)
        q[:, first + 2] = th
        qd = np.zeros_like(q)
    else:
        from box2d_sim_env_b200 import envgen
        q, qd = envgen.clutter_states(n, seed=seed)
    tg = rng.uniform(vel_limits[:nd, 0], vel_limits[:nd, 1], (N_TARGETS, n, nd)).astype(np.float32)
    if cfg["env"] in BASE_TOP_Y:
        tg[:, :, 1] = np.abs(tg[:, :, 1])   # the robot pushes forward
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
    place_obstacle(batch, cfg, n_worlds)
    times = []
    for r in range(repeats):
        q, qd, tg = synth_inputs(limits, n_worlds, seed0 + r, cfg)
        # a fresh batch per repeat would also time construction; the reference's clone() is not part of the metric,
        # so worlds are reused and only set_state + stepping is timed (setWorldState on a live world)
        t0 = time.perf_counter()
        batch.rollout(q, qd, 0, tg, STEPS_PER_TARGET, threads)
        times.append(time.perf_counter() - t0)
    return n_worlds * ROLLOUT_STEPS, times


def _touching_sets(gcounts, gio, oracle_batch, m):
    """Per world: the sorted (fixtureA, fixtureB, manifold point count) tuples of the touching contacts, CUDA vs oracle."""
    same, touching = 0, 0
    for w in range(m):
        oio, _ = oracle_batch.world(w).contacts()
        k = int(gcounts[w])
        g = sorted((int(r[0]), int(r[1]), int(r[4])) for r in gio[w, :k] if r[2])
        o = sorted((int(r[0]), int(r[1]), int(r[4])) for r in oio if r[2])
        same += int(g == o)
        touching += len(o)
    return same, touching


def parity_gate(cfg, batch, limits, n, seed, threads, sample=1024):
    """Correctness gate that accompanies the measurement (BASELINE.json).  The first `sample` worlds of one
    bench step's inputs go through the CUDA path and through the CPU oracle (here as the checker):

      single_step  one step from the identical fresh-cache state: DOF positions / velocities (max |diff|, worlds that are
                   bit-identical) and the touching-pair sets with their manifold point counts;
      rollout      the same comparison at every checkpoint of the 100-step rollout (after each 10-step target segment);
      vs_libm      the divergence statistics the north star asks for: the same rollout on the oracle built with libm
                   sinf / cosf in b2Rot (what upstream Box2D calls; the CUDA path and the oracle proper share one fixed
                   sin / cos sequence, so their difference is 0 by construction) -- per checkpoint max / mean |dq|, |dqd|
                   and the number of worlds whose touching-pair sets still agree.

    The gate FAILS when it is vacuous: no touching contact after the single step, or fewer than one touching contact per
    8 worlds summed over the rollout's checkpoints."""
    from oracle import b2o
    m = min(sample, n)
    q, qd, tg = synth_inputs(limits, n, seed, cfg)
    env_desc = oracle_env(cfg)
    env = b2o.OracleEnv(env_desc)
    out = {"worlds": m, "physics_steps": ROLLOUT_STEPS}

    def compare(gq, gqd, oq, oqd):
        dq = np.abs(gq[:m].astype(np.float64) - oq)
        dv = np.abs(gqd[:m].astype(np.float64) - oqd)
        same = np.all(gq[:m] == oq, axis=1) & np.all(gqd[:m] == oqd, axis=1)
        return {"max_abs_dq": float(dq.max()), "max_abs_dqd": float(dv.max()), "mean_abs_dq": float(dq.mean()),
                "mean_abs_dqd": float(dv.mean()), "bit_identical_worlds": int(same.sum())}

    # ---- one step from the fresh-cache state
    ob = b2o.OracleBatch(env, m)   # freshly loaded worlds: the CUDA side starts from fresh clones (reset_caches) too
    place_obstacle(ob, cfg, m)
    gq, gqd = batch.rollout(0, tg[:1], 1, q, qd, reset_caches=True)
    counts, cio, _ = batch.contacts(max_per_world=64)
    oq, oqd = ob.rollout(q[:m], qd[:m], 0, np.ascontiguousarray(tg[:1, :m]), 1, threads)
    single = compare(gq, gqd, oq, oqd)
    single["touching_pair_sets_identical"], single["touching_contacts"] = _touching_sets(counts, cio, ob, m)
    out["single_step"] = single

    # ---- the rollout, compared at every checkpoint; the libm-trig oracle runs alongside
    lm = b2o.load_variant("_libm")
    ob = b2o.OracleBatch(env, m)
    ol = lm.OracleBatch(lm.OracleEnv(env_desc), m)
    place_obstacle(ob, cfg, m)
    place_obstacle(ol, cfg, m)
    roll = {"checkpoint_steps": [], "touching_contacts_at_checkpoint": [], "touching_pair_sets_identical_at_checkpoint": [],
            "max_abs_dq": 0.0, "max_abs_dqd": 0.0, "bit_identical_worlds": m}
    libm = {"checkpoint_steps": [], "max_abs_dq": [], "mean_abs_dq": [], "max_abs_dqd": [], "mean_abs_dqd": [],
            "touching_pair_sets_identical": []}
    mid = None
    for t in range(N_TARGETS):
        first = t == 0
        gq, gqd = batch.rollout(0, tg[t:t + 1], STEPS_PER_TARGET, q if first else None, qd if first else None,
                                reset_caches=first)
        counts, cio, _ = batch.contacts(max_per_world=64)
        seg = np.ascontiguousarray(tg[t:t + 1, :m])
        oq, oqd = ob.rollout(q[:m] if first else None, qd[:m] if first else None, 0, seg, STEPS_PER_TARGET, threads)
        lq, lqd = ol.rollout(q[:m] if first else None, qd[:m] if first else None, 0, seg, STEPS_PER_TARGET, threads)
        c = compare(gq, gqd, oq, oqd)
        if first:
            mid = (oq.copy(), oqd.copy())
        same, touching = _touching_sets(counts, cio, ob, m)
        roll["checkpoint_steps"].append((t + 1) * STEPS_PER_TARGET)
        roll["touching_contacts_at_checkpoint"].append(touching)
        roll["touching_pair_sets_identical_at_checkpoint"].append(same)
        roll["max_abs_dq"] = max(roll["max_abs_dq"], c["max_abs_dq"])
        roll["max_abs_dqd"] = max(roll["max_abs_dqd"], c["max_abs_dqd"])
        roll["bit_identical_worlds"] = min(roll["bit_identical_worlds"], c["bit_identical_worlds"])
        d = compare(gq, gqd, lq, lqd)
        libm["checkpoint_steps"].append((t + 1) * STEPS_PER_TARGET)
        for k in ("max_abs_dq", "mean_abs_dq", "max_abs_dqd", "mean_abs_dqd"):
            libm[k].append(d[k])
        libm["touching_pair_sets_identical"].append(_touching_sets(counts, cio, ol, m)[0])
    roll["touching_contacts"] = int(sum(roll["touching_contacts_at_checkpoint"]))
    roll["touching_pair_sets_identical"] = int(min(roll["touching_pair_sets_identical_at_checkpoint"]))
    stats = [ob.world(w).stats() for w in range(m)]
    roll["toi_events"] = int(sum(x["toi_events"] for x in stats))
    roll["toi_calls"] = int(sum(x["toi_calls"] for x in stats))
    out["rollout"] = roll

    # single step against the libm-trig oracle: the north star's 1e-5 bound against reference-like trigonometry
    ol1 = lm.OracleBatch(lm.OracleEnv(env_desc), m)
    place_obstacle(ol1, cfg, m)
    gq, gqd = batch.rollout(0, tg[:1], 1, q, qd, reset_caches=True)
    counts, cio, _ = batch.contacts(max_per_world=64)
    lq, lqd = ol1.rollout(q[:m], qd[:m], 0, np.ascontiguousarray(tg[:1, :m]), 1, threads)
    l1 = compare(gq, gqd, lq, lqd)
    l1["touching_pair_sets_identical"] = _touching_sets(counts, cio, ol1, m)[0]
    libm["single_step"] = l1
    libm["what"] = "CUDA path vs the oracle rebuilt with libm sinf/cosf in b2Rot (upstream Box2D's calls)"
    out["rollout_vs_libm"] = libm

    # A config whose synthetic states start without overlaps (the clutter scene) has no touching contact after one step from
    # the inputs: there the one-step comparison is repeated from the oracle's state after the first target segment (fresh
    # caches on both sides, bodies already pressed together), so that the single-step bound is checked on live manifolds too.
    if single["touching_contacts"] == 0 and mid is not None:
        q1, qd1 = q.copy(), qd.copy()
        q1[:m], qd1[:m] = mid
        ob = b2o.OracleBatch(env, m)
        place_obstacle(ob, cfg, m)
        gq, gqd = batch.rollout(0, tg[1:2], 1, q1, qd1, reset_caches=True)
        counts, cio, _ = batch.contacts(max_per_world=64)
        oq, oqd = ob.rollout(q1[:m], qd1[:m], 0, np.ascontiguousarray(tg[1:2, :m]), 1, threads)
        single = compare(gq, gqd, oq, oqd)
        single["touching_pair_sets_identical"], single["touching_contacts"] = _touching_sets(counts, cio, ob, m)
        single["from"] = "the oracle's state after the first %d steps (the input states have no overlaps)" % STEPS_PER_TARGET
        out["single_step"] = single
    nonvacuous = single["touching_contacts"] > 0 and roll["touching_contacts"] * 8 >= m
    out["nonvacuous"] = bool(nonvacuous)
    # BASELINE.json: single-step state error <= 1e-5 and bit-exact contact-pair sets / manifold point counts
    out["pass"] = bool(nonvacuous and single["max_abs_dq"] <= 1e-5 and single["max_abs_dqd"] <= 1e-5 and
                       single["touching_pair_sets_identical"] == m and roll["touching_pair_sets_identical"] == m and
                       l1["max_abs_dq"] <= 1e-5 and l1["max_abs_dqd"] <= 1e-5)
    return out


def workload_text(cfg, n):
    return cfg["text"] % n


def config_dict(cfg, name, worlds_per_gpu_cfg, worlds_total):
    """The `config` object both arms print (identical keys and values for the same command line)."""
    return {"workload": workload_text(cfg, worlds_per_gpu_cfg), "name": name, "worlds_total": worlds_total,
            "physics_steps_per_step": ROLLOUT_STEPS, "dt": 0.01, "velocity_iterations": 15, "position_iterations": 15,
            "l2": "flushed between timed iterations (384 MB write)",
            "inputs": "box in the robot's palm (1/4 resting against the base), static obstacle ahead, forward-pushing random velocity targets"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference cannot be compiled here
    (Box2D / sim_env / yaml-cpp / Eigen / Boost are absent), so this times the oracle port on all host
    threads.  Each step is a bounded sample of the arm's workload: the same worlds/inputs generator, at most
    `--ref-worlds` worlds per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    limits = robot_velocity_limits(oracle_env(cfg))   # from the environment description: the CUDA library is never loaded here
    threads = host_threads()
    full = args.worlds or cfg["worlds"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
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
        "config": config_dict(cfg, args.config, full, full if cfg["scaling"] == "strong" else full * world),
        "cpu_baseline": {"value": value, "unit": "world-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "sampled": {"worlds_per_step_sampled": n},
    }
    print(json.dumps(line))


def measure(args, cfg_name, n_cfg, steps, warmup, ctx, with_cpu):
    """One benchmark record of config `cfg_name` on this rank's GPU (all ranks call it together).  Returns the JSON-line dict on
    rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    import box2d_sim_env_b200 as b2g
    from box2d_sim_env_b200 import sharding

    world, rank, local, dev = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"]
    cfg = CONFIGS[cfg_name]
    strong = cfg["scaling"] == "strong"
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
    place_obstacle(batch, cfg, n)
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
    flush = ctx["flush"]
    # pinned host staging for the end-to-end path
    h_q = torch.empty((n, nq), dtype=torch.float32).pin_memory()
    h_qd = torch.empty_like(h_q).pin_memory()
    h_tg = torch.empty((N_TARGETS, n, nd), dtype=torch.float32).pin_memory()
    h_oq = torch.empty((n, nq), dtype=torch.float32).pin_memory()
    h_oqd = torch.empty((n, nq), dtype=torch.float32).pin_memory()
    d_oq = torch.empty((n, nq), dtype=torch.float32, device=dev)
    d_oqd = torch.empty_like(d_oq)
    # several GPUs, end to end: the ONE call a C++ planner makes on a multi-GPU node, b2g_multi_rollout (include/b2g.h), issued
    # by rank 0 for all `world` devices of the node -- host buffers for the whole job in, all-gathered states out; the other
    # ranks wait at the barrier
    multi = None
    if world > 1 and rank == 0:
        multi = b2g.MultiGpuBatch(model, n_total, list(range(world)))
        place_obstacle(multi, cfg, n_total)
        m_q = torch.empty((n_total, nq), dtype=torch.float32).pin_memory()
        m_qd = torch.empty_like(m_q).pin_memory()
        m_tg = torch.empty((N_TARGETS, n_total, nd), dtype=torch.float32).pin_memory()
        m_oq = torch.empty((n_total, nq), dtype=torch.float32).pin_memory()
        m_oqd = torch.empty_like(m_oq).pin_memory()
        m_stream = torch.cuda.ExternalStream(L.b2g_batch_stream(multi.shard(0)[0]), device=torch.device("cuda", 0))
    gather_events = []

    def rows(step_idx, first, last):
        # weak scaling: every rank draws its own batch; strong scaling: rows of the one global batch
        if strong:
            return synth_inputs(limits, n_total, 1234 + step_idx, cfg, first, last)
        parts = [synth_inputs(limits, n_cfg, 1234 + 1000 * r + step_idx, cfg) for r in range(first // n_cfg, (last + n_cfg - 1) // n_cfg)]
        return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts], axis=1))

    def load_inputs(step_idx, to_device):
        q, qd, tg = rows(step_idx, lo, hi)
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
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
                sharding.gather_states(d_out, n_total, out=d_all)
                g1.record(stream)
                gather_events.append((g0, g1))

    def barrier(host_only=False):
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            # host_only: a gloo barrier.  While rank 0 drives every GPU of the node through b2g_multi_rollout the other ranks
            # must not park an NCCL barrier kernel on their GPU (it would time-slice with rank 0's rollout kernel there)
            dist.barrier(group=ctx["host_pg"]) if host_only else dist.barrier()
        torch.cuda.synchronize()

    def timed_pass(n_steps, e2e, first_idx):
        total_ms = 0.0
        for s in range(n_steps):
            if e2e and world > 1:
                if rank == 0:
                    q, qd, tg = rows(first_idx + s, 0, n_total)
                    m_q.copy_(torch.from_numpy(q))
                    m_qd.copy_(torch.from_numpy(qd))
                    m_tg.copy_(torch.from_numpy(tg))
            else:
                load_inputs(first_idx + s, to_device=not e2e)
            with torch.cuda.stream(stream):
                flush.zero_()  # evict L2 between timed iterations
            barrier(host_only=e2e and world > 1)
            if e2e and world > 1:
                ms = 0.0
                if rank == 0:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(m_stream)
                    multi.rollout_host(0, m_q.data_ptr(), m_qd.data_ptr(), m_tg.data_ptr(), N_TARGETS, STEPS_PER_TARGET,
                                       m_oq.data_ptr(), m_oqd.data_ptr())
                    e1.record(m_stream)
                    m_stream.synchronize()
                    ms = e0.elapsed_time(e1)
                barrier(host_only=True)
                total_ms += ms
                continue
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if e2e:
                # the call a planner makes: ONE host-pointer C-ABI call (b2g_batch_rollout) on pinned host buffers --
                # H2D of states and targets, fresh-clone set_state, the rollout kernel, get_state, D2H of the results
                batch.rollout_host(0, h_q.data_ptr(), h_qd.data_ptr(), h_tg.data_ptr(), N_TARGETS, STEPS_PER_TARGET,
                                   h_oq.data_ptr(), h_oqd.data_ptr())
            else:
                rollout_device()
            e1.record(stream)
            barrier()
            total_ms += e0.elapsed_time(e1)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # warm-up (nothing is JIT-compiled: the library is prebuilt), then the timed region
    warm = max(warmup, 3) if cfg_name != "cfg5" or args.config == "cfg5" else max(warmup, 1)
    timed_pass(warm, False, 0)
    timed_pass(1, True, 50)   # the end-to-end path once, untimed: first use allocates its staging buffers
    batch.take_kernel_ms()
    del gather_events[:]
    launches0 = L.b2g_launch_count()
    with ClockSampler(local) as clocks:
        dev_ms = timed_pass(steps, False, 100)
        kernel_ms, kernel_launches = batch.take_kernel_ms()
        launches = int(L.b2g_launch_count() - launches0)
        counts, cio_end, _ = batch.contacts(max_per_world=16)
        touching_per_world = float(cio_end[..., 2].sum()) / max(n, 1)
        gather_ms = sum(a.elapsed_time(b) for a, b in gather_events) / max(len(gather_events), 1) if gather_events else None
        e2e_ms = timed_pass(steps, True, 100)
    status = batch.status()
    multi_info = None
    if world > 1:
        # where the step's time goes on N GPUs: the rollout kernel of every rank (min / max over ranks) and the all-gather,
        # which also waits for the slowest rank's kernel
        k = torch.tensor([kernel_ms / max(kernel_launches, 1)], dtype=torch.float64, device=dev)
        kmin, kmax, g = k.clone(), k.clone(), torch.tensor([gather_ms or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(kmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(g, op=dist.ReduceOp.MAX)
        multi_info = {"kernel_ms_per_launch_min_over_ranks": float(kmin.item()), "kernel_ms_per_launch_max_over_ranks": float(kmax.item()),
                      "gather_ms_max_over_ranks": float(g.item()), "gather_bytes": int(n_total * 2 * nq * 4),
                      "collective": "ncclAllGather of [worlds_total, 2, nq] f32 on the batch stream (torch.distributed, one process per GPU)",
                      "e2e_path": "b2g_multi_rollout (C ABI, one process drives all GPUs: H2D of every slice, kernels, ncclAllGather, D2H)"}
        if multi is not None:
            multi_info["c_abi_last_timing"] = multi.last_timing()
            multi.close()
    extras = {}
    if cfg_name == "cfg4":   # batched checkCollision on the rolled-out worlds, device-timed
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
    value = units_per_step * steps / (dev_ms * 1e-3)
    e2e_value = units_per_step * steps / (e2e_ms * 1e-3)
    mean_contacts = float(counts.mean())
    peak, peak_kind = measured_peak()
    bytes_per_launch = algorithmic_bytes_per_world_step(mean_contacts, cfg) * n * ROLLOUT_STEPS
    achieved = bytes_per_launch * kernel_launches / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0

    line = None
    if rank == 0:
        prof = profile_metrics(cfg_name, n)
        clock_summary = clocks.summary()
        config = config_dict(cfg, cfg_name, n_cfg, n_total)
        run_info = {"worlds_per_gpu": n, "status_nonzero_worlds": int((status != 0).sum()), "mean_live_contacts": mean_contacts,
                    "touching_contacts_per_world": touching_per_world}
        run_info.update(extras)
        if world > 1:
            h2d = int((2 * n_total * nq + N_TARGETS * n_total * nd) * 4)
            d2h = int(2 * n_total * nq * 4)
        else:
            h2d = int(h_q.numel() * 4 * 2 + h_tg.numel() * 4)
            d2h = int((h_oq.numel() + h_oqd.numel()) * 4)
        line = {
            "metric": "world-steps/sec", "value": value, "unit": "world-steps/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "world-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / steps},
            "gpu_launches": launches, "run": run_info,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": prof["traffic"] if prof else None, "kernel": "k_step", "peak_kind": peak_kind,
                         "source": prof["source"] if prof else None,
                         "sm_issue_active_pct": prof["sm_issue_active_pct"] if prof else None,
                         "kernel_ms_per_launch": kernel_ms / max(kernel_launches, 1),
                         "algorithmic_bytes_per_world_step": algorithmic_bytes_per_world_step(mean_contacts, cfg),
                         "note": "FP32 issue / dependent-latency bound, not bandwidth bound: DESIGN.md §3, profiles/"},
            "clocks": clock_summary,
        }
        if multi_info:
            line["multi_gpu"] = multi_info
        if prof:
            # the roof this path actually sits under: FP32 lane-issue slots.  thread instructions per world-step (ncu capture
            # named in `source`) x measured world-steps/s of the kernel, against 148 SMs x 128 lanes x SM clock under load
            mhz = clock_summary.get("sm_mhz") or clock_summary.get("sm_max_mhz") or 1965.0
            lane_peak = 148 * 128 * mhz * 1e6
            kernel_rate = n * ROLLOUT_STEPS * kernel_launches / (kernel_ms * 1e-3) if kernel_ms > 0 else 0.0
            line["instruction_roofline"] = {
                "thread_instructions_per_world_step": prof["thread_instructions_per_world_step"],
                "lanes_active_per_instruction": prof["lanes_active_per_instruction"],
                "achieved_thread_instructions_per_s": prof["thread_instructions_per_world_step"] * kernel_rate,
                "peak_thread_instructions_per_s": lane_peak, "sm_mhz": mhz,
                "frac": prof["thread_instructions_per_world_step"] * kernel_rate / lane_peak, "source": prof["source"]}
        if with_cpu and world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            nref = min(n, args.ref_worlds)
            units, times = cpu_oracle_throughput(cfg, limits, nref, 1 + args.cpu_repeats, threads)
            timed = times[1:]
            line["cpu_baseline"] = {"value": units * len(timed) / sum(timed), "unit": "world-steps/s", "cores": threads,
                                    "kind": "port",
                                    "sample": "%d worlds x %d steps x %d repeats of the same workload, one world per thread"
                                              % (nref, ROLLOUT_STEPS, len(timed))}
            # correctness gate run with the measurement (outside every timed region; the oracle is the checker)
            line["parity"] = parity_gate(cfg, batch, limits, n, 1234 + 100 + steps - 1, threads)
    # the batch (and its stream, which torch's allocator has seen) stays alive until the process ends
    ctx.setdefault("keep", []).append((batch, stream))
    return line


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the batched backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_pg = dist.new_group(backend="gloo")
    ctx = {"world": world, "rank": rank, "local": local, "dev": dev, "host_pg": host_pg,
           "flush": torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)}  # > 126 MB L2
    cfg = CONFIGS[args.config]
    line = measure(args, args.config, args.worlds or cfg["worlds"], args.steps, args.warmup, ctx, with_cpu=True)
    if args.config == "cfg2" and not args.no_cfg5 and not args.worlds:
        # BASELINE.json configs[4] rides along: the planner sweep (1 Mi worlds in total, strong scaling, states all-gathered),
        # 2 timed steps, so that the driver's 1 / 2 / 4 / 8-GPU runs record it next to the weak-scaling headline
        sub = measure(args, "cfg5", CONFIGS["cfg5"]["worlds"], 2, 1, ctx, with_cpu=False)
        if rank == 0 and sub:
            keep = ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config", "e2e", "run", "roofline", "multi_gpu")
            line["cfg5"] = {k: sub[k] for k in keep if k in sub}
    if args.config == "cfg2" and world == 1 and not args.worlds and not args.no_cfg5:
        sub = measure(args, "cfg2r1", CONFIGS["cfg2r1"]["worlds"], 3, 3, ctx, with_cpu=False)
        if sub:
            line["r1_inputs"] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "run", "e2e") if k in sub}
            line["r1_inputs"]["inputs"] = "round-1 placement: box 1.3 m ahead of the robot, obstacle at its YAML pose (BENCH_r01: 5.06e7 world-steps/s)"
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def profile_metrics(config_name, worlds):
    """Counters of the stepping kernel from the newest committed `ncu --set full` summary of this config and world count
    (profiles/r<round><letter>_<config>_<worlds>_k_step_ncu_full.txt, written by tools/ncu_profile_summary.py from ONE rollout
    launch = 100 physics steps per world).  Returns None when no capture matches -- the bench line then prints null instead
    of a stale constant."""
    import glob
    import re
    pat = os.path.join(ROOT, "profiles", "r*_%s_%d_k_step_ncu_full.txt" % (config_name, worlds))

    def order(path):  # r1e < r1f < r2a < r10a: (round number, letters)
        m = re.match(r"r(\d+)([a-z]*)_", os.path.basename(path))
        return (int(m.group(1)), m.group(2)) if m else (0, "")

    files = sorted(glob.glob(pat), key=order)
    if not files:
        return None
    vals = {}
    for line in open(files[-1]):
        f = line.split()
        if len(f) >= 2 and ("__" in f[0]):
            try:
                v = float(f[-1])
            except ValueError:
                continue
            vals[f[0]] = v * _UNIT.get(f[1], 1.0) if len(f) == 3 else v
    need = ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active")
    if any(k not in vals for k in need):
        return None
    world_steps = float(worlds) * ROLLOUT_STEPS
    return {"source": os.path.relpath(files[-1], ROOT),
            "traffic": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
            "sm_issue_active_pct": vals["smsp__issue_active.avg.pct_of_peak_sustained_active"],
            "lanes_active_per_instruction": vals["smsp__thread_inst_executed_per_inst_executed.ratio"],
            "thread_instructions_per_world_step": vals["smsp__inst_executed.sum"] *
                                                  vals["smsp__thread_inst_executed_per_inst_executed.ratio"] / world_steps,
            "warp_instructions_per_launch": vals["smsp__inst_executed.sum"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(k for k in CONFIGS if k != "cfg2r1"), help="BASELINE.json config (default: the one the "
                    "metric is quoted on, test_env.yaml x 4096 worlds)")
    ap.add_argument("--worlds", type=int, default=0, help="override the config's world count (per GPU; total for cfg5)")
    ap.add_argument("--ref-worlds", type=int, default=4096, help="worlds per step of the CPU oracle sample")
    ap.add_argument("--cpu-repeats", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the cfg5 (1 Mi-world sweep) sub-record of the default run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
