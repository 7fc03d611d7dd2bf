#!/usr/bin/env python
"""bench.py — world-steps/s of the batched Box2D hot path on N B200s (BASELINE.json metric).

One bench "step" = one planner rollout over a batch of synthetic worlds:
    set_state (fresh caches)  ->  10 x [setTargetVelocity ; stepPhysics(10)]  ->  get_state
on `--worlds` copies of test_env.yaml (BASELINE.json configs[1]: test_env.yaml x 4096 with random robot velocity
controls, 1 B200).  With --gpus N every rank owns its own batch of `--worlds` worlds (weak scaling) and the final
[n, 22] states are all-gathered over NCCL, which is the only collective the path has.

    value     device-timed throughput with inputs resident in HBM (CUDA events on the batch's stream)
    e2e       same metric through the host-pointer path: pinned host buffers, H2D of the step's inputs and D2H of
              its results inside the timed region
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

ENV_NAME = "test_env.yaml"
STEPS_PER_TARGET = 10
N_TARGETS = 10
ROLLOUT_STEPS = STEPS_PER_TARGET * N_TARGETS
BASE_Q = np.array([0, 0, 0, 0, 0, 0, 0, 0.9, 1.0, -1.0, 0.0], np.float32)


def synth_inputs(limits, n, seed):
    """cfg2 inputs: obj1 pose jittered U(+-0.2 m, +-pi), robot targets ~ U(velocity limits)
    resampled every 10 steps; counter-based so CPU and GPU arms draw identical controls."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    q = np.tile(BASE_Q, (n, 1))
    q[:, 7:9] += rng.uniform(-0.2, 0.2, (n, 2)).astype(np.float32)
    q[:, 9] = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    qd = np.zeros_like(q)
    tg = rng.uniform(limits[:7, 2], limits[:7, 3], (N_TARGETS, n, 7)).astype(np.float32)
    return q, qd, tg


def algorithmic_bytes_per_world_step(mean_contacts):
    # SURVEY.md §8(d), shape A: 6 dynamic bodies x 64 B + 6 active friction joints x 24 B + 4 active revolute
    # joints x 40 B + 7 fixtures on dynamic bodies x 32 B + 28 B of controls + 64 B per live contact
    return 940.0 + 64.0 * mean_contacts


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


def cpu_oracle_throughput(limits, n_worlds, repeats, threads, seed0=9000):
    """CPU oracle stepping the same workload, one world per thread-pool slot."""
    from oracle import b2o
    env = b2o.load_env_yaml(os.path.join(ROOT, "box2d_sim_env_b200", "data", ENV_NAME))
    batch = b2o.OracleBatch(env, n_worlds)
    times = []
    for r in range(repeats):
        q, qd, tg = synth_inputs(limits, n_worlds, seed0 + r)
        # a fresh batch per repeat would also time construction; the reference's clone() is not part of the metric,
        # so worlds are reused and only set_state + stepping is timed (setWorldState on a live world)
        t0 = time.perf_counter()
        batch.rollout(q, qd, 0, tg, STEPS_PER_TARGET, threads)
        times.append(time.perf_counter() - t0)
    return n_worlds * ROLLOUT_STEPS, times


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference cannot be compiled here
    (Box2D / sim_env / yaml-cpp / Eigen / Boost are absent), so this times the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import box2d_sim_env_b200 as b2g
    limits = b2g.Model(b2g.data_path(ENV_NAME)).dof_limits()
    threads = host_threads()
    n = args.ref_worlds
    units, times = cpu_oracle_throughput(limits, n, args.warmup + args.steps, threads)
    timed = times[args.warmup:]
    total = sum(timed)
    value = units * len(timed) / total
    line = {
        "impl": "reference", "metric": "world-steps/sec", "value": value, "unit": "world-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "test_env.yaml x %d worlds x %d physics steps per step (bounded sample of the %d-world "
                               "config), random robot velocity targets resampled every 10 steps" % (n, ROLLOUT_STEPS, args.worlds),
                   "worlds_per_step": n, "physics_steps_per_step": ROLLOUT_STEPS},
        "cpu_baseline": {"value": value, "unit": "world-steps/s", "cores": threads, "kind": "port",
                         "sample": "%d worlds x %d steps x %d repeats, one world per thread" % (n, ROLLOUT_STEPS, len(timed))},
        "e2e": {"value": value, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import box2d_sim_env_b200 as b2g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the batched backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.worlds
    L = b2g.lib()
    batch = b2g.BatchBox2DWorld(n, device=local).loadWorld(b2g.data_path(ENV_NAME))
    model = batch.model
    limits = model.dof_limits()
    nq = model.nq
    stream = torch.cuda.ExternalStream(batch.stream(), device=dev)

    # device-resident inputs / outputs
    d_q = torch.empty((n, nq), dtype=torch.float32, device=dev)
    d_qd = torch.empty_like(d_q)
    d_tg = torch.empty((N_TARGETS, n, 7), dtype=torch.float32, device=dev)
    d_out = torch.empty((2, n, nq), dtype=torch.float32, device=dev)
    d_all = torch.empty((world, 2, n, nq), dtype=torch.float32, device=dev) if world > 1 else None
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # pinned host staging for the end-to-end path
    h_q = torch.empty((n, nq), dtype=torch.float32).pin_memory()
    h_qd = torch.empty_like(h_q).pin_memory()
    h_tg = torch.empty((N_TARGETS, n, 7), dtype=torch.float32).pin_memory()
    h_out = torch.empty((2, n, nq), dtype=torch.float32).pin_memory()

    def load_inputs(step_idx, to_device):
        q, qd, tg = synth_inputs(limits, n, 1234 + 1000 * rank + step_idx)
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
        batch.get_state_dev(d_out[0].data_ptr(), d_out[1].data_ptr())
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(d_all.view(-1), d_out.view(-1))

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
                flush.zero_()  # evict L2 between timed iterations (state of 4096 worlds is smaller than L2)
            barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
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

    # warm-up (also JIT-free: everything is precompiled), then the timed region
    timed_pass(max(args.warmup, 3), False, 0)
    batch.take_kernel_ms()
    launches0 = L.b2g_launch_count()
    with ClockSampler(local) as clocks:
        dev_ms = timed_pass(args.steps, False, 100)
        kernel_ms, kernel_launches = batch.take_kernel_ms()
        launches = int(L.b2g_launch_count() - launches0)
        counts, _, _ = batch.contacts()
        e2e_ms = timed_pass(args.steps, True, 100)
    status = batch.status()

    units_per_step = n * ROLLOUT_STEPS
    value = world * units_per_step * args.steps / (dev_ms * 1e-3)
    e2e_value = world * units_per_step * args.steps / (e2e_ms * 1e-3)
    mean_contacts = float(counts.mean())
    peak, peak_kind = measured_peak()
    bytes_per_launch = algorithmic_bytes_per_world_step(mean_contacts) * units_per_step
    achieved = bytes_per_launch * kernel_launches / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0

    if rank == 0:
        line = {
            "metric": "world-steps/sec", "value": value, "unit": "world-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "test_env.yaml x %d worlds per GPU x %d physics steps per step, random robot velocity "
                                   "targets resampled every 10 steps (BASELINE.json configs[1])" % (n, ROLLOUT_STEPS),
                       "worlds_per_gpu": n, "physics_steps_per_step": ROLLOUT_STEPS, "dt": 0.01, "velocity_iterations": 15,
                       "position_iterations": 15, "l2": "flushed between timed iterations (384 MB write)",
                       "status_nonzero_worlds": int((status != 0).sum()), "mean_live_contacts": mean_contacts},
            "e2e": {"value": e2e_value, "unit": "world-steps/s",
                    "h2d_bytes_per_step": int(h_q.numel() * 4 * 2 + h_tg.numel() * 4), "d2h_bytes_per_step": int(h_out.numel() * 4)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "k_step", "peak_kind": peak_kind,
                         "kernel_ms_per_launch": kernel_ms / max(kernel_launches, 1),
                         "note": "latency/issue bound, not bandwidth bound: see DESIGN.md"},
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            nref = args.ref_worlds
            units, times = cpu_oracle_throughput(limits, nref, 1 + args.cpu_repeats, threads)
            timed = times[1:]
            line["cpu_baseline"] = {"value": units * len(timed) / sum(timed), "unit": "world-steps/s", "cores": threads,
                                    "kind": "port",
                                    "sample": "%d worlds x %d steps x %d repeats of the same workload, one world per thread"
                                              % (nref, ROLLOUT_STEPS, len(timed))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--worlds", type=int, default=4096, help="worlds per GPU")
    ap.add_argument("--ref-worlds", type=int, default=8192, help="worlds per step of the CPU oracle sample")
    ap.add_argument("--cpu-repeats", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
