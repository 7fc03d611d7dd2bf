"""Workload for compute-sanitizer (memcheck / racecheck / synccheck, one tool per run): every stepping kernel on shapes A
and C at a ragged small batch (95 worlds) and at 2049 worlds, plus checkCollision and guarded stepping.
Usage: compute-sanitizer --tool <tool> python tools/sanitize_case.py [steps]
Kernels reached: k_step_narrow (default small-batch path), k_step (B2G_LANES=32: CTA-wide phase barriers with parked padding
worlds in the last CTA), k_step_multi<4> (B2G_LPW=4: group barriers, shared-memory lists), k_step_guarded, k_check_collision."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import box2d_sim_env_b200 as b2g
from box2d_sim_env_b200 import envgen

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6


def run(tag, model, n, q, qd, env):
    for k, v in env.items():
        os.environ[k] = v
    batch = b2g.BatchBox2DWorld(n).loadWorld(model)
    for k in env:
        os.environ.pop(k)
    rng = np.random.default_rng(7)
    lim = model.dof_limits()
    nd = model.objects[0]["n_dofs"]
    batch.setWorldState(q, qd, reset_caches=True)
    batch.setTargetVelocity(0, rng.uniform(lim[:nd, 2], lim[:nd, 3], (n, nd)).astype(np.float32))
    batch.stepPhysics(steps)
    counts, _, _ = batch.checkCollision(max_per_world=16)
    allowed = np.ones((model.n_bodies, model.n_bodies), np.uint8)
    allowed[1, :] = 0
    done, nsteps = batch.stepPhysicsGuarded(allowed, 2)
    print("%-28s n=%5d steps=%d touching=%d guarded-complete=%d status-nonzero=%d" %
          (tag, n, steps, int(counts.sum()), int(done.sum()), int((batch.status() != 0).sum())), flush=True)


def states_a(model, n):
    rng = np.random.default_rng(3)
    q = np.zeros((n, model.nq), np.float32)
    q[:, 7:9] = rng.uniform(-0.5, 0.5, (n, 2))
    q[:, 9] = rng.uniform(-3, 3, n)
    return q, np.zeros_like(q)


model_a = b2g.Model(b2g.data_path("test_env.yaml"))
model_c = b2g.Model(envgen.clutter_env())
for n in (95, 2049):
    qa, qda = states_a(model_a, n)
    qc, qdc = envgen.clutter_states(n, seed=4, speed=3.0)
    run("A narrow", model_a, n, qa, qda, {})
    run("A k_step (barriers)", model_a, n, qa, qda, {"B2G_LANES": "32", "B2G_STEP_BLOCK": "224"})
    run("A k_step_multi<4>", model_a, n, qa, qda, {"B2G_LPW": "4"})
    run("C narrow", model_c, n, qc, qdc, {})
    run("C k_step (barriers)", model_c, n, qc, qdc, {"B2G_LANES": "32", "B2G_STEP_BLOCK": "96"})
    run("C k_step_multi<4>", model_c, n, qc, qdc, {"B2G_LPW": "4"})
