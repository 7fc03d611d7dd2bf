"""bench.py command line on a machine without a GPU: the reference arm (CPU oracle) prints the contract's JSON line, the
product arm refuses to run instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT)


def test_reference_arm_prints_contract_line():
    out = run("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-worlds", "64")
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "world-steps/sec" and line["unit"] == "world-steps/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "test_env.yaml" in line["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = run("--steps", "1")
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)


def test_reference_arm_never_loads_the_product_library():
    """The CPU arm's DOF limits come from the environment description: libb2g.so must not be mapped in that process."""
    code = ("import sys, json; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-worlds', '16'];"
            "import bench; bench.main(); maps = open('/proc/self/maps').read(); assert 'libb2g' not in maps, 'product library mapped';"
            "assert 'box2d_sim_env_b200' not in sys.modules")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr


def test_both_arms_print_the_same_config_object():
    sys.path.insert(0, ROOT)
    import bench
    cfg = bench.CONFIGS["cfg2"]
    a = bench.config_dict(cfg, "cfg2", 4096, 4096)
    out = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-worlds", "16")
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["config"] == a


def test_velocity_limits_from_description_match_the_model(b2g):
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    for name in ("cfg2", "cfg3", "cfg4"):
        cfg = bench.CONFIGS[name]
        lim = bench.robot_velocity_limits(bench.oracle_env(cfg))
        model = bench.load_model(cfg)
        assert np.array_equal(lim, model.dof_limits()[:cfg["robot_dofs"], 2:4])


class _OracleBackedBatch:
    """Stand-in for BatchBox2DWorld built on the oracle: lets the bench's parity gate run where there is no GPU."""

    def __init__(self, b2o, env_desc, n):
        self.b2o, self.desc, self.n = b2o, env_desc, n
        self.env = b2o.OracleEnv(env_desc)
        self.ob = b2o.OracleBatch(self.env, n)

    def rollout(self, robot, tg, spt, q=None, qd=None, reset_caches=True):
        if reset_caches:
            self.ob = self.b2o.OracleBatch(self.env, self.n)
            if getattr(self, "pose", None):   # reset_caches keeps the poses of static objects
                for w in range(self.n):
                    self.ob.world(w).set_object_pose(2, *self.pose)
        import numpy as np
        return self.ob.rollout(q, qd, robot, np.ascontiguousarray(tg), spt, 4)

    def setObjectPose(self, obj, pose):
        self.pose = [float(x) for x in pose[0]]

    def contacts(self, max_per_world=64):
        import numpy as np
        counts = np.zeros(self.n, np.int32)
        io = np.zeros((self.n, max_per_world, 8), np.int32)
        for w in range(self.n):
            oio, _ = self.ob.world(w).contacts()
            counts[w] = len(oio)
            io[w, :len(oio)] = oio[:max_per_world]
        return counts, io, None


def test_parity_gate_is_not_vacuous(oracle):
    """The bench workloads must exercise contacts: the gate counts touching contacts after one step and over the rollout's
    checkpoints, reports TOI activity and the divergence against libm trigonometry, and fails when it compares nothing."""
    sys.path.insert(0, ROOT)
    import bench
    for name, min_live in (("cfg2", 1.0), ("cfg3", 0.5)):
        cfg = bench.CONFIGS[name]
        n = 128
        desc = bench.oracle_env(cfg)
        stand_in = _OracleBackedBatch(oracle, desc, n)
        bench.place_obstacle(stand_in, cfg, n)
        gate = bench.parity_gate(cfg, stand_in, bench.robot_velocity_limits(desc), n, 1234, 4, sample=n)
        if name == "cfg2":
            assert gate["rollout"]["toi_events"] > 0, "the headline workload must exercise the TOI sub-steps"
        assert gate["pass"] and gate["nonvacuous"], gate
        assert gate["single_step"]["touching_contacts"] >= n // 8
        assert gate["rollout"]["touching_contacts"] * 8 >= n
        assert gate["single_step"]["bit_identical_worlds"] == n and gate["rollout"]["bit_identical_worlds"] == n
        lm = gate["rollout_vs_libm"]
        assert len(lm["max_abs_dq"]) == bench.N_TARGETS and lm["single_step"]["max_abs_dq"] <= 1e-5
        assert max(lm["max_abs_dq"]) > 0.0   # a real divergence curve, not zeros by construction
