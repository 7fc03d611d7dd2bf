"""The header-only C++ facade (include/b2g_world.hpp) compiles as plain C++11 against the C ABI and behaves like the
Python mirror.  The CPU test checks the host side and the loud failure without a device; the GPU test steps worlds."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_demo(tmp_path, b2g):
    b2g.lib()  # makes sure libb2g.so exists
    exe = os.path.join(str(tmp_path), "facade_demo")
    libdir = os.path.join(ROOT, "box2d_sim_env_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp"), "-o", exe, "-L", libdir, "-lb2g",
                           "-Wl,-rpath," + libdir])
    return exe


def test_facade_compiles_and_fails_loudly_without_gpu(tmp_path, b2g):
    import torch
    exe = build_demo(tmp_path, b2g)
    out = subprocess.run([exe, b2g.data_path("test_env.yaml"), "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "nq=11 bodies=9 objects=3 first=my_robot" in out.stdout
    if not torch.cuda.is_available():
        assert "cuda_named=1 lookup=1" in out.stdout


@pytest.mark.gpu
def test_facade_matches_python_mirror(tmp_path, b2g):
    exe = build_demo(tmp_path, b2g)
    out = subprocess.run([exe, b2g.data_path("test_env.yaml"), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("GPU")][0]
    q0_cpp = np.array([float.fromhex(t) for t in line.split("q0=")[1].split()], np.float32)
    n = 64
    batch = b2g.BatchBox2DWorld(n).loadWorld(b2g.data_path("test_env.yaml"))
    q, qd = batch.getWorldState()
    q[:, 7] = np.float32(0.2) + np.float32(0.01) * np.arange(n, dtype=np.float32)
    batch.setWorldState(q, qd, reset_caches=True)
    u = np.zeros((n, 7), np.float32)
    u[:, 0], u[:, 2] = 0.5, 0.3
    batch.setTargetVelocity("my_robot", u)
    batch.stepPhysics(10)
    batch.saveState()
    counts, _, _ = batch.stepPhysicsRecord(20)
    assert batch.restoreState()
    batch.stepPhysics(20)
    q1, _ = batch.getWorldState()
    assert np.array_equal(q1[0], q0_cpp)
    assert "recorded=%d " % int(np.minimum(counts, 32).sum()) in line and "restored=1" in line
    view = [l for l in out.stdout.splitlines() if l.startswith("VIEW")][0]
    assert "bad=0" in view and "threw=1" in view, view


@pytest.mark.gpu
def test_cpp_planner_on_several_gpus(tmp_path, b2g):
    """tests/cpp/multi_demo.cpp: a C++ caller shards a batch over the node's GPUs through b2g_multi_* (one process, one host
    thread per device, ncclAllGather of the states) and gets the single-batch result bit for bit."""
    import torch
    b2g.lib()
    exe = os.path.join(str(tmp_path), "multi_demo")
    libdir = os.path.join(ROOT, "box2d_sim_env_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "multi_demo.cpp"), "-o", exe, "-L", libdir, "-lb2g",
                           "-Wl,-rpath," + libdir])
    ndev = min(torch.cuda.device_count(), 8)
    for nd in sorted({1, ndev}):
        out = subprocess.run([exe, b2g.data_path("test_env.yaml"), str(nd), "3001"], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "devices=%d worlds=3001 covered=3001 identical=1" % nd in out.stdout, out.stdout
