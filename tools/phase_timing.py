"""Cycles per phase of the stepping kernel (B2G_PHASE_TIMING build, b2g_device.cuh: tick).  Build the variant first:
  B2G_DEFINES=-DB2G_PHASE_TIMING python -c "from box2d_sim_env_b200 import build; build.build(force=True, out='build/libb2g_pt.so')"
Usage (GPU box): B2G_LIB=build/libb2g_pt.so python tools/phase_timing.py <config> <worlds>"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import box2d_sim_env_b200 as b2g
import bench
NAMES = ["control", "step head / newFixture", "collide (narrow phase)", "islands (DFS + reorder)", "integrate velocities",
         "contact init + warm start", "joint init", "velocity iterations", "store impulses + simple islands", "integrate positions",
         "position iterations", "sync transforms / fixtures / sleep", "findNewContacts", "TOI", "clear forces / tail"]
name, n = sys.argv[1], int(sys.argv[2])
cfg = bench.CONFIGS[name]
model = bench.load_model(cfg)
batch = b2g.BatchBox2DWorld(n).loadWorld(model)
bench.place_obstacle(batch, cfg, n)
q, qd, tg = bench.synth_inputs(model.dof_limits(), n, 1, cfg)
d_tg = torch.from_numpy(tg).cuda()
lib = b2g.lib()
out = (ctypes.c_ulonglong * 32)()
for r in range(2):
    batch.setWorldState(q, qd, reset_caches=True)
    torch.cuda.synchronize()
    lib.b2g_debug_phase_cycles(out, 1)
    batch.take_kernel_ms()
    batch.rollout_dev(0, d_tg.data_ptr(), bench.N_TARGETS, bench.STEPS_PER_TARGET)
    ms, nl = batch.take_kernel_ms()
    torch.cuda.synchronize()
    lib.b2g_debug_phase_cycles(out, 0)
cyc = np.array(out[:15], np.float64)
print("%s n=%d: %.3f ms (instrumented build); share of warp cycles per phase:" % (name, n, ms / nl))
for i, nm in enumerate(NAMES):
    print("  %-36s %6.2f %%" % (nm, 100 * cyc[i] / cyc.sum()))
