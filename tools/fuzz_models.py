"""Exploration: random-model parity (tests/test_gpu_parity.py::test_random_models_parity) over many seeds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import box2d_sim_env_b200 as b2g
import test_gpu_parity as t
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = []
for seed in range(lo, hi):
    try:
        t.test_random_models_parity(b2g, seed)
    except b2g.B2GError as e:
        print('seed', seed, 'B2GError:', str(e)[:120])
    except AssertionError as e:
        if "too few touching" not in str(e):
            bad.append((seed, str(e)[:300]))
print("seeds %d..%d: %d failures" % (lo, hi, len(bad)))
for b in bad[:5]:
    print(b)
