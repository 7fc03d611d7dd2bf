"""Exploration: mutate the shipped YAML files (deletions, stray tokens, truncation, indentation) and load them: the loader /
model builder must either build a finite model or refuse with B2GError -- never crash.  Usage: fuzz_yaml.py <seed> <count>"""
import sys, os, random, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import box2d_sim_env_b200 as b2g
src_env = open(b2g.data_path('test_env.yaml')).read()
src_robot_path = None
import re, glob
data_dir = os.path.dirname(b2g.data_path('test_env.yaml'))
files = {os.path.relpath(p, data_dir): open(p).read() for p in glob.glob(data_dir+'/**/*.yaml', recursive=True)}
random.seed(int(sys.argv[1]))
n_err = n_ok = 0
tmp = tempfile.mkdtemp()
for it in range(int(sys.argv[2])):
    target = random.choice(list(files))
    mutated = dict(files)
    s = mutated[target]
    for _ in range(random.randint(1, 4)):
        k = random.random()
        if not s: break
        pos = random.randrange(len(s))
        if k < 0.3: s = s[:pos] + s[pos + random.randint(1, 20):]
        elif k < 0.6: s = s[:pos] + random.choice(['[', ']', ':', '-', ' ', '\n', '  ', '#', '{', '}', ',', '1e999', 'nan', '-', '"', "'", '\t']) + s[pos:]
        elif k < 0.8: s = s[:pos]
        else:
            lines = s.split('\n'); i = random.randrange(len(lines)); lines[i] = lines[i].replace(' ', '', 1) if random.random() < 0.5 else ' ' + lines[i]; s = '\n'.join(lines)
    mutated[target] = s
    for rel, txt in mutated.items():
        p = os.path.join(tmp, rel); os.makedirs(os.path.dirname(p), exist_ok=True); open(p, 'w').write(txt)
    try:
        m = b2g.Model(os.path.join(tmp, 'test_env.yaml'))
        n_ok += 1
        import numpy as np
        bm=m.body_model(); fm=m.fixture_model()[1]; jm=m.joint_model()[1]; dl=m.dof_limits()
        if not (np.isfinite(bm).all() and np.isfinite(fm).all() and np.isfinite(jm).all()):
            print('NONFINITE model from mutation of', target); open(os.path.join(tmp, 'bad_%d.yaml' % it), 'w').write(s)
    except b2g.B2GError:
        n_err += 1
print("ok", n_ok, "rejected", n_err)
