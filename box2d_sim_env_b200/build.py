"""Build libb2g.so (CUDA kernels + C ABI) in-tree for sm_100a.

nvcc cross-compiles without a GPU.  --fmad=false keeps device arithmetic identical to the scalar SSE code the
reference is compiled to (no FMA contraction); -lineinfo lets ncu map stalls to source lines.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb2g.so")
SOURCES = ["b2g_kernels.cu", "b2g_capi.cpp", "yaml_env.cpp", "model_build.cpp"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    if os.path.getmtime(os.path.join(HERE, "..", "include", "b2g.h")) > t:
        return True
    return False


def build(force=False, verbose=False, out=None):
    """The kernels (b2g_kernels.cu) are compiled twice -- -DB2G_SETS=0: one model per batch, the hot path; -DB2G_SETS=1: several
    parameter sets per batch, every warp stages the model variant of its tile (b2g_launch.h) -- in parallel, then linked with
    the host sources."""
    global LIB
    if out:  # tuning experiments: build a variant next to the product library
        LIB = out
    if not force and not needs_build():
        return LIB
    common = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--fmad=false", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2"]
    if os.environ.get("B2G_DEFINES"):  # tuning experiments only, e.g. B2G_DEFINES="-DB2G_PHASE_TIMING"
        common[1:1] = os.environ["B2G_DEFINES"].split()
    if os.environ.get("B2G_MAXREG"):  # tuning experiments only; the product build uses the launch bounds in the source
        common[1:1] = ["-maxrregcount", os.environ["B2G_MAXREG"]]
    if verbose:
        common[1:1] = ["-Xptxas", "-v"]
    objdir = os.path.join(HERE, "..", "build", "obj_" + os.path.basename(LIB).replace(".", "_"))
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for sets in (0, 1):
        obj = os.path.join(objdir, "b2g_kernels_sets%d.o" % sets)
        objs.append(obj)
        procs.append(subprocess.Popen(common + ["-DB2G_SETS=%d" % sets, "-c", "b2g_kernels.cu", "-o", obj], cwd=CSRC))
    codes = [p.wait() for p in procs]
    if any(codes):
        raise subprocess.CalledProcessError(max(codes), "nvcc -c b2g_kernels.cu")
    subprocess.check_call(common + ["-shared", "-o", LIB] + objs + [x for x in SOURCES if x != "b2g_kernels.cu"] + ["-ldl"], cwd=CSRC)
    shutil.rmtree(objdir, ignore_errors=True)   # the objects are not needed again (and would travel to the GPU box with the tree)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
