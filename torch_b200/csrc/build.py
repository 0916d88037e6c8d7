"""Build libtltorch_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.abspath(os.path.join(HERE, "..", "_cuda"))
LIB = os.path.join(OUT_DIR, "libtltorch_cuda.so")
SOURCES = ["api.cu", "contract.cu", "conv.cu", "gemm_tc.cu", "blocktt.cu", "dwconv.cu", "pointwise_tc.cu", "modedot.cu", "mmd16.cu", "rows.cu", "skinny.cu", "dwconv_rows.cu", "dwconv_tma.cu", "khatri_rao.cu", "dwsep_rows.cu", "dwsep_staged.cu", "cpchain.cu", "cplinear.cu", "trl_tail.cu", "split3.cu", "op_level.cu", "cpconv_rows.cu", "cpconv_rows2.cu", "spatial_project.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".py"))]
    deps.append(os.path.abspath(os.path.join(HERE, "..", "..", "include", "tltorch_cuda.h")))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    # headers every translation unit depends on: a change there rebuilds everything, a change in one .cu only that object
    shared = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    shared += [os.path.abspath(os.path.join(HERE, "..", "..", "include", "tltorch_cuda.h")), os.path.abspath(__file__)]
    newest_shared = max(os.path.getmtime(d) for d in shared)
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        path = os.path.join(HERE, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(newest_shared, os.path.getmtime(path)):
            continue
        cmd = [nvcc, *FLAGS, "-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f"==== {src}\n{out}")
        if pr.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.check_call(cmd)
    with open(os.path.join(OUT_DIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
