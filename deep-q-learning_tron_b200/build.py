"""Build libtron_b200.so (sm_100a only) in-tree with nvcc.  `python -m tron_b200.build` or build()."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libtron_b200.so")
LIB_DEBUG = os.path.join(HERE, "libtron_b200_debug.so")  # -DTRON_DEBUG: device-side range checks, tests only
SOURCES = ["abi.cu", "host_env.cu", "step_c144.cu", "step_generic.cu", "step_sparse.cu", "step_bits.cu", "step_trail.cu", "minimax.cu", "misc_kernels.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC,-O2,-Wall", "-Xptxas", "-v"]


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "tron_b200.h"))
    return hdrs


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False, debug=False):
    """debug=True builds libtron_b200_debug.so (every computed index range-checked on the device) next to the release library."""
    obj_dir = OBJ + ("_debug" if debug else "")
    lib = LIB_DEBUG if debug else LIB
    os.makedirs(obj_dir, exist_ok=True)
    deps = _deps()
    flags = NVCC_FLAGS + (["-DTRON_DEBUG"] if debug else [])

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, src)] + deps):
            cmd = [_nvcc()] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(obj + ".log", "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
            if verbose:
                print(r.stderr[-2000:])
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(lib, objs):
        cmd = [_nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
