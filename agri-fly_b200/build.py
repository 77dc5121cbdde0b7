"""In-tree build of libagrifly_b200.so (nvcc, sm_100a only) and of the oracle libraries.

`python agri-fly_b200/build.py` or `__graft_entry__.build()`.  The .so files are git-ignored but
travel to the GPU box with the gpurun snapshot.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libagrifly_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
          "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]

# (source, extra flags).  The parity kernels must not be FMA-contracted.
UNITS = [
    ("agf_kernels_parity.cu", ["-fmad=false"]),
    ("agf_kernels_fast_f64.cu", ["-prec-div=false", "-prec-sqrt=false", "-ftz=true"]),
    ("agf_kernels_fast_f32.cu", ["-prec-div=false", "-prec-sqrt=false", "-ftz=true"]),
    ("agf_batch.cu", ["-fmad=false"]),
    ("agf_config.cpp", []),
]
HEADERS = ["agf_step.cuh", "agf_types.h", "agf_math.h", "agf_launch.h", "agf_host_params.h", os.path.join(ROOT, "include", "agrifly_b200.h")]


def _digest(paths, flags):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _compile(src, extra, verbose):
    os.makedirs(BUILD, exist_ok=True)
    obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".sha"
    deps = [os.path.join(CSRC, src)] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    flags = ARCH + COMMON + extra
    dig = _digest(deps, flags)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False, ""
    cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, True, r.stderr


def build_native(verbose=False, force=False):
    if force:
        for f in os.listdir(BUILD) if os.path.isdir(BUILD) else []:
            os.remove(os.path.join(BUILD, f))
    with cf.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        futs = [ex.submit(_compile, s, e, verbose) for s, e in UNITS]
        res = [f.result() for f in futs]
    objs = [r[0] for r in res]
    if any(r[1] for r in res) or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        for r in res:
            if r[2]:
                sys.stderr.write(r[2])
    return LIB


def build_oracle():
    """oracle/ is test infrastructure; building the checker is not using it."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port", "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return r.stdout


if __name__ == "__main__":
    print(build_native(verbose="-v" in sys.argv, force="-f" in sys.argv))
    if "--oracle" in sys.argv:
        print(build_oracle())
