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

# (source, object stem, extra flags).  The parity kernels must not be FMA-contracted.
FAST = ["-prec-div=false", "-prec-sqrt=false", "-ftz=true"]
UNITS = [
    ("agf_kernels_parity.cu", "agf_kernels_parity", ["-fmad=false"]),
    ("agf_kernels_fast.cu", "agf_kernels_fast_f32_uwb", FAST + ["-DAGF_FAST_F64=0", "-DAGF_FAST_UWB=1"]),
    ("agf_kernels_fast.cu", "agf_kernels_fast_f32_rates", FAST + ["-DAGF_FAST_F64=0", "-DAGF_FAST_UWB=0"]),
    ("agf_kernels_fast.cu", "agf_kernels_fast_f64_uwb", FAST + ["-DAGF_FAST_F64=1", "-DAGF_FAST_UWB=1"]),
    ("agf_kernels_fast.cu", "agf_kernels_fast_f64_rates", FAST + ["-DAGF_FAST_F64=1", "-DAGF_FAST_UWB=0"]),
    ("agf_batch.cu", "agf_batch", ["-fmad=false"]),
    # the batched RAPPIDS planner (include/agrifly_b200_rappids.h): parity and throughput variants of the kernel
    ("agf_rappids_plan.cu", "agf_rappids_plan_parity", ["-fmad=false", "-DAGF_RAPPIDS_PARITY=1"]),
    ("agf_rappids_plan.cu", "agf_rappids_plan_fast", ["-DAGF_RAPPIDS_PARITY=0"]),
    ("agf_rappids.cu", "agf_rappids", ["-fmad=false"]),
    ("agf_config.cpp", "agf_config", []),
    ("agf_nccl.cpp", "agf_nccl", []),
]
HEADERS = ["agf_step.cuh", "agf_types.h", "agf_math.h", "agf_launch.h", "agf_host_params.h", "agf_rappids_plan.cuh", "agf_nccl.h",
           os.path.join(ROOT, "include", "agrifly_b200.h"), os.path.join(ROOT, "include", "agrifly_b200_rappids.h")]


def _digest(paths, flags):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _compile(src, stem, extra, verbose):
    os.makedirs(BUILD, exist_ok=True)
    obj = os.path.join(BUILD, stem + ".o")
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
    keep = {o + ".o" for _, o, _ in UNITS} | {o + ".o.sha" for _, o, _ in UNITS}
    keep |= {f for f in (os.listdir(BUILD) if os.path.isdir(BUILD) else []) if f.startswith("var_")}
    for f in os.listdir(BUILD) if os.path.isdir(BUILD) else []:
        if f not in keep:
            os.remove(os.path.join(BUILD, f))  # objects of translation units that no longer exist
    with cf.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        futs = [ex.submit(_compile, s, o, e, verbose) for s, o, e in UNITS]
        res = [f.result() for f in futs]
    objs = [r[0] for r in res]
    if any(r[1] for r in res) or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        for r in res:
            if r[2]:
                sys.stderr.write(r[2])
    return LIB


def build_variant(name, defines, units=("agf_kernels_fast_f32_uwb",)):
    """Tuning aid (profiles/): a copy of the library with extra -D flags on the named units, written to
    agri-fly_b200/variants/libagrifly_b200_<name>.so; load it with AGF_LIB_PATH=<that file>."""
    build_native()
    vdir = os.path.join(HERE, "variants")
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for src, stem, extra in UNITS:
        if stem in units:
            obj = os.path.join(BUILD, "var_%s_%s.o" % (name, stem))
            cmd = [NVCC] + ARCH + COMMON + extra + list(defines) + ["-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for variant %s:\n%s\n%s" % (name, r.stdout, r.stderr))
            objs.append(obj)
        else:
            objs.append(os.path.join(BUILD, stem + ".o"))
    out = os.path.join(vdir, "libagrifly_b200_%s.so" % name)
    r = subprocess.run([NVCC] + ARCH + ["-shared", "-o", out] + objs + ["-cudart", "static", "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    for o in objs:
        if os.path.basename(o).startswith("var_"):
            os.remove(o)
    return out


def build_oracle():
    """oracle/ is test infrastructure; building the checker is not using it."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port", "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return r.stdout


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python build.py --variant <name> -DX=1 -DY=2
        k = sys.argv.index("--variant")
        print(build_variant(sys.argv[k + 1], [a for a in sys.argv[k + 2:] if a.startswith("-D")]))
        sys.exit(0)
    print(build_native(verbose="-v" in sys.argv, force="-f" in sys.argv))
    if "--oracle" in sys.argv:
        print(build_oracle())
