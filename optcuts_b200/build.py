"""Build liboptcuts_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m optcuts_b200.build [--force]

The library has no torch dependency (static cudart); it is the C-ABI declared in
include/optcuts_b200.h.  The element-kernel unit is compiled with -fmad=false so per-element values
are bit-identical to the CPU reference's (see csrc/ocb_kernels.cu); the solver unit keeps FMAs.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "liboptcuts_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
UNITS = [
    ("ocb_api.cu", []),
    ("ocb_kernels.cu", ["-fmad=false"]),
    ("ocb_pcg.cu", []),
    ("ocb_mas.cu", []),
    ("ocb_stencils.cu", ["-fmad=false"]),
    ("ocb_direct.cu", []),
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "optcuts_b200.h"))
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [NVCC] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = r.stdout + r.stderr
            with open(o + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + log)
            if verbose or r.returncode != 0:
                sys.stderr.write(log)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % src)
    if force or _newer(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
