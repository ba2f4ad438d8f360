"""Build the sm_100a shared library in-tree: ssme_b200/lib/libssme_b200.so.

nvcc cross-compiles without a GPU.  One object per CTA size (pf_inst.cu with -DSSME_NT=...)
plus the C-ABI translation unit, compiled in parallel, linked with the static CUDA runtime so
the library has no dependency on which libcudart the host process already loaded.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libssme_b200.so")
NT_LIST = (32, 64, 128, 256, 512, 1024)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # the canonical arithmetic spells out every fused multiply-add
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    inc = os.path.join(HERE, "..", "include")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(inc, "ssme_b200.h"), __file__]
    deps += [os.path.join(inc, "ssme_b200", f) for f in os.listdir(os.path.join(inc, "ssme_b200"))]
    return max(os.path.getmtime(d) for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    global LIB, OBJDIR
    extra = os.environ.get("SSME_NVCC_EXTRA", "").split()  # experiments only, e.g. -DSSME_STAGGER_NS=400
    if extra:
        tag = "_".join(e.replace("-D", "").replace("=", "") for e in extra)
        LIB = os.path.join(LIBDIR, "libssme_b200_%s.so" % tag)
        OBJDIR = os.path.join(HERE, "build", tag)
        NVCC_FLAGS.extend(extra)
        force = True
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    newest = _sources()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    nvcc = _nvcc()
    jobs = []
    for nt in NT_LIST:
        obj = os.path.join(OBJDIR, "pf_inst_nt%d.o" % nt)
        jobs.append((obj, [nvcc, *NVCC_FLAGS, "-DSSME_NT=%d" % nt, "-c", os.path.join(CSRC, "pf_inst.cu"), "-o", obj]))
    obj = os.path.join(OBJDIR, "capi.o")
    jobs.append((obj, [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, "capi.cu"), "-o", obj]))
    obj = os.path.join(OBJDIR, "spill_capi.o")
    jobs.append((obj, [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, "spill_capi.cu"), "-o", obj]))
    obj = os.path.join(OBJDIR, "pmmh_capi.o")  # host-only C++ (the PMMH loop above the C ABI)
    jobs.append((obj, [nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-c", os.path.join(CSRC, "pmmh_capi.cpp"), "-o", obj]))
    todo = [(o, c) for o, c in jobs if force or not os.path.exists(o) or os.path.getmtime(o) < newest]
    if verbose:
        for _, c in todo:
            c.insert(1, "-Xptxas=-v")
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        outs = list(ex.map(lambda oc: _run(oc[1]), todo))
    if verbose:
        for o in outs:
            sys.stderr.write(o)
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
          "-o", LIB, *[o for o, _ in jobs]])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
