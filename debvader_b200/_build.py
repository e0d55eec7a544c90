"""Build libdebvader_b200.so in-tree with nvcc for sm_100a (no torch, no cmake).

    python -m debvader_b200._build [--force] [--verbose] [--ablate]

Two variants of the same sources:
  libdebvader_b200.so         the product: no environment switches, no debug hooks in the kernels
  libdebvader_b200_ablate.so  -DDBV_ABLATE: DBV_* environment switches (kernel A/B selection, timing ablations, tuning
                              knobs), the tcgen05 descriptor probes (tc_probe.cu, include/debvader_b200_debug.h) and
                              the clock64 instrumentation of the halo kernel; used by tools/ and by the cross-check
                              tests (DEBVADER_B200_LIB=... selects it)

The library links cudart statically and resolves cuTensorMapEncodeTiled through
cudaGetDriverEntryPoint, so it loads (and exports its symbols) on a machine
without a GPU or libcuda — compute entry points then fail with DBV_ERR_CUDA.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdebvader_b200.so")
LIB_ABLATE = os.path.join(HERE, "libdebvader_b200_ablate.so")
SOURCES = ["api.cu", "field_kernels.cu", "simt_kernels.cu", "tc_conv.cu", "tc_pair.cu", "tc_pairh.cu", "tc_halo.cu", "tc_halo2.cu", "host_stage.cu", "detect_kernels.cu"]
SOURCES_ABLATE = SOURCES + ["tc_probe.cu"]
HEADERS = ["common.cuh", "epilogue.cuh", "kernels.h", "tc_ptx.cuh", "tc_pair_ptx.cuh", "host_stage.h", os.path.join("..", "..", "include", "debvader_b200.h"),
           os.path.join("..", "..", "include", "debvader_b200_debug.h")]
# detect_kernels.cu is compared bit for bit with oracle/detect_numpy.py: no fused multiply-adds (numpy rounds every product and sum)
EXTRA_FLAGS = {"detect_kernels.cu": ["-fmad=false"]}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build(ablate: bool = False) -> bool:
    lib = LIB_ABLATE if ablate else LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in (SOURCES_ABLATE if ablate else SOURCES) + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, ablate: bool = False) -> str:
    """Compile one variant (parallel nvcc per source, then link).  Returns the library path."""
    lib = LIB_ABLATE if ablate else LIB
    if not force and not needs_build(ablate):
        return lib
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build", "ablate" if ablate else "product")
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + (["-DDBV_ABLATE"] if ablate else [])
    procs = []
    objs = []
    for s in SOURCES_ABLATE if ablate else SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *flags, *EXTRA_FLAGS.get(s, []), "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    tmp = lib + ".tmp"
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs, "-ldl", "-lpthread", "-lrt"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, lib)
    return lib


def build_all(force: bool = False, verbose: bool = False):
    return build(force, verbose, ablate=False), build(force, verbose, ablate=True)


if __name__ == "__main__":
    if "--ablate" in sys.argv:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ablate=True))
    else:
        for l in build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv):
            print(l)
