"""Build libssdx.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  -fmad=false: the parity contract needs one fp32 rounding per
operation (the kernels also use explicit __f*_rn intrinsics on the IoU / decode expressions).
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libssdx.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ["capi.cu", "decode.cu", "encode.cu", "encode_grid.cu", "nms.cu", "eval.cu", "seg.cu", "metric.cu", "ap.cu", "loss.cu", "comm.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden,-ffp-contract=off",
    "--shared", "-cudart", "static",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "ssdx.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, debug_timers: bool = False, variant=None) -> str:
    """debug_timers: also build lib/libssdx_dbg.so with -DSSDX_PHASE_TIMERS (tools/phase_timers.py).
    variant: (name, [-D flags]) builds lib/libssdx_<name>.so for tuning sweeps (tools/time_stage.py)."""
    if variant is not None:
        return _compile(os.path.join(LIBDIR, f"libssdx_{variant[0]}.so"), verbose, list(variant[1]))
    if debug_timers:
        return _compile(os.path.join(LIBDIR, "libssdx_dbg.so"), verbose, ["-DSSDX_PHASE_TIMERS"])
    if not force and not needs_build():
        return LIB
    return _compile(LIB, verbose, [])


def _compile(out: str, verbose: bool, extra) -> str:
    """One nvcc per translation unit (in parallel, objects cached under build/ by source + flag hash), then one link."""
    import hashlib
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(os.path.dirname(HERE), "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = find_nvcc()
    cflags = [f for f in NVCC_FLAGS if f != "--shared"] + list(extra) + ["-I", INCLUDE, "-I", CSRC]
    if verbose:
        cflags += ["-Xptxas", "-v"]
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(INCLUDE, "ssdx.h"))
    hdr_hash = hashlib.sha1()
    for h in headers:
        with open(h, "rb") as f:
            hdr_hash.update(f.read())
    hdr_hash.update(" ".join(cflags).encode())

    def one(src_name):
        src = os.path.join(CSRC, src_name)
        hsh = hashlib.sha1(hdr_hash.digest())
        with open(src, "rb") as f:
            hsh.update(f.read())
        obj = os.path.join(objdir, f"{os.path.splitext(src_name)[0]}.{hsh.hexdigest()[:16]}.o")
        if os.path.exists(obj) and not verbose:
            return obj, 0, ""
        r = subprocess.run([nvcc] + cflags + ["-c", "-o", obj, src], capture_output=True, text=True)
        if r.returncode != 0 and os.path.exists(obj):
            os.remove(obj)
        return obj, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(one, SOURCES))
    log = "".join(r[2] for r in results)
    if verbose or any(r[1] != 0 for r in results):
        sys.stderr.write(log)
    if any(r[1] != 0 for r in results):
        raise RuntimeError("nvcc failed building " + os.path.basename(out))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static", "-Xcompiler", "-fPIC",
           "-ldl", "-o", out] + [r[0] for r in results]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking " + os.path.basename(out))
    keep = {os.path.basename(r[0]) for r in results}
    for f in os.listdir(objdir):  # drop objects of older sources of the same unit (variants keep their own hashes)
        stem = f.split(".")[0]
        if f not in keep and not extra and any(k.startswith(stem + ".") for k in keep):
            try:
                os.remove(os.path.join(objdir, f))
            except OSError:
                pass
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--debug-timers", action="store_true")
    ap.add_argument("--variant", help="NAME: build lib/libssdx_NAME.so with the --define values (tuning sweeps)")
    ap.add_argument("--define", action="append", default=[], help="KEY=VALUE passed to nvcc as -DKEY=VALUE")
    a = ap.parse_args()
    var = (a.variant, ["-D" + d for d in a.define]) if a.variant else None
    print(build(force=a.force, verbose=a.verbose, debug_timers=a.debug_timers, variant=var))
