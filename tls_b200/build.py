"""Builds libtlsb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in (
    "tlsb_host.cu", "tlsb_resident.cu", "tlsb_tiled.cu", "tlsb_aux_kernels.cu", "tlsb_spectra.cu")]
HEADERS = [os.path.join(os.path.dirname(HERE), "include", "tlsb200.h"), os.path.join(HERE, "csrc", "tlsb_internal.h"),
           os.path.join(HERE, "csrc", "tlsb_device.cuh")]
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libtlsb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libtlsb200.so cannot be built")


def source_hash():
    """sha256 over the kernel and host sources of the library (profiles/ncu_*.json record it: a profile of another
    build is reported as stale by bench.py and refused by tests/test_bench_contract.py)."""
    import hashlib

    h = hashlib.sha256()
    for p in sorted(SRC + HEADERS):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def is_stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > built for p in SRC + HEADERS + [os.path.abspath(__file__)])


def _run(cmd):
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a gcc wrapper that nvcc must not pick up
    env.pop("CXX", None)
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return proc.stderr


def build(force=False, verbose=False, defines=(), out=None):
    """Compile if the library is missing or older than its sources; returns the path.
    One nvcc -c per translation unit (in parallel), then one link.
    `defines` / `out` build an experimental variant next to the product library (scripts/gpu_variants.sh)."""
    if out is None and not force and not is_stale():
        return LIB
    out = out or LIB
    tag = os.path.splitext(os.path.basename(out))[0]
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = nvcc_path()
    objs = [os.path.join(OBJ_DIR, "%s_%s.o" % (tag, os.path.splitext(os.path.basename(src))[0])) for src in SRC]
    cmds = [[nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + ["-c", "-o", obj, src]
            for src, obj in zip(SRC, objs)]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(cmds)) as pool:
        logs = list(pool.map(_run, cmds))
    _run([nvcc] + LINK_FLAGS + ["-o", out] + objs)
    if verbose:
        print("".join(logs))
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
