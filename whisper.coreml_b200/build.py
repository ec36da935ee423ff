"""Build libwhisper_b200.so in-tree with nvcc for sm_100a (the analogue of coreml/Makefile:13-19).

    python whisper.coreml_b200/build.py [--force]

Every .cu under csrc/ is compiled to build/<name>.o (in parallel, skipped when up to date) and
linked into whisper.coreml_b200/libwhisper_b200.so.  The ABI smoke driver (csrc/abi_smoke.cpp,
analogue of coreml/coremlTest.cpp) is linked against it as build/abi_smoke.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libwhisper_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(os.path.dirname(HERE), "include")]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.basename(src).rsplit(".", 1)[0] + ".o")
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))
    if force or _newer(obj, [src] + headers):
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose or r.stderr.strip():
            sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    if force or _newer(LIB, objs):
        subprocess.check_call([NVCC, "-arch=sm_100a", "-shared", "-o", LIB] + objs)
    smoke_src = os.path.join(CSRC, "abi_smoke.cpp")
    smoke_bin = os.path.join(OBJ, "abi_smoke")
    if os.path.exists(smoke_src) and (force or _newer(smoke_bin, [smoke_src, LIB])):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(os.path.dirname(HERE), "include"),
                               smoke_src, "-o", smoke_bin, "-L", HERE, "-lwhisper_b200", "-Wl,-rpath,$ORIGIN/.."])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
