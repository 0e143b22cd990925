"""Builds libeno.so (the C-ABI shared library of include/eno.h) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc
cross-compiles without a GPU.  Every translation unit is compiled to an object under
build/ (recompiled only when one of the headers it includes, transitively, is newer) and the
objects are linked into easy-neural-ode_b200/libeno.so.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libeno.so")
SOURCES = [os.path.join(CSRC, "eno_abi.cu"), os.path.join(CSRC, "css_engine.cu"), os.path.join(CSRC, "lat_engine.cu")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr", "--extended-lambda",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps(path, seen=None):
    """The file and every header it includes with quotes, transitively."""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for inc in re.findall(r'^\s*#include\s+"([^"]+)"', open(path).read(), flags=re.M):
        _deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _obj(src, tag=""):
    return os.path.join(OBJ, os.path.basename(src).replace(".cu", tag + ".o"))


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in _deps(src))


def _compile(src, obj, extra, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed compiling {os.path.basename(src)}")
    if verbose:
        sys.stderr.write(res.stderr)


def build(force=False, verbose=False, phases=False):
    """Compile what is stale and link; returns the library path.
    phases=True builds the development variant libeno_phases.so with in-kernel phase clocks."""
    os.makedirs(OBJ, exist_ok=True)
    tag, extra, out = ("_phases", ["-DENO_PHASES"], os.path.join(HERE, "libeno_phases.so")) if phases else ("", [], LIB)
    todo = [(s, _obj(s, tag)) for s in SOURCES if force or _stale(s, _obj(s, tag))]
    if todo:
        with ThreadPoolExecutor(len(todo)) as ex:
            list(ex.map(lambda so: _compile(so[0], so[1], extra, verbose), todo))
    objs = [_obj(s, tag) for s in SOURCES]
    if todo or not os.path.exists(out) or any(os.path.getmtime(o) > os.path.getmtime(out) for o in objs):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + objs + ["-o", out]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed linking " + os.path.basename(out))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, phases="--phases" in sys.argv))
