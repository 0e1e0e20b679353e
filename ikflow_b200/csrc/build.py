"""Builds libikflow_b200.so (sm_100a only) in-tree with nvcc.  Used by ``__graft_entry__.build()`` and ``make``-less
developer builds:  ``python -m ikflow_b200.csrc.build``.

The shared library is the product; there is no CPU fallback.  It links the CUDA runtime statically so that the only
run-time dependency is the driver.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB_DIR = os.path.join(os.path.dirname(HERE), "lib")
LIB_PATH = os.path.join(LIB_DIR, "libikflow_b200.so")
SOURCES = ["api.cu", "robot.cu", "flow.cu"]
HEADERS = ["common.h", "flow_common.cuh", "flow_mma.cuh", "flow_umma.cuh", os.path.join(ROOT, "include", "ikflow_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [
        _nvcc(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "-O3", "-std=c++17",
        "--shared", "-Xcompiler", "-fPIC",
        "-cudart", "static",
        "-o", LIB_PATH,
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(HERE, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
