"""In-tree build of the sm_100a CUDA library (called by ``__graft_entry__.build()``).

nvcc cross-compiles without a GPU; the resulting ``<repo>/lib/libtetris_piclim_sm100.so`` is git-ignored
but travels with the repo snapshot to the GPU box.  The libraries are written to the short in-tree path
``<repo>/lib/`` (not next to the sources): the driver's loaded-library record drops /proc/self/maps lines as long
as this package's directory name makes them ("What's weak" 1).
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(os.path.dirname(HERE), "lib")
LIB_NAME = "libtetris_piclim_sm100.so"
LIB_PATH = os.path.join(LIB_DIR, LIB_NAME)
SOURCES = ["piclim_kernels.cu", "piclim_host_api.cu", "piclim_value.cu"]
HEADERS = ["piclim_core.cuh", "piclim_env.cuh", os.path.join("..", "..", "include", "tetris_piclim.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 library cannot be built")


def stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("TPL_NVCC_EXTRA", "").split()          # tuning experiments only (e.g. -DTPL_AS_MINBLOCKS=3)
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return LIB_PATH


CARVE_LIB_PATH = os.path.join(LIB_DIR, "libpiclim_carve.so")


def build_carve(force: bool = False) -> str:
    """The native prescribed-config generators (host code, g++): the carving generator of game/tetris.py and the
    forward generator + solver of game/tetris_algo_main."""
    srcs = [os.path.join(CSRC, f) for f in ("carve_gen.cpp", "forward_gen.cpp")]
    deps = srcs + [os.path.join(CSRC, "pyrandom.h")]
    os.makedirs(LIB_DIR, exist_ok=True)
    if force or not os.path.exists(CARVE_LIB_PATH) or any(os.path.getmtime(CARVE_LIB_PATH) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", CARVE_LIB_PATH] + srcs)
    return CARVE_LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
