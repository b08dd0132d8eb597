"""In-tree build of libaps_b200.so (sm_100a only) with plain nvcc — no torch extension machinery.

    python -m aps_b200.build [--force]

Objects go to <repo>/build/, the library to aps_b200/libaps_b200.so (git-ignored, but it travels
to the GPU box with the gpurun snapshot).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# tuning builds: APS_B200_VARIANT=<tag> + APS_B200_NVCC_EXTRA="-DFOO=1" -> libaps_b200_<tag>.so, build/<tag>/
VARIANT = os.environ.get("APS_B200_VARIANT", "")
OBJ = os.path.join(ROOT, "build", VARIANT) if VARIANT else os.path.join(ROOT, "build")
LIB = os.path.join(HERE, f"libaps_b200_{VARIANT}.so" if VARIANT else "libaps_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"
              ] + os.environ.get("APS_B200_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libaps_b200.so cannot be built")
    return exe


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str) -> str:
    h = hashlib.sha1()
    deps = [path] + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
            os.path.join(ROOT, "include", "aps_b200.h")
        ]
    for d in deps:
        with open(d, "rb") as fd:
            h.update(fd.read())
    h.update(" ".join(ARCH + NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, force: bool) -> str:
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp = obj + ".sha1"
    dig = _digest(path)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj
    cmd = [_nvcc()] + ARCH + NVCC_FLAGS + ["-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as fd:
        fd.write(dig)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        objs = list(pool.map(lambda s: _compile(s, force), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(f"built {LIB} from {len(srcs)} sources")
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
