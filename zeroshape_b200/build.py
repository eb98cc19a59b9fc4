"""Build recipe for libzeroshape_b200.so (plain nvcc, sm_100a only, in-tree output).

    python -m zeroshape_b200.build            # build if stale
    python -m zeroshape_b200.build --force

The library has no torch dependency (C ABI, raw pointers); it links the static CUDA runtime.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libzeroshape_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps_mtime():
    files = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(HERE, "..", "include", "zeroshape_b200.h")]
    return max(os.path.getmtime(f) for f in files)


def is_stale():
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _deps_mtime()


def _compile(src, force_all):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    hdr_m = max([os.path.getmtime(f) for f in glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))]
                + [os.path.getmtime(os.path.join(HERE, "..", "include", "zeroshape_b200.h"))])
    if not force_all and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_m):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(s, force), _sources()))
    objs = [o for o, _ in results]
    log = "\n".join(l for _, l in results if l)
    with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
        f.write(log)
    if verbose:
        print(log)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
