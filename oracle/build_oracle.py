"""Build the C part of the oracle (TEST INFRASTRUCTURE) and, when /root/reference is present, the
reference's own chamfer CUDA extension into oracle/_ref/ (compiled from the sources where they lie;
no reference source is copied into this repo).

    python -m oracle.build_oracle          # C oracle (gcc)
    python -m oracle.build_oracle --ref    # + reference chamfer3D.cu as a torch extension (needs nvcc, ~40 s)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
REF_OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("ZEROSHAPE_REFERENCE", "/root/reference")
LIB = os.path.join(OUT, "liboracle_chamfer.so")


def build_c(force=False):
    src = os.path.join(HERE, "chamfer_ref.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, src, "-lm"], check=True)
    return LIB


def ref_chamfer_path():
    return os.path.join(REF_OUT, "chamfer_3D.so")


def build_ref_chamfer(force=False):
    """Compile the UNMODIFIED reference extension (external/chamfer3D/{chamfer_cuda.cpp,chamfer3D.cu})
    for sm_100a into oracle/_ref/chamfer_3D.so.  Only possible where /root/reference exists."""
    srcdir = os.path.join(REF_ROOT, "external", "chamfer3D")
    if not os.path.isdir(srcdir):
        return None
    out = ref_chamfer_path()
    if os.path.exists(out) and not force:
        return out
    os.makedirs(REF_OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name="chamfer_3D", sources=[os.path.join(srcdir, "chamfer_cuda.cpp"), os.path.join(srcdir, "chamfer3D.cu")],
         build_directory=REF_OUT, verbose=False,
         extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"])
    return out if os.path.exists(out) else None


if __name__ == "__main__":
    print(build_c(force="--force" in sys.argv))
    if "--ref" in sys.argv:
        print(build_ref_chamfer(force="--force" in sys.argv))
