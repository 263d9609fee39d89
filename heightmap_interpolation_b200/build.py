"""Build the CUDA shared library in-tree with nvcc for sm_100a (B200).

``python -m heightmap_interpolation_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting ``csrc/libhmi_b200.so`` is git-ignored but travels with the tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libhmi_b200.so")
SOURCES = ["hmi_capi.cu", "hmi_halo.cu", "hmi_group.cu", "hmi_ops.cu", "hmi_areas.cu", "hmi_generic.cu", "hmi_stream.cu", "hmi_pipe.cu", "hmi_stream_nl.cu", "hmi_init.cu",
           "hmi_xfer.cu", "hmi_host.cpp"]
HEADERS = ["hmi_common.cuh", "hmi_internal.cuh", "hmi_stream.cuh", "hmi_stream_nl.cuh", "hmi_stream_common.cuh", "hmi_init.cuh", "hmi_xfer.cuh", os.path.join("..", "..", "include", "hmi_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile every source to an object file (in parallel — hmi_stream.cu alone takes about a minute)
    and link them into csrc/libhmi_b200.so."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    if os.path.exists(LIB):          # a failed build must not leave the previous library behind
        os.remove(LIB)
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc()] + compile_flags + ["-c", "-o", obj, src]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        subprocess.check_call(cmd, cwd=CSRC)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                           "-o", LIB] + objs, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
