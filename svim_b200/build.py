"""Build the native pieces in-tree (they travel to the GPU box with the snapshot):

  svim_b200/libsvimgpu.so          CUDA kernels + C ABI (include/svimgpu.h), sm_100a only
  svim_b200/synth/libsvimsynth.so  synthetic-input generator (host C++)
  svim_b200/libsvimio.so           BAM reader / writer (host C++)
  svim_b200/_svimfastobj.so        CPython extension: Signature / SignatureCluster objects built in one C loop (csrc_host/fastobj.c)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libsvimgpu.so")
SRC = os.path.join(HERE, "csrc")


def nvcc_path():
    for c in ("/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def run_atomic(cmd, out):
    """Run a compiler command whose output argument is the literal string "@OUT@", writing beside `out` and renaming into place:
    a concurrent rank (torchrun) or a gpurun snapshot never sees a half-written library."""
    tmp = "%s.tmp%d" % (out, os.getpid())
    try:
        subprocess.check_call([tmp if a == "@OUT@" else a for a in cmd])
        os.replace(tmp, out)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    return out


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "..", "include", "svimgpu.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_gpu(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
           "-Xcompiler", "-fPIC,-Wno-deprecated-declarations", "-diag-suppress", "177,550,1444", "-shared", "-cudart", "static", "-o", "@OUT@", os.path.join(SRC, "api.cu"), "-lnccl", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    return run_atomic(cmd, SO)


def build_all(force=False):
    from . import synth, io
    build_gpu(force)
    synth.build(force)
    io.build_bamio(force)
    from . import fastobj
    fastobj.build(force)


if __name__ == "__main__":
    build_gpu(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
