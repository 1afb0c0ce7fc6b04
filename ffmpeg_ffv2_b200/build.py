"""Builds ffmpeg_ffv2_b200/libffgpu.so in-tree: gcc for the host C, nvcc (sm_100a only) for
the kernels and the C-ABI glue.  The CUDA runtime is linked statically so that the library
has no dependency on which libcudart a host application ships."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libffgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + \
           [os.path.join(HERE, "..", "include", "ffgpu.h")]


def build(force=False, verbose=False, variant=None, defines=(), csrc=None):
    """variant/defines/csrc: development builds for A/B runs (build/libffgpu_<variant>.so,
    loaded through the FFGPU_LIB environment variable); the product is the plain build"""
    global OUT, CSRC
    if variant:
        saved = OUT, CSRC
        OUT = os.path.join(HERE, "build", "libffgpu_%s.so" % variant)
        CSRC = csrc or CSRC
        try:
            return _build(True, verbose, "build/" + variant, list(defines))
        finally:
            OUT, CSRC = saved
    return _build(force, verbose, "build", [])


def _build(force, verbose, objsub, defines):
    if not force and not _newer(OUT, sources()):
        return OUT
    objdir = os.path.join(HERE, objsub)
    os.makedirs(objdir, exist_ok=True)
    run = lambda cmd: subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.PIPE,
                                     stderr=subprocess.STDOUT)
    host_o = os.path.join(objdir, "ffv1_host.o")
    run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-Wall", "-Wextra", "-c",
         os.path.join(CSRC, "ffv1_host.c"), "-o", host_o])
    objs = [host_o]
    for cu in ("ffv1_kernels.cu", "ffgpu_api.cu"):
        o = os.path.join(objdir, cu.replace(".cu", ".o"))
        run([NVCC] + ARCH + defines + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                             "-c", os.path.join(CSRC, cu), "-o", o])
        objs.append(o)
    run([NVCC] + ARCH + ["-shared", "-cudart", "static", "-o", OUT] + objs)
    return OUT


def build_e2e_driver(force=False):
    """tools/libffgpu_e2e.so: the C host loop bench.py times for the end-to-end number
    (two threads on the public C ABI, tools/e2e_driver.c)"""
    root = os.path.join(HERE, "..")
    src = os.path.join(root, "tools", "e2e_driver.c")
    out = os.path.join(root, "tools", "libffgpu_e2e.so")
    build(force=force)
    if not force and not _newer(out, [src, os.path.join(root, "include", "ffgpu.h"), OUT]):
        return out
    subprocess.run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra",
                    "-I", os.path.join(root, "include"), src, "-o", out,
                    "-L", HERE, "-lffgpu", "-Wl,-rpath,$ORIGIN/../ffmpeg_ffv2_b200", "-lpthread"],
                   check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_e2e_driver(force="--force" in sys.argv))
