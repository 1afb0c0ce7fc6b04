import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")
    # the CPU checkers are test infrastructure; build them on demand (gcc, seconds)
    oracle_so = os.path.join(ROOT, "oracle", "libffv1_oracle.so")
    src = os.path.join(ROOT, "oracle", "ffv1_oracle.c")
    if (not os.path.exists(oracle_so) or os.path.getmtime(oracle_so) < os.path.getmtime(src)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "libffv1_oracle.so"],
                       check=True, stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/libavcodec") and not os.path.exists(
            os.path.join(ROOT, "oracle", "_ref", "libffv1ref.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"],
                       check=True, stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/libavcodec") and os.path.exists(
            os.path.join(ROOT, "ffmpeg_ffv2_b200", "libffgpu.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "glue"],
                       check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


    # the product library and the CPU emulation of its device functions: rebuild when stale
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        from ffmpeg_ffv2_b200 import build as b
        b.build(force=False)
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emul")], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    # tests/test_host_pipeline_cpu.py re-runs GPU tests in a child process against the
    # library's host side over a stand-in device (FFGPU_LIB=tests/emul/cpu/libffgpu_cpu.so)
    if os.environ.get("FFGPU_HOST_PIPELINE_RUN") == "1":
        return
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
