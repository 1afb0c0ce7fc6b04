"""CPU-side checks of the product library: libffgpu.so loads, exports every symbol that
include/ffgpu.h declares, and its host layer (encode_init / extradata / decode_init, pure C)
agrees with the oracle.  No GPU compute is attempted here."""
import ctypes as C
import os
import re

import pytest

import cpucodec as cc

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def F():
    import ffmpeg_ffv2_b200 as f
    return f


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ffgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ffgpu_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = C.CDLL(F().lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), name
    bound = {s[0] for s in F().codec.SYMBOLS}
    assert declared == bound, declared ^ bound
    assert lib.ffgpu_abi_version() == 2


FORMATS = ["yuv420p", "yuv444p", "yuv411p", "gray", "gray9le", "ya8", "yuva422p", "yuv420p9le",
           "yuv420p10le", "yuv440p10le", "yuv422p12le", "yuv444p14le", "yuv444p16le", "gray16le",
           "yuva420p10le", "bgr0", "bgra", "gbrp9le", "gbrp12le", "gbrp16le", "gbrap10le",
           "rgb48le", "rgba64le", "nosuchfmt"]
OPTIONS = [dict(), dict(slices=4), dict(slices=1023), dict(slices=9, coder=2), dict(coder=-2, context=1),
           dict(level=3, coder=1), dict(level=1), dict(level=0), dict(level=2), dict(level=4, strict=-2),
           dict(slices=7), dict(slicecrc=0, slices=4), dict(slicecrc=1), dict(context=2),
           dict(gop_size=1, slices=30)]


@pytest.mark.parametrize("size", [(3840, 2160), (1920, 1080), (352, 288), (97, 61)])
def test_encode_init_matches_oracle(size):
    w, h = size
    f = F()
    n = 0
    for fmt in FORMATS:
        for kw in OPTIONS:
            if kw.get("level") == 4:
                # version 4 is not restated by the oracle port: the compiled reference is the
                # checker, for the RGB layouts the product codes (see ff_stream_from_options)
                rgb = fmt in ("bgr0", "bgra") or fmt.startswith("gbr")
                if rgb:
                    g = f.FFV1Encoder(w, h, fmt, **kw)
                    assert g.info["version"] == 4 and g.info["micro_version"] == 2
                    if cc.available("ref"):
                        r = cc.Encoder("ref", w, h, fmt, **kw)
                        assert g.info == r.info and g.extradata == r.extradata, (fmt, kw)
                    d = f.FFV1Decoder(w, h, g.extradata)
                    assert d.pix_fmt is not None
                    n += 1
                else:
                    with pytest.raises(f.FFGpuError):
                        f.FFV1Encoder(w, h, fmt, **kw)
                continue
            try:
                o = cc.Encoder("oracle", w, h, fmt, **kw)
            except cc.CodecError as e:
                with pytest.raises(f.FFGpuError) as ei:
                    f.FFV1Encoder(w, h, fmt, **kw)
                assert ei.value.code == e.code, (fmt, kw)
                continue
            g = f.FFV1Encoder(w, h, fmt, **kw)
            assert g.info == o.info, (fmt, kw)
            assert g.extradata == o.extradata, (fmt, kw)
            if g.extradata:
                d = f.FFV1Decoder(w, h, g.extradata)
                od = cc.Decoder("oracle", w, h, g.extradata)
                assert d.pix_fmt is not None
                d.close()
                od.close()
            g.close()
            n += 1
    assert n > 100


def test_decode_init_rejects_bad_extradata():
    f = F()
    enc = f.FFV1Encoder(352, 288, "yuv420p", slices=4)
    ex = bytearray(enc.extradata)
    ex[5] ^= 0x40
    with pytest.raises(f.FFGpuError) as ei:
        f.FFV1Decoder(352, 288, bytes(ex))
    assert ei.value.code == f.INVALIDDATA
    with pytest.raises(f.FFGpuError):
        f.FFV1Decoder(0, 288, enc.extradata)


def test_frame_layout():
    f = F()
    n, planes = f.frame_layout("yuv420p10le", 3840, 2160)
    assert n == 3840 * 2160 * 2 * 3 // 2
    assert [p[3] for p in planes] == [7680, 3840, 3840] and [p[2] for p in planes] == [2160, 1080, 1080]
    n, planes = f.frame_layout("bgr0", 97, 61)
    assert planes[0][1] % 256 == 0 and planes[0][3] == 97 * 4
    with pytest.raises(ValueError):
        f.frame_layout("bogus", 16, 16)


def test_pixel_path_fails_loudly_without_a_gpu():
    """no CPU fallback: without a CUDA device the pixel path must error out"""
    import numpy as np
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    f = F()
    enc = f.FFV1Encoder(64, 48, "yuv420p", slices=4)
    planes = [np.zeros((48, 64), np.uint8), np.zeros((24, 32), np.uint8), np.zeros((24, 32), np.uint8)]
    with pytest.raises(f.FFGpuError) as ei:
        enc.encode(planes)
    assert ei.value.code == f.EXTERNAL


def test_e2e_driver_builds_and_links_against_the_abi():
    """tools/e2e_driver.c (bench.py's host loop) uses only include/ffgpu.h entry points"""
    import ctypes
    from ffmpeg_ffv2_b200 import build as b
    lib = ctypes.CDLL(b.build_e2e_driver())
    assert hasattr(lib, "ffgpu_e2e_run") and hasattr(lib, "ffgpu_e2e_free_packets")
    src = open(os.path.join(os.path.dirname(__file__), "..", "tools", "e2e_driver.c")).read()
    assert "oracle" not in src and "ffv1_" not in src.replace("ffgpu_ffv1_", "")


def test_staging_copies_of_the_library_without_a_gpu():
    """tests/emul/host_staging_test.cu includes ffgpu_api.cu and calls its static host
    functions: pictures between caller planes and the pinned staging layout, top-down and
    bottom-up (negative linesize), through the copy-thread pool and inline"""
    import shutil
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
    if not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        pytest.skip("nvcc not available")
    b = subprocess.run(["make", "-C", here, "host_staging_test"], capture_output=True, text=True)
    assert b.returncode == 0, b.stderr[-2000:]
    for threads in ("1", "5"):
        r = subprocess.run([os.path.join(here, "host_staging_test")], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, FFGPU_COPY_THREADS=threads))
        assert r.returncode == 0 and "host copy ok" in r.stdout, (threads, r.stdout[-500:], r.stderr[-1000:])
