"""The drop-in proof with the reference's own program: oracle/_ref/ffmpeg is the reference
tree's `ffmpeg` built by oracle/build_ffmpeg.sh with integration/ffv1_gpu.c added exactly as
INTEGRATION.md describes and linked against libffgpu.so.  `-c:v ffv1_gpu` must give the
packets (`-f framemd5` of the coded stream) and the decoded pictures (`-f framemd5` after
decoding) of `-c:v ffv1`, the way tests/fate-run.sh:188-210 (enc_dec) compares codecs, on the
lavfi sources BASELINE.json names."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FFMPEG = os.path.join(ROOT, "oracle", "_ref", "ffmpeg")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(FFMPEG), reason="oracle/_ref/ffmpeg not built")]

SOURCES = {
    "testsrc2": "testsrc2=s={w}x{h}:r=25",
    "mandelbrot": "mandelbrot=s={w}x{h}:r=25",
    "noise": "testsrc2=s={w}x{h}:r=25,noise=alls=100:allf=t+u:all_seed=1234",
}


def ffmpeg(*args, stdin=None):
    env = dict(os.environ)
    import ffmpeg_ffv2_b200 as F
    env["LD_LIBRARY_PATH"] = os.path.dirname(F.lib_path()) + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([FFMPEG, "-hide_banner", "-loglevel", "error", "-nostdin"] + list(args),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=150)
    assert r.returncode == 0, r.stderr.decode(errors="replace")[-2000:]
    return r.stdout


def md5_lines(out):
    return [l for l in out.decode().splitlines() if l and not l.startswith("#")]


CASES = [
    # BASELINE C1: every default (v3, 2x2 slices, Golomb-Rice, -g 12: carried states)
    ("C1-default", "testsrc2", 1920, 1080, "yuv420p", [], 25),
    # BASELINE C2 at its full size: reference maximum slice count, range coder, intra
    ("C2", "testsrc2", 3840, 2160, "yuv420p10le", ["-slices", "1023", "-g", "1"], 16),
    ("C2-mandelbrot", "mandelbrot", 3840, 2160, "yuv420p10le", ["-slices", "1023", "-g", "1"], 6),
    ("C2-noise", "noise", 1920, 1080, "yuv420p10le", ["-slices", "255", "-g", "1"], 6),
    # BASELINE C3: RGB with the large context model
    ("C3", "testsrc2", 1280, 720, "bgr0", ["-coder", "range_tab", "-context", "1", "-g", "1"], 6),
    ("C4-444p16", "testsrc2", 1280, 720, "yuv444p16le", ["-g", "1"], 6),
]


@pytest.mark.parametrize("name,src,w,h,fmt,opts,n", CASES)
def test_ffmpeg_ffv1_gpu_is_a_drop_in(tmp_path, name, src, w, h, fmt, opts, n):
    lavfi = SOURCES[src].format(w=w, h=h)
    common = ["-f", "lavfi", "-i", lavfi, "-frames:v", str(n), "-pix_fmt", fmt]
    # 1. encoder: byte-identical packets (framemd5 of the coded stream hashes the packets)
    cpu = md5_lines(ffmpeg(*common, "-c:v", "ffv1", *opts, "-f", "framemd5", "-"))
    gpu = md5_lines(ffmpeg(*common, "-c:v", "ffv1_gpu", *opts, "-f", "framemd5", "-"))
    assert len(cpu) == n and gpu == cpu, name
    # 2. decoder: the reference-encoded file decodes to identical pictures
    nut = str(tmp_path / "ref.nut")
    ffmpeg(*common, "-c:v", "ffv1", *opts, "-y", nut)
    want = md5_lines(ffmpeg("-c:v", "ffv1", "-i", nut, "-f", "framemd5", "-"))
    got = md5_lines(ffmpeg("-c:v", "ffv1_gpu", "-i", nut, "-f", "framemd5", "-"))
    assert len(want) == n and got == want, name
    # 3. transcode through both GPU codecs at once, like `ffmpeg -c:v ffv1_gpu -i in -c:v ffv1_gpu out`
    both = md5_lines(ffmpeg("-c:v", "ffv1_gpu", "-i", nut, "-c:v", "ffv1_gpu", *opts, "-f", "framemd5", "-"))
    assert both == cpu, name


def test_ffmpeg_two_gpus_option(tmp_path):
    """-gpus 2: the routing handle inside the product (two sub-handles; on a one-GPU box both
    sit on GPU 0 -- the ordering logic is the same)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs (the routing itself is covered by test_gpu_parity)")
    common = ["-f", "lavfi", "-i", SOURCES["testsrc2"].format(w=1920, h=1080), "-frames:v", "40",
              "-pix_fmt", "yuv420p10le"]
    opts = ["-slices", "255", "-g", "1"]
    cpu = md5_lines(ffmpeg(*common, "-c:v", "ffv1", *opts, "-f", "framemd5", "-"))
    gpu = md5_lines(ffmpeg(*common, "-c:v", "ffv1_gpu", "-gpus", "2", *opts, "-f", "framemd5", "-"))
    assert gpu == cpu


def test_ffmpeg_two_pass(tmp_path):
    """`-pass 1` / `-pass 2` through fftools: the pass log the GPU codec leaves (stats_out,
    written by ffmpeg.c:1326,1953) and the second pass's packets equal the CPU codec's"""
    common = ["-f", "lavfi", "-i", SOURCES["testsrc2"].format(w=640, h=360), "-frames:v", "10",
              "-pix_fmt", "yuv420p"]
    opts = ["-coder", "range_tab", "-slices", "4", "-g", "1"]
    logs = {}
    for codec in ("ffv1", "ffv1_gpu"):
        log = str(tmp_path / codec)
        ffmpeg(*common, "-c:v", codec, *opts, "-pass", "1", "-passlogfile", log, "-f", "null", "-")
        logs[codec] = open(log + "-0.log").read()
    assert logs["ffv1"] == logs["ffv1_gpu"] and len(logs["ffv1"]) > 1000
    out = {}
    for codec in ("ffv1", "ffv1_gpu"):
        out[codec] = md5_lines(ffmpeg(*common, "-c:v", codec, *opts, "-pass", "2", "-passlogfile",
                                      str(tmp_path / "ffv1"), "-f", "framemd5", "-"))
    assert out["ffv1"] == out["ffv1_gpu"] and len(out["ffv1"]) == 10
