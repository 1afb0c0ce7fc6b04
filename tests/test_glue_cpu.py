"""The FFmpeg side of the drop-in boundary without a GPU.

oracle/_ref/ffmpeg is the reference's own program with integration/ffv1_gpu.c linked in.  On a
GPU box tests/test_ffmpeg_dropin.py runs it against the real libffgpu.so.  Here the same
binary is given tests/emul/mock/libffgpu.so (tests/emul/mock_ffgpu.cpp): the C ABI's
launch-group contract -- planes held by pointer and touched only when a group is launched,
results in order, FFGPU_EAGAIN / FFGPU_EOF where ffgpu_api.cu returns them -- with the CPU
emulation of the product's device functions as the codec.  What is under test is the glue:
frame ownership while pictures are in flight, packet order and timestamps, drain, flush,
option plumbing, two-pass logs.  A second build of the stand-in with the address sanitizer
(LD_PRELOAD=libasan, leak check on) catches planes freed while the library still holds them
and anything the glue forgets to free.

`-c:v ffv1_gpu` must equal `-c:v ffv1` the way tests/fate-run.sh:188-210 (enc_dec) compares
codecs: framemd5 of the coded stream (packets) and of the decoded pictures."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FFMPEG = os.path.join(ROOT, "oracle", "_ref", "ffmpeg")
EMUL = os.path.join(HERE, "emul")

pytestmark = pytest.mark.skipif(not os.path.exists(FFMPEG), reason="oracle/_ref/ffmpeg not built")


def build(target):
    r = subprocess.run(["make", "-C", EMUL, target], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cannot build %s: %s" % (target, r.stderr[-300:]))


@pytest.fixture(scope="module")
def mock_dir():
    build("mock/libffgpu.so")
    return os.path.join(EMUL, "mock")


@pytest.fixture(scope="module")
def asan_env():
    build("mock-asan/libffgpu.so")
    lib = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(lib) or not os.path.exists(lib):
        pytest.skip("libasan.so not found")
    return {"LD_LIBRARY_PATH": os.path.join(EMUL, "mock-asan"), "LD_PRELOAD": lib,
            "ASAN_OPTIONS": "detect_leaks=1:abort_on_error=0:exitcode=97"}


def ffmpeg(libdir_or_env, *args, extra_env=None):
    env = dict(os.environ)
    if isinstance(libdir_or_env, dict):
        env.update(libdir_or_env)
    else:
        env["LD_LIBRARY_PATH"] = libdir_or_env
    env.update(extra_env or {})
    r = subprocess.run([FFMPEG, "-hide_banner", "-loglevel", "error", "-nostdin"] + list(args),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
    assert r.returncode == 0, r.stderr.decode(errors="replace")[-3000:]
    return r.stdout


def md5_lines(out):
    return [l for l in out.decode().splitlines() if l and not l.startswith("#")]


def src(w, h, n, fmt, graph="testsrc2=s={w}x{h}:r=25"):
    return ["-f", "lavfi", "-i", graph.format(w=w, h=h), "-frames:v", str(n), "-pix_fmt", fmt]


CASES = [
    # name, w, h, fmt, codec options, frames
    ("defaults-g12", 352, 288, "yuv420p", [], 27),                                # Golomb-Rice, carried states
    ("C2-shape", 320, 240, "yuv420p10le", ["-slices", "30", "-g", "1"], 19),
    ("rgb-ctx1", 160, 120, "bgr0", ["-coder", "range_tab", "-context", "1", "-g", "1"], 9),
    ("444p16", 160, 120, "yuv444p16le", ["-g", "1"], 9),
    ("gbrp10-g4", 160, 120, "gbrp10le", ["-slices", "4", "-g", "4"], 10),
    ("v4-bgra", 160, 120, "bgra", ["-level", "4", "-strict", "-2", "-g", "1"], 6),
    ("gray-nocrc", 161, 99, "gray", ["-slicecrc", "0", "-coder", "range_def", "-g", "1"], 7),
]


@pytest.mark.parametrize("name,w,h,fmt,opts,n", CASES, ids=[c[0] for c in CASES])
def test_ffmpeg_with_the_glue_equals_the_reference_codec(tmp_path, mock_dir, name, w, h, fmt, opts, n):
    common = src(w, h, n, fmt)
    cpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1", *opts, "-f", "framemd5", "-"))
    gpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1_gpu", *opts, "-f", "framemd5", "-"))
    assert len(cpu) == n and gpu == cpu                   # packets, their order, pts/dts, key flags
    nut = str(tmp_path / "ref.nut")
    ffmpeg(mock_dir, *common, "-c:v", "ffv1", *opts, "-y", nut)
    want = md5_lines(ffmpeg(mock_dir, "-c:v", "ffv1", "-i", nut, "-f", "framemd5", "-"))
    got = md5_lines(ffmpeg(mock_dir, "-c:v", "ffv1_gpu", "-i", nut, "-f", "framemd5", "-"))
    assert len(want) == n and got == want                 # pictures, their order, timestamps
    both = md5_lines(ffmpeg(mock_dir, "-c:v", "ffv1_gpu", "-i", nut, "-c:v", "ffv1_gpu", *opts, "-f", "framemd5", "-"))
    assert both == cpu


@pytest.mark.parametrize("batch,depth", [(1, 1), (1, 3), (3, 2), (5, 8), (1024, 1), (7, 0)])
def test_every_pipeline_shape_returns_the_stream_in_order(tmp_path, mock_dir, batch, depth):
    """groups of 1 .. more pictures than the stream has (everything comes out in the flush),
    1 .. 8 groups in flight: send_frame's EAGAIN, receive's EAGAIN while a group runs, the
    blocking receive when every group is busy, the drain"""
    n, opts = 23, ["-slices", "4", "-g", "1"]
    common = src(128, 96, n, "yuv420p10le")
    shape = ["-gpu_batch", str(batch), "-gpu_depth", str(depth)]
    cpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1", *opts, "-f", "framemd5", "-"))
    gpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1_gpu", *shape, *opts, "-f", "framemd5", "-"))
    assert len(cpu) == n and gpu == cpu
    nut = str(tmp_path / "ref.nut")
    ffmpeg(mock_dir, *common, "-c:v", "ffv1", *opts, "-y", nut)
    want = md5_lines(ffmpeg(mock_dir, "-c:v", "ffv1", "-i", nut, "-f", "framemd5", "-"))
    got = md5_lines(ffmpeg(mock_dir, "-c:v", "ffv1_gpu", *shape, "-i", nut, "-f", "framemd5", "-"))
    assert len(want) == n and got == want


def test_more_pictures_in_flight_than_the_ring_starts_with(tmp_path, mock_dir):
    """600 pictures held until the flush: the glue's frame ring (256 entries to begin with)
    grows, nothing is answered with EAGAIN on both sides"""
    n, opts = 600, ["-slices", "4", "-g", "1"]
    common = src(32, 32, n, "gray", "testsrc2=s={w}x{h}:r=25")
    shape = ["-gpu_batch", "1024", "-gpu_depth", "1"]
    cpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1", *opts, "-f", "framemd5", "-"))
    gpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1_gpu", *shape, *opts, "-f", "framemd5", "-"))
    assert len(cpu) == n and gpu == cpu
    nut = str(tmp_path / "ref.nut")
    ffmpeg(mock_dir, *common, "-c:v", "ffv1", *opts, "-y", nut)
    want = md5_lines(ffmpeg(mock_dir, "-c:v", "ffv1", "-i", nut, "-f", "framemd5", "-"))
    got = md5_lines(ffmpeg(mock_dir, "-c:v", "ffv1_gpu", *shape, "-i", nut, "-f", "framemd5", "-"))
    assert len(want) == n and got == want


def test_bottom_up_pictures_from_a_filter(mock_dir):
    """`-vf vflip` hands the encoder AVFrames with a negative linesize: the glue passes the
    planes of its own reference on as they are"""
    n, opts = 7, ["-slices", "4", "-g", "1"]
    common = src(128, 96, n, "yuv420p10le") + ["-vf", "vflip"]
    cpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1", *opts, "-f", "framemd5", "-"))
    gpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1_gpu", *opts, "-f", "framemd5", "-"))
    assert len(cpu) == n and gpu == cpu


@pytest.mark.parametrize("level", ["0", "1"])
def test_streams_that_announce_their_format_in_the_first_key_frame(tmp_path, mock_dir, level):
    """version 0 / 1: no extradata, the glue probes the first packet before it allocates"""
    n = 8
    common = src(176, 144, n, "yuv420p")
    avi = str(tmp_path / "v.avi")
    ffmpeg(mock_dir, *common, "-c:v", "ffv1", "-level", level, "-y", avi)
    want = md5_lines(ffmpeg(mock_dir, "-c:v", "ffv1", "-i", avi, "-f", "framemd5", "-"))
    got = md5_lines(ffmpeg(mock_dir, "-c:v", "ffv1_gpu", "-i", avi, "-f", "framemd5", "-"))
    assert len(want) == n and got == want
    cpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1", "-level", level, "-f", "framemd5", "-"))
    gpu = md5_lines(ffmpeg(mock_dir, *common, "-c:v", "ffv1_gpu", "-level", level, "-f", "framemd5", "-"))
    assert gpu == cpu


def test_input_played_twice_flushes_the_decoder_in_between(tmp_path, mock_dir):
    """-stream_loop 1: ffmpeg.c:4264-4273 drains the decoder at the end of the file, calls
    avcodec_flush_buffers and feeds it the file again -- the handle has to take packets after
    its EOF"""
    n, opts = 11, ["-slices", "4", "-g", "1"]
    nut = str(tmp_path / "ref.nut")
    ffmpeg(mock_dir, *src(128, 96, n, "yuv420p"), "-c:v", "ffv1", *opts, "-y", nut)
    want = md5_lines(ffmpeg(mock_dir, "-stream_loop", "1", "-c:v", "ffv1", "-i", nut, "-f", "framemd5", "-"))
    got = md5_lines(ffmpeg(mock_dir, "-stream_loop", "1", "-c:v", "ffv1_gpu", "-gpu_batch", "4", "-i", nut,
                           "-f", "framemd5", "-"))
    assert len(want) == 2 * n and got == want


def test_two_pass_logs_and_second_pass(tmp_path, mock_dir):
    """-pass 1 / -pass 2 through fftools: stats_out of the glue (ffmpeg.c:1326,1953 writes it
    to the pass log) and the packets of the second pass"""
    common = src(160, 120, 8, "yuv420p")
    opts = ["-coder", "range_tab", "-slices", "4", "-g", "1"]
    logs = {}
    for codec in ("ffv1", "ffv1_gpu"):
        log = str(tmp_path / codec)
        ffmpeg(mock_dir, *common, "-c:v", codec, *opts, "-pass", "1", "-passlogfile", log, "-f", "null", "-")
        logs[codec] = open(log + "-0.log").read()
    assert logs["ffv1"] == logs["ffv1_gpu"] and len(logs["ffv1"]) > 1000
    out = {}
    for codec in ("ffv1", "ffv1_gpu"):
        out[codec] = md5_lines(ffmpeg(mock_dir, *common, "-c:v", codec, *opts, "-pass", "2", "-passlogfile",
                                      str(tmp_path / "ffv1"), "-f", "framemd5", "-"))
    assert out["ffv1"] == out["ffv1_gpu"] and len(out["ffv1"]) == 8


def test_rejected_options_fail_like_the_reference(mock_dir):
    """encode_init's refusals reach the user as a failed ffmpeg run for both codecs (and the
    half-made context is cleaned up: FF_CODEC_CAP_INIT_CLEANUP)"""
    for opts in (["-level", "4"],                          # version 4 needs -strict experimental
                 ["-slices", "5000"]):
        for codec in ("ffv1", "ffv1_gpu"):
            env = dict(os.environ, LD_LIBRARY_PATH=mock_dir)
            r = subprocess.run([FFMPEG, "-hide_banner", "-loglevel", "error", "-nostdin",
                                *src(64, 48, 2, "yuv420p"), "-c:v", codec, *opts, "-f", "null", "-"],
                               capture_output=True, env=env, timeout=120)
            assert r.returncode != 0, (codec, opts)


def test_nothing_is_freed_early_or_left_behind(tmp_path, asan_env):
    """the same runs under the address sanitizer with the leak check on: the stand-in reads
    the planes only when a group is launched and writes the decoded pictures only then, so a
    frame the glue released too early is a use-after-free here, and whatever the glue or the
    stand-in forgot at close is a leak"""
    n, opts = 13, ["-slices", "4", "-g", "1"]
    common = src(128, 96, n, "yuv420p10le")
    shape = ["-gpu_batch", "5", "-gpu_depth", "2"]
    nut = str(tmp_path / "a.nut")
    ffmpeg(asan_env, *common, "-c:v", "ffv1_gpu", *shape, *opts, "-y", nut)
    got = md5_lines(ffmpeg(asan_env, "-c:v", "ffv1_gpu", *shape, "-i", nut, "-c:v", "ffv1_gpu", *shape, *opts,
                           "-f", "framemd5", "-"))
    assert len(got) == n
    # first pass (stats_out is allocated and released by the glue), and a refused open
    ffmpeg(asan_env, *common, "-c:v", "ffv1_gpu", "-coder", "range_tab", *opts, "-pass", "1", "-passlogfile",
           str(tmp_path / "p"), "-f", "null", "-")
    env = dict(os.environ)
    env.update(asan_env)
    r = subprocess.run([FFMPEG, "-hide_banner", "-loglevel", "error", "-nostdin", *common, "-c:v", "ffv1_gpu",
                        "-level", "4", "-f", "null", "-"], capture_output=True, env=env, timeout=120)
    assert r.returncode not in (0, 97), r.stderr.decode(errors="replace")[-2000:]
    assert b"AddressSanitizer" not in r.stderr and b"LeakSanitizer" not in r.stderr


def test_the_cuda_frames_half_of_the_glue_compiles():
    """`#if CONFIG_CUDA`: AV_PIX_FMT_CUDA frames into the encoder and out of the decoder
    (hw_configs, get_format negotiation, AVHWFramesContext, the context bracket).  This tree has
    no hwcontext_cuda to link against, so the code is type-checked against the reference's
    headers with CONFIG_CUDA=1; its library side (device pointers in ffgpu_picture /
    ffgpu_picture_out) is what test_gpu_parity and test_host_pipeline_cpu exercise"""
    if not (os.path.isdir("/root/reference/libavcodec") and os.path.exists("/usr/local/cuda/include/cuda.h")):
        pytest.skip("reference headers or cuda.h not available")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "glue-cuda-check"], capture_output=True, text=True)
    assert r.returncode == 0 and "warning" not in r.stderr, r.stderr[-3000:]
