"""The HOST side of libffgpu.so without a GPU.

tests/emul/cpu/libffgpu.so is the unmodified csrc/ffgpu_api.cu (handles, launch groups,
routing over several devices, staging, packet arenas, damage bookkeeping, error paths)
compiled as C++ over a stand-in CUDA runtime (tests/emul/fake_cuda.cpp: host memory, no
asynchrony) and CPU launchers (tests/emul/fake_kernels.cpp: the product's own
__host__ __device__ slice functions, one loop iteration per work item, the coder form chosen
from the launch shape like the real launchers).  It is test infrastructure: only this file
loads it (through FFGPU_LIB, in child processes); the product has no CPU path.

With it the `-m gpu` tests themselves -- written against the public C ABI -- run here: what
they then check is everything in the library except the CUDA kernels' own wrappers, which
only the B200 run covers."""
import os
import subprocess
import sys
import textwrap

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CPU_DIR = os.path.join(HERE, "emul", "cpu")
CPU_LIB = os.path.join(CPU_DIR, "libffgpu.so")


@pytest.fixture(scope="module")
def cpu_env():
    if not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("CUDA headers not available")
    r = subprocess.run(["make", "-C", os.path.join(HERE, "emul"), "cpu/libffgpu.so"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    env = dict(os.environ)
    env.update(FFGPU_HOST_PIPELINE_RUN="1", FFGPU_LIB=CPU_LIB, FAKE_CUDA_DEVICES="2",
               LD_LIBRARY_PATH=CPU_DIR + ":" + env.get("LD_LIBRARY_PATH", ""))
    return env


def test_the_gpu_tests_through_the_host_pipeline(cpu_env):
    """every `-m gpu` test that fits a CPU (no 4K/8K pictures, no torch.cuda tensors): parity
    with the oracle through send/receive and the synchronous calls, every coder form, two
    "devices" behind one routing handle, two-pass, version 4, damaged packets, wide slice
    headers, prefix-cache overflow, incompressible pictures that outgrow the buffers,
    bottom-up pictures, the libavcodec glue under the vtable harness and inside the
    reference's ffmpeg (defaults with carried states, two-pass logs)"""
    select = ("not full_size and not resident and not C2 and not C3 and not C4 and not two_gpus_option")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_parity.py"),
                        os.path.join(HERE, "test_ffmpeg_dropin.py"), "-m", "gpu", "-q", "-n", "4", "-k", select,
                        "-p", "no:cacheprovider"],
                       capture_output=True, text=True, env=cpu_env, timeout=1500, cwd=ROOT)
    tail = r.stdout[-4000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


ALLOC_FAILURES = textwrap.dedent("""
    import ctypes as C, os, sys
    sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
    import numpy as np
    import ffmpeg_ffv2_b200 as F, cpucodec as cc, synth
    lib = F.lib()
    live = lib.fake_cuda_live_blocks; live.restype = C.c_long
    reset = lib.fake_cuda_reset_alloc_counter
    w, h, fmt = 96, 64, "yuv420p10le"
    kw = dict(slices=4, gop_size=1, max_batch=2, pipeline_depth=2)
    src = synth.testsrc2_like(fmt, w, h, 0)
    want = cc.Encoder("oracle", w, h, fmt, slices=4, gop_size=1).encode(src)
    assert live() == 0
    failed = ok = 0
    for k in range(1, 400):
        enc = F.FFV1Encoder(w, h, fmt, **kw)
        dec = F.FFV1Decoder(w, h, enc.extradata, max_batch=2, pipeline_depth=2)
        os.environ["FAKE_CUDA_FAIL_ALLOC"] = str(k); reset()
        hit = False
        try:
            pkt = enc.encode(src)
        except F.FFGpuError as e:
            assert e.code in (F.EXTERNAL, -12), e
            hit = True
            # a failure while the device side is being created releases all of it; later ones
            # (the staging of the first pageable picture) leave a working handle
            assert live() == 0 or enc.launches == 0 and k > 20, ("device memory kept after a failed start", k, live())
        if not hit:
            try:
                out = dec.decode(want)
            except F.FFGpuError as e:
                assert e.code in (F.EXTERNAL, -12), e
                hit = True
        os.environ["FAKE_CUDA_FAIL_ALLOC"] = "0"
        # the next call starts over on the same handles
        assert enc.encode(src) == want, k
        out = dec.decode(want)
        assert all(np.array_equal(a, b) for a, b in zip(out, src)), k
        enc.close(); dec.close()
        assert live() == 0, ("leak after close", k, live())
        failed += hit; ok += not hit
        if not hit:
            break
    assert failed >= 20 and ok == 1, (failed, ok)
    print("alloc failures ok", failed)
""")


def test_a_failed_allocation_leaves_nothing_behind(cpu_env):
    """the k-th device or pinned allocation of a starting handle fails, for every k: the call
    reports the CUDA error, nothing stays allocated, the NEXT call on the same handle starts
    over and succeeds, and close() returns every block (the stand-in runtime counts them)"""
    r = subprocess.run([sys.executable, "-c", ALLOC_FAILURES % dict(root=ROOT, tests=HERE)],
                       capture_output=True, text=True, env=cpu_env, timeout=600)
    assert r.returncode == 0 and "alloc failures ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_the_host_pipeline_under_the_sanitizers(cpu_env):
    """the same library built with -fsanitize=address,undefined: the stand-in runtime hands out
    ordinary heap blocks, so every "device" array (token and bitstream arenas, line scratch,
    state rows, packet arenas, result tables) has red zones, and the device functions as well
    as the host code around them are checked for out-of-bounds accesses and undefined
    behaviour on damaged, incompressible, wide-header and routed streams"""
    emul = os.path.join(HERE, "emul")
    r = subprocess.run(["make", "-C", emul, "cpu-asan/libffgpu.so"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available: " + r.stderr[-300:])
    libs = [subprocess.run(["gcc", "-print-file-name=" + n], capture_output=True, text=True).stdout.strip()
            for n in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(p) and os.path.exists(p) for p in libs):
        pytest.skip("sanitizer runtimes not found")
    env = dict(cpu_env)
    d = os.path.join(emul, "cpu-asan")
    env.update(FFGPU_LIB=os.path.join(d, "libffgpu.so"), LD_PRELOAD=" ".join(libs),
               ASAN_OPTIONS="detect_leaks=0:abort_on_error=0",
               LD_LIBRARY_PATH=d + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    select = ("every_form or damaged or wider or prefix_sets or incompressible or several_gpus or pipelined "
              "or bottom_up or version4_ycbcr or (version4_rgb and (bgra or gbrp16le)) or "
              "((encoder_packets or decoder_pictures) and (yuv420p10le or ya8 or rgb48le or yuva420p))")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_parity.py"), "-m", "gpu", "-q",
                        "-n", "4", "-k", select, "-p", "no:cacheprovider"],
                       capture_output=True, text=True, env=env, timeout=1500, cwd=ROOT)
    tail = r.stdout[-4000:] + r.stderr[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout and "Sanitizer" not in tail, tail


UNRESET_STATES = textwrap.dedent("""
    import sys
    sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
    import ffmpeg_ffv2_b200 as F, cpucodec as cc, synth
    w, h = 96, 64
    hits = 0
    for fmt, kw in (("ya8", dict(slices=6, gop_size=1)), ("yuv420p", dict(slices=4, gop_size=1)),
                    ("yuv420p", dict(slices=4, gop_size=1, level=4, strict=-2))):
        ref = cc.Encoder("oracle" if kw.get("level") != 4 else "ref", w, h, fmt, **kw)
        pk = [ref.encode(synth.noise(fmt, w, h, i)) for i in range(2)]
        for b in range(256):
            # the first byte of a packet carries the key-frame bit of the range coder
            # both pictures in one launch group: the second one's state slot has never been used
            dec = F.FFV1Decoder(w, h, ref.extradata, max_batch=3, pipeline_depth=2)
            try:
                for f, p in enumerate((pk[0], bytes([b]) + pk[1][1:])):
                    while not dec.send_packet(p, pts=f, dst=dec.alloc_picture()):
                        dec.receive_frame()
                dec.send_packet(None)
                while True:
                    r = dec.receive_frame()
                    if r == F.EOF:
                        break
                    hits += r is not None and not r[0].key_frame
            except F.FFGpuError:
                pass
            dec.close()
    assert hits > 0, "no packet was taken for a non-key frame"
    print("unreset states ok", hits)
""")


def test_states_that_no_key_frame_has_reset(cpu_env):
    """a packet whose key-frame bit is cleared in an intra-only Golomb-Rice stream decodes with
    states no key frame initialised.  The arenas start zeroed like the reference's av_mallocz'ed
    ones: on the stand-in runtime's 0xA5-filled allocations get_vlc_symbol's k search
    (ffv1dec.c:77-81) met count == 0 and did not end (found by tests/emul/api_fuzz.py)"""
    emul = os.path.join(HERE, "emul")
    r = subprocess.run(["make", "-C", emul, "cpu-asan/libffgpu.so"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available: " + r.stderr[-300:])
    libs = [subprocess.run(["gcc", "-print-file-name=" + n], capture_output=True, text=True).stdout.strip()
            for n in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(p) and os.path.exists(p) for p in libs):
        pytest.skip("sanitizer runtimes not found")
    env = dict(cpu_env)
    d = os.path.join(emul, "cpu-asan")
    env.update(FFGPU_LIB=os.path.join(d, "libffgpu.so"), LD_PRELOAD=" ".join(libs),
               ASAN_OPTIONS="detect_leaks=0:abort_on_error=0",
               LD_LIBRARY_PATH=d + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([sys.executable, "-c", UNRESET_STATES % dict(root=ROOT, tests=HERE)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "unreset states ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_damaged_streams_through_the_public_api_under_the_sanitizers(cpu_env):
    """tests/emul/api_fuzz.py: mutated packets and extradata into decode_frame, send_packet /
    receive_frame and a routing handle over two stand-in devices, sanitizer build"""
    emul = os.path.join(HERE, "emul")
    r = subprocess.run(["make", "-C", emul, "cpu-asan/libffgpu.so"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available: " + r.stderr[-300:])
    libs = [subprocess.run(["gcc", "-print-file-name=" + n], capture_output=True, text=True).stdout.strip()
            for n in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(p) and os.path.exists(p) for p in libs):
        pytest.skip("sanitizer runtimes not found")
    env = dict(cpu_env)
    env.update(FFGPU_LIB=os.path.join(emul, "cpu-asan", "libffgpu.so"), LD_PRELOAD=" ".join(libs),
               ASAN_OPTIONS="detect_leaks=0:abort_on_error=0")
    for seed in ("1", "9"):
        r = subprocess.run([sys.executable, os.path.join(emul, "api_fuzz.py"), seed, "40"],
                           capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0 and "api fuzz ok" in r.stdout, (seed, r.stdout[-1000:], r.stderr[-3000:])


def test_two_caller_threads_under_the_thread_sanitizer():
    """tests/emul/tsan_e2e.c: bench.py's C host loop (an encoder thread and a decoder thread on
    the public C ABI, tools/e2e_driver.c) over the host-pipeline sources built with
    -fsanitize=thread: routing handles over two stand-in devices, pageable pictures through
    the shared copy-thread pool; 64 pictures must come back identical and race-free"""
    emul = os.path.join(HERE, "emul")
    if not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("CUDA headers not available")
    b = subprocess.run(["make", "-C", emul, "tsan_e2e"], capture_output=True, text=True)
    if b.returncode != 0:
        pytest.skip("thread sanitizer build not available: " + b.stderr[-300:])
    r = subprocess.run([os.path.join(emul, "tsan_e2e")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, FAKE_CUDA_DEVICES="2", FFGPU_COPY_THREADS="4"))
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "decoded=64" in out and "mismatching planes: 0" in out, out[-3000:]
    assert "ThreadSanitizer" not in out, out[-3000:]


DEVICE_BATCHES = textwrap.dedent("""
    import ctypes as C, sys
    sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
    import numpy as np
    import ffmpeg_ffv2_b200 as F, cpucodec as cc, synth
    lib = F.lib()
    lib.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    lib.cudaFree.argtypes = [C.c_void_p]
    def dev_alloc(n):
        p = C.c_void_p()
        assert lib.cudaMalloc(C.byref(p), n) == 0
        return p.value
    w, h, fmt = 320, 180, "yuv420p10le"
    kw = dict(slices=20, gop_size=1)
    n = 5
    frame_bytes, planes = F.frame_layout(fmt, w, h)
    host = np.zeros((n, frame_bytes), np.uint8)
    srcs = [synth.testsrc2_like(fmt, w, h, i) for i in range(n)]
    for i, src in enumerate(srcs):
        for (off, pitch, rows, rb), a in zip(planes, src):
            host[i, off:off + pitch * rows].reshape(rows, pitch)[:, :rb] = a
    # --- ffgpu_ffv1_encode_device / _device_fetch / ffgpu_ffv1_decode_device / _device_status
    d_in = dev_alloc(host.nbytes)
    C.memmove(d_in, host.ctypes.data, host.nbytes)
    enc = F.FFV1Encoder(w, h, fmt, max_batch=8, **kw)
    enc.encode_device(d_in, n)
    ref = cc.Encoder("oracle", w, h, fmt, **kw)
    want = [ref.encode(s) for s in srcs]
    pkts = [enc.device_fetch(i) for i in range(n)]
    assert pkts == want
    dec = F.FFV1Decoder(w, h, enc.extradata, max_batch=8)
    d_out = dev_alloc(host.nbytes)
    dec.decode_device(pkts, d_out)
    assert dec.device_status() == [0] * n
    back = np.zeros_like(host)
    C.memmove(back.ctypes.data, d_out, host.nbytes)
    for i in range(n):
        for (off, pitch, rows, rb), a in zip(planes, srcs[i]):
            assert np.array_equal(back[i, off:off + pitch * rows].reshape(rows, pitch)[:, :rb], a)
    # a damaged picture in the batch is reported, the others are clean
    bad = bytearray(pkts[2]); bad[len(bad) // 2] ^= 0x40
    dec.decode_device(pkts[:2] + [bytes(bad)] + pkts[3:], d_out)
    st = dec.device_status()
    assert st[2] >= 1 and st[:2] == [0, 0] and st[3:] == [0, 0], st
    # more pictures than the handle was opened for
    try:
        dec.decode_device(pkts * 3, d_out)
        raise SystemExit("oversized batch accepted")
    except F.FFGpuError as e:
        assert e.code == F.EINVAL
    # --- device pointers in ffgpu_picture / ffgpu_picture_out (AV_PIX_FMT_CUDA frames), own pitch
    enc2 = F.FFV1Encoder(w, h, fmt, max_batch=2, pipeline_depth=2, **kw)
    dec2 = F.FFV1Decoder(w, h, enc2.extradata, max_batch=2, pipeline_depth=2)
    got = []
    for i, src in enumerate(srcs):
        pic = F.codec.Picture()
        for k, a in enumerate(src):
            padded = np.ascontiguousarray(np.pad(a, ((0, 0), (0, 64))))
            p = dev_alloc(padded.nbytes)
            C.memmove(p, padded.ctypes.data, padded.nbytes)
            pic.data[k] = p
            pic.linesize[k] = padded.strides[0]
        pic.sar_num, pic.sar_den, pic.pts = 0, 1, i
        while True:
            r = lib.ffgpu_ffv1_encode_send_frame(enc2.h, C.byref(pic))
            if r == 0:
                break
            assert r == F.EAGAIN, r
            got.append(enc2.receive_packet()[0])
    enc2.send_frame(None)
    while True:
        r = enc2.receive_packet()
        if r == F.EOF:
            break
        if r is not None:
            got.append(r[0])
    assert got == want
    outs = []
    for i, p in enumerate(want):
        po = F.codec.PictureOut()
        shapes = []
        for k, a in enumerate(srcs[i]):
            pitch = a.shape[1] + 48
            ptr = dev_alloc(pitch * a.shape[0])
            po.data[k] = ptr
            po.linesize[k] = pitch
            shapes.append((ptr, pitch, a.shape))
        outs.append(shapes)
        while not dec2.send_packet(p, pts=i, dst=(po, None)):
            assert dec2.receive_frame() not in (None, F.EOF)
    dec2.send_packet(None)
    while dec2.receive_frame() != F.EOF:
        pass
    for i, shapes in enumerate(outs):
        for (ptr, pitch, shape), a in zip(shapes, srcs[i]):
            buf = np.zeros((shape[0], pitch), np.uint8)
            C.memmove(buf.ctypes.data, ptr, buf.nbytes)
            assert np.array_equal(buf[:, :shape[1]], a), i
    print("device batches ok")
""")


def test_device_resident_batches_and_device_pointer_planes(cpu_env):
    """the entry points for pictures that stay in device memory (encode_device, device_fetch,
    decode_device, decode_device_status incl. a damaged picture and an oversized batch) and
    device pointers in ffgpu_picture / ffgpu_picture_out with a pitch of their own, on
    "device" memory of the stand-in runtime"""
    r = subprocess.run([sys.executable, "-c", DEVICE_BATCHES % dict(root=ROOT, tests=HERE)],
                       capture_output=True, text=True, env=cpu_env, timeout=600)
    assert r.returncode == 0 and "device batches ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_random_call_sequences_keep_the_stream_in_order(cpu_env):
    """tests/emul/api_sequences.py: pictures / packets sent, results received and streams
    flushed (also mid-stream) in random order, one device and routing handles, every group
    size and depth, carried-state streams routed by whole GOPs: complete, ordered,
    oracle-identical output, never EAGAIN on both sides, input accepted again after an EOF"""
    for seed in ("1", "7"):
        r = subprocess.run([sys.executable, os.path.join(HERE, "emul", "api_sequences.py"), seed, "40"],
                           capture_output=True, text=True, env=cpu_env, timeout=900)
        assert r.returncode == 0 and "api sequences ok" in r.stdout, (seed, r.stdout[-1000:], r.stderr[-3000:])
