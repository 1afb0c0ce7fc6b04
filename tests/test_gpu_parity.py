"""GPU parity tests proper: the CUDA path, called through the C ABI (libffgpu.so via
ffmpeg_ffv2_b200), against the oracle on the same seeded inputs -- bit-exact packets for
the encoder, identical pictures for the decoder -- plus the FATE-chained golden packets
and size-independent round-trip properties at the BASELINE.json sizes."""
import hashlib
import json
import os

import numpy as np
import pytest

import cpucodec as cc
import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def gpu():
    import ffmpeg_ffv2_b200 as F
    return F


def md5(b):
    return hashlib.md5(b).hexdigest()


FORMATS_SMALL = [
    "yuv420p", "yuv444p", "yuv422p", "yuv410p", "gray", "ya8", "yuva420p",
    "yuv420p10le", "yuv422p10le", "yuv444p16le", "yuv420p16le", "gray16le", "gray10le",
    "yuva444p10le", "bgr0", "bgra", "gbrp10le", "gbrp16le", "gbrap12le", "rgb48le", "rgba64le",
]
OPTIONS = [
    dict(),
    dict(slices=4),
    dict(slices=9, coder=2),
    dict(slices=4, coder=-2, context=1),
    dict(level=3, coder=0, context=1),
    dict(level=1, coder=1),
    dict(level=3, slicecrc=0, gop_size=1, slices=12),
    dict(level=3, slices=30, context=1, gop_size=1),
]


@pytest.mark.parametrize("fmt", FORMATS_SMALL)
def test_encoder_packets_match_oracle(fmt):
    F = gpu()
    w, h = 97, 61
    for kw in OPTIONS:
        try:
            ref = cc.Encoder("oracle", w, h, fmt, **kw)
        except cc.CodecError as e:
            with pytest.raises(F.FFGpuError) as ei:
                F.FFV1Encoder(w, h, fmt, **kw)
            assert ei.value.code == e.code
            continue
        enc = F.FFV1Encoder(w, h, fmt, **kw)
        assert enc.info == ref.info and enc.extradata == ref.extradata
        for kind in ("smooth", "noise", "extremes", "testsrc2"):
            e2 = F.FFV1Encoder(w, h, fmt, **kw)
            r2 = cc.Encoder("oracle", w, h, fmt, **kw)
            for fr in range(3):
                planes = synth.GENERATORS[kind](fmt, w, h, fr)
                assert e2.encode(planes) == r2.encode(planes), (fmt, kw, kind, fr)
            e2.close()
        enc.close()


# The slice coders exist in several forms that the library picks by launch shape (slices per
# launch): one slice per warp ("lone", the straight-line coders), one slice per warp with the
# warp-wide loops, stage B in two halves, and one slice per lane.  The small pictures of this
# file all take the first by default; the tuning hooks force the others.
CODER_FORMS = [
    dict(),
    dict(FFGPU_LONE="0"),
    dict(FFGPU_LONE="0", FFGPU_SPLIT="0"),
    dict(FFGPU_LANE_STRIDE="1"),
    dict(FFGPU_LANE_STRIDE="4"),
]


@pytest.mark.parametrize("env", CODER_FORMS, ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()) or "default")
def test_every_form_of_the_slice_coders_matches_the_oracle(env, monkeypatch):
    F = gpu()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    w, h = 211, 83
    for fmt, kw in [("yuv420p10le", dict(slices=9, gop_size=1)), ("yuv444p", dict(slices=4, coder=1, gop_size=1)),
                    ("yuv420p", dict(slices=4, coder=-2, context=1)), ("gray16le", dict(slices=4)),
                    ("bgr0", dict(slices=4, coder=2, context=1, gop_size=1)), ("yuv422p", dict(coder=0, slices=4))]:
        for kind in ("testsrc2", "noise", "smooth"):
            enc = F.FFV1Encoder(w, h, fmt, **kw)
            ref = cc.Encoder("oracle", w, h, fmt, **kw)
            rdec = cc.Decoder("oracle", w, h, ref.extradata)
            dec = F.FFV1Decoder(w, h, ref.extradata)
            for fr in range(3):
                planes = synth.GENERATORS[kind](fmt, w, h, fr)
                pkt = ref.encode(planes)
                assert enc.encode(planes) == pkt, (env, fmt, kind, fr)
                want = rdec.decode(pkt)
                got = dec.decode(pkt, fmt_hint=rdec.pix_fmt)
                for a, b in zip(want, got):
                    assert np.array_equal(a, b), (env, fmt, kind, fr)
            enc.close()
            dec.close()


@pytest.mark.parametrize("fmt", FORMATS_SMALL)
def test_decoder_pictures_match_oracle(fmt):
    F = gpu()
    w, h = 97, 61
    for kw in OPTIONS:
        try:
            enc = cc.Encoder("oracle", w, h, fmt, **kw)
        except cc.CodecError:
            continue
        for kind in ("smooth", "noise", "testsrc2"):
            enc = cc.Encoder("oracle", w, h, fmt, **kw)
            ref = cc.Decoder("oracle", w, h, enc.extradata)
            dec = F.FFV1Decoder(w, h, enc.extradata)
            for fr in range(3):
                pkt = enc.encode(synth.GENERATORS[kind](fmt, w, h, fr))
                want = ref.decode(pkt)
                got = dec.decode(pkt, fmt_hint=ref.pix_fmt)
                assert dec.pix_fmt == ref.pix_fmt
                for a, b in zip(want, got):
                    assert np.array_equal(a, b), (fmt, kw, kind, fr)
            dec.close()


def test_fate_golden_packets():
    """the reference's own FATE vectors (tests/golden/make_fate_golden.py)"""
    F = gpu()
    full = json.load(open(os.path.join(HERE, "golden", "fate_full.json")))
    inputs = np.load(os.path.join(HERE, "golden", "fate_inputs.npz"))
    for test in sorted(inputs.files):
        g = full[test]
        w, h, fmt = g["width"], g["height"], g["pix_fmt"]
        kw = dict(g["options"])
        if "strict" in kw:
            kw["strict"] = kw.pop("strict")
        enc = F.FFV1Encoder(w, h, fmt, **kw)
        assert enc.extradata.hex() == g["extradata"]
        dec = F.FFV1Decoder(w, h, enc.extradata)
        for i, flat in enumerate(inputs[test]):
            planes, off = [], 0
            for bw, rows in cc.plane_geometry(fmt, w, h):
                planes.append(np.ascontiguousarray(flat[off:off + bw * rows].reshape(rows, bw)))
                off += bw * rows
            pkt = enc.encode(planes)
            assert md5(pkt) == g["packet_md5"][i], (test, i)
            out = dec.decode(pkt, fmt_hint="gbrp16le" if fmt == "rgb48le" else fmt)
            if fmt == "rgb48le":
                px = planes[0].view("<u2").reshape(h, w, 3)
                want = [px[:, :, 1], px[:, :, 2], px[:, :, 0]]
                got = [o.view("<u2") for o in out]
            elif fmt == "bgr0":
                want = [planes[0].reshape(h, w, 4)[:, :, :3]]
                got = [out[0].reshape(h, w, 4)[:, :, :3]]
            else:
                want, got = planes, out
            for a, b in zip(want, got):
                assert np.array_equal(a, b), (test, i)
        enc.close()
        dec.close()


def test_pipelined_matches_synchronous():
    """send/receive (launch groups over several streams) gives the same packets, in order"""
    F = gpu()
    w, h, fmt = 320, 240, "yuv420p10le"
    kw = dict(slices=30, gop_size=1)
    frames = [synth.testsrc2_like(fmt, w, h, i) for i in range(23)]
    ref = cc.Encoder("oracle", w, h, fmt, **kw)
    want = [ref.encode(f) for f in frames]
    enc = F.FFV1Encoder(w, h, fmt, max_batch=4, pipeline_depth=3, **kw)
    got, i = [], 0
    while True:
        if i < len(frames):
            if enc.send_frame(frames[i], pts=i):
                i += 1
                continue
        elif i == len(frames):
            enc.send_frame(None)
            i += 1
        r = enc.receive_packet()
        if r == F.EOF:
            break
        if r is not None:
            got.append(r)
    assert [g[2] for g in got] == list(range(len(frames)))
    assert [g[0] for g in got] == want
    # and decode them back, pipelined, straight into caller-provided pictures
    dec = F.FFV1Decoder(w, h, enc.extradata, max_batch=4, pipeline_depth=3)
    out, i = [], 0
    while True:
        if i < len(want):
            if dec.send_packet(want[i], pts=i, dst=dec.alloc_picture()):
                i += 1
                continue
        elif i == len(want):
            dec.send_packet(None)
            i += 1
        r = dec.receive_frame()
        if r == F.EOF:
            break
        if r is not None:
            out.append(r)
    assert len(out) == len(frames)
    for (po, arrs), src in zip(out, frames):
        assert po.damaged_slices == 0
        for a, b in zip(arrs, src):
            assert np.array_equal(a, b)


CONFIGS_FULL = [
    # BASELINE.json configs[0..4] (frames reduced; geometry and options as stated)
    ("C1", 1920, 1080, "yuv420p", dict(), 2),
    ("C2", 3840, 2160, "yuv420p10le", dict(slices=1023, gop_size=1), 2),
    ("C3", 3840, 2160, "bgr0", dict(coder=2, context=1, gop_size=1), 1),
    ("C4", 3840, 2160, "yuv444p16le", dict(gop_size=1), 1),
    ("C5", 7680, 4320, "yuv420p10le", dict(gop_size=1), 1),
]


def test_incompressible_pictures_outgrow_the_staging_buffers():
    """noise compresses to ~65 % of raw: the packets of a group exceed the encoder's pinned
    packet buffer (sized for raw/4 per picture -> the SM copy is skipped and the plain copy
    fallback runs) and the decoder's packet arena (raw/2 per picture -> a group is launched
    early); both must still give the oracle's packets and the input pictures"""
    F = gpu()
    w, h, fmt = 1280, 720, "yuv420p10le"
    kw = dict(slices=16, gop_size=1)
    frames = [synth.noise(fmt, w, h, i) for i in range(7)]
    ref = cc.Encoder("oracle", w, h, fmt, **kw)
    want = [ref.encode(f) for f in frames]
    raw = sum(a.nbytes for a in frames[0])
    assert min(len(p) for p in want) > raw // 2 + 65536
    enc = F.FFV1Encoder(w, h, fmt, max_batch=3, pipeline_depth=2, **kw)
    got, i = [], 0
    while True:
        if i < len(frames):
            if enc.send_frame(frames[i], pts=i):
                i += 1
                continue
        elif i == len(frames):
            enc.send_frame(None)
            i += 1
        r = enc.receive_packet()
        if r == F.EOF:
            break
        if r is not None:
            got.append(r)
    assert [g[0] for g in got] == want
    dec = F.FFV1Decoder(w, h, enc.extradata, max_batch=3, pipeline_depth=2)
    out, i = [], 0
    while True:
        if i < len(want):
            if dec.send_packet(want[i], pts=i, dst=dec.alloc_picture()):
                i += 1
                continue
        elif i == len(want):
            dec.send_packet(None)
            i += 1
        r = dec.receive_frame()
        if r == F.EOF:
            break
        if r is not None:
            out.append(r)
    assert [po.pts for po, _ in out] == list(range(len(frames)))
    for (po, arrs), src in zip(out, frames):
        assert po.damaged_slices == 0
        for a, b in zip(arrs, src):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("name,w,h,fmt,kw,nframes", CONFIGS_FULL)
def test_full_size_roundtrip_and_oracle(name, w, h, fmt, kw, nframes):
    """at BASELINE sizes: packet md5 == oracle (the oracle finishes a 4K frame in seconds),
    and decode(encode(x)) == x through the GPU decoder"""
    F = gpu()
    enc = F.FFV1Encoder(w, h, fmt, **kw)
    ref = cc.Encoder("oracle", w, h, fmt, threads=8, **kw)
    assert enc.extradata == ref.extradata
    dec = F.FFV1Decoder(w, h, enc.extradata)
    for i in range(nframes):
        src = synth.testsrc2_like(fmt, w, h, i)
        pkt = enc.encode(src)
        assert md5(pkt) == md5(ref.encode(src)), (name, i)
        out = dec.decode(pkt)
        if fmt == "bgr0":
            assert np.array_equal(out[0].reshape(h, w, 4)[:, :, :3], src[0].reshape(h, w, 4)[:, :, :3])
        else:
            for a, b in zip(out, src):
                assert np.array_equal(a, b), (name, i)


def _corrupt_cases(p1):
    """(name, packet): a payload byte of slice 1 (CRC mismatch -> concealment) and a byte of a
    size trailer (slice pointer chain shortened, ffv1dec.c:746-756)"""
    sizes, end = [], len(p1)
    while end > 0:
        size = int.from_bytes(p1[end - 8:end - 5], "big")
        sizes.append((end - 8 - size, size))
        end -= size + 8
    sizes = sizes[::-1]
    a = bytearray(p1)
    a[sizes[1][0] + sizes[1][1] // 2] ^= 0x55
    b = bytearray(p1)
    b[sizes[1][0] + sizes[1][1] + 1] ^= 0x55
    return [("payload", bytes(a)), ("size-trailer", bytes(b))]


@pytest.mark.parametrize("fmt,kw", [("yuv420p", dict(slices=4, gop_size=1)),
                                    ("yuv420p10le", dict(slices=4, gop_size=1))])
def test_damaged_packets_behave_like_the_reference(fmt, kw):
    F = gpu()
    w, h = 128, 96
    enc = cc.Encoder("oracle", w, h, fmt, **kw)
    f0, f1 = synth.smooth(fmt, w, h, 0), synth.smooth(fmt, w, h, 1)
    p0, p1 = enc.encode(f0), enc.encode(f1)
    for name, bad in _corrupt_cases(p1):
        ref = cc.Decoder("oracle", w, h, enc.extradata)
        dec = F.FFV1Decoder(w, h, enc.extradata)
        if name == "payload":
            # concealment needs a previous picture; in the other case the rectangles no slice
            # covers are unspecified in the reference (uninitialised buffer): start from zero
            ref.decode(p0)
            dec.decode(p0)
        want = ref.decode(bad)
        got = dec.decode(bad)
        if name == "payload":
            assert dec.last.damaged_slices >= 1
        for a, b in zip(want, got):
            assert np.array_equal(a, b), (fmt, name)
        dec.close()


def test_libavcodec_glue_is_a_drop_in():
    """integration/ffv1_gpu.c (AVCodec objects forwarding to libffgpu.so) and the UNMODIFIED
    reference codec objects, driven through the same AVCodec init/send/receive/decode/close
    boundary by the same harness: identical extradata, packets and pictures"""
    if not (cc.available("glue") and cc.available("ref")):
        pytest.skip("oracle/_ref (reference + glue harness) not built")
    w, h = 352, 288
    for fmt, kw in (("yuv420p", dict(slices=4)), ("yuv420p10le", dict(slices=30, gop_size=1)),
                    ("bgr0", dict(level=3, coder=2, context=1)), ("yuv444p16le", dict(level=3)),
                    ("yuv420p", dict())):
        ref = cc.Encoder("ref", w, h, fmt, **kw)
        glue = cc.Encoder("glue", w, h, fmt, **kw)
        assert glue.extradata == ref.extradata
        dref = cc.Decoder("ref", w, h, ref.extradata)
        dglue = cc.Decoder("glue", w, h, ref.extradata)
        for i in range(3):
            planes = synth.testsrc2_like(fmt, w, h, i)
            a, b = ref.encode(planes), glue.encode(planes)
            assert a == b, (fmt, kw, i)
            fa, fb = dref.decode(a), dglue.decode(a)
            assert dref.pix_fmt == dglue.pix_fmt
            for x, y in zip(fa, fb):
                assert np.array_equal(x, y), (fmt, kw, i)


def test_device_resident_batch():
    """pictures already in HBM -> packets in HBM (the path bench.py's `value` times)"""
    import torch
    F = gpu()
    w, h, fmt = 640, 360, "yuv420p10le"
    kw = dict(slices=60, gop_size=1)
    n = 5
    frame_bytes, planes = F.frame_layout(fmt, w, h)
    host = np.zeros((n, frame_bytes), np.uint8)
    srcs = []
    for i in range(n):
        src = synth.testsrc2_like(fmt, w, h, i)
        srcs.append(src)
        for (off, pitch, rows, rb), a in zip(planes, src):
            host[i, off:off + pitch * rows].reshape(rows, pitch)[:, :rb] = a
    dev = torch.from_numpy(host).cuda()
    enc = F.FFV1Encoder(w, h, fmt, max_batch=8, **kw)
    enc.encode_device(dev.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = cc.Encoder("oracle", w, h, fmt, **kw)
    pkts = [enc.device_fetch(i) for i in range(n)]
    assert pkts == [ref.encode(s) for s in srcs]
    # decode the batch back into device memory
    dec = F.FFV1Decoder(w, h, enc.extradata, max_batch=8)
    out = torch.zeros((n, frame_bytes), dtype=torch.uint8, device="cuda")
    dec.decode_device(pkts, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    back = out.cpu().numpy()
    for i in range(n):
        for (off, pitch, rows, rb), a in zip(planes, srcs[i]):
            assert np.array_equal(back[i, off:off + pitch * rows].reshape(rows, pitch)[:, :rb], a)


def _wide_slice_packet(w, h, fmt):
    """a 2-slice packet on a 6x4-slice stream: the real slice 0 plus a hand-coded slice whose
    header claims 4 of the 6 grid columns of row 1 (ffv1dec.c:176-184 accepts any rectangle
    inside the picture)"""
    import random
    import ffv1_bits as fb
    kw = dict(slices=24, coder=-2, level=3, gop_size=1)
    which = "ref" if cc.available("ref") else "oracle"
    enc = cc.Encoder(which, w, h, fmt, **kw)
    assert (enc.info["num_h_slices"], enc.info["num_v_slices"]) == (6, 4)
    p0 = enc.encode(synth.smooth(fmt, w, h, 0))
    p1 = enc.encode(synth.smooth(fmt, w, h, 1))
    sl = fb.split_v3_packet(p1)
    rc = fb.RangeEncoder(*fb.default_tables())
    st = [128] * 32
    for v in (0, 1, 3, 0, 0, 0, 3, 0, 1):      # sx sy sw-1 sh-1 | qidx x2 | ps | sar
        rc.put_symbol(st, v)
    rnd = random.Random(5)
    st2 = [[128] * 32 for _ in range(8)]
    for _ in range(6000):                      # arbitrary but well-formed range-coded payload
        rc.put_symbol(st2[rnd.randrange(8)], rnd.choice((0, 0, 0, 1, -1, 2, -3, 7, -20)), True)
    return enc.extradata, p0, p1[:sl[0][1] + 8] + fb.wrap_slice(rc.terminate(1)), which


def test_slice_header_wider_than_its_grid_cell():
    """with more than 16 slices the per-slice line scratch is one grid cell
    wide; a header that names a wider rectangle must decode like the reference (picture-wide
    scratch from the pool), not write past the scratch"""
    F = gpu()
    w, h, fmt = 192, 96, "yuv420p"
    xd, p0, bad, which = _wide_slice_packet(w, h, fmt)
    ref = cc.Decoder(which, w, h, xd)
    dec = F.FFV1Decoder(w, h, xd)
    ref.decode(p0)
    dec.decode(p0)
    # Only what the two slices of the packet cover is defined: the cell of slice 0 and the
    # wide rectangle (4 of 6 columns of grid row 1), which fails its end-of-slice check and is
    # concealed from the previous picture.  Everything else is whatever the output buffer
    # held (the reference harness and the decoder rotate their buffers differently).
    for _ in range(3):                         # a few times: the pool counter is per launch
        want = [a.copy() for a in ref.decode(bad)]
        got = dec.decode(bad)
        for p, (a, b) in enumerate(zip(want, got)):
            sh = 1 if p else 0
            assert np.array_equal(a[:24 >> sh, :32 >> sh], b[:24 >> sh, :32 >> sh])
            assert np.array_equal(a[24 >> sh:48 >> sh, :128 >> sh], b[24 >> sh:48 >> sh, :128 >> sh])
        assert dec.last.damaged_slices >= 1
    # 20 such packets in ONE launch group exhaust the 16-entry pool: still no overrun, the
    # slices that found no scratch are reported damaged
    dec2 = F.FFV1Decoder(w, h, xd, max_batch=24, pipeline_depth=1)
    n, i, got = 20, 0, 0
    while True:
        if i < n:
            if dec2.send_packet(bad, pts=i, dst=dec2.alloc_picture()):
                i += 1
                continue
        elif i == n:
            dec2.send_packet(None)
            i += 1
        r = dec2.receive_frame()
        if r == F.EOF:
            break
        if r is not None:
            got += 1
    assert got == n


def test_more_prefix_sets_than_cache_slots_in_one_group():
    """> 8 distinct (key, picture structure, SAR) combinations inside one
    launch group used to recycle prefix slots that queued pictures still referenced"""
    F = gpu()
    w, h, fmt = 160, 120, "yuv420p10le"
    kw = dict(slices=4, gop_size=1)
    frames = [synth.testsrc2_like(fmt, w, h, i) for i in range(24)]
    sars = [(i + 1, 1 + (i % 3)) for i in range(12)]
    which = "ref" if cc.available("ref") else "oracle"
    enc = F.FFV1Encoder(w, h, fmt, max_batch=32, pipeline_depth=2, **kw)
    got, i = [], 0
    while True:
        if i < len(frames):
            if enc.send_frame(frames[i], pts=i, sar=sars[i % len(sars)]):
                i += 1
                continue
        elif i == len(frames):
            enc.send_frame(None)
            i += 1
        r = enc.receive_packet()
        if r == F.EOF:
            break
        if r is not None:
            got.append(r[0])
    assert len(got) == len(frames)
    # every packet decodes to its picture and carries ITS sample aspect ratio
    dec = F.FFV1Decoder(w, h, enc.extradata)
    for i, p in enumerate(got):
        out = dec.decode(p)
        for a, b in zip(out, frames[i]):
            assert np.array_equal(a, b)
        assert (dec.last.sar_num, dec.last.sar_den) == sars[i % len(sars)], i


def _pump_encoder(F, enc, frames):
    got, i = [], 0
    while True:
        if i < len(frames):
            if enc.send_frame(frames[i], pts=i):
                i += 1
                continue
        elif i == len(frames):
            enc.send_frame(None)
            i += 1
        r = enc.receive_packet()
        if r == F.EOF:
            return got
        if r is not None:
            got.append(r)


def _pump_decoder(F, dec, pkts):
    out, i = [], 0
    while True:
        if i < len(pkts):
            if dec.send_packet(pkts[i], pts=i, dst=dec.alloc_picture()):
                i += 1
                continue
        elif i == len(pkts):
            dec.send_packet(None)
            i += 1
        r = dec.receive_frame()
        if r == F.EOF:
            return out
        if r is not None:
            out.append(r)


def _device_list(n):
    """n device ordinals: distinct GPUs when the box has them, else GPU 0 repeated (the
    routing and reordering logic is the same)"""
    import torch
    have = torch.cuda.device_count()
    return tuple(i % max(have, 1) for i in range(n))


@pytest.mark.parametrize("ndev,kw,nframes", [
    (2, dict(slices=16, gop_size=1), 23),              # picture i -> GPU i mod 2
    (3, dict(slices=9, gop_size=1, coder=1), 17),
    (2, dict(slices=4, gop_size=5), 23),               # carried states: whole GOPs round-robin
])
def test_stream_spread_over_several_gpus(ndev, kw, nframes):
    """SURVEY 8e inside the product: one handle, ndevices > 1, gives the single-GPU packet
    sequence in presentation order; the decoder restores the pictures in packet order"""
    F = gpu()
    w, h, fmt = 352, 288, "yuv420p10le"
    frames = [synth.GENERATORS[("testsrc2", "noise", "smooth")[i % 3]](fmt, w, h, i) for i in range(nframes)]
    ref = cc.Encoder("oracle", w, h, fmt, **kw)
    want = [ref.encode(f) for f in frames]
    devs = _device_list(ndev)
    # small groups, shallow pipeline: the in-order hand-back has to launch partly filled
    # groups of the GPU whose packet is due (the send side is blocked on another GPU)
    for batch, depth in ((4, 1), (3, 2), (0, 0)):
        enc = F.FFV1Encoder(w, h, fmt, devices=devs, max_batch=batch, pipeline_depth=depth, **kw)
        assert enc.extradata == ref.extradata
        got = _pump_encoder(F, enc, frames)
        assert [g[2] for g in got] == list(range(nframes))
        assert [g[0] for g in got] == want, (ndev, kw, batch, depth)
        # the handle accepts pictures again after the flush
        assert [g[0] for g in _pump_encoder(F, enc, frames[:ndev + 1])][0] == want[0] or kw["gop_size"] != 1
        enc.close()
        dec = F.FFV1Decoder(w, h, ref.extradata, devices=devs, max_batch=batch, pipeline_depth=depth)
        out = _pump_decoder(F, dec, want)
        assert [po.pts for po, _ in out] == list(range(nframes))
        for (po, arrs), src in zip(out, frames):
            assert po.damaged_slices == 0
            for a, b in zip(arrs, src):
                assert np.array_equal(a, b)
        dec.close()
    # synchronous calls on a routing handle
    enc = F.FFV1Encoder(w, h, fmt, devices=devs, **kw)
    dec = F.FFV1Decoder(w, h, ref.extradata, devices=devs)
    for i in range(7):
        pkt = enc.encode(frames[i])
        assert pkt == want[i]
        for a, b in zip(dec.decode(pkt), frames[i]):
            assert np.array_equal(a, b)
    assert enc.launches > 0 and dec.launches > 0


@pytest.mark.parametrize("fmt", ["bgr0", "bgra", "gbrp10le", "gbrap12le", "gbrp16le"])
def test_version4_rgb_matches_the_reference(fmt):
    """FFV1 version 4 on the GPU: k_rct_stat / k_rct_pick choose the per-slice RCT
    coefficients like choose_rct_params (ffv1enc.c:963-1043), the slice coders finish the
    longer slice header on the device; packets == the compiled reference, pictures restored"""
    if not cc.available("ref"):
        pytest.skip("oracle/_ref not built")
    F = gpu()
    w, h = 160, 96
    for kw in (dict(slices=4, coder=2), dict(slices=9, coder=-2, context=1), dict(coder=0),
               dict(slices=12, coder=1, gop_size=1)):
        kw = dict(kw, level=4, strict=-2)
        ref = cc.Encoder("ref", w, h, fmt, **kw)
        enc = F.FFV1Encoder(w, h, fmt, **kw)
        assert enc.info == ref.info and enc.extradata == ref.extradata
        dref = cc.Decoder("ref", w, h, ref.extradata)
        dec = F.FFV1Decoder(w, h, ref.extradata)
        for i, kind in enumerate(("smooth", "testsrc2", "extremes", "smooth", "testsrc2")):
            planes = synth.GENERATORS[kind](fmt, w, h, i)
            pkt = ref.encode(planes)
            assert enc.encode(planes) == pkt, (fmt, kw, kind)
            want = dref.decode(pkt)
            got = dec.decode(pkt, fmt_hint=dref.pix_fmt)
            for a, b in zip(want, got):
                assert np.array_equal(a, b), (fmt, kw, kind)
        enc.close()
        dec.close()


@pytest.mark.parametrize("fmt", ["yuv420p", "yuv444p10le", "gray", "yuva420p", "yuv420p16le"])
def test_version4_ycbcr_streams_decode(fmt):
    """version 4 YCbCr streams of the reference encoder decode to the reference's pictures"""
    if not cc.available("ref"):
        pytest.skip("oracle/_ref not built")
    F = gpu()
    w, h = 160, 96
    for kw in (dict(slices=4, coder=2), dict(coder=0), dict(slices=12, coder=1, gop_size=1)):
        ref = cc.Encoder("ref", w, h, fmt, level=4, strict=-2, **kw)
        dref = cc.Decoder("ref", w, h, ref.extradata)
        dec = F.FFV1Decoder(w, h, ref.extradata)
        for i, kind in enumerate(("smooth", "noise", "testsrc2", "smooth")):
            pkt = ref.encode(synth.GENERATORS[kind](fmt, w, h, i))
            want = dref.decode(pkt)
            got = dec.decode(pkt, fmt_hint=dref.pix_fmt)
            for a, b in zip(want, got):
                assert np.array_equal(a, b), (fmt, kw, kind)
        dec.close()


def test_pictures_resident_on_the_gpu_through_send_receive():
    """SURVEY 8f-1 at the C ABI: ffgpu_picture.data[] / ffgpu_picture_out.data[] may be CUDA
    device pointers (what AV_PIX_FMT_CUDA frames carry); the copy kind is taken from the
    pointer.  Also the context bracket the glue puts around every call for such frames."""
    import ctypes as C
    import torch
    F = gpu()
    w, h, fmt = 352, 288, "yuv420p10le"
    kw = dict(slices=30, gop_size=1)
    frames = [synth.testsrc2_like(fmt, w, h, i) for i in range(9)]
    ref = cc.Encoder("oracle", w, h, fmt, **kw)
    want = [ref.encode(f) for f in frames]
    # the bracket: the primary context the runtime (torch) uses, pushed and popped again
    torch.zeros(1, device="cuda")
    cu = C.CDLL("libcuda.so.1")
    ctx = C.c_void_p()
    assert cu.cuCtxGetCurrent(C.byref(ctx)) == 0 and ctx.value
    lib = F.lib()
    assert lib.ffgpu_cuda_push_context(ctx) == 0
    try:
        # device-resident planes with a pitch of their own
        dev = []
        for planes in frames:
            dev.append([torch.from_numpy(np.pad(p, ((0, 0), (0, 64)))).cuda() for p in planes])
        enc = F.FFV1Encoder(w, h, fmt, max_batch=4, pipeline_depth=2, **kw)
        got, i = [], 0
        while True:
            if i < len(dev):
                pic = F.codec.Picture()
                for k, t in enumerate(dev[i]):
                    pic.data[k] = t.data_ptr()
                    pic.linesize[k] = t.stride(0)
                pic.sar_num, pic.sar_den, pic.pts = 0, 1, i
                r = lib.ffgpu_ffv1_encode_send_frame(enc.h, C.byref(pic))
                assert r in (0, F.EAGAIN)
                if r == 0:
                    i += 1
                    continue
            elif i == len(dev):
                enc.send_frame(None)
                i += 1
            r = enc.receive_packet()
            if r == F.EOF:
                break
            if r is not None:
                got.append(r[0])
        assert got == want
        # decode into device-resident destination planes
        dec = F.FFV1Decoder(w, h, enc.extradata, max_batch=4, pipeline_depth=2)
        outs = [[torch.zeros_like(t) for t in dev[0]] for _ in want]
        sent = done = 0
        flushed = False
        while True:
            if sent < len(want):
                po = F.codec.PictureOut()
                for k, t in enumerate(outs[sent]):
                    po.data[k] = t.data_ptr()
                    po.linesize[k] = t.stride(0)
                if dec.send_packet(want[sent], pts=sent, dst=(po, outs[sent])):
                    sent += 1
                    continue
            elif not flushed:
                dec.send_packet(None)
                flushed = True
            r = dec.receive_frame()
            if r == F.EOF:
                break
            if r is not None:
                done += 1
        torch.cuda.synchronize()
        assert done == len(want)
        for o, src in zip(outs, frames):
            for t, p in zip(o, src):
                assert np.array_equal(t.cpu().numpy()[:, :p.shape[1]], p)
    finally:
        assert lib.ffgpu_cuda_pop_context() == 0


@pytest.mark.parametrize("fmt", ["yuv420p", "yuv420p10le", "bgr0"])
def test_two_pass_matches_the_reference(fmt):
    """SURVEY 8f-3 on the GPU: decision counters of the slice coders -> stats text == the
    reference's stats_out; second pass (initial states on the device, sorted transition
    table) -> extradata and packets == the reference's; decoder with initial states"""
    if not cc.available("ref"):
        pytest.skip("oracle/_ref not built")
    F = gpu()
    w, h = 160, 96
    kinds = ("smooth", "noise", "testsrc2", "smooth", "testsrc2", "noise", "smooth")
    for kw in (dict(slices=4, coder=2), dict(slices=12, coder=-2, context=1, gop_size=1), dict(coder=1, slices=9)):
        frames = [synth.GENERATORS[k](fmt, w, h, i) for i, k in enumerate(kinds)]
        r1 = cc.Encoder("ref", w, h, fmt, pass1=1, **kw)
        g1 = F.FFV1Encoder(w, h, fmt, pass1=1, **kw)
        for f in frames:
            assert g1.encode(f) == r1.encode(f)
        stats = r1.stats_out()
        assert g1.stats_out() == stats, (fmt, kw)
        # the pipelined path over two sub-handles adds its counters up
        g1b = F.FFV1Encoder(w, h, fmt, pass1=1, devices=_device_list(2), max_batch=3, **kw)
        _pump_encoder(F, g1b, frames)
        assert g1b.stats_out() == stats or kw.get("gop_size", 12) != 1
        r2 = cc.Encoder("ref", w, h, fmt, pass2=1, stats_in=stats, **kw)
        g2 = F.FFV1Encoder(w, h, fmt, pass2=1, stats_in=stats, **kw)
        assert g2.extradata == r2.extradata, (fmt, kw)
        dref = cc.Decoder("ref", w, h, r2.extradata)
        dec = F.FFV1Decoder(w, h, r2.extradata)
        for f in frames:
            pkt = r2.encode(f)
            assert g2.encode(f) == pkt, (fmt, kw)
            try:
                want = dref.decode(pkt)
            except cc.CodecError:
                break
            got = dec.decode(pkt, fmt_hint=dref.pix_fmt)
            for a, b in zip(want, got):
                assert np.array_equal(a, b), (fmt, kw)


def test_bottom_up_pictures():
    """AVFrames with a negative linesize (what `-vf vflip` hands an encoder): the reference
    walks them by pointer arithmetic (ffv1enc.c:283), the library stages them row by row --
    pageable and pinned host memory alike -- and the decoder writes into such a picture"""
    import torch
    from ffmpeg_ffv2_b200.codec import PictureOut
    F = gpu()
    w, h = 176, 98
    for fmt, kw in (("yuv420p10le", dict(slices=12, gop_size=1)), ("bgr0", dict(coder=1, gop_size=1)),
                    ("yuv420p", dict())):
        ref = cc.Encoder("oracle", w, h, fmt, **kw)
        enc = F.FFV1Encoder(w, h, fmt, **kw)
        dec = F.FFV1Decoder(w, h, enc.extradata)
        for i, pinned in enumerate((False, True, False)):
            src = synth.testsrc2_like(fmt, w, h, i)
            want = ref.encode(src)
            # the same picture stored bottom-up: row 0 is the LAST row in memory
            if pinned and torch.cuda.is_available():
                hold = [torch.from_numpy(np.ascontiguousarray(a[::-1])).pin_memory() for a in src]
                stored = [t.numpy() for t in hold]
            else:
                stored = [np.ascontiguousarray(a[::-1]) for a in src]
            views = [a[::-1] for a in stored]
            assert all(v.strides[0] < 0 for v in views)
            assert enc.encode(views) == want, (fmt, i)
            # decoder: destination planes given up front, bottom-up as well
            dst_store = [np.zeros_like(a) for a in src]
            out = PictureOut()
            for k, a in enumerate(dst_store):
                v = a[::-1]
                out.data[k] = v.ctypes.data
                out.linesize[k] = v.strides[0]
            assert dec.send_packet(want, pts=i, dst=(out, dst_store))
            assert dec.send_packet(None)
            got = dec.receive_frame()
            assert got not in (None, F.EOF)
            assert dec.receive_frame() == F.EOF
            for a, b in zip(dst_store, src):
                assert np.array_equal(a[::-1], b), (fmt, i)
    # the same through the reference's own ffmpeg: `-vf vflip` in front of the encoder
    ROOT = os.path.dirname(HERE)
    ffmpeg = os.path.join(ROOT, "oracle", "_ref", "ffmpeg")
    if os.path.exists(ffmpeg):
        import subprocess
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = os.path.dirname(F.lib_path()) + ":" + env.get("LD_LIBRARY_PATH", "")
        out = {}
        for codec in ("ffv1", "ffv1_gpu"):
            r = subprocess.run([ffmpeg, "-hide_banner", "-loglevel", "error", "-nostdin", "-f", "lavfi", "-i",
                                "testsrc2=s=640x360:r=25", "-frames:v", "9", "-pix_fmt", "yuv420p10le", "-vf", "vflip",
                                "-c:v", codec, "-slices", "12", "-g", "1", "-f", "framemd5", "-"],
                               capture_output=True, env=env, timeout=150)
            if b"No such filter" in r.stderr:
                pytest.skip("this build of oracle/_ref/ffmpeg has no vflip filter")
            assert r.returncode == 0, r.stderr.decode(errors="replace")[-2000:]
            out[codec] = [l for l in r.stdout.decode().splitlines() if l and not l.startswith("#")]
        assert len(out["ffv1"]) == 9 and out["ffv1_gpu"] == out["ffv1"]
