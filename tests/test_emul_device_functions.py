"""The product's __host__ __device__ slice functions (csrc/ffv1_slice.cuh: symbolize, range /
Golomb coders, packet assembly, slice header parser, decoder) compiled for the CPU by
tests/emul and checked bit-exactly against the oracle.  This is how the device logic is
verified in the GPU-less build container; the CUDA kernels wrap the same functions and are
checked again on the B200 by test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

import cpucodec as cc
import synth

pytestmark = pytest.mark.skipif(not cc.available("emul"), reason="tests/emul not built")

# the forms a launch can take: warp-wide loops, the straight-line coders of
# one-slice-per-warp launches, the encoder's stage B in two halves
FORMS = [None, "FFV1_EMUL_LONE", "FFV1_EMUL_SPLIT"]


def use_form(monkeypatch, form):
    for f in FORMS[1:]:
        monkeypatch.delenv(f, raising=False)
    if form:
        monkeypatch.setenv(form, "1")


FORMATS = ["yuv420p", "yuv410p", "gray", "ya8", "yuva420p", "yuv420p10le", "yuv444p16le", "gray16le",
           "yuva444p10le", "bgr0", "bgra", "gbrp10le", "gbrp16le", "gbrap12le", "rgb48le"]
OPTIONS = [dict(), dict(slices=4), dict(slices=9, coder=2), dict(slices=4, coder=-2, context=1),
           dict(level=3, coder=0, context=1), dict(level=1, coder=1),
           dict(level=3, slicecrc=0, gop_size=1, slices=12)]


@pytest.mark.parametrize("fmt", FORMATS)
def test_device_functions_match_oracle(fmt):
    w, h = 64, 48
    for kw in OPTIONS:
        try:
            orc = cc.Encoder("oracle", w, h, fmt, **kw)
        except cc.CodecError as e:
            with pytest.raises(cc.CodecError) as ei:
                cc.Encoder("emul", w, h, fmt, **kw)
            assert ei.value.code == e.code
            continue
        emu = cc.Encoder("emul", w, h, fmt, **kw)
        assert emu.info == orc.info and emu.extradata == orc.extradata
        do = cc.Decoder("oracle", w, h, orc.extradata)
        de = cc.Decoder("emul", w, h, orc.extradata)
        for i, kind in enumerate(("smooth", "noise", "extremes", "testsrc2", "smooth")):
            planes = synth.GENERATORS[kind](fmt, w, h, i)
            po, pe = orc.encode(planes), emu.encode(planes)
            assert po == pe, (fmt, kw, kind)
            fo, fe = do.decode(po), de.decode(po)
            assert do.pix_fmt == de.pix_fmt
            for a, b in zip(fo, fe):
                assert np.array_equal(a, b), (fmt, kw, kind)


@pytest.mark.parametrize("form", ["FFV1_EMUL_SPLIT", "FFV1_EMUL_LONE"])
def test_the_other_forms_of_the_slice_coders(form, monkeypatch):
    """stage B in two halves (ff_chain_token + ff_encode_slice_records) and the straight-line
    coders of one-slice-per-warp launches (ff_encode_slice_range_lone,
    ff_decode_slice_range_planar_lone) give the same bytes and pictures as the oracle"""
    monkeypatch.setenv(form, "1")
    w, h = 80, 56
    for fmt, kw in [("yuv420p", dict(slices=4, coder=1)), ("yuv420p10le", dict(slices=4, gop_size=1)),
                    ("yuv444p16le", dict(coder=2, context=1)), ("gray", dict(coder=1, gop_size=1, context=1)),
                    ("ya8", dict(coder=1, slices=4)), ("bgr0", dict(coder=1, slices=4)),
                    ("gbrp16le", dict(slices=4, gop_size=1)), ("bgra", dict(level=4, strict=-2, coder=1, slices=4))]:
        checker = "ref" if kw.get("level") == 4 else "oracle"    # the port does not restate version 4
        if not cc.available(checker):
            continue
        orc = cc.Encoder(checker, w, h, fmt, **kw)
        emu = cc.Encoder("emul", w, h, fmt, **kw)
        do = cc.Decoder(checker, w, h, orc.extradata)
        de = cc.Decoder("emul", w, h, orc.extradata)
        for i, kind in enumerate(("testsrc2", "noise", "extremes", "smooth")):
            planes = synth.GENERATORS[kind](fmt, w, h, i)
            po = orc.encode(planes)
            assert po == emu.encode(planes), (form, fmt, kw, kind)
            for a, b in zip(do.decode(po), de.decode(po)):
                assert np.array_equal(a, b), (form, fmt, kw, kind)


@pytest.mark.parametrize("form", ["FFV1_EMUL_SPLIT", "FFV1_EMUL_LONE"])
def test_the_other_forms_over_the_whole_format_matrix(form, monkeypatch):
    """the GPU tests' small pictures all land in one-slice-per-warp launches, i.e. in the
    straight-line coders: every format x option set of the main parity test through them (and
    through the two-halves stage B), at a size whose slices are ragged"""
    monkeypatch.setenv(form, "1")
    w, h = 131, 67
    extra = [dict(slices=4, gop_size=1), dict(coder=1, gop_size=1, slices=9), dict(coder=2, context=1, slices=4)]
    for fmt in FORMATS:
        for kw in OPTIONS + extra:
            try:
                orc = cc.Encoder("oracle", w, h, fmt, **kw)
            except cc.CodecError:
                continue
            emu = cc.Encoder("emul", w, h, fmt, **kw)
            do = cc.Decoder("oracle", w, h, orc.extradata)
            de = cc.Decoder("emul", w, h, orc.extradata)
            for i, kind in enumerate(("testsrc2", "noise", "extremes")):
                planes = synth.GENERATORS[kind](fmt, w, h, i)
                po = orc.encode(planes)
                assert po == emu.encode(planes), (form, fmt, kw, kind)
                for a, b in zip(do.decode(po), de.decode(po)):
                    assert np.array_equal(a, b), (form, fmt, kw, kind)


@pytest.mark.parametrize("form", [None, "FFV1_EMUL_LONE"])
def test_pictures_narrower_than_the_look_ahead(form, monkeypatch):
    """the decoders read the previous line several samples ahead of their use, clamped to the
    line's last sample: lines of 1..5 samples (and their subsampled chroma) exercise every clamp"""
    if form:
        monkeypatch.setenv(form, "1")
    for (w, h) in [(2, 2), (3, 5), (4, 3), (5, 1), (7, 2), (9, 4)]:
        for fmt, kw in [("yuv420p", dict(coder=1)), ("yuv444p10le", dict()), ("gray16le", dict(context=1)),
                        ("yuv410p", dict(coder=1, context=1)), ("ya8", dict(coder=1)), ("bgr0", dict(coder=1, context=1)),
                        ("gbrp16le", dict())]:
            orc = cc.Encoder("oracle", w, h, fmt, **kw)
            emu = cc.Encoder("emul", w, h, fmt, **kw)
            do = cc.Decoder("oracle", w, h, orc.extradata)
            de = cc.Decoder("emul", w, h, orc.extradata)
            for i, kind in enumerate(("noise", "testsrc2", "extremes")):
                planes = synth.GENERATORS[kind](fmt, w, h, i)
                po = orc.encode(planes)
                assert po == emu.encode(planes), (form, w, h, fmt, kind)
                for a, b in zip(do.decode(po), de.decode(po)):
                    assert np.array_equal(a, b), (form, w, h, fmt, kind)


def test_msb_aligned_sample_depth():
    """bits_per_raw_sample below the container depth in a 16-bit format (ffv1dec.c:158)"""
    w, h, fmt = 48, 32, "gray16le"
    orc = cc.Encoder("oracle", w, h, fmt, bits_per_raw_sample=11)
    emu = cc.Encoder("emul", w, h, fmt, bits_per_raw_sample=11)
    planes = synth.noise(fmt, w, h, 0)
    po = orc.encode(planes)
    assert po == emu.encode(planes)
    a = cc.Decoder("oracle", w, h).decode(po)
    b = cc.Decoder("emul", w, h).decode(po)
    assert np.array_equal(a[0], b[0])


@pytest.mark.parametrize("form", FORMS[:2])
def test_damaged_packets(form, monkeypatch):
    use_form(monkeypatch, form)
    w, h = 128, 96
    for fmt in ("yuv420p", "yuv420p10le"):
        enc = cc.Encoder("oracle", w, h, fmt, slices=4, gop_size=1)
        p0 = enc.encode(synth.smooth(fmt, w, h, 0))
        p1 = enc.encode(synth.smooth(fmt, w, h, 1))
        sizes, end = [], len(p1)
        while end > 0:
            size = int.from_bytes(p1[end - 8:end - 5], "big")
            sizes.append((end - 8 - size, size))
            end -= size + 8
        sizes = sizes[::-1]
        for off, prev in ((sizes[1][0] + sizes[1][1] // 2, True), (sizes[1][0] + sizes[1][1] + 1, False),
                          (sizes[2][0] + 3, True)):
            bad = bytearray(p1)
            bad[off] ^= 0x55
            out = {}
            for which in ("oracle", "emul"):
                d = cc.Decoder(which, w, h, enc.extradata)
                if prev:
                    d.decode(p0)
                fn = getattr(cc.api(which).lib, cc.PATHS[which][1] + "decoder_damaged")
                frames = d.decode(bytes(bad))
                out[which] = (frames, fn(C.c_void_p(d.h)))
            assert out["oracle"][1] == out["emul"][1]
            for a, b in zip(out["oracle"][0], out["emul"][0]):
                assert np.array_equal(a, b)


@pytest.mark.parametrize("form", FORMS[:2])
def test_wide_slice_header_decodes_like_the_reference(form, monkeypatch):
    """a slice header naming a rectangle wider than its grid cell (hand-coded with
    tests/ffv1_bits.py): product device functions == reference == oracle"""
    import numpy as np
    import ffv1_bits as fb
    import random
    if not cc.available("ref"):
        pytest.skip("oracle/_ref not built")
    use_form(monkeypatch, form)
    w, h, fmt = 192, 96, "yuv420p"
    enc = cc.Encoder("ref", w, h, fmt, slices=24, coder=-2, level=3, gop_size=1)
    p0 = enc.encode(synth.smooth(fmt, w, h, 0))
    p1 = enc.encode(synth.smooth(fmt, w, h, 1))
    sl = fb.split_v3_packet(p1)
    assert len(sl) == 24 and fb.wrap_slice(p1[sl[3][0]:sl[3][0] + sl[3][1]]) == \
        p1[sl[3][0]:sl[3][0] + sl[3][1] + 8]          # the helper's trailer/CRC == the reference's
    rc = fb.RangeEncoder(*fb.default_tables())
    st = [128] * 32
    for v in (0, 1, 3, 0, 0, 0, 3, 0, 1):
        rc.put_symbol(st, v)
    rnd = random.Random(5)
    st2 = [[128] * 32 for _ in range(8)]
    for _ in range(6000):
        rc.put_symbol(st2[rnd.randrange(8)], rnd.choice((0, 0, 0, 1, -1, 2, -3, 7, -20)), True)
    bad = p1[:sl[0][1] + 8] + fb.wrap_slice(rc.terminate(1))
    outs = {}
    for which in ("ref", "oracle", "emul"):
        d = cc.Decoder(which, w, h, enc.extradata)
        d.decode(p0)
        outs[which] = [a.copy() for a in d.decode(bad)]
    for which in ("oracle", "emul"):
        for a, b in zip(outs["ref"], outs[which]):
            assert np.array_equal(a, b), which


V4_RGB = ["bgr0", "bgra", "gbrp9le", "gbrp10le", "gbrap12le", "gbrp14le", "gbrp16le", "gbrap16le"]
V4_OPTIONS = [dict(slices=4, coder=2), dict(slices=9, coder=-2, context=1), dict(coder=0),
              dict(slices=4, coder=1, gop_size=1)]


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("fmt", V4_RGB)
def test_version4_rgb_matches_the_reference(fmt, form, monkeypatch):
    """FFV1 version 4: per-slice RCT coefficients (choose_rct_params), the longer
    slice header, Golomb-Rice slices closed on the device.  The oracle port does not restate
    version 4; the checker is the compiled reference (oracle/_ref)."""
    if not cc.available("ref"):
        pytest.skip("oracle/_ref not built")
    use_form(monkeypatch, form)
    w, h = 96, 64
    for kw in V4_OPTIONS:
        kw = dict(kw, level=4, strict=-2)
        ref = cc.Encoder("ref", w, h, fmt, **kw)
        emu = cc.Encoder("emul", w, h, fmt, **kw)
        assert emu.info == ref.info and emu.extradata == ref.extradata
        dr = cc.Decoder("ref", w, h, ref.extradata)
        de = cc.Decoder("emul", w, h, ref.extradata)
        for i, kind in enumerate(("smooth", "noise", "testsrc2", "extremes", "noise")):
            planes = synth.GENERATORS[kind](fmt, w, h, i)
            pr = ref.encode(planes)
            try:
                pe = emu.encode(planes)
            except cc.CodecError as e:
                # incompressible 16-bit pictures outgrow version 4's small slice buffers: the
                # reference then writes raw "PCM" slices, the product says "encoded frame too
                # large" (it never writes PCM slices, see ff_encode_slice_range)
                assert e.code == -1094995529 and kind == "noise" and "16" in fmt, (fmt, kw, kind)
                emu = cc.Encoder("emul", w, h, fmt, **kw)
                ref = cc.Encoder("ref", w, h, fmt, **kw)
                pe = None
            assert pe is None or pr == pe, (fmt, kw, kind)
            for a, b in zip(dr.decode(pr), de.decode(pr)):     # PCM slices included
                assert np.array_equal(a, b), (fmt, kw, kind)


@pytest.mark.parametrize("fmt", ["yuv420p", "yuv444p10le", "yuv422p10le", "gray", "ya8", "yuva420p",
                                 "yuv420p16le", "yuv444p16le"])
@pytest.mark.parametrize("form", FORMS[:2])
def test_version4_ycbcr_streams_decode(fmt, form, monkeypatch):
    """the reference also writes version 4 YCbCr streams (their RCT coefficients come from
    reads outside the planes, so only the decoder can be compared): slice_reset_contexts,
    the plane-context numbering of gray+alpha, coefficients that are parsed and ignored"""
    if not cc.available("ref"):
        pytest.skip("oracle/_ref not built")
    use_form(monkeypatch, form)
    w, h = 96, 64
    for kw in (dict(slices=4, coder=2), dict(coder=0), dict(slices=9, coder=1, gop_size=1)):
        ref = cc.Encoder("ref", w, h, fmt, level=4, strict=-2, **kw)
        dr = cc.Decoder("ref", w, h, ref.extradata)
        de = cc.Decoder("emul", w, h, ref.extradata)
        for i, kind in enumerate(("smooth", "noise", "testsrc2")):
            pkt = ref.encode(synth.GENERATORS[kind](fmt, w, h, i))
            for a, b in zip(dr.decode(pkt), de.decode(pkt)):
                assert np.array_equal(a, b), (fmt, kw, kind)


@pytest.mark.parametrize("form", FORMS[:2])
def test_version4_pcm_slice_decodes_like_the_reference(form, monkeypatch):
    """slice_coding_mode == 1 (raw bits, ffv1dec_template.c:37-47): the reference encoder only
    writes such slices when a packet buffer overflows, so one is hand-coded here"""
    import ffv1_bits as fb
    import random
    if not cc.available("ref"):
        pytest.skip("oracle/_ref not built")
    use_form(monkeypatch, form)
    w, h = 64, 32
    for fmt, bits, nsym in (("bgr0", 8, 3), ("gbrp10le", 10, 3), ("yuv444p10le", 10, 3)):
        enc = cc.Encoder("ref", w, h, fmt, level=4, strict=-2, slices=4, coder=-2, gop_size=1)
        p1 = enc.encode(synth.smooth(fmt, w, h, 1))
        sl = fb.split_v3_packet(p1)
        rc = fb.RangeEncoder(*fb.default_tables())
        rc.put([128], 0, 1)                          # key frame bit (slice 0 carries it)
        st = [128] * 32
        qn = 2 if enc.info["colorspace"] == 0 else 2
        for v in [0, 0, 0, 0] + [0] * qn + [3, 0, 1]:   # sx sy sw-1 sh-1 | qidx | ps | sar
            rc.put_symbol(st, v)
        rc.put(st, 0, 0)                             # slice_reset_contexts
        rc.put_symbol(st, 1)                         # slice_coding_mode = 1
        rnd = random.Random(7)
        sw, sh = w // 2, h // 2
        for _ in range(sw * sh * nsym):
            v = rnd.randrange(1 << bits)
            for i in range(bits - 1, -1, -1):
                rc.put([128], 0, (v >> i) & 1)
        first = fb.wrap_slice(rc.terminate(1))
        rest = p1[sl[1][0]:]
        pkt = first + rest
        outs = {}
        for which in ("ref", "emul"):
            d = cc.Decoder(which, w, h, enc.extradata)
            outs[which] = [a.copy() for a in d.decode(pkt)]
        for a, b in zip(outs["ref"], outs["emul"]):
            assert np.array_equal(a, b), fmt


@pytest.mark.parametrize("fmt", ["yuv420p", "yuv420p10le", "bgr0", "gbrp16le", "gray"])
def test_two_pass_matches_the_reference(fmt):
    """the first pass's statistics text (AVCodecContext.stats_out), the second pass's
    sorted transition table and initial states (extradata), its packets, and the decoder with
    initial states, all against the compiled reference.  (For more than 8 bits the reference's
    second-pass streams do not decode to the source, and with the large context model not at
    all -- by its own decoder; the product reproduces exactly that.)"""
    if not cc.available("ref"):
        pytest.skip("oracle/_ref not built")
    w, h = 96, 64
    kinds = ("smooth", "noise", "testsrc2", "smooth", "testsrc2", "noise", "smooth")
    for kw in (dict(slices=4, coder=2), dict(slices=4, coder=-2, context=1),
               dict(coder=1, gop_size=1, slices=9), dict(coder=0)):
        frames = [synth.GENERATORS[k](fmt, w, h, i) for i, k in enumerate(kinds)]
        r1 = cc.Encoder("ref", w, h, fmt, pass1=1, **kw)
        e1 = cc.Encoder("emul", w, h, fmt, pass1=1, **kw)
        assert r1.extradata == e1.extradata and r1.info == e1.info
        for f in frames:
            assert r1.encode(f) == e1.encode(f)
        stats = r1.stats_out()
        assert e1.stats_out() == stats, (fmt, kw)
        r2 = cc.Encoder("ref", w, h, fmt, pass2=1, stats_in=stats, **kw)
        e2 = cc.Encoder("emul", w, h, fmt, pass2=1, stats_in=stats, **kw)
        assert r2.extradata == e2.extradata, (fmt, kw)
        dr = cc.Decoder("ref", w, h, r2.extradata)
        de = cc.Decoder("emul", w, h, r2.extradata)
        for f in frames:
            pkt = r2.encode(f)
            assert e2.encode(f) == pkt, (fmt, kw)
            try:
                want = dr.decode(pkt)
            except cc.CodecError:
                break
            for a, b in zip(want, de.decode(pkt)):
                assert np.array_equal(a, b), (fmt, kw)


@pytest.mark.parametrize("form", FORMS[1:])
@pytest.mark.parametrize("fmt,kw", [("yuv420p", dict(slices=4, coder=2)), ("bgr0", dict(slices=4, coder=-2, gop_size=1)),
                                    ("gray", dict(coder=1, gop_size=1, slices=9))])
def test_second_pass_through_the_other_forms(fmt, kw, form, monkeypatch):
    """a second pass (sorted transition table, initial states from the extradata instead of
    128) through the straight-line coders and the two-halves stage B: packets and pictures of
    the compiled reference"""
    if not cc.available("ref"):
        pytest.skip("oracle/_ref not built")
    w, h = 80, 48
    frames = [synth.GENERATORS[k](fmt, w, h, i) for i, k in enumerate(("smooth", "testsrc2", "noise", "smooth"))]
    r1 = cc.Encoder("ref", w, h, fmt, pass1=1, **kw)
    for f in frames:
        r1.encode(f)
    stats = r1.stats_out()
    use_form(monkeypatch, form)
    r2 = cc.Encoder("ref", w, h, fmt, pass2=1, stats_in=stats, **kw)
    e2 = cc.Encoder("emul", w, h, fmt, pass2=1, stats_in=stats, **kw)
    assert r2.extradata == e2.extradata
    dr = cc.Decoder("ref", w, h, r2.extradata)
    de = cc.Decoder("emul", w, h, r2.extradata)
    for f in frames:
        pkt = r2.encode(f)
        assert e2.encode(f) == pkt, (fmt, kw, form)
        for a, b in zip(dr.decode(pkt), de.decode(pkt)):
            assert np.array_equal(a, b), (fmt, kw, form)


def test_damaged_input_under_the_sanitizers():
    """tests/emul/fuzz: the parsers and both forms of the slice decoders over mutated packets
    and mutated extradata, built with -fsanitize=address,undefined -- any out-of-bounds access
    on hostile input (a device fault on the GPU) fails here"""
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
    b = subprocess.run(["make", "-C", here, "fuzz"], capture_output=True, text=True)
    if b.returncode != 0:
        pytest.skip("sanitizer build not available: " + b.stderr[-200:])
    for seed in (1, 7):
        r = subprocess.run([os.path.join(here, "fuzz"), "40", str(seed)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "fuzz ok" in r.stdout, (seed, r.stdout[-500:], r.stderr[-3000:])
