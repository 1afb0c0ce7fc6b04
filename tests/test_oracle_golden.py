"""The CPU restatement (oracle/ffv1_oracle.c) against the FATE-chained golden vectors.

tests/golden/fate_full.json and fate_inputs.npz were produced by
tests/golden/make_fate_golden.py: the reference's own ffmpeg reproduced the md5/size of
all 21 tests/ref/vsynth/vsynth{1,2,3}-ffv1* goldens, and the per-packet md5s below were
read from those AVI files.  Bit-exact is the bar (integer/byte work).
"""
import hashlib
import json
import os

import numpy as np
import pytest

import cpucodec as cc

HERE = os.path.dirname(os.path.abspath(__file__))
FULL = json.load(open(os.path.join(HERE, "golden", "fate_full.json")))
INPUTS = np.load(os.path.join(HERE, "golden", "fate_inputs.npz"))


def split(flat, fmt, w, h):
    planes, off = [], 0
    for bw, rows in cc.plane_geometry(fmt, w, h):
        planes.append(np.ascontiguousarray(flat[off:off + bw * rows].reshape(rows, bw)))
        off += bw * rows
    return planes


@pytest.mark.parametrize("which", ["oracle", "ref"])
@pytest.mark.parametrize("test", sorted(INPUTS.files))
def test_encoder_reproduces_fate_packets(test, which):
    if not cc.available(which):
        pytest.skip("%s library not built" % which)
    g = FULL[test]
    enc = cc.Encoder(which, g["width"], g["height"], g["pix_fmt"], **g["options"])
    assert enc.extradata.hex() == g["extradata"]
    frames = INPUTS[test]
    for i, flat in enumerate(frames):
        planes = split(flat, g["pix_fmt"], g["width"], g["height"])
        assert hashlib.md5(flat.tobytes()).hexdigest() == g["input_md5"][i]
        pkt = enc.encode(planes)
        assert hashlib.md5(pkt).hexdigest() == g["packet_md5"][i], (test, i)


@pytest.mark.parametrize("which", ["oracle", "ref"])
@pytest.mark.parametrize("test", sorted(INPUTS.files))
def test_decoder_restores_fate_inputs(test, which):
    """decode(packet) must give back the encoder's input (framemd5 equality)."""
    if not cc.available(which):
        pytest.skip("%s library not built" % which)
    g = FULL[test]
    w, h, fmt = g["width"], g["height"], g["pix_fmt"]
    enc = cc.Encoder("oracle", w, h, fmt, **g["options"])
    dec = cc.Decoder(which, w, h, bytes.fromhex(g["extradata"]))
    for i, flat in enumerate(INPUTS[test]):
        planes = split(flat, fmt, w, h)
        pkt = enc.encode(planes)
        out = dec.decode(pkt)
        if fmt == "rgb48le":          # decoded as planar gbrp16le (ffv1dec.c:722)
            assert dec.pix_fmt == "gbrp16le"
            px = planes[0].view("<u2").reshape(h, w, 3)
            want = [px[:, :, 1], px[:, :, 2], px[:, :, 0]]
            got = [o.view("<u2") for o in out]
        elif fmt == "bgr0":           # the X byte is not coded
            want = [planes[0].reshape(h, w, 4)[:, :, :3]]
            got = [out[0].reshape(h, w, 4)[:, :, :3]]
        else:
            want, got = planes, out
        for a, b in zip(want, got):
            assert np.array_equal(a, b), (test, i)


def test_crc_and_state_table_known_answers():
    lib = cc.api("oracle").lib
    import ctypes as C
    lib.ffv1o_crc32.restype = C.c_uint32
    # CRC over data followed by its own stored CRC is zero (the property ffv1 relies on)
    data = bytes(range(97))
    v = lib.ffv1o_crc32(0, data, len(data))
    both = data + int(v).to_bytes(4, "little")
    assert lib.ffv1o_crc32(0, both, len(both)) == 0
    # textbook CRC-32/MPEG-2 style register with init 0: "123456789" -> 0x89A1897F
    v = lib.ffv1o_crc32(0, b"123456789", 9)
    assert int(v).to_bytes(4, "little").hex() == "89a1897f"
    one = (C.c_uint8 * 256)()
    lib.ffv1o_default_state_transition(one)
    t = list(one)
    assert t[0] == 0 and t[128] > 128 and t[247] == 248 and t[248] == 248
    assert all(t[i] > i for i in range(8, 248))
