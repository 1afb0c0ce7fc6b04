"""TEST INFRASTRUCTURE: random call sequences on the PUBLIC send/receive API of the host-pipeline
build (tests/emul/cpu*/libffgpu.so, selected by FFGPU_LIB): pictures and packets sent, results
received and streams flushed in random order, on one device and behind routing handles, with
every group size and depth.  Whatever the order of the calls, the packets / pictures must come
back complete, in order and identical to the oracle's, EAGAIN must never be answered on both
sides at once, and a flushed handle must accept input again after its EOF.
Usage: api_sequences.py <seed> <trials>"""
import os, sys, random
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..')); sys.path.insert(0, os.path.join(HERE, '..'))
import numpy as np
import ffmpeg_ffv2_b200 as F, cpucodec as cc, synth

seed = int(sys.argv[1]); trials = int(sys.argv[2])
rnd = random.Random(seed)
w, h = 64, 48
STREAMS = [("yuv420p10le", dict(slices=4, gop_size=1)), ("yuv420p", dict(slices=4, gop_size=3)),
           ("bgr0", dict(coder=1, gop_size=1)), ("gray", dict(coder=1, gop_size=5, slices=4)),
           ("yuv444p16le", dict(slices=9, gop_size=1, coder=2, context=1))]
stats = dict(enc=0, dec=0, flushes=0, eagain=0)

for trial in range(trials):
    fmt, kw = rnd.choice(STREAMS)
    n = rnd.randrange(0, 18)
    frames = [synth.GENERATORS[rnd.choice(("smooth", "noise", "testsrc2"))](fmt, w, h, i) for i in range(n)]
    ref = cc.Encoder("oracle", w, h, fmt, **kw)
    want = [ref.encode(f) for f in frames]
    shape = dict(max_batch=rnd.choice((0, 1, 2, 3, 5)), pipeline_depth=rnd.choice((0, 1, 2, 3)),
                 devices=rnd.choice(((), (), (0, 1), (0, 1, 0))))
    # a mid-stream flush restarts nothing in the bitstream only for intra-only streams
    may_flush = kw["gop_size"] == 1
    # ---------------- encoder ----------------
    enc = F.FFV1Encoder(w, h, fmt, **kw, **shape)
    got, sent, flushing, stalled = [], 0, False, 0
    while True:
        act = rnd.random()
        if not flushing and sent < n and act < 0.55:
            if enc.send_frame(frames[sent], pts=sent):
                sent += 1; stalled = 0
                continue
            stats["eagain"] += 1                     # full: the next receive must deliver
            r = enc.receive_packet()
            assert r is not None and r != F.EOF, ("EAGAIN on both sides", trial, fmt, shape)
            got.append(r)
            continue
        if not flushing and (sent == n or (may_flush and act > 0.93)):
            enc.send_frame(None); flushing = True; stats["flushes"] += 1
            continue
        r = enc.receive_packet()
        if r == F.EOF:
            assert flushing, "EOF without a flush"
            flushing = False
            assert len(got) == sent, ("packets missing at EOF", len(got), sent, trial, fmt, shape)
            if sent == n:
                break
        elif r is not None:
            got.append(r); stalled = 0
        else:
            assert not flushing, "EAGAIN while flushing"
            stalled += 1
            assert stalled < 10000, "no progress"
    assert [g[0] for g in got] == want, ("packets differ", trial, fmt, kw, shape)
    assert [g[2] for g in got] == list(range(n)), ("pts order", trial)
    enc.close()
    stats["enc"] += n
    # ---------------- decoder ----------------
    dec = F.FFV1Decoder(w, h, ref.extradata, max_batch=shape["max_batch"], pipeline_depth=shape["pipeline_depth"],
                        devices=shape["devices"])
    outs, sent, flushing, stalled, dsts = [], 0, False, 0, []
    up_front = rnd.random() < 0.7                    # destination planes named at send time
    while True:
        act = rnd.random()
        if not flushing and sent < n and act < 0.55:
            dst = dec.alloc_picture(fmt) if up_front else None
            if dec.send_packet(want[sent], pts=sent, dst=dst):
                sent += 1; stalled = 0
                continue
            stats["eagain"] += 1
            r = dec.receive_frame()
            assert r is not None and r != F.EOF, ("EAGAIN on both sides", trial, fmt, shape)
            outs.append(([a.copy() for a in r[1]], r[0].pts))
            continue
        if not flushing and (sent == n or (may_flush and act > 0.93)):
            dec.send_packet(None); flushing = True; stats["flushes"] += 1
            continue
        r = dec.receive_frame()
        if r == F.EOF:
            assert flushing, "EOF without a flush"
            flushing = False
            assert len(outs) == sent, ("pictures missing at EOF", len(outs), sent, trial, fmt, shape)
            if sent == n:
                break
        elif r is not None:
            outs.append(([a.copy() for a in r[1]], r[0].pts)); stalled = 0
        else:
            assert not flushing, "EAGAIN while flushing"
            stalled += 1
            assert stalled < 10000, "no progress"
    for i, (planes, pts) in enumerate(outs):
        assert pts == i, ("picture order", trial, i, pts)
        assert all(np.array_equal(a, b) for a, b in zip(planes, frames[i])), ("picture differs", trial, i, fmt, shape)
    dec.close()
    stats["dec"] += n
print("api sequences ok", seed, stats)
