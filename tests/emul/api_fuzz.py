"""TEST INFRASTRUCTURE: damaged packets and extradata through the PUBLIC decoder API of the
host-pipeline build (tests/emul/cpu*/libffgpu.so, selected by FFGPU_LIB): synchronous, pipelined
and routed over two stand-in devices.  Usage: api_fuzz.py <seed> <iterations per stream>"""
import os, sys, random
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..')); sys.path.insert(0, os.path.join(HERE, '..'))
import numpy as np
import ffmpeg_ffv2_b200 as F, cpucodec as cc, synth
seed = int(sys.argv[1]); iters = int(sys.argv[2])
rnd = random.Random(seed)
CASES = [("yuv420p", dict(slices=4)), ("yuv420p10le", dict(slices=12, gop_size=1)), ("bgr0", dict(slices=4, coder=1, context=1, gop_size=1)),
         ("gray", dict(level=1, coder=1)), ("yuv410p", dict(level=0)), ("bgra", dict(level=4, strict=-2, coder=1, gop_size=1)),
         ("yuv444p16le", dict(coder=2, gop_size=1, slices=24)), ("gbrp16le", dict(slices=4)), ("ya8", dict(slices=6, gop_size=1))]
def mutate(p):
    p = bytearray(p); n = len(p); k = rnd.randrange(8)
    if k == 0:
        for _ in range(1 + rnd.randrange(4)): p[rnd.randrange(n)] ^= 1 << rnd.randrange(8)
    elif k == 1:
        for _ in range(1 + rnd.randrange(8)): p[rnd.randrange(n)] = rnd.randrange(256)
    elif k == 2: p = p[:rnd.randrange(n)]
    elif k == 3:
        for _ in range(1 + rnd.randrange(3)): p[n - 1 - rnd.randrange(min(n, 64))] = rnd.randrange(256)
    elif k == 4:
        for _ in range(1 + rnd.randrange(3)): p[rnd.randrange(min(n, 24))] = rnd.randrange(256)
    elif k == 5:
        a = rnd.randrange(n); l = rnd.randrange(n - a) % 64; p[a:a + l] = bytes([rnd.choice((0, 255))]) * l
    elif k == 6: p += bytes(rnd.randrange(256) for _ in range(rnd.randrange(32)))
    else:
        a = rnd.randrange(n); l = rnd.randrange(n - a); p[a:a + l] = bytes(rnd.randrange(256) for _ in range(l))
    return bytes(p)
w, h = 96, 64
stats = dict(ok=0, err=0, open_err=0)
for fmt, kw in CASES:
    ref = cc.Encoder("oracle" if kw.get("level") != 4 else "ref", w, h, fmt, **kw)
    pk = [ref.encode(synth.GENERATORS[k](fmt, w, h, i)) for i, k in enumerate(("smooth", "noise", "testsrc2"))]
    for it in range(iters):
        ex = ref.extradata
        if ex and it % 5 == 4: ex = mutate(ex)
        try:
            dec = F.FFV1Decoder(w, h, ex, max_batch=rnd.choice((1, 2, 3)), pipeline_depth=rnd.choice((1, 2)),
                                devices=[0, 1] if it % 4 == 3 else ())
        except F.FFGpuError:
            stats["open_err"] += 1; continue
        pipelined = it % 2
        try:
            for f, p in enumerate(pk):
                m = p if (f == 0 and it % 3 == 0) else mutate(p)
                if not m: m = b"\0"
                try:
                    if pipelined:
                        while not dec.send_packet(m, pts=f, dst=dec.alloc_picture() if dec.pix_fmt else None):
                            dec.receive_frame()
                    else:
                        dec.decode(m, fmt_hint=fmt)
                    stats["ok"] += 1
                except F.FFGpuError:
                    stats["err"] += 1
            if pipelined:
                dec.send_packet(None)
                for _ in range(10):
                    try:
                        r = dec.receive_frame()
                    except F.FFGpuError:
                        stats["err"] += 1; continue
                    if r == F.EOF: break
        finally:
            dec.close()
print("api fuzz ok", seed, stats)
