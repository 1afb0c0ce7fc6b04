"""Deterministic synthetic pictures for the parity tests and the bench (numpy only).

``testsrc2_like`` / ``mandelbrot`` / ``noise`` stand in for the lavfi sources named in
BASELINE.json (the GPU box has no ffmpeg binary): flat colour bars with sharp edges,
smooth gradients and a moving diagonal (testsrc2-like, highly compressible), a smooth
fractal (mandelbrot), and full-depth white noise (incompressible).
"""
import numpy as np

import cpucodec as cc

# (layout, depth, hshift, vshift, alpha); layout: planar / ya8 / bgr32 / gbrp / rgb48
_FMT = {}


def _add(name, layout, depth, hs=0, vs=0, alpha=0, chroma=1):
    _FMT[name] = dict(layout=layout, depth=depth, hs=hs, vs=vs, alpha=alpha, chroma=chroma)


for _d, _s in ((8, ""), (9, "9le"), (10, "10le"), (12, "12le"), (14, "14le"), (16, "16le")):
    for _n, _h, _v in (("444", 0, 0), ("440", 0, 1), ("422", 1, 0), ("420", 1, 1), ("411", 2, 0),
                       ("410", 2, 2)):
        _add("yuv%sp%s" % (_n, _s), "planar", _d, _h, _v)
        _add("yuva%sp%s" % (_n, _s), "planar", _d, _h, _v, alpha=1)
    _add("gray%s" % _s, "planar", _d, chroma=0)
    _add("gbrp%s" % _s, "gbrp", _d)
    _add("gbrap%s" % _s, "gbrp", _d, alpha=1)
_add("ya8", "ya8", 8, alpha=1, chroma=0)
_add("bgr0", "bgr32", 8)
_add("bgra", "bgr32", 8, alpha=1)
_add("rgb48le", "rgb48", 16)
_add("rgba64le", "rgb48", 16, alpha=1)


def fmt_info(pix_fmt):
    return _FMT[pix_fmt]


def _components(pix_fmt, w, h):
    """list of (width, height) of the logical components in coding order"""
    f = _FMT[pix_fmt]
    cw, ch = -(-w >> f["hs"]), -(-h >> f["vs"])
    if f["layout"] == "planar":
        comps = [(w, h)] + ([(cw, ch), (cw, ch)] if f["chroma"] else [])
        return comps + ([(w, h)] if f["alpha"] else [])
    if f["layout"] == "ya8":
        return [(w, h), (w, h)]
    return [(w, h)] * (3 + f["alpha"])


def pack(pix_fmt, comps):
    """logical component arrays (int, coding order Y,U,V,A or G,B,R,A... see below) ->
    list of 2-D uint8 memory planes in the pix_fmt's layout.
    For RGB layouts comps are given as R,G,B[,A]."""
    f = _FMT[pix_fmt]
    d = f["depth"]
    if f["layout"] == "planar":
        if d <= 8:
            return [np.ascontiguousarray(c.astype(np.uint8)) for c in comps]
        return [np.ascontiguousarray(c.astype("<u2")).view(np.uint8).reshape(c.shape[0], -1)
                for c in comps]
    if f["layout"] == "ya8":
        a = np.stack([comps[0], comps[1]], axis=-1).astype(np.uint8)
        return [np.ascontiguousarray(a.reshape(a.shape[0], -1))]
    if f["layout"] == "bgr32":
        r, g, b = comps[:3]
        a = comps[3] if f["alpha"] else np.zeros_like(r)
        px = np.stack([b, g, r, a], axis=-1).astype(np.uint8)
        return [np.ascontiguousarray(px.reshape(px.shape[0], -1))]
    if f["layout"] == "gbrp":
        r, g, b = comps[:3]
        order = [g, b, r] + ([comps[3]] if f["alpha"] else [])
        return [np.ascontiguousarray(c.astype("<u2")).view(np.uint8).reshape(c.shape[0], -1)
                for c in order]
    if f["layout"] == "rgb48":
        px = np.stack(list(comps), axis=-1).astype("<u2")
        return [np.ascontiguousarray(px).view(np.uint8).reshape(px.shape[0], -1)]
    raise ValueError(pix_fmt)


def _scale(a, depth):
    """a in [0,1) float -> integer samples of the given depth"""
    return np.clip((a * (1 << depth)).astype(np.int64), 0, (1 << depth) - 1)


def testsrc2_like(pix_fmt, w, h, frame=0):
    f = _FMT[pix_fmt]
    out = []
    for ci, (cw, ch) in enumerate(_components(pix_fmt, w, h)):
        y, x = np.mgrid[0:ch, 0:cw]
        xs, ys = x / max(cw, 1), y / max(ch, 1)
        bars = ((xs * 8).astype(np.int64) * (37 + 11 * ci) % 256) / 256.0       # colour bars
        grad = (xs + 0.13 * ci) % 1.0                                          # horizontal ramp
        ramp2 = ((x + y + 3 * frame) % 256) / 256.0                            # moving diagonal
        checker = (((x >> 3) + (y >> 3) + frame) & 1) * 0.75
        img = np.where(ys < 0.45, bars, np.where(ys < 0.6, grad, np.where(ys < 0.8, ramp2, checker)))
        # a moving box and a thin line pattern, like testsrc2's animated elements
        bx = (frame * 7) % max(cw - cw // 8, 1)
        box = (x >= bx) & (x < bx + cw // 8) & (ys > 0.2) & (ys < 0.35)
        img = np.where(box, 0.9 - 0.2 * ci, img)
        out.append(_scale(img, f["depth"]))
    return pack(pix_fmt, out)


def mandelbrot(pix_fmt, w, h, frame=0, iters=48):
    f = _FMT[pix_fmt]
    out = []
    zoom = 1.5 / (1.0 + 0.05 * frame)
    for ci, (cw, ch) in enumerate(_components(pix_fmt, w, h)):
        y, x = np.mgrid[0:ch, 0:cw]
        c = (-0.745 + (x / max(cw, 1) - 0.5) * 2 * zoom) + 1j * (0.113 + (y / max(ch, 1) - 0.5) * 2 * zoom * ch / max(cw, 1))
        z = np.zeros_like(c)
        n = np.zeros(c.shape, np.float64)
        for _ in range(iters):
            m = np.abs(z) <= 2.0
            z = np.where(m, z * z + c, z)
            n += m
        img = ((n / iters) * (1.0 + 0.5 * ci)) % 1.0
        out.append(_scale(img, f["depth"]))
    return pack(pix_fmt, out)


def noise(pix_fmt, w, h, frame=0, seed=1234):
    f = _FMT[pix_fmt]
    rng = np.random.default_rng(seed + 7919 * frame)
    out = [rng.integers(0, 1 << f["depth"], size=(ch, cw)) for (cw, ch) in _components(pix_fmt, w, h)]
    return pack(pix_fmt, out)


def smooth(pix_fmt, w, h, frame=0, seed=99):
    """random-walk texture: moderate entropy, exercises many contexts"""
    f = _FMT[pix_fmt]
    rng = np.random.default_rng(seed + 31 * frame)
    out = []
    for (cw, ch) in _components(pix_fmt, w, h):
        a = (np.cumsum(rng.integers(-3, 4, size=(ch, cw)), axis=1) +
             np.cumsum(rng.integers(-2, 3, size=(ch, 1)), axis=0) + (1 << (f["depth"] - 1)))
        out.append(a % (1 << f["depth"]))
    return pack(pix_fmt, out)


def extremes(pix_fmt, w, h, frame=0):
    """max/min alternations: largest residuals, sign/wrap corner cases"""
    f = _FMT[pix_fmt]
    out = []
    for (cw, ch) in _components(pix_fmt, w, h):
        a = np.full((ch, cw), (1 << f["depth"]) - 1, np.int64)
        a[:, ::7] = 0
        a[::5] = 3
        a[(frame % 3)::11, ::2] = (1 << (f["depth"] - 1))
        out.append(a)
    return pack(pix_fmt, out)


GENERATORS = dict(testsrc2=testsrc2_like, mandelbrot=mandelbrot, noise=noise, smooth=smooth,
                  extremes=extremes)


def frame_bytes(planes):
    return sum(p.size for p in planes)
