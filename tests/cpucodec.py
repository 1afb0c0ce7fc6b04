"""ctypes bindings to the two CPU checkers (TEST INFRASTRUCTURE):

  * ``ref``    -- oracle/_ref/libffv1ref.so: the UNMODIFIED reference FFV1 codec,
                  compiled from /root/reference by oracle/Makefile and driven through
                  its AVCodec vtable by oracle/ref_harness.c
  * ``oracle`` -- oracle/libffv1_oracle.so: the from-scratch CPU restatement
                  (oracle/ffv1_oracle.c)

Both export the same small C API (prefix ffv1ref_ / ffv1o_), so one wrapper serves
both.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.abspath(os.path.join(_HERE, ".."))
PATHS = {
    "ref": (os.path.join(_ROOT, "oracle", "_ref", "libffv1ref.so"), "ffv1ref_"),
    "oracle": (os.path.join(_ROOT, "oracle", "libffv1_oracle.so"), "ffv1o_"),
    # the product's libavcodec glue (integration/ffv1_gpu.c -> libffgpu.so) behind the SAME
    # AVCodec-vtable harness as "ref": needs a GPU to encode/decode
    "glue": (os.path.join(_ROOT, "oracle", "_ref", "libffv1glue.so"), "ffv1glue_"),
    # the product's device functions compiled for the CPU (tests/emul, test-only)
    "emul": (os.path.join(_ROOT, "tests", "emul", "libffv1_emul.so"), "ffv1emul_"),
}


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("pix_fmt", C.c_char_p),
        ("slices", C.c_int), ("level", C.c_int), ("gop_size", C.c_int),
        ("coder", C.c_int), ("context", C.c_int), ("slicecrc", C.c_int),
        ("strict", C.c_int), ("threads", C.c_int), ("bits_per_raw_sample", C.c_int),
    ]


class _Api:
    def __init__(self, path, prefix):
        L = C.CDLL(path)
        g = lambda n: getattr(L, prefix + n)
        self.encoder_open = g("encoder_open")
        self.encoder_open.restype = C.c_void_p
        self.encoder_open.argtypes = [C.POINTER(Params), C.POINTER(C.c_int)]
        self.encoder_extradata = g("encoder_extradata")
        self.encoder_extradata.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8))]
        self.encoder_info = g("encoder_info")
        self.encoder_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        self.encode = g("encode")
        self.encode.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        self.encoder_close = g("encoder_close")
        self.encoder_close.argtypes = [C.c_void_p]
        self.decoder_open = g("decoder_open")
        self.decoder_open.restype = C.c_void_p
        self.decoder_open.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                      C.POINTER(C.c_int)]
        self.decode = g("decode")
        self.decode.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p),
                                C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.POINTER(C.c_int)]
        self.decoder_close = g("decoder_close")
        self.decoder_close.argtypes = [C.c_void_p]
        self.plane_geometry = g("plane_geometry")
        self.plane_geometry.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.lib = L


_apis = {}


def available(which):
    return os.path.exists(PATHS[which][0])


def api(which):
    if which not in _apis:
        _apis[which] = _Api(*PATHS[which])
    return _apis[which]


def plane_geometry(pix_fmt, w, h, which="oracle"):
    """[(bytes per row, rows)] for every memory plane of pix_fmt."""
    a = api(which)
    bw, rows = C.c_int(), C.c_int()
    n = a.plane_geometry(pix_fmt.encode(), w, h, 0, C.byref(bw), C.byref(rows))
    if n < 0:
        raise ValueError("unknown pix_fmt %r" % pix_fmt)
    out = []
    for p in range(n):
        a.plane_geometry(pix_fmt.encode(), w, h, p, C.byref(bw), C.byref(rows))
        out.append((bw.value, rows.value))
    return out


class CodecError(RuntimeError):
    def __init__(self, what, code):
        RuntimeError.__init__(self, "%s failed: %d" % (what, code))
        self.code = code


class Encoder:
    def __init__(self, which, width, height, pix_fmt, slices=0, level=-99, gop_size=12, coder=0,
                 context=0, slicecrc=-1, strict=0, threads=1, bits_per_raw_sample=0,
                 pass1=0, pass2=0, stats_in=None):
        self.a = api(which)
        self._fmt = pix_fmt.encode()
        self.p = Params(width, height, self._fmt, slices, level, gop_size, coder, context,
                        slicecrc, strict, threads, bits_per_raw_sample)
        err = C.c_int()
        if pass1 or pass2 or stats_in:
            # two-pass coding: only the checkers that have it ("ref", "emul")
            self._stats = stats_in.encode() if isinstance(stats_in, str) else stats_in
            fn = getattr(self.a.lib, PATHS[which][1] + "encoder_open2")
            fn.restype = C.c_void_p
            fn.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_int)]
            self.h = fn(C.byref(self.p), pass1, pass2, self._stats, C.byref(err))
        else:
            self.h = self.a.encoder_open(C.byref(self.p), C.byref(err))
        self.which = which
        if not self.h:
            raise CodecError("%s encoder init" % which, err.value)
        self.width, self.height, self.pix_fmt = width, height, pix_fmt
        self._out = np.empty(16, np.uint8)

    @property
    def extradata(self):
        ptr = C.POINTER(C.c_uint8)()
        n = self.a.encoder_extradata(self.h, C.byref(ptr))
        return bytes(bytearray(ptr[:n])) if n > 0 else b""

    def stats_out(self):
        """AVCodecContext.stats_out after the flush of a first pass"""
        fn = getattr(self.a.lib, PATHS[self.which][1] + "encoder_stats_out")
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        buf = C.create_string_buffer(6 << 20)
        n = fn(self.h, buf, len(buf))
        if n < 0:
            raise CodecError("stats_out", n)
        return buf.raw[:n].decode()

    @property
    def info(self):
        v = (C.c_int * 8)()
        self.a.encoder_info(self.h, v)
        k = ["version", "micro_version", "ac", "num_h_slices", "num_v_slices", "ec",
             "bits_per_raw_sample", "colorspace"]
        return dict(zip(k, list(v)))

    def encode(self, planes, want_bytes=True):
        """planes: list of 2-D uint8 arrays (one row per picture line)."""
        ptrs = (C.c_void_p * 4)()
        ls = (C.c_int * 4)()
        for i, pl in enumerate(planes):
            assert pl.dtype == np.uint8 and pl.ndim == 2 and pl.strides[1] == 1
            ptrs[i] = pl.ctypes.data
            ls[i] = pl.strides[0]
        key = C.c_int()
        cap = sum(p.size for p in planes) * 2 + (1 << 20)
        if self._out.size < cap:
            self._out = np.empty(cap, np.uint8)
        n = self.a.encode(self.h, ptrs, ls, self._out.ctypes.data, cap, C.byref(key))
        if n < 0:
            raise CodecError("encode", n)
        self.last_key = key.value
        return self._out[:n].tobytes() if want_bytes else n

    def close(self):
        if getattr(self, "h", None):
            self.a.encoder_close(self.h)
            self.h = None

    def __del__(self):
        self.close()


class Decoder:
    def __init__(self, which, width, height, extradata=b"", threads=1):
        self.a = api(which)
        self.which = which
        err = C.c_int()
        self.h = self.a.decoder_open(width, height, extradata, len(extradata), threads,
                                     C.byref(err))
        if not self.h:
            raise CodecError("%s decoder init" % which, err.value)
        self.width, self.height = width, height

    def decode(self, pkt, copy=True):
        ptrs = (C.c_void_p * 4)()
        ls = (C.c_int * 4)()
        name = C.c_char_p()
        key = C.c_int()
        r = self.a.decode(self.h, pkt, len(pkt), ptrs, ls, C.byref(name), C.byref(key))
        if r < 0:
            raise CodecError("decode", r)
        self.pix_fmt = name.value.decode()
        self.last_key = key.value
        if not copy:
            return None
        planes = []
        for p, (bw, rows) in enumerate(plane_geometry(self.pix_fmt, self.width, self.height)):
            buf = (C.c_uint8 * (ls[p] * rows)).from_address(ptrs[p])
            planes.append(np.frombuffer(buf, np.uint8).reshape(rows, ls[p])[:, :bw].copy())
        return planes

    def close(self):
        if getattr(self, "h", None):
            self.a.decoder_close(self.h)
            self.h = None

    def __del__(self):
        self.close()
