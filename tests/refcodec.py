"""ctypes binding to oracle/_ref/libffv1ref.so (the UNMODIFIED reference FFV1
codec driven through its AVCodec vtable by oracle/ref_harness.c).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "..", "oracle", "_ref", "libffv1ref.so")


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("pix_fmt", C.c_char_p),
        ("slices", C.c_int), ("level", C.c_int), ("gop_size", C.c_int),
        ("coder", C.c_int), ("context", C.c_int), ("slicecrc", C.c_int),
        ("strict", C.c_int), ("threads", C.c_int), ("bits_per_raw_sample", C.c_int),
    ]


_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        L.ffv1ref_encoder_open.restype = C.c_void_p
        L.ffv1ref_encoder_open.argtypes = [C.POINTER(Params), C.POINTER(C.c_int)]
        L.ffv1ref_encoder_extradata.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8))]
        L.ffv1ref_encoder_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ffv1ref_encode.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                     C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ffv1ref_encoder_close.argtypes = [C.c_void_p]
        L.ffv1ref_decoder_open.restype = C.c_void_p
        L.ffv1ref_decoder_open.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                           C.POINTER(C.c_int)]
        L.ffv1ref_decode.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.POINTER(C.c_int)]
        L.ffv1ref_decoder_close.argtypes = [C.c_void_p]
        L.ffv1ref_plane_geometry.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int,
                                             C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ffv1ref_last_error.restype = C.c_char_p
        L.ffv1ref_set_log_level.argtypes = [C.c_int]
        _lib = L
    return _lib


def plane_geometry(pix_fmt, w, h):
    """[(bytewidth, rows)] per plane according to libavutil pixdesc."""
    out = []
    bw, rows = C.c_int(), C.c_int()
    n = lib().ffv1ref_plane_geometry(pix_fmt.encode(), w, h, 0, C.byref(bw), C.byref(rows))
    if n < 0:
        raise ValueError(pix_fmt)
    for p in range(n):
        lib().ffv1ref_plane_geometry(pix_fmt.encode(), w, h, p, C.byref(bw), C.byref(rows))
        out.append((bw.value, rows.value))
    return out


class RefEncoder:
    def __init__(self, width, height, pix_fmt, slices=0, level=-99, gop_size=12, coder=0,
                 context=0, slicecrc=-1, strict=0, threads=1, bits_per_raw_sample=0):
        self._fmt = pix_fmt.encode()
        self.p = Params(width, height, self._fmt, slices, level, gop_size, coder, context,
                        slicecrc, strict, threads, bits_per_raw_sample)
        err = C.c_int()
        self.h = lib().ffv1ref_encoder_open(C.byref(self.p), C.byref(err))
        if not self.h:
            raise RuntimeError("ref encoder init failed: %d %s" % (
                err.value, lib().ffv1ref_last_error().decode(errors="replace")))
        self.err = err.value
        self.width, self.height, self.pix_fmt = width, height, pix_fmt
        self._out = np.empty(16, np.uint8)

    @property
    def extradata(self):
        ptr = C.POINTER(C.c_uint8)()
        n = lib().ffv1ref_encoder_extradata(self.h, C.byref(ptr))
        return bytes(bytearray(ptr[:n])) if n > 0 else b""

    @property
    def info(self):
        a = (C.c_int * 8)()
        lib().ffv1ref_encoder_info(self.h, a)
        k = ["version", "micro_version", "ac", "num_h_slices", "num_v_slices", "ec",
             "bits_per_raw_sample", "colorspace"]
        return dict(zip(k, list(a)))

    def encode(self, planes, want_bytes=True):
        """planes: list of 2-D uint8 arrays (row = linesize bytes)."""
        ptrs = (C.c_void_p * 4)()
        ls = (C.c_int * 4)()
        for i, pl in enumerate(planes):
            assert pl.dtype == np.uint8 and pl.ndim == 2 and pl.strides[1] == 1
            ptrs[i] = pl.ctypes.data
            ls[i] = pl.strides[0]
        key = C.c_int()
        if want_bytes:
            cap = sum(p.size for p in planes) * 2 + (1 << 20)
            if self._out.size < cap:
                self._out = np.empty(cap, np.uint8)
            n = lib().ffv1ref_encode(self.h, ptrs, ls, self._out.ctypes.data, cap, C.byref(key))
        else:
            n = lib().ffv1ref_encode(self.h, ptrs, ls, None, 0, C.byref(key))
        if n < 0:
            raise RuntimeError("ref encode failed: %d %s" % (
                n, lib().ffv1ref_last_error().decode(errors="replace")))
        self.last_key = key.value
        return self._out[:n].tobytes() if want_bytes else n

    def close(self):
        if self.h:
            lib().ffv1ref_encoder_close(self.h)
            self.h = None

    def __del__(self):
        self.close()


class RefDecoder:
    def __init__(self, width, height, extradata=b"", threads=1):
        err = C.c_int()
        self.h = lib().ffv1ref_decoder_open(width, height, extradata, len(extradata), threads,
                                            C.byref(err))
        if not self.h:
            raise RuntimeError("ref decoder init failed: %d" % err.value)
        self.width, self.height = width, height

    def decode(self, pkt, copy=True):
        ptrs = (C.c_void_p * 4)()
        ls = (C.c_int * 4)()
        name = C.c_char_p()
        key = C.c_int()
        r = lib().ffv1ref_decode(self.h, pkt, len(pkt), ptrs, ls, C.byref(name), C.byref(key))
        if r < 0:
            raise RuntimeError("ref decode failed: %d %s" % (
                r, lib().ffv1ref_last_error().decode(errors="replace")))
        self.pix_fmt = name.value.decode()
        self.last_key = key.value
        if not copy:
            return None
        planes = []
        for p, (bw, rows) in enumerate(plane_geometry(self.pix_fmt, self.width, self.height)):
            buf = (C.c_uint8 * (ls[p] * rows)).from_address(ptrs[p])
            a = np.frombuffer(buf, np.uint8).reshape(rows, ls[p])[:, :bw].copy()
            planes.append(a)
        return planes

    def close(self):
        if self.h:
            lib().ffv1ref_decoder_close(self.h)
            self.h = None

    def __del__(self):
        self.close()
