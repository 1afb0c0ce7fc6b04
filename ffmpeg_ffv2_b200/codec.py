"""ctypes mirror of include/ffgpu.h (the C ABI of libffgpu.so)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

EAGAIN, ENOMEM, EINVAL, ENOSPC, ENOSYS = -11, -12, -22, -28, -38
EOF = -541478725
INVALIDDATA = -1094995529
EXTERNAL = -542398533


def lib_path():
    # FFGPU_LIB: development hook for A/B runs of differently built libraries
    return os.environ.get("FFGPU_LIB") or os.path.join(_HERE, "libffgpu.so")


class FFGpuError(RuntimeError):
    def __init__(self, what, code, text=""):
        RuntimeError.__init__(self, "%s failed: %d %s" % (what, code, text))
        self.code = code


class EncOptions(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("pix_fmt", C.c_char_p),
        ("slices", C.c_int), ("level", C.c_int), ("gop_size", C.c_int),
        ("coder", C.c_int), ("context", C.c_int), ("slicecrc", C.c_int),
        ("strict_std_compliance", C.c_int), ("bits_per_raw_sample", C.c_int),
        ("device", C.c_int), ("max_batch", C.c_int), ("pipeline_depth", C.c_int),
        ("ndevices", C.c_int), ("devices", C.c_int * 16),
        ("pass1", C.c_int), ("pass2", C.c_int), ("stats_in", C.c_char_p),
    ]


class Picture(C.Structure):
    _fields_ = [
        ("data", C.c_void_p * 4), ("linesize", C.c_int * 4),
        ("interlaced_frame", C.c_int), ("top_field_first", C.c_int),
        ("sar_num", C.c_int), ("sar_den", C.c_int), ("pts", C.c_int64),
    ]


class DecOptions(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("extradata", C.c_char_p),
        ("extradata_size", C.c_int), ("device", C.c_int), ("max_batch", C.c_int),
        ("pipeline_depth", C.c_int), ("ndevices", C.c_int), ("devices", C.c_int * 16),
    ]


class PictureOut(C.Structure):
    _fields_ = [
        ("data", C.c_void_p * 4), ("linesize", C.c_int * 4),
        ("key_frame", C.c_int), ("interlaced_frame", C.c_int), ("top_field_first", C.c_int),
        ("sar_num", C.c_int), ("sar_den", C.c_int), ("damaged_slices", C.c_int),
        ("pts", C.c_int64),
    ]


# every symbol include/ffgpu.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("ffgpu_ffv1_encode_init", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(EncOptions)]),
    ("ffgpu_ffv1_encoder_extradata", C.c_int, [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8))]),
    ("ffgpu_ffv1_encoder_info", None, [C.c_void_p, C.POINTER(C.c_int)]),
    ("ffgpu_ffv1_encoder_max_packet", C.c_size_t, [C.c_void_p]),
    ("ffgpu_ffv1_encode_frame", C.c_int, [C.c_void_p, C.POINTER(Picture), C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    ("ffgpu_ffv1_encode_send_frame", C.c_int, [C.c_void_p, C.POINTER(Picture)]),
    ("ffgpu_ffv1_encode_receive_packet", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t,
                                                   C.POINTER(C.c_size_t), C.POINTER(C.c_int),
                                                   C.POINTER(C.c_int64)]),
    ("ffgpu_ffv1_encode_packet_ready", C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    ("ffgpu_ffv1_encode_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("ffgpu_ffv1_encode_device_result", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p),
                                                  C.POINTER(C.c_size_t)]),
    ("ffgpu_ffv1_encode_device_fetch", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                                 C.POINTER(C.c_size_t)]),
    ("ffgpu_ffv1_encoder_stats_out", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    ("ffgpu_ffv1_encode_close", C.c_int, [C.c_void_p]),
    ("ffgpu_ffv1_decode_init", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(DecOptions)]),
    ("ffgpu_ffv1_decoder_pix_fmt", C.c_char_p, [C.c_void_p]),
    ("ffgpu_ffv1_decoder_info", None, [C.c_void_p, C.POINTER(C.c_int)]),
    ("ffgpu_ffv1_decoder_probe", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    ("ffgpu_ffv1_decode_frame", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t,
                                          C.POINTER(PictureOut), C.POINTER(C.c_int)]),
    ("ffgpu_ffv1_decode_send_packet", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int64,
                                                C.POINTER(PictureOut)]),
    ("ffgpu_ffv1_decode_receive_frame", C.c_int, [C.c_void_p, C.POINTER(PictureOut)]),
    ("ffgpu_ffv1_decode_device", C.c_int, [C.c_void_p, C.POINTER(C.c_char_p),
                                           C.POINTER(C.c_size_t), C.c_int, C.c_void_p, C.c_void_p]),
    ("ffgpu_ffv1_decode_device_status", C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int]),
    ("ffgpu_ffv1_decode_close", C.c_int, [C.c_void_p]),
    ("ffgpu_ffv1_frame_layout", C.c_size_t, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_int), C.POINTER(C.c_int),
                                             C.POINTER(C.c_int)]),
    ("ffgpu_ffv1_encoder_launches", C.c_uint64, [C.c_void_p]),
    ("ffgpu_ffv1_decoder_launches", C.c_uint64, [C.c_void_p]),
    ("ffgpu_ffv1_encoder_profile", C.c_int, [C.c_void_p, C.c_int]),
    ("ffgpu_ffv1_encoder_kernel_ms", C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    ("ffgpu_ffv1_decoder_profile", C.c_int, [C.c_void_p, C.c_int]),
    ("ffgpu_ffv1_decoder_kernel_ms", C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    ("ffgpu_ffv1_encoder_decisions", C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    ("ffgpu_cuda_push_context", C.c_int, [C.c_void_p]),
    ("ffgpu_cuda_pop_context", C.c_int, []),
    ("ffgpu_last_error", C.c_char_p, []),
    ("ffgpu_abi_version", C.c_int, []),
]

_lib = None


def lib():
    """Loads libffgpu.so.  Fails loudly if the CUDA library has not been built: there is no
    CPU implementation to fall back to."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(
                "%s is missing: build it with `python -m ffmpeg_ffv2_b200.build` (nvcc, sm_100a). "
                "The FFV1 pixel path has no CPU fallback." % path)
        L = C.CDLL(path)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _err():
    return (lib().ffgpu_last_error() or b"").decode(errors="replace")


def frame_layout(pix_fmt, width, height):
    """device picture layout of the *_device entry points:
    (frame_bytes, [(offset, pitch, rows, rowbytes) per plane])"""
    off = (C.c_size_t * 4)()
    pitch = (C.c_int * 4)()
    rows = (C.c_int * 4)()
    rb = (C.c_int * 4)()
    n = lib().ffgpu_ffv1_frame_layout(pix_fmt.encode(), width, height, off, pitch, rows, rb)
    if not n:
        raise ValueError("unknown pix_fmt %r" % pix_fmt)
    return n, [(off[i], pitch[i], rows[i], rb[i]) for i in range(4) if rows[i]]


def _picture(planes, pts=0, interlaced=0, tff=0, sar=(0, 1)):
    p = Picture()
    for i, a in enumerate(planes):
        assert a.dtype == np.uint8 and a.ndim == 2 and a.strides[1] == 1
        p.data[i] = a.ctypes.data
        p.linesize[i] = a.strides[0]
    p.interlaced_frame, p.top_field_first = interlaced, tff
    p.sar_num, p.sar_den = sar
    p.pts = pts
    return p


class FFV1Encoder:
    """Mirror of ff_ffv1_encoder (ffv1enc.c:1323): options are the AVOptions / AVCodecContext
    fields encode_init reads; errors are FFGpuError carrying the AVERROR code."""

    def __init__(self, width, height, pix_fmt, slices=0, level=-99, gop_size=12, coder=0, context=0,
                 slicecrc=-1, strict=0, bits_per_raw_sample=0, device=0, max_batch=0,
                 pipeline_depth=0, devices=(), pass1=0, pass2=0, stats_in=None):
        self._fmt = pix_fmt.encode()
        self._stats = stats_in.encode() if isinstance(stats_in, str) else stats_in
        self.opt = EncOptions(width, height, self._fmt, slices, level, gop_size, coder, context,
                              slicecrc, strict, bits_per_raw_sample, device, max_batch,
                              pipeline_depth, len(devices), (C.c_int * 16)(*devices),
                              pass1, pass2, self._stats)
        self.h = C.c_void_p()
        r = lib().ffgpu_ffv1_encode_init(C.byref(self.h), C.byref(self.opt))
        if r < 0:
            self.h = None
            raise FFGpuError("encode_init", r, _err())
        self.width, self.height, self.pix_fmt = width, height, pix_fmt
        self._buf = np.empty(int(lib().ffgpu_ffv1_encoder_max_packet(self.h)) + 64, np.uint8)
        self._keep = []

    @property
    def extradata(self):
        ptr = C.POINTER(C.c_uint8)()
        n = lib().ffgpu_ffv1_encoder_extradata(self.h, C.byref(ptr))
        return bytes(bytearray(ptr[:n])) if n > 0 else b""

    @property
    def info(self):
        v = (C.c_int * 8)()
        lib().ffgpu_ffv1_encoder_info(self.h, v)
        k = ["version", "micro_version", "ac", "num_h_slices", "num_v_slices", "ec",
             "bits_per_raw_sample", "colorspace"]
        return dict(zip(k, list(v)))

    @property
    def launches(self):
        return int(lib().ffgpu_ffv1_encoder_launches(self.h))

    def encode(self, planes, **kw):
        """AVCodec.encode2: one picture (list of 2-D uint8 planes) -> packet bytes"""
        pic = _picture(planes, **kw)
        n = C.c_size_t()
        key = C.c_int()
        r = lib().ffgpu_ffv1_encode_frame(self.h, C.byref(pic), self._buf.ctypes.data, self._buf.size,
                                          C.byref(n), C.byref(key))
        if r < 0:
            raise FFGpuError("encode_frame", r, _err())
        self.last_key = key.value
        return self._buf[:n.value].tobytes()

    def send_frame(self, planes, pts=0, **kw):
        """returns False on EAGAIN (receive first)"""
        if planes is None:
            r = lib().ffgpu_ffv1_encode_send_frame(self.h, None)
        else:
            pic = _picture(planes, pts=pts, **kw)
            r = lib().ffgpu_ffv1_encode_send_frame(self.h, C.byref(pic))
        if r == EAGAIN:
            return False
        if r < 0:
            raise FFGpuError("send_frame", r, _err())
        return True

    def receive_packet(self, copy=True):
        """(bytes, key, pts) | None on EAGAIN | EOF constant after flush"""
        n = C.c_size_t()
        key = C.c_int()
        pts = C.c_int64()
        r = lib().ffgpu_ffv1_encode_receive_packet(self.h, self._buf.ctypes.data, self._buf.size,
                                                   C.byref(n), C.byref(key), C.byref(pts))
        if r == EAGAIN:
            return None
        if r == EOF:
            return EOF
        if r < 0:
            raise FFGpuError("receive_packet", r, _err())
        return (self._buf[:n.value].tobytes() if copy else n.value), key.value, pts.value

    def encode_device(self, d_ptr, nframes, stream=0):
        r = lib().ffgpu_ffv1_encode_device(self.h, C.c_void_p(d_ptr), nframes, C.c_void_p(stream))
        if r < 0:
            raise FFGpuError("encode_device", r, _err())

    def device_fetch(self, frame):
        n = C.c_size_t()
        r = lib().ffgpu_ffv1_encode_device_fetch(self.h, frame, self._buf.ctypes.data, self._buf.size,
                                                 C.byref(n))
        if r < 0:
            raise FFGpuError("encode_device_fetch", r, _err())
        return self._buf[:n.value].tobytes()

    ENC_KERNELS = ["symbolize", "fill_state", "sort", "code", "pack_slice_scan",
                   "pack_frame_scan", "pack_gather"]

    def profile(self, enable=True):
        r = lib().ffgpu_ffv1_encoder_profile(self.h, int(enable))
        if r < 0:
            raise FFGpuError("encoder_profile", r, _err())

    def kernel_ms(self):
        ms = (C.c_float * 8)()
        n = lib().ffgpu_ffv1_encoder_kernel_ms(self.h, ms, 8)
        if n < 0:
            raise FFGpuError("encoder_kernel_ms", n, _err())
        return dict(zip(self.ENC_KERNELS, list(ms)[:n]))

    def stats_out(self):
        """AVCodecContext.stats_out of a first pass (text)"""
        buf = C.create_string_buffer(6 << 20)
        n = lib().ffgpu_ffv1_encoder_stats_out(self.h, buf, len(buf))
        if n < 0:
            raise FFGpuError("encoder_stats_out", n, _err())
        return buf.raw[:n].decode()

    def decisions(self):
        """(total binary decisions, decisions of the heaviest slice) of the last device batch"""
        tot = C.c_uint64()
        mx = C.c_uint32()
        r = lib().ffgpu_ffv1_encoder_decisions(self.h, C.byref(tot), C.byref(mx))
        if r < 0:
            raise FFGpuError("encoder_decisions", r, _err())
        return tot.value, mx.value

    def device_result(self, frame):
        p = C.c_void_p()
        n = C.c_size_t()
        r = lib().ffgpu_ffv1_encode_device_result(self.h, frame, C.byref(p), C.byref(n))
        if r < 0:
            raise FFGpuError("encode_device_result", r, _err())
        return p.value, n.value

    def close(self):
        if getattr(self, "h", None):
            lib().ffgpu_ffv1_encode_close(self.h)
            self.h = None

    def __del__(self):
        self.close()


class FFV1Decoder:
    """Mirror of ff_ffv1_decoder (ffv1dec.c:1087)."""

    def __init__(self, width, height, extradata=b"", device=0, max_batch=0, pipeline_depth=0,
                 devices=()):
        self._ex = bytes(extradata)
        self.opt = DecOptions(width, height, self._ex, len(self._ex), device, max_batch,
                              pipeline_depth, len(devices), (C.c_int * 16)(*devices))
        self.h = C.c_void_p()
        r = lib().ffgpu_ffv1_decode_init(C.byref(self.h), C.byref(self.opt))
        if r < 0:
            self.h = None
            raise FFGpuError("decode_init", r, _err())
        self.width, self.height = width, height
        self._pending = []
        self._idle = (PictureOut(), [])

    @property
    def pix_fmt(self):
        v = lib().ffgpu_ffv1_decoder_pix_fmt(self.h)
        return v.decode() if v else None

    @property
    def launches(self):
        return int(lib().ffgpu_ffv1_decoder_launches(self.h))

    def _alloc_out(self, fmt):
        from . import codec  # noqa: F401
        _, planes = frame_layout(fmt, self.width, self.height)
        arrs = [np.zeros((rows, rb), np.uint8) for (_o, _p, rows, rb) in planes]
        out = PictureOut()
        for i, a in enumerate(arrs):
            out.data[i] = a.ctypes.data
            out.linesize[i] = a.strides[0]
        return out, arrs

    def decode(self, pkt, fmt_hint=None):
        """AVCodec.decode: packet bytes -> list of 2-D uint8 planes.  For v0/v1 streams the
        output format is only known after the header; pass fmt_hint or rely on a retry."""
        if self.pix_fmt is None:
            r = lib().ffgpu_ffv1_decoder_probe(self.h, pkt, len(pkt))
            if r < 0:
                raise FFGpuError("decoder_probe", r, _err())
        fmt = self.pix_fmt or fmt_hint
        out, arrs = self._alloc_out(fmt)
        got = C.c_int()
        r = lib().ffgpu_ffv1_decode_frame(self.h, pkt, len(pkt), C.byref(out), C.byref(got))
        if r < 0:
            raise FFGpuError("decode_frame", r, _err())
        self.last = out
        return arrs

    def send_packet(self, pkt, pts=0, dst=None):
        """dst: optional (PictureOut, arrays) from alloc_picture() to decode straight into"""
        if pkt is None:
            r = lib().ffgpu_ffv1_decode_send_packet(self.h, None, 0, 0, None)
        else:
            r = lib().ffgpu_ffv1_decode_send_packet(self.h, pkt, len(pkt), pts,
                                                    C.byref(dst[0]) if dst else None)
        if r == EAGAIN:
            return False
        if r < 0:
            raise FFGpuError("send_packet", r, _err())
        if pkt is not None:
            self._pending.append(dst)          # None: destination supplied at receive time
        return True

    def alloc_picture(self, fmt=None):
        return self._alloc_out(fmt or self.pix_fmt)

    def receive_frame(self, out=None):
        """returns (PictureOut, arrays) | None on EAGAIN | EOF"""
        if self._pending and self._pending[0] is not None:
            o = self._pending[0]               # decoded straight into the picture given at send
        elif self._pending:
            o = out if out is not None else self._alloc_out(self.pix_fmt)
        else:
            o = self._idle                     # nothing in flight: only EAGAIN / EOF can come back
        r = lib().ffgpu_ffv1_decode_receive_frame(self.h, C.byref(o[0]))
        if r == 0 and self._pending:
            self._pending.pop(0)
        if r == EAGAIN:
            return None
        if r == EOF:
            return EOF
        if r < 0:
            raise FFGpuError("receive_frame", r, _err())
        return o

    DEC_KERNELS = ["init_state", "sort", "decode"]

    def profile(self, enable=True):
        r = lib().ffgpu_ffv1_decoder_profile(self.h, int(enable))
        if r < 0:
            raise FFGpuError("decoder_profile", r, _err())

    def kernel_ms(self):
        ms = (C.c_float * 4)()
        n = lib().ffgpu_ffv1_decoder_kernel_ms(self.h, ms, 4)
        if n < 0:
            raise FFGpuError("decoder_kernel_ms", n, _err())
        return dict(zip(self.DEC_KERNELS, list(ms)[:n]))

    def decode_device(self, pkts, d_ptr, stream=0):
        n = len(pkts)
        arr = (C.c_char_p * n)(*pkts)
        sizes = (C.c_size_t * n)(*[len(p) for p in pkts])
        r = lib().ffgpu_ffv1_decode_device(self.h, arr, sizes, n, C.c_void_p(d_ptr), C.c_void_p(stream))
        if r < 0:
            raise FFGpuError("decode_device", r, _err())
        self._n_device = n

    def device_status(self):
        """damaged-slice count per picture of the last decode_device() batch (waits for it)"""
        n = getattr(self, "_n_device", 0)
        arr = (C.c_int * max(n, 1))()
        r = lib().ffgpu_ffv1_decode_device_status(self.h, arr, n)
        if r < 0:
            raise FFGpuError("decode_device_status", r, _err())
        return list(arr)[:n]

    def close(self):
        if getattr(self, "h", None):
            lib().ffgpu_ffv1_decode_close(self.h)
            self.h = None

    def __del__(self):
        self.close()
