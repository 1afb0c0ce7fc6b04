"""ffmpeg_ffv2_b200 -- B200-native (sm_100a) FFV1 slice pixel path behind a C ABI.

The product is ``libffgpu.so`` (CUDA kernels + host C, see ``csrc/`` and
``include/ffgpu.h``).  This package is a thin ctypes mirror of the reference's
``ff_ffv1_encoder`` / ``ff_ffv1_decoder`` operator interface (same option names,
same error codes) used by the tests and the bench.  There is no CPU pixel path:
importing works anywhere, encoding/decoding needs a CUDA device.
"""
from .codec import (FFV1Decoder, FFV1Encoder, FFGpuError, frame_layout, lib, lib_path,  # noqa: F401
                    EAGAIN, EOF, EINVAL, ENOSYS, ENOSPC, INVALIDDATA, EXTERNAL)

__all__ = ["FFV1Encoder", "FFV1Decoder", "FFGpuError", "frame_layout", "lib", "lib_path"]
