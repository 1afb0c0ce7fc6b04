"""Test helper: a few lines of FFV1's range ENCODER in Python (rangecoder.h:71-121,
ffv1enc.c:185-231), used to hand-craft slice headers the real encoders never write."""
import ctypes as C


def default_tables():
    """one_state/zero_state of ff_build_rac_states(0.05*2^32, 248), from the host C layer"""
    import ffmpeg_ffv2_b200 as F
    lib = C.CDLL(F.lib_path())
    buf = (C.c_uint8 * 512)()
    lib.ff_default_tables(buf)
    b = bytes(buf)
    return list(b[:256]), list(b[256:])


def crc32_mpeg(data, crc=0):
    """AV_CRC_32_IEEE as FFV1 uses it: poly 0x04C11DB7, MSB first, init 0, no final xor"""
    for byte in data:
        crc ^= byte << 24
        for _ in range(8):
            crc = ((crc << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if crc & 0x80000000 else (crc << 1) & 0xFFFFFFFF
    return crc


class RangeEncoder:
    def __init__(self, one, zero):
        self.one, self.zero = one, zero
        self.low, self.range = 0, 0xFF00
        self.outstanding_count, self.outstanding_byte = 0, -1
        self.out = bytearray()

    def _renorm(self):
        while self.range < 0x100:
            if self.outstanding_byte < 0:
                self.outstanding_byte = self.low >> 8
            elif self.low <= 0xFF00:
                self.out.append(self.outstanding_byte)
                self.out.extend(b"\xff" * self.outstanding_count)
                self.outstanding_count = 0
                self.outstanding_byte = self.low >> 8
            elif self.low >= 0x10000:
                self.out.append((self.outstanding_byte + 1) & 0xFF)
                self.out.extend(b"\x00" * self.outstanding_count)
                self.outstanding_count = 0
                self.outstanding_byte = (self.low >> 8) & 0xFF
            else:
                self.outstanding_count += 1
            self.low = (self.low & 0xFF) << 8
            self.range <<= 8

    def put(self, state, i, bit):
        r1 = (self.range * state[i]) >> 8
        if not bit:
            self.range -= r1
            state[i] = self.zero[state[i]]
        else:
            self.low += self.range - r1
            self.range = r1
            state[i] = self.one[state[i]]
        self._renorm()

    def put_symbol(self, state, v, signed=False):
        if v == 0:
            self.put(state, 0, 1)
            return
        a = abs(v)
        e = a.bit_length() - 1
        self.put(state, 0, 0)
        for i in range(e):
            self.put(state, 1 + min(i, 9), 1)
        self.put(state, 1 + min(e, 9), 0)
        for i in range(e - 1, -1, -1):
            self.put(state, 22 + min(i, 9), (a >> i) & 1)
        if signed:
            self.put(state, 11 + min(e, 10), v < 0)

    def terminate(self, version=1):
        if version == 1:
            st = [129]
            self.put(st, 0, 0)
        self.range = 0xFF
        self.low += 0xFF
        self._renorm()
        self.range = 0xFF
        self._renorm()
        return bytes(self.out)


def split_v3_packet(pkt, ec=True):
    """[(start, payload_size)] of the slices of a v3 packet, walked from the tail
    (ffv1dec.c:890-903)"""
    trailer = 8 if ec else 3
    out, end = [], len(pkt)
    while end > 0:
        size = int.from_bytes(pkt[end - trailer:end - trailer + 3], "big")
        out.append((end - trailer - size, size))
        end -= size + trailer
    return out[::-1]


def wrap_slice(payload, ec=True):
    """payload + 24-bit size [+ 0x00 + CRC-32 LE], ffv1enc.c:1248-1261"""
    b = bytes(payload) + len(payload).to_bytes(3, "big")
    if ec:
        b += b"\x00"
        b += crc32_mpeg(b).to_bytes(4, "big")      # av_crc value stored with AV_WL32 == these bytes
    return b
