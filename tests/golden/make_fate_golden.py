#!/usr/bin/env python3
"""Generates tests/golden/fate_*.{npz,json} -- run ONCE in the build container.

Chain of custody for the parity pin (TEST INFRASTRUCTURE):

  1. the reference's own FATE sources are generated with the reference's test tools
     tests/videogen.c / tests/rotozoom.c (compiled directly with gcc);
  2. a scratch build of the reference ffmpeg (in /tmp, never in the repo) runs the
     exact enc_dec command of tests/fate-run.sh:188-210 for the 7 FFV1 variants of
     tests/fate/vcodec.mak:168-185 on vsynth1/2/3, and the AVI md5 + size must equal
     the reference's committed goldens tests/ref/vsynth/vsynth{1,2,3}-ffv1*;
  3. the per-packet md5s and the extradata are read back from those very AVI files;
  4. the direct-gcc build of the reference codec (oracle/_ref/libffv1ref.so) must
     reproduce the same packets from the same input frames (closing the loop between
     the FATE goldens and the library the GPU tests compare against);
  5. what travels to the GPU box: the small inputs (vsynth3 all variants, first frames
     of vsynth1/2 for the yuv420p variants) with their packet md5s, and the full
     per-packet md5 lists of all 21 tests (fate_full.json).

Usage: python tests/golden/make_fate_golden.py /tmp/ffbuild/ffmpeg
"""
import hashlib
import json
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpucodec as cc  # noqa: E402

FLAGS = "-flags +bitexact -sws_flags +accurate_rnd+bitexact -fflags +bitexact".split()
DEC_OPTS = "-threads 1 -idct simple".split() + FLAGS
ENC_OPTS = "-threads 1 -idct simple -dct fastint".split()

VARIANTS = {   # tests/fate/vcodec.mak:168-185
    "ffv1": dict(encopts="-slices 4", fmt="yuv420p", kw=dict(slices=4)),
    "ffv1-v0": dict(encopts="", fmt="yuv420p", kw=dict()),
    "ffv1-v3-yuv420p": dict(encopts="-level 3 -pix_fmt yuv420p", fmt="yuv420p", kw=dict(level=3)),
    "ffv1-v3-yuv422p10": dict(encopts="-level 3 -pix_fmt yuv422p10 -sws_flags neighbor+bitexact",
                              fmt="yuv422p10le", kw=dict(level=3)),
    "ffv1-v3-yuv444p16": dict(encopts="-level 3 -pix_fmt yuv444p16 -sws_flags neighbor+bitexact",
                              fmt="yuv444p16le", kw=dict(level=3)),
    "ffv1-v3-bgr0": dict(encopts="-level 3 -pix_fmt bgr0 -sws_flags neighbor+bitexact",
                         fmt="bgr0", kw=dict(level=3)),
    "ffv1-v3-rgb48": dict(encopts="-level 3 -pix_fmt rgb48 -strict -2 -sws_flags neighbor+bitexact",
                          fmt="rgb48le", kw=dict(level=3, strict=-2)),
}
SOURCES = {"vsynth1": (352, 288), "vsynth2": (352, 288), "vsynth3": (34, 34)}


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


def md5(b):
    return hashlib.md5(b).hexdigest()


def avi_extradata(avi):
    i = avi.find(b"strf")
    size = struct.unpack("<I", avi[i + 4:i + 8])[0]
    return avi[i + 8 + 40:i + 8 + size]


def split_frames(raw, fmt, w, h):
    geo = cc.plane_geometry(fmt, w, h)
    fsz = sum(bw * rows for bw, rows in geo)
    assert len(raw) % fsz == 0, (len(raw), fsz)
    frames = []
    for f in range(len(raw) // fsz):
        off = f * fsz
        planes = []
        for bw, rows in geo:
            planes.append(np.frombuffer(raw, np.uint8, bw * rows, off).reshape(rows, bw).copy())
            off += bw * rows
        frames.append(planes)
    return frames


def main():
    ffmpeg = sys.argv[1]
    tmp = "/tmp/fate_golden"
    os.makedirs(tmp, exist_ok=True)
    for tool in ("videogen", "rotozoom"):
        run(["gcc", "-O2", "-o", f"{tmp}/{tool}", f"{REF}/tests/{tool}.c", "-lm"])
    run([f"{tmp}/videogen", f"{tmp}/vsynth1.yuv"])
    run([f"{tmp}/rotozoom", f"{REF}/tests/reference.pnm", f"{tmp}/vsynth2.yuv"])
    run([f"{tmp}/videogen", f"{tmp}/vsynth3.yuv", "34", "34"])

    full = {}
    small = {}
    for src, (w, h) in SOURCES.items():
        for name, v in VARIANTS.items():
            test = f"{src}-{name}"
            ref_lines = open(f"{REF}/tests/ref/vsynth/{test}").read().split("\n")
            want_md5, want_size = ref_lines[0].split()[0], int(ref_lines[1].split()[0])
            avi_path = f"{tmp}/{test}.avi"
            run([ffmpeg, "-f", "rawvideo", "-s", f"{w}x{h}", "-pix_fmt", "yuv420p"] + DEC_OPTS +
                ["-i", f"{tmp}/{src}.yuv"] + ENC_OPTS + ["-c", "ffv1"] + v["encopts"].split() +
                FLAGS + ["-f", "avi", "-y", avi_path])
            avi = open(avi_path, "rb").read()
            got_md5, got_size = md5(avi), len(avi)
            assert (got_md5, got_size) == (want_md5, want_size), (test, got_md5, want_md5)
            fm = run([ffmpeg, "-i", avi_path, "-c", "copy", "-f", "framemd5", "-"]).stdout.decode()
            pkt_md5 = [l.split(",")[-1].strip() for l in fm.splitlines() if l and l[0] != "#"]
            extradata = avi_extradata(avi)
            raw = run([ffmpeg, "-i", avi_path, "-f", "rawvideo", "-pix_fmt", v["fmt"],
                       "-sws_flags", "neighbor+bitexact+accurate_rnd", "-"]).stdout
            frames = split_frames(raw, v["fmt"], w, h)
            assert len(frames) == len(pkt_md5) == 50, (len(frames), len(pkt_md5))
            # close the loop with the direct-gcc build of the reference codec
            enc = cc.Encoder("ref", w, h, v["fmt"], **v["kw"])
            if len(extradata) == len(enc.extradata) + 1 and extradata[-1] == 0:
                extradata = extradata[:-1]          # riff chunk padding of an odd-sized strf
            assert enc.extradata == extradata, test
            mine = [md5(enc.encode(f)) for f in frames]
            assert mine == pkt_md5, (test, "ref harness packets differ from FATE AVI packets")
            enc.close()
            full[test] = dict(avi_md5=got_md5, avi_size=got_size, width=w, height=h,
                              pix_fmt=v["fmt"], options=v["kw"], extradata=extradata.hex(),
                              packet_md5=pkt_md5,
                              input_md5=[md5(b"".join(p.tobytes() for p in f)) for f in frames])
            keep = 13 if src == "vsynth3" else (2 if v["fmt"] == "yuv420p" else 0)
            if keep:
                small[test] = np.stack([np.concatenate([p.reshape(-1) for p in f])
                                        for f in frames[:keep]])
            print("ok", test, got_md5, got_size, "packets", len(pkt_md5))
    json.dump(full, open(os.path.join(HERE, "fate_full.json"), "w"), indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "fate_inputs.npz"), **small)
    print("wrote", len(full), "tests,", len(small), "with inputs")


if __name__ == "__main__":
    main()
