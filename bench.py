#!/usr/bin/env python3
"""bench.py -- FFV1 encode & decode frames/s on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic 4K 10-bit pictures:
the batch is encoded (pictures -> packets) and the packets are decoded again
(packets -> pictures).  frames/s = pictures that went through BOTH directions per second.

  value  : inputs already resident in HBM (ffgpu_ffv1_encode_device / _decode_device),
           timed with CUDA events on the launching stream
  e2e    : the same work through the reference-facing C ABI with HOST buffers
           (send_frame/receive_packet, send_packet/receive_frame), pinned host memory,
           H2D + D2H inside the timed region, wall clock bracketed by synchronize()
  --impl reference : the reference's own CPU FFV1 (oracle/_ref, slice-threaded over all
           host threads), same metric / config, bounded sample per step

One process per GPU (torchrun for N > 1); pictures are partitioned across ranks
(rank r codes pictures r, r+N, ...) with no data-path collective: "scaling": "weak".
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "C2": dict(desc="4K 3840x2160 yuv420p10le, -slices 1023 (33x31, reference max), range coder "
                    "(forced by >8 bit), context 0, -g 1, slice CRC",
               w=3840, h=2160, fmt="yuv420p10le", opts=dict(slices=1023, gop_size=1), batch=192),
    # the other BASELINE configs: few, large slices -- the path's parallelism is slices x pictures
    # in flight, so these run with many pictures per launch group
    "C1": dict(desc="1080p yuv420p 8-bit, default 2x2 slices, Golomb-Rice, -g 1",
               w=1920, h=1080, fmt="yuv420p", opts=dict(gop_size=1), batch=768),
    "C3": dict(desc="4K bgr0 RCT, -coder range_tab -context 1, default 2x2 slices, -g 1",
               w=3840, h=2160, fmt="bgr0", opts=dict(coder=2, context=1, gop_size=1), batch=144),
    "C4": dict(desc="4K yuv444p16le, default 2x2 slices, range coder, -g 1: DECODE ONLY of a "
                    "reference-compatible stream",
               w=3840, h=2160, fmt="yuv444p16le", opts=dict(gop_size=1), batch=144, decode_only=True),
    "C5": dict(desc="8K 7680x4320 yuv420p10le, default 3x3 slices, range coder, -g 1",
               w=7680, h=4320, fmt="yuv420p10le", opts=dict(gop_size=1), batch=96),
    "few": dict(desc="1080p yuv420p10le, default 3x3 slices, range coder (debug: one slice per warp)",
                w=1920, h=1080, fmt="yuv420p10le", opts=dict(gop_size=1), batch=1),
    "small": dict(desc="640x360 yuv420p10le 60 slices (debug)", w=640, h=360, fmt="yuv420p10le",
                  opts=dict(slices=60, gop_size=1), batch=16),
}

# the lavfi sources BASELINE.json names, as filter graphs of the reference's own
# ffmpeg (oracle/_ref/ffmpeg, built by oracle/build_ffmpeg.sh); tests/synth.py stands in, and
# says so in `config.source`, where that binary is missing
LAVFI = {
    "testsrc2": "testsrc2=s={w}x{h}:r=25",
    "mandelbrot": "mandelbrot=s={w}x{h}:r=25",
    "noise": "testsrc2=s={w}x{h}:r=25,noise=alls=100:allf=t+u:all_seed=1234",
}
FFMPEG = os.path.join(ROOT, "oracle", "_ref", "ffmpeg")


def md5(b):
    return hashlib.md5(b).hexdigest()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def halt(self):
        """stop sampling, keep the samples (a later start() adds to them)"""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.t.join(timeout=2)
            self.proc = None

    def stop(self):
        if not self.proc and not self.lines:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.halt()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def lavfi_frames(wl, source, count, first):
    """`count` consecutive frames of a lavfi source, starting at frame `first`, converted to
    the workload's pixel format by the reference's ffmpeg: list of lists of 2-D uint8 planes"""
    import cpucodec as cc
    w, h, fmt = wl["w"], wl["h"], wl["fmt"]
    graph = LAVFI[source].format(w=w, h=h) + ",trim=start_frame=%d:end_frame=%d" % (first, first + count)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "ffmpeg_ffv2_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([FFMPEG, "-hide_banner", "-loglevel", "error", "-nostdin", "-f", "lavfi", "-i", graph,
                        "-frames:v", str(count), "-pix_fmt", fmt, "-f", "rawvideo", "-"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=900)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.decode(errors="replace")[-500:])
    geo = cc.plane_geometry(fmt, w, h)
    per = sum(bw * rows for bw, rows in geo)
    raw = np.frombuffer(r.stdout, np.uint8)
    if raw.size != per * count:
        raise RuntimeError("lavfi source gave %d bytes, expected %d" % (raw.size, per * count))
    out = []
    for i in range(count):
        planes, off = [], i * per
        for bw, rows in geo:
            planes.append(np.ascontiguousarray(raw[off:off + bw * rows].reshape(rows, bw)))
            off += bw * rows
        out.append(planes)
    return out


def make_frames(wl, count, part=0, source="testsrc2"):
    """(pictures, description of where they came from).  Rank r of N takes the frames
    r*count .. r*count+count-1 of the source (its own part of the stream)."""
    import synth
    first = part * count
    if os.path.exists(FFMPEG) and source in LAVFI and not make_frames.force_synth:
        try:
            fr = lavfi_frames(wl, source, count, first)
            digest = hashlib.md5(b"".join(p.tobytes() for p in fr[0])).hexdigest()
            return fr, "lavfi %s (reference ffmpeg), frames %d..%d, md5(frame %d)=%s" % (
                LAVFI[source].format(w=wl["w"], h=wl["h"]), first, first + count - 1, first, digest)
        except Exception as e:               # noqa: BLE001
            print("lavfi source unavailable (%s): falling back to tests/synth.py" % e, file=sys.stderr)
    gen = {"testsrc2": synth.testsrc2_like, "mandelbrot": synth.mandelbrot, "noise": synth.noise}[source]
    return [gen(wl["fmt"], wl["w"], wl["h"], first + i) for i in range(count)], \
        "%s-like (tests/synth.py; oracle/_ref/ffmpeg not available)" % source


def enc_samples(fmt, w, h):
    """coded samples per picture"""
    import synth
    f = synth.fmt_info(fmt)
    cw, ch = -(-w >> f["hs"]), -(-h >> f["vs"])
    if f["layout"] in ("planar", "ya8"):
        n = w * h + (2 * cw * ch if f["chroma"] and f["layout"] == "planar" else 0)
        return n + (w * h if f["alpha"] else 0)
    return (3 + f["alpha"]) * w * h


make_frames.force_synth = False


def metric_name(name, wl):
    if name == "C2":
        return "ffv1_encode_decode_fps_4k10"
    return "ffv1_%s_fps_%s" % ("decode" if wl.get("decode_only") else "encode_decode", name)


def cpu_codec_kind():
    import cpucodec as cc
    return ("ref", "reference") if cc.available("ref") else ("oracle", "port")


def run_cpu(wl, frames, threads, budget_s, max_frames):
    """the reference's CPU encoder+decoder (slice threads), bounded sample"""
    import cpucodec as cc
    which, kind = cpu_codec_kind()
    enc = cc.Encoder(which, wl["w"], wl["h"], wl["fmt"], threads=threads, **wl["opts"])
    dec = cc.Decoder(which, wl["w"], wl["h"], enc.extradata, threads=threads)
    enc.encode(frames[0])                       # first-touch of the packet buffer, untimed
    enc.close()
    enc = cc.Encoder(which, wl["w"], wl["h"], wl["fmt"], threads=threads, **wl["opts"])
    enc.encode(frames[0])
    pkts, n = [], 0
    t0 = time.perf_counter()
    while n < max_frames and (n < 2 or time.perf_counter() - t0 < budget_s / 2):
        pkts.append(enc.encode(frames[n % len(frames)]))
        n += 1
    t_enc = time.perf_counter() - t0
    dec.decode(pkts[0], copy=False)
    t0 = time.perf_counter()
    for p in pkts:
        dec.decode(p, copy=False)
    t_dec = time.perf_counter() - t0
    return dict(kind=kind, frames=n, t_enc=t_enc, t_dec=t_dec,
                enc_fps=n / t_enc, dec_fps=n / t_dec, fps=n / (t_enc + t_dec))


def picture_bytes(wl):
    import cpucodec as cc
    return sum(bw * rows for bw, rows in cc.plane_geometry(wl["fmt"], wl["w"], wl["h"]))


def distinct_pictures(wl, B):
    """how many different pictures a step cycles through"""
    return min(8 if picture_bytes(wl) < (64 << 20) else 4, B)


def run_config(args, wl, src_desc):
    """`config` of the JSON line: the workload both arms are quoted on, key for key the same in
    the GPU arm and in --impl reference (whose bounded sample of it is described in its
    `cpu_baseline.sample`)"""
    B = args.batch or wl["batch"]
    return {"workload": args.workload + ": " + wl["desc"], "frames_per_step_per_gpu": B,
            "distinct_pictures": distinct_pictures(wl, B), "source": src_desc,
            "l2": "inputs larger than L2 (%.0f MB per step)" % (B * picture_bytes(wl) / 1e6),
            "partition": "pictures round-robin over ranks, no collective"}


def run_cpu_instances(wl, frames, instances, threads, per_instance):
    """`instances` independent encoder+decoder pairs of the reference codec, each with
    `threads` slice threads, side by side (the calls release the GIL): frame-level on top of
    slice-level parallelism, what N ffmpeg processes on one box give.  Every instance codes
    `per_instance` pictures; the phases (all encode, then all decode) are timed as a whole."""
    import threading
    import cpucodec as cc
    which, kind = cpu_codec_kind()
    encs, decs = [], []
    for _ in range(instances):
        warm = cc.Encoder(which, wl["w"], wl["h"], wl["fmt"], threads=threads, **wl["opts"])
        warm.encode(frames[0])                  # first-touch of the packet buffer, untimed
        warm.close()
        e = cc.Encoder(which, wl["w"], wl["h"], wl["fmt"], threads=threads, **wl["opts"])
        e.encode(frames[0])
        encs.append(e)
        decs.append(cc.Decoder(which, wl["w"], wl["h"], e.extradata, threads=threads))
    pkts = [[] for _ in range(instances)]
    errors = []

    def phase(fn):
        ts = [threading.Thread(target=fn, args=(i,)) for i in range(instances)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    def enc_fn(i):
        try:
            for k in range(per_instance):
                pkts[i].append(encs[i].encode(frames[(i + k) % len(frames)]))
        except Exception as e:                  # noqa: BLE001
            errors.append(e)

    def dec_fn(i):
        try:
            for pkt in pkts[i]:
                decs[i].decode(pkt, copy=False)
        except Exception as e:                  # noqa: BLE001
            errors.append(e)

    for i in range(instances):
        decs[i].decode(encs[i].encode(frames[0]), copy=False)        # decoder warm-up, untimed
    t_enc = phase(enc_fn)
    t_dec = phase(dec_fn)
    for e in encs:
        e.close()
    if errors:
        raise errors[0]
    n = instances * per_instance
    return dict(kind=kind, frames=n, t_enc=t_enc, t_dec=t_dec,
                enc_fps=n / t_enc, dec_fps=n / t_dec, fps=n / (t_enc + t_dec))


def best_cpu_split(wl, frames, dec_only):
    """untimed calibration: which split of the box's threads into codec instances x slice
    threads codes this workload fastest -> ((instances, threads), {split: pictures/s})"""
    cores = os.cpu_count() or 1
    splits, calib = [(1, cores)], {}
    if picture_bytes(wl) <= (64 << 20):         # 8K pictures: one instance (memory, time)
        splits += [(n, max(1, cores // n)) for n in (2, 4) if cores >= 2 * n]
    best = splits[0]
    if len(splits) > 1:
        for sp in splits:
            try:
                r = run_cpu_instances(wl, frames, sp[0], sp[1], max(2, 4 // sp[0]))
                calib[sp] = r["dec_fps"] if dec_only else r["fps"]
            except Exception as e:              # noqa: BLE001
                print("cpu baseline: %d instances failed in calibration (%s)" % (sp[0], e), file=sys.stderr)
        if calib:
            best = max(calib, key=calib.get)
    return best, calib


def reference_arm(args, wl, rank, world, real_stdout):
    """--impl reference: rank 0 alone times the reference CPU implementation, on every host
    thread.  The reference's own parallelism is slice threads (pthread_slice.c); independent
    codec instances side by side add frame-level parallelism on top.  An untimed calibration
    picks the split of the box's threads (1, 2 or 4 instances) that codes the workload
    fastest, and the timed steps run that split."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames, src_desc = make_frames(wl, distinct_pictures(wl, args.batch or wl["batch"]), 0, args.source)
    which, kind = cpu_codec_kind()
    dec_only = bool(wl.get("decode_only"))
    sample = 8

    best, calib = best_cpu_split(wl, frames, dec_only)
    inst, threads = best
    per = max(1, sample // inst)
    sample = per * inst
    res = []
    for step in range(args.warmup + args.steps):
        r = run_cpu_instances(wl, frames, inst, threads, per)
        if step >= args.warmup:
            res.append(r)
    tot_frames = sum(r["frames"] for r in res)
    tot_t = sum((0.0 if dec_only else r["t_enc"]) + r["t_dec"] for r in res)
    fps = tot_frames / tot_t
    out = {
        "impl": "reference", "metric": metric_name(args.workload, wl), "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(len(res), 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": run_config(args, wl, src_desc), "frames_per_step": sample,
        "encode_fps": tot_frames / sum(r["t_enc"] for r in res),
        "decode_fps": tot_frames / sum(r["t_dec"] for r in res),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "instances": inst, "slice_threads_per_instance": threads,
                         "calibration_fps": {"%dx%d" % k: round(v, 2) for k, v in calib.items()},
                         "sample": "%d pictures %s per step, %d steps; %d codec instance(s) x %d slice threads "
                                   "(the reference's own ffv1enc.c/ffv1dec.c, oracle/_ref), the fastest split of "
                                   "the box's %d threads in an untimed calibration" % (
                             sample, "decoded" if dec_only else "encoded+decoded", len(res), inst, threads, cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), file=real_stdout, flush=True)


def bind_to_gpu_numa(gpu_index):
    """run this process (and its first-touch allocations: pinned buffers, staging) on the host
    cores next to the GPU.  With all ranks on the default CPU set, eight ranks' pinned buffers
    land on one socket and every DMA of the far GPUs crosses the socket interconnect."""
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id",
                              "--format=csv,noheader"], stdout=subprocess.PIPE, text=True,
                             timeout=20).stdout.strip().lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]                      # 00000000:1b:00.0 -> 0000:1b:00.0
        base = "/sys/bus/pci/devices/" + bdf
        node = open(base + "/numa_node").read().strip()
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        bind_to_gpu_numa.all_cpus = allowed
        use = (cpus & allowed) or allowed
        os.sched_setaffinity(0, use)
        return {"gpu": gpu_index, "pci": bdf, "numa_node": int(node), "cpus": len(use),
                "bound": bool(cpus & allowed)}
    except Exception as e:                     # noqa: BLE001
        return {"gpu": gpu_index, "bound": False, "why": repr(e)[:100]}


def main():
    # the contract is ONE JSON line on stdout: keep the real stdout for it and send whatever
    # libraries print to fd 1 (e.g. the NCCL version banner) to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--source", default="testsrc2", choices=sorted(LAVFI),
                    help="lavfi source of the pictures (BASELINE.json: testsrc2, mandelbrot, noise)")
    ap.add_argument("--batch", type=int, default=0, help="pictures per step and GPU (0 = default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--route", type=int, default=0,
                    help="e2e leg through ONE routing handle over this many GPUs of this process "
                         "(ffgpu_enc_options.ndevices), instead of one rank per GPU")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-memory e2e leg")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to the GPU's NUMA node")
    ap.add_argument("--e2e-groups", type=int, default=4, help="launch groups a step's pictures are split into")
    ap.add_argument("--e2e-depth", type=int, default=6, help="launch groups in flight (e2e leg)")
    ap.add_argument("--e2e-host", choices=("c", "python"), default="c",
                    help="host loop of the e2e measurement: tools/e2e_driver.c or Python threads")
    ap.add_argument("--e2e-repeat", type=int, default=6,
                    help="the e2e leg streams the step's pictures this many times back to back, so "
                         "pipeline fill and drain are amortised like in a long transcode")
    ap.add_argument("--synth", action="store_true",
                    help="pictures from tests/synth.py even where the reference ffmpeg is available")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    make_frames.force_synth = args.synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, wl, rank, world, real_stdout)
        return 0

    numa = bind_to_gpu_numa(local) if not args.no_numa else {"bound": False, "why": "--no-numa"}
    import torch
    import torch.distributed as dist
    import ffmpeg_ffv2_b200 as F
    import cpucodec as cc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the FFV1 pixel path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    w, h, fmt, opts = wl["w"], wl["h"], wl["fmt"], wl["opts"]
    frame_bytes, planes = F.frame_layout(fmt, w, h)
    B = args.batch or wl["batch"]
    distinct = distinct_pictures(wl, B)
    dec_only = bool(wl.get("decode_only"))

    # ---- synthetic input: `distinct` pictures (this rank's share of the stream), cycled ----
    srcs, src_desc = make_frames(wl, distinct, rank, args.source)
    # pinned host copies of the distinct pictures only (picture i of a step is srcs[i % distinct]):
    # keeps the pinned footprint small when 8 ranks share one host
    h_frames = torch.empty((distinct, frame_bytes), dtype=torch.uint8).pin_memory()
    hf = h_frames.numpy()
    hf[:] = 0
    for i in range(distinct):
        for (off, pitch, rows, rb), a in zip(planes, srcs[i]):
            hf[i, off:off + pitch * rows].reshape(rows, pitch)[:, :rb] = a
    host_planes = [[hf[i % distinct, off:off + pitch * rows].reshape(rows, pitch)[:, :rb]
                    for (off, pitch, rows, rb) in planes] for i in range(B)]
    d_frames = h_frames.cuda(non_blocking=False)[torch.arange(B, device="cuda") % distinct].contiguous()
    d_out = torch.zeros((B, frame_bytes), dtype=torch.uint8, device="cuda")
    h_out = torch.empty((B, frame_bytes), dtype=torch.uint8).pin_memory()
    ho = h_out.numpy()
    ho[:] = 0

    # the device-resident path uses one launch group at a time: depth 1 keeps memory for big batches
    enc = F.FFV1Encoder(w, h, fmt, device=local, max_batch=B, pipeline_depth=1, **opts)
    dec = F.FFV1Decoder(w, h, enc.extradata, device=local, max_batch=B, pipeline_depth=1)
    # a real (non-default) stream: the C ABI treats a NULL stream as "use the handle's own"
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    # ---- parity gate: packets byte-identical to the CPU reference, pictures restored ----
    which, kind = cpu_codec_kind()
    enc.encode_device(d_frames.data_ptr(), B, stream)
    torch.cuda.synchronize()
    pkts = [enc.device_fetch(i) for i in range(B)]
    cpu_enc = cc.Encoder(which, w, h, fmt, threads=min(os.cpu_count() or 1, 64), **opts)
    for i in range(min(2, distinct)):
        want = cpu_enc.encode(srcs[i])
        if pkts[i] != want:
            raise SystemExit("PARITY FAILURE: GPU packet %d differs from the CPU %s" % (i, kind))
    cpu_enc.close()
    dec.decode_device(pkts, d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    def same_pictures(a, b):
        if fmt == "bgr0":                     # the X byte of bgr0 is not coded (ffv1enc.c: transparency 0)
            return torch.equal(a.view(torch.int32) & 0x00FFFFFF, b.view(torch.int32) & 0x00FFFFFF)
        return torch.equal(a, b)

    if not same_pictures(d_out, d_frames):
        raise SystemExit("PARITY FAILURE: decoded pictures differ from the input")
    pkt_bytes = sum(len(p) for p in pkts)
    raw_bytes = sum(rb * rows for (_o, _p, rows, rb) in planes)
    decisions_total, decisions_heaviest = enc.decisions()

    # ---- value: inputs resident in HBM, CUDA events on the launching stream ----
    enc.profile(True)
    dec.profile(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    kern = {}
    t_enc_ms = t_dec_ms = 0.0
    launches0 = enc.launches + dec.launches
    sampler = ClockSampler(local)
    for step in range(args.warmup + args.steps):
        if step == args.warmup:
            barrier()
            sampler.start()
            launches0 = enc.launches + dec.launches
            wall0 = time.perf_counter()
        ev[0].record()
        if not dec_only:
            enc.encode_device(d_frames.data_ptr(), B, stream)
        ev[1].record()
        dec.decode_device(pkts, d_out.data_ptr(), stream)
        ev[2].record()
        torch.cuda.synchronize()
        if step >= args.warmup:
            t_enc_ms += ev[0].elapsed_time(ev[1])
            t_dec_ms += ev[1].elapsed_time(ev[2])
            for k, v in ([] if dec_only else list(enc.kernel_ms().items())) + list(dec.kernel_ms().items()):
                kern[k] = kern.get(k, 0.0) + v
    barrier()
    wall = time.perf_counter() - wall0
    sampler.halt()                     # resumed around the timed e2e steps below
    gpu_launches = enc.launches + dec.launches - launches0
    enc.profile(False)
    dec.profile(False)
    K = args.steps
    dev_time = max_over_ranks(max((t_enc_ms + t_dec_ms) / 1e3, 0.0))
    total_frames = B * K * world
    value = total_frames / dev_time
    enc_fps = None if dec_only else B * K * world / max_over_ranks(t_enc_ms / 1e3)
    dec_fps = B * K * world / max_over_ranks(t_dec_ms / 1e3)
    for k in kern:
        kern[k] /= K

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peaks()
    dom = max(("code", "decode"), key=lambda k: kern.get(k, 0.0))
    alg_bytes = (raw_bytes + pkt_bytes / B) * B           # raw + packet bytes per picture
    achieved = alg_bytes / (kern[dom] / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_code_range" if dom == "code" else "k_decode",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kern[dom]}

    # ---- the serial side of the path: binary decisions (SURVEY 8d "cycles per bin") ----
    nsamples = enc_samples(fmt, w, h) * B
    decisions = None
    if decisions_total:
        mhz = 1965.0
        slots = 148 * 4 * mhz * 1e3            # warp-instruction issue slots per millisecond
        decisions = {
            "per_sample": decisions_total / nsamples,
            "per_step": int(decisions_total),
            "heaviest_slice": int(decisions_heaviest),
            "slices_per_step": int(B * enc.info["num_h_slices"] * enc.info["num_v_slices"]),
            # throughput view: issue slots the whole GPU spends per decision (1 slot = one warp
            # instruction on one of the 148 x 4 schedulers at the maximum SM clock)
            "code_issue_slots_per_decision": kern.get("code", 0.0) * slots / decisions_total,
            "decode_issue_slots_per_decision": kern.get("decode", 0.0) * slots / decisions_total,
            # latency view: the launch cannot be shorter than its heaviest slice, a strictly
            # serial chain of decisions on one lane
            "code_cycles_per_decision_if_bound_by_heaviest": kern.get("code", 0.0) * mhz * 1e3 / max(decisions_heaviest, 1),
            "decode_cycles_per_decision_if_bound_by_heaviest": kern.get("decode", 0.0) * mhz * 1e3 / max(decisions_heaviest, 1),
        }

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e and not dec_only:
        # smaller launch groups, more of them in flight: H2D, kernels and D2H overlap
        vb = max(B // args.e2e_groups, 1)
        enc.close()
        dec.close()
        # --route N: ONE process, the product's own round-robin over N GPUs (ndevices in the
        # options of the C ABI) instead of one torchrun rank per GPU
        routed = tuple(range(args.route)) if args.route > 1 and world == 1 else ()
        enc = F.FFV1Encoder(w, h, fmt, device=local, max_batch=vb, pipeline_depth=args.e2e_depth,
                            devices=routed, **opts)
        dec = F.FFV1Decoder(w, h, enc.extradata, device=local, max_batch=vb, pipeline_depth=args.e2e_depth,
                            devices=routed)

        # The host loop is C (tools/e2e_driver.c): an encoder thread and a decoder thread on the
        # public C ABI, like the codec threads of a transcoder.  Packets go to the decoder as
        # soon as the encoder returns them, so H2D of pictures, kernels of both directions and
        # D2H of decoded pictures overlap.  --e2e-host python keeps the same loop in Python.
        import ctypes as C
        import queue

        class E2ERun(C.Structure):
            _fields_ = [("enc", C.c_void_p), ("dec", C.c_void_p), ("nframes", C.c_int), ("nsrc", C.c_int),
                        ("src", C.POINTER(F.codec.Picture)), ("ndst", C.c_int),
                        ("dst", C.POINTER(F.codec.PictureOut)), ("timeout_s", C.c_double),
                        ("pkt", C.POINTER(C.POINTER(C.c_uint8))), ("pkt_size", C.POINTER(C.c_size_t)),
                        ("decoded", C.c_int), ("damaged", C.c_int), ("error", C.c_int),
                        ("message", C.c_char * 256), ("produced", C.c_int), ("enc_finished", C.c_int),
                        ("failed", C.c_int), ("t0", C.c_double)]

        def e2e_step_c(src=None, dst=None, repeat=None):
            NE = B * (repeat or args.e2e_repeat)
            run = E2ERun()
            run.enc, run.dec = enc.h, dec.h
            run.nframes, run.nsrc, run.ndst = NE, B, B
            run.src, run.dst = src or c_src, dst or c_dst
            run.timeout_s = 120.0
            pk = (C.POINTER(C.c_uint8) * NE)()
            sz = (C.c_size_t * NE)()
            run.pkt, run.pkt_size = pk, sz
            r = e2e_lib.ffgpu_e2e_run(C.byref(run))
            if r < 0:
                raise SystemExit("e2e pipeline failed: %s" % run.message.decode(errors="replace"))

            def collect():                       # untimed: packets of the first and last pass
                out = [C.string_at(pk[i], sz[i]) for i in list(range(B)) + list(range(NE - B, NE))]
                e2e_lib.ffgpu_e2e_free_packets(C.byref(run))
                return out
            return collect, run.decoded, NE


        def e2e_step():
            """encoder and decoder run concurrently on two host threads, like the codec threads
            of a transcoder: packets go to the decoder as soon as the encoder returns them, so
            H2D of pictures, kernels of both directions and D2H of decoded pictures overlap and
            a blocking wait in one codec does not stall the other (ctypes releases the GIL)"""
            NE = B * args.e2e_repeat
            q = queue.Queue()
            out_pk = []
            state = {"decoded": 0, "err": None}
            deadline = time.perf_counter() + 120.0

            def enc_thread():
                try:
                    sent, done = 0, False
                    while not done:
                        if time.perf_counter() > deadline:
                            raise RuntimeError("encoder stalled at %d" % sent)
                        if sent < NE and enc.send_frame(host_planes[sent % B], pts=sent):
                            sent += 1
                            if sent == NE:
                                enc.send_frame(None)
                        while True:
                            r = enc.receive_packet()
                            if r == F.EOF:
                                done = True
                                break
                            if r is None:
                                break
                            out_pk.append(r[0])
                            q.put(r[0])
                except Exception as e:           # noqa: BLE001
                    state["err"] = e
                q.put(None)

            def dec_thread():
                try:
                    sent, got, done, eos = 0, 0, False, False
                    pend = None
                    while not done:
                        if time.perf_counter() > deadline:
                            raise RuntimeError("decoder stalled at %d/%d" % (got, sent))
                        if pend is None and not eos:
                            try:
                                pend = q.get(timeout=0.0005) if sent > got else q.get(timeout=0.05)
                                if pend is None:
                                    eos = True
                                    dec.send_packet(None)
                            except queue.Empty:
                                pend = None
                        if pend is not None and dec.send_packet(pend, pts=sent, dst=dsts[sent % B]):
                            sent += 1
                            pend = None
                        while True:
                            r = dec.receive_frame()
                            if r == F.EOF:
                                done = True
                                break
                            if r is None:
                                break
                            got += 1
                    state["decoded"] = got
                except Exception as e:           # noqa: BLE001
                    state["err"] = e

            te = threading.Thread(target=enc_thread)
            td = threading.Thread(target=dec_thread)
            te.start()
            td.start()
            te.join()
            td.join()
            if state["err"] is not None:
                raise SystemExit("e2e pipeline failed: %r" % (state["err"],))
            return out_pk[:B] + out_pk[-B:], state["decoded"], len(out_pk)

        dsts = []
        for i in range(B):
            po = F.codec.PictureOut()
            arrs = []
            for k, (off, pitch, rows, rb) in enumerate(planes):
                a = ho[i, off:off + pitch * rows].reshape(rows, pitch)[:, :rb]
                po.data[k] = a.ctypes.data
                po.linesize[k] = pitch
                arrs.append(a)
            dsts.append((po, arrs))
        c_dst = (F.codec.PictureOut * B)(*[po for po, _a in dsts])
        c_src = (F.codec.Picture * B)()
        for i in range(B):
            for k, a in enumerate(host_planes[i]):
                c_src[i].data[k] = a.ctypes.data
                c_src[i].linesize[k] = a.strides[0]
            c_src[i].sar_num, c_src[i].sar_den = 0, 1
        use_c = args.e2e_host == "c"
        if use_c:
            from ffmpeg_ffv2_b200 import build as _b
            e2e_lib = C.CDLL(_b.build_e2e_driver())
            e2e_lib.ffgpu_e2e_run.restype = C.c_int
            e2e_lib.ffgpu_e2e_run.argtypes = [C.POINTER(E2ERun)]
            e2e_lib.ffgpu_e2e_free_packets.restype = None
            e2e_lib.ffgpu_e2e_free_packets.argtypes = [C.POINTER(E2ERun)]
        e2e_t = 0.0
        e2e_launch0 = 0
        for step in range(args.warmup + args.steps):
            if step == args.warmup:
                e2e_launch0 = enc.launches + dec.launches
                sampler.start()
            barrier()
            t0 = time.perf_counter()
            ends_pk, done, npk = e2e_step_c() if use_c else e2e_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            if step >= args.warmup:
                e2e_t += max_over_ranks(dt)
            if callable(ends_pk):
                ends_pk = ends_pk()
            assert done == B * args.e2e_repeat and npk == done
        def same_host(a, b):
            if fmt == "bgr0":
                return np.array_equal(a.view(np.uint32) & 0x00FFFFFF, b.view(np.uint32) & 0x00FFFFFF)
            return np.array_equal(a, b)

        if ends_pk[:B] != pkts or ends_pk[-B:] != pkts or not all(
                same_host(ho[i], hf[i % distinct]) for i in range(B)):
            raise SystemExit("PARITY FAILURE: e2e path differs from the device path / the input")
        R = args.e2e_repeat
        e2e = {"value": B * R * K * world / e2e_t, "unit": "frames/s",
               "frames_per_step_per_gpu": B * R,
               "h2d_bytes_per_step": int(R * (B * raw_bytes + pkt_bytes)),
               "d2h_bytes_per_step": int(R * (pkt_bytes + B * raw_bytes)),
               "ms_per_step": 1e3 * e2e_t / K,
               "api": "ffgpu_ffv1_encode_send_frame/receive_packet + "
                      "ffgpu_ffv1_decode_send_packet/receive_frame, pinned host buffers",
               "host_loop": "tools/e2e_driver.c (2 threads)" if use_c else "python (2 threads)",
               "frames_per_launch_group": vb, "launch_groups_in_flight": args.e2e_depth,
               "routed_devices": len(routed) or 1,
               "gpu_launches": int(enc.launches + dec.launches - e2e_launch0)}

        # ---- the same through ordinary (pageable) memory, what AVFrames are: the library
        # stages the pictures through its own pinned buffers with a pool of copy threads ----
        if use_c and not args.no_pageable:
            pg_src = [[np.array(a, copy=True) for a in host_planes[i]] for i in range(distinct)]
            pg_out = np.zeros((B, frame_bytes), np.uint8)
            p_src = (F.codec.Picture * B)()
            p_dst = (F.codec.PictureOut * B)()
            for i in range(B):
                for k, a in enumerate(pg_src[i % distinct]):
                    p_src[i].data[k] = a.ctypes.data
                    p_src[i].linesize[k] = a.strides[0]
                p_src[i].sar_num, p_src[i].sar_den = 0, 1
                for k, (off, pitch, rows, rb) in enumerate(planes):
                    p_dst[i].data[k] = pg_out[i, off:].ctypes.data
                    p_dst[i].linesize[k] = pitch
            R2 = args.e2e_repeat if args.e2e_repeat <= 3 else args.e2e_repeat // 2
            t_pg = 0.0
            for step in range(2):
                barrier()
                t0 = time.perf_counter()
                ends, done, npk = e2e_step_c(p_src, p_dst, R2)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                barrier()
                if step:
                    t_pg = max_over_ranks(dt)
                ends = ends()
                assert done == B * R2
            if ends[:B] != pkts or not all(same_host(pg_out[i], hf[i % distinct]) for i in range(B)):
                raise SystemExit("PARITY FAILURE: pageable e2e path differs from the device path / the input")
            e2e["pageable"] = {"value": B * R2 * world / t_pg, "unit": "frames/s",
                               "buffers": "malloc'd (numpy) pictures in and out, staged by the library "
                                          "(FFGPU_COPY_THREADS=%s)" % os.environ.get("FFGPU_COPY_THREADS", "8")}
            del pg_src, pg_out

    clocks = sampler.stop()            # samples of both timed regions (device-resident + e2e)

    # ---- PCIe context: plain pinned copies of 1 GiB, both directions ----
    pcie = None
    try:
        nb = 1 << 30
        hp = torch.empty(nb, dtype=torch.uint8).pin_memory()
        dp = torch.empty(nb, dtype=torch.uint8, device="cuda")
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        dp.copy_(hp, non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        dp.copy_(hp, non_blocking=True)
        e1.record()
        hp.copy_(dp, non_blocking=True)
        e2.record()
        torch.cuda.synchronize()
        pcie = {"h2d_GBps": nb / e0.elapsed_time(e1) / 1e6, "d2h_GBps": nb / e1.elapsed_time(e2) / 1e6}
        # both directions at once (what the e2e pipeline asks of the link and of host DRAM)
        hq = torch.empty(nb, dtype=torch.uint8).pin_memory()
        dq = torch.empty(nb, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        d0, d1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        d0.record()
        s1.wait_event(d0)
        s2.wait_event(d0)
        with torch.cuda.stream(s1):
            for _ in range(2):
                dp.copy_(hp, non_blocking=True)
        with torch.cuda.stream(s2):
            for _ in range(2):
                hq.copy_(dq, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        d1.record()
        torch.cuda.synchronize()
        pcie["duplex_GBps_each_way"] = 2 * nb / d0.elapsed_time(d1) / 1e6
        # the same duplex copies on ALL ranks at once: what the host fabric (PCIe roots, host
        # DRAM, the socket interconnect) gives the whole job -- the ceiling of e2e at N > 1
        barrier()
        d0.record()
        s1.wait_event(d0)
        s2.wait_event(d0)
        with torch.cuda.stream(s1):
            for _ in range(3):
                dp.copy_(hp, non_blocking=True)
        with torch.cuda.stream(s2):
            for _ in range(3):
                hq.copy_(dq, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        d1.record()
        torch.cuda.synchronize()
        mine = 3 * nb / d0.elapsed_time(d1) / 1e6
        pcie["all_ranks_duplex_GBps_each_way_this_rank"] = mine
        pcie["all_ranks_duplex_GBps_each_way_aggregate"] = sum_over_ranks(mine)
        if e2e:
            # bytes the e2e leg moved each way per second, against that ceiling
            e2e_GBps = e2e["h2d_bytes_per_step"] * world / (e2e["ms_per_step"] / 1e3) / 1e9
            e2e["host_link_GBps_each_way"] = e2e_GBps
            e2e["fraction_of_all_ranks_duplex_probe"] = e2e_GBps / pcie["all_ranks_duplex_GBps_each_way_aggregate"]
        del hp, dp, hq, dq
    except Exception as ex:                # noqa: BLE001
        print("pcie probe failed: %r" % (ex,), file=sys.stderr)
        if world > 1:
            raise
        pcie = None

    # ---- CPU baseline beside it (rank 0, N == 1): the reference's slice-threaded CPU codec ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if getattr(bind_to_gpu_numa, "all_cpus", None):      # the CPU baseline gets every host core
            os.sched_setaffinity(0, bind_to_gpu_numa.all_cpus)
        threads = os.cpu_count() or 1
        inst, calib = 1, {}
        try:
            # the same choice as `--impl reference`: the fastest split of the host threads into
            # codec instances x slice threads, then about 12 s of work on it
            (inst, per_thr), calib = best_cpu_split(wl, srcs, dec_only)
            rate = calib.get((inst, per_thr), 0.0)
            per = int(min(400, max(4, rate * 12.0)) // inst) or 1
            r = run_cpu_instances(wl, srcs, inst, per_thr, per)
        except Exception as e:                   # noqa: BLE001
            print("cpu baseline: instance split failed (%s): one instance, slice threads only" % e, file=sys.stderr)
            inst, per_thr = 1, threads
            r = run_cpu(wl, srcs, threads, 16.0, 400)
        cpu = {"value": r["dec_fps"] if dec_only else r["fps"], "unit": "frames/s", "cores": threads,
               "kind": r["kind"], "encode_fps": r["enc_fps"], "decode_fps": r["dec_fps"],
               "instances": inst, "slice_threads_per_instance": per_thr,
               "calibration_fps": {"%dx%d" % k: round(v, 2) for k, v in calib.items()},
               "sample": "%d pictures encoded+decoded once (%.1f s), %d codec instance(s) x %d slice threads%s" % (
                   r["frames"], r["t_enc"] + r["t_dec"], inst, per_thr,
                   "; value = decode only" if dec_only else "")}

    out = {
        "metric": metric_name(args.workload, wl), "value": value, "unit": "frames/s",
        "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_time / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": run_config(args, wl, src_desc),
        "encode_fps": enc_fps, "decode_fps": dec_fps,
        "pixel_GBps": value * raw_bytes / 1e9,
        "packet_bytes_per_picture": pkt_bytes / B, "raw_bytes_per_picture": raw_bytes,
        "kernel_ms_per_step": kern, "wall_ms_per_step": 1e3 * wall / K,
        "decisions": decisions,
        "roofline": roofline, "gpu_launches": int(gpu_launches), "clocks": clocks,
        "e2e": e2e, "cpu_baseline": cpu, "pcie": pcie, "numa": numa,
    }
    if rank == 0:
        print(json.dumps(out), file=real_stdout, flush=True)
    enc.close()
    dec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
