"""The N > 1 path on CPU: two processes over gloo partition a stream round-robin
(ffmpeg_ffv2_b200.partition), each codes its share (with the oracle standing in for the GPU
codec), rank 0 gathers the packets and puts them back in presentation order; the result must
equal the single-process stream.  Also checks the max-over-ranks timing reduction bench.py
uses."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))

W, H, FMT, NFRAMES = 64, 48, "yuv420p10le", 11
OPTS = dict(slices=4, gop_size=1)


def _encode_share(indices):
    import cpucodec as cc
    import synth
    enc = cc.Encoder("oracle", W, H, FMT, **OPTS)
    return [(i, enc.encode(synth.testsrc2_like(FMT, W, H, i))) for i in indices]


def _worker(rank, world, port, q):
    from ffmpeg_ffv2_b200.partition import frames_for_rank
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = _encode_share(frames_for_rank(NFRAMES, rank, world))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((gathered, float(t.item())))
    dist.destroy_process_group()


def test_round_robin_partition_over_two_ranks():
    from ffmpeg_ffv2_b200.partition import ReorderQueue, frames_for_rank, owner
    assert frames_for_rank(7, 1, 3) == [1, 4] and owner(5, 4) == 1
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    rq = ReorderQueue()
    out = []
    # packets arrive rank by rank; the queue releases them in presentation order
    for share in gathered[::-1]:
        for pts, pkt in share:
            rq.push(pts, pkt)
            out += [p for _, p in rq.pop_ready()]
    assert len(rq) == 0 and len(out) == NFRAMES
    want = [p for _, p in _encode_share(range(NFRAMES))]
    assert out == want
