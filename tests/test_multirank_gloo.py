"""The N > 1 path on CPU: two processes over gloo split a stream by the rule the product's
routing handle uses (ffgpu_api.cu menc_owner: picture k belongs to worker (k / chunk) % N,
chunk = 1 when every frame is a key frame, else the GOP length), each codes its share with
its own encoder instance -- the product's device functions compiled for the CPU (tests/emul)
standing in for a GPU -- rank 0 gathers the packets and hands them on in picture order; the
result must equal the single-process stream of the oracle.  With a GOP > 1 this only holds
because whole GOPs stay on one worker (the adaptive states carry from frame to frame,
ffv1enc.c:1071).  Also checks the max-over-ranks timing reduction bench.py uses."""
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

W, H, FMT, NFRAMES = 64, 48, "yuv420p10le", 23


def owner(k, chunk, world):
    return (k // chunk) % world


def _encode(which, opts, indices):
    import cpucodec as cc
    import synth
    enc = cc.Encoder(which, W, H, FMT, **opts)
    return [(i, enc.encode(synth.testsrc2_like(FMT, W, H, i))) for i in indices]


def _worker(rank, world, port, opts, chunk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = _encode("emul", opts, [k for k in range(NFRAMES) if owner(k, chunk, world) == rank])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((gathered, float(t.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("opts,chunk", [(dict(slices=4, gop_size=1), 1),
                                        (dict(slices=4, gop_size=5, coder=1), 5)])
def test_stream_split_over_two_ranks(opts, chunk):
    import cpucodec as cc
    if not cc.available("emul"):
        pytest.skip("tests/emul not built")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, opts, chunk, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    # in-order hand-over: the packet that is due comes from the worker that owns picture k
    cursor = [0, 0]
    out = []
    for k in range(NFRAMES):
        r = owner(k, chunk, 2)
        pts, pkt = gathered[r][cursor[r]]
        cursor[r] += 1
        assert pts == k
        out.append(pkt)
    assert [len(g) for g in gathered] == cursor
    assert out == [p for _, p in _encode("oracle", opts, range(NFRAMES))]
