"""N>1 path on CPU: two ranks (gloo) each encode their range of closed-GOP
chunks with the host-emulated kernels; rank 0 concatenates the gathered byte
strings and must get exactly the single-process sharded stream."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ops
import util

W, H, N, CHUNK = 176, 144, 8, 2


def _clip():
    _, _, fr = util.read_y4m(util.clip("dist", W, H, N, "420"))
    return b"".join(ops.yuv_bytes(f) for f in fr)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = util.pkg()
    yuv = _clip()
    o = P.enc_opts(W, H, P.SUBSAMP_420, (30, 1), emu=True, qp=60, gop=CHUNK, noeos=1)
    part = P.encode_rank_shard(o, yuv, N, CHUNK, rank, world, emu=True)
    parts = [None] * world
    dist.gather_object(part, parts if rank == 0 else None, dst=0)
    # the timing reduction bench.py uses: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((b"".join(parts), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_ranges_cover_all_chunks():
    P = util.pkg()
    for n in (1, 2, 5, 100):
        for world in (1, 2, 4, 8):
            r = [P.rank_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_sharded_encode_gloo():
    util.ensure_emu()
    P = util.pkg()
    yuv = _clip()
    o = P.enc_opts(W, H, P.SUBSAMP_420, (30, 1), emu=True, qp=60, gop=CHUNK, noeos=1)
    want = P.encode_frames(o, yuv, N, emu=True, chunk=CHUNK, threads=1)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    assert got == want
    meta, nfr, _ = P.decode_frames(got, emu=True)
    assert nfr == N and meta.width == W
