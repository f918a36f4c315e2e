"""CPU test of the N>1 path: two gloo ranks shard the work with the plans bench.py / the multi-GPU
driver use (srcnn_cpp_b200/shard.py), compute their shares with the CPU checker, exchange the results
and check that the union equals the unsplit result bit for bit.  No collective exists on the data path;
gloo is only used here to gather results for the comparison."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import natural_like
    from oracle.oracle import Oracle
    from srcnn_cpp_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    rng = np.random.default_rng(123)           # same data on every rank
    frames = np.stack([natural_like(rng, 24, 30) for _ in range(5)])
    scale = 2.0
    # ---- frame sharding ----
    mine = shard.frames_for_rank(len(frames), rank, world)
    out = torch.zeros((len(frames), 48, 60, 3), dtype=torch.uint8)
    for f in mine:
        out[f] = torch.from_numpy(orc.pipeline(frames[f], scale))
    dist.all_reduce(out, op=dist.ReduceOp.SUM)  # disjoint frames: the sum is the union
    # ---- row-band sharding of one image: each rank only looks at the source rows its bands need ----
    img = natural_like(rng, 40, 36)
    oh = 80
    ofs, _ = orc.cubic_taps(40, oh)
    band_out = torch.zeros((oh, 72, 3), dtype=torch.uint8)
    for (r0, r1) in shard.bands_for_rank(oh, rank, world, bands_per_rank=2):
        s0, s1 = shard.band_src_rows_py(40, scale, r0, r1, ofs)
        # emulate "this rank only holds rows [s0,s1)": poison everything else
        poisoned = np.full_like(img, 255 if rank else 0)
        poisoned[s0:s1] = img[s0:s1]
        full = orc.pipeline(poisoned, scale)
        band_out[r0:r1] = torch.from_numpy(full[r0:r1])
    dist.all_reduce(band_out, op=dist.ReduceOp.SUM)
    if rank == 0:
        want_frames = np.stack([orc.pipeline(f, scale) for f in frames])
        want_img = orc.pipeline(img, scale)
        q.put((bool(np.array_equal(out.numpy(), want_frames)), bool(np.array_equal(band_out.numpy(), want_img))))
    dist.destroy_process_group()


def test_two_rank_sharding_equals_unsplit():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_frames, ok_bands = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_frames, "frame-sharded union differs from the unsplit batch"
    assert ok_bands, "row-band union (with only the planned source rows visible) differs from the unsplit image"


def test_plans_cover_everything_once():
    sys.path.insert(0, ROOT)
    from srcnn_cpp_b200 import shard
    for world in (1, 2, 4, 8):
        got = sorted(sum((shard.frames_for_rank(1024, r, world) for r in range(world)), []))
        assert got == list(range(1024))
        for bpr in (1, 3):
            rows = []
            for r in range(world):
                for (a, b) in shard.bands_for_rank(65536, r, world, bpr):
                    rows.append((a, b))
            rows.sort()
            assert rows[0][0] == 0 and rows[-1][1] == 65536
            assert all(rows[i][1] == rows[i + 1][0] for i in range(len(rows) - 1))
