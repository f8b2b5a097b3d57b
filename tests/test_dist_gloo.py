"""World-size-2/3 gloo runs of the multi-GPU host logic (tile deal, all-gather order, de-interleave indexing)
on CPU.  The render and de-interleave kernels are replaced by numpy stand-ins that follow include/rt.h's
layout contract; the CUDA versions are checked against the same contract in test_gpu_parity.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

TILE = 8


def _expected_frame(W, H):
    y, x = np.mgrid[0:H, 0:W]
    f = np.zeros((H, W, 4), dtype=np.uint8)
    f[..., 0] = (x * 7 + y * 3) % 251
    f[..., 1] = (x // TILE + 5 * (y // TILE)) % 253
    f[..., 2] = (x ^ y) % 256
    f[..., 3] = 255
    return f


def _fake_render(scene, cam, p, shard, stream):
    """Fills this rank's compact tile buffer from the expected frame (what the render kernel does)."""
    import petershirleyraytracer_b200 as rt
    W, H = p.width, p.height
    L = rt.tile_layout(p)
    exp = _expected_frame(W, H)
    buf = np.zeros((L.tiles_per_shard, TILE * TILE, 4), dtype=np.uint8)
    for l, t in enumerate(range(p.shard_rank, L.tiles_total, p.shard_count)):
        ty, tx = divmod(t, L.tiles_x)
        for ly in range(min(TILE, H - ty * TILE)):
            for lx in range(min(TILE, W - tx * TILE)):
                buf[l, ly * TILE + lx] = exp[ty * TILE + ly, tx * TILE + lx]
    shard.copy_(torch.from_numpy(buf.reshape(-1)))


def _np_deinterleave(p, gathered, frame, stream):
    """numpy mirror of deinterleave_kernel's indexing (rt_kernels.cuh)."""
    import petershirleyraytracer_b200 as rt
    W, H = p.width, p.height
    L = rt.tile_layout(p)
    g = gathered.numpy().reshape(p.shard_count, L.tiles_per_shard, TILE * TILE, 4)
    y, x = np.mgrid[0:H, 0:W]
    t = (y // TILE) * L.tiles_x + (x // TILE)
    out = g[t % p.shard_count, t // p.shard_count, (y % TILE) * TILE + (x % TILE)]
    frame.copy_(torch.from_numpy(np.ascontiguousarray(out).reshape(-1)))


def _worker(rank, world, port, W, H, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import petershirleyraytracer_b200 as rt
    from petershirleyraytracer_b200 import dist as rdist
    p = rt.make_params(W, H, 4)
    tiles = rdist.tiles_of_rank(p, rank, world)
    frame = rdist.render_sharded(None, None, p, rank, world, device=torch.device("cpu"), render_fn=_fake_render,
                                 deinterleave_fn=_np_deinterleave)
    ok = np.array_equal(frame.numpy(), _expected_frame(W, H))
    counts = torch.tensor([len(tiles)], dtype=torch.int64)
    dist.all_reduce(counts)
    L = rt.tile_layout(rdist.shard_params(p, rank, world))
    ret[rank] = bool(ok) and counts.item() == L.tiles_total and len(tiles) <= L.tiles_per_shard
    dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H", [(2, 64, 40), (3, 50, 37), (2, 9, 9)])
def test_sharded_assembly_gloo(world, W, H):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, ret)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


# ------------------------------------------------------------------ sample-split deal: all tiles, a range of samples per rank
def _fake_pass(scene, cam, p, begin, accum, stream):
    """Adds a deterministic integer 'radiance' per (pixel, absolute sample index, channel) -- what rt_render_pass_device
    does with real samples."""
    W, H = p.width, p.height
    a = accum.numpy().reshape(H, W, 3)
    y, x = np.mgrid[0:H, 0:W]
    for s in range(begin, begin + p.spp):
        for c in range(3):
            a[..., c] += ((x * 131 + y * 71 + s * 17 + c * 5) % 1009).astype(np.int64) << 20


def _expected_sums(W, H, spp):
    y, x = np.mgrid[0:H, 0:W]
    out = np.zeros((H, W, 3), dtype=np.int64)
    for s in range(spp):
        for c in range(3):
            out[..., c] += ((x * 131 + y * 71 + s * 17 + c * 5) % 1009).astype(np.int64) << 20
    return out


def _np_to_frame(p, accum, total, frame, stream):
    a = accum.numpy().reshape(p.height, p.width, 3)
    f = np.zeros((p.height, p.width, 4), dtype=np.uint8)
    f[..., :3] = ((a // total) >> 22) % 256
    f[..., 3] = 255
    frame.copy_(torch.from_numpy(f.reshape(-1)))


def _split_worker(rank, world, port, W, H, spp, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import petershirleyraytracer_b200 as rt
    from petershirleyraytracer_b200 import dist as rdist
    p = rt.make_params(W, H, spp)
    frame = rdist.render_sample_split(None, None, p, rank, world, device=torch.device("cpu"), pass_fn=_fake_pass,
                                      to_frame_fn=_np_to_frame)
    exp = np.zeros((H, W, 4), dtype=np.uint8)
    exp[..., :3] = ((_expected_sums(W, H, spp) // spp) >> 22) % 256
    exp[..., 3] = 255
    ranges = [rdist.sample_range(spp, r, world) for r in range(world)]
    covered = ranges[0][0] == 0 and ranges[-1][1] == spp and all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
    ret[rank] = bool(np.array_equal(frame.numpy(), exp)) and covered
    dist.destroy_process_group()


@pytest.mark.parametrize("world,spp", [(2, 7), (3, 3), (2, 1)])
def test_sample_split_assembly_gloo(world, spp):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_split_worker, args=(r, world, port, 40, 24, spp, ret)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
