"""Multi-GPU driver: one process per GPU (torch.distributed), tile-sharded render, one all-gather.

The frame's 8x8 tiles are dealt round-robin to the ranks (tile t -> rank t % world).  Every rank renders its
tiles into a compact buffer (tile slot l holds frame tile l*world + rank), the buffers are all-gathered and
de-interleaved into the frame on every rank.  Integer radiance sums and the (pixel, sample)-keyed RNG make
the result bit-identical to a single-GPU render.  No other collective is on the path.
"""
from __future__ import annotations

import copy

import numpy as np
import torch
import torch.distributed as dist

import petershirleyraytracer_b200 as rt


def shard_params(params: rt.RtParams, rank: int, world: int) -> rt.RtParams:
    p = copy.copy(params)
    p.shard_rank, p.shard_count = rank, world
    return p


def tiles_of_rank(params: rt.RtParams, rank: int, world: int) -> np.ndarray:
    """Frame tile indices (row-major over the tile grid) rendered by `rank`, in shard-slot order."""
    L = rt.tile_layout(shard_params(params, 0, world))
    return np.arange(rank, L.tiles_total, world, dtype=np.int64)


def _cuda_render(scene, cam, p, shard: torch.Tensor, stream: int) -> None:
    rt.render_device(scene, cam, p, shard.data_ptr(), 0, stream)


def _cuda_deinterleave(p, gathered: torch.Tensor, frame: torch.Tensor, stream: int) -> None:
    rt.deinterleave(p, gathered.data_ptr(), frame.data_ptr(), frame.device.index or 0, stream)


def render_sharded(scene, cam, params: rt.RtParams, rank: int, world: int, device: torch.device | None = None,
                   render_fn=_cuda_render, deinterleave_fn=_cuda_deinterleave, group=None) -> torch.Tensor:
    """Returns the assembled (H, W, 4) uint8 frame on `device` (identical on every rank)."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    p = shard_params(params, rank, world)
    L = rt.tile_layout(p)
    stream = torch.cuda.current_stream().cuda_stream if device.type == "cuda" else 0
    shard = torch.zeros(L.shard_bytes, dtype=torch.uint8, device=device)
    frame = torch.empty(params.height * params.width * 4, dtype=torch.uint8, device=device)
    render_fn(scene, cam, p, shard, stream)
    if world > 1:
        gathered = torch.empty(world * L.shard_bytes, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(gathered, shard, group=group)
    else:
        gathered = shard
    if world > 1 or render_fn is not _cuda_render:
        deinterleave_fn(p, gathered, frame, stream)
    else:
        frame = shard  # shard_count == 1 renders straight into frame order
    if device.type == "cuda" and render_fn is _cuda_render:
        rt.render_finish(scene)
    return frame[: params.height * params.width * 4].view(params.height, params.width, 4)


# ------------------------------------------------------------------ sample-split deal (every rank: all tiles, a sample range)
def sample_range(spp: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of the samples-per-pixel rank `rank` traces: contiguous ranges, sizes differing by at most one."""
    return rank * spp // world, (rank + 1) * spp // world


def _cuda_pass(scene, cam, p, begin, accum: torch.Tensor, stream: int) -> None:
    rt.render_pass_device(scene, cam, p, begin, accum.data_ptr(), 0, stream)


def _cuda_to_frame(p, accum: torch.Tensor, total: int, frame: torch.Tensor, stream: int) -> None:
    rt.accum_to_frame(p, accum.data_ptr(), total, frame.data_ptr(), frame.device.index or 0, stream)


def render_sample_split(scene, cam, params: rt.RtParams, rank: int, world: int, device: torch.device | None = None,
                        pass_fn=_cuda_pass, to_frame_fn=_cuda_to_frame, group=None) -> torch.Tensor:
    """The other way to share a frame: every rank traces ALL tiles for its range of the samples (the Philox streams are
    keyed on the absolute sample index), the integer radiance sums are added up with one all-reduce (order-independent,
    W*H*3 int64), and write_color runs on the total.  Every rank sees the same spatial mix of cheap and expensive
    pixels, so the load balance does not depend on the picture; the frame is the single-GPU frame bit for bit.
    Needs spp >= world.  Returns the (H, W, 4) uint8 frame on `device` (identical on every rank)."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    W, H = params.width, params.height
    begin, end = sample_range(params.spp, rank, world)
    stream = torch.cuda.current_stream().cuda_stream if device.type == "cuda" else 0
    accum = torch.zeros(H * W * 3, dtype=torch.int64, device=device)
    if end > begin:
        p = copy.copy(params)
        p.spp = end - begin
        p.shard_rank, p.shard_count = 0, 1
        pass_fn(scene, cam, p, begin, accum, stream)
    if world > 1:
        dist.all_reduce(accum, group=group)      # int64 sums of 20.44 fixed-point values: exact, any order
    frame = torch.empty(H * W * 4, dtype=torch.uint8, device=device)
    to_frame_fn(params, accum, params.spp, frame, stream)
    if device.type == "cuda" and pass_fn is _cuda_pass and end > begin:
        rt.render_finish(scene)
    return frame.view(H, W, 4)
