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
