"""Multi-GPU rendering: one process per GPU, BVH replicated, camera samples sharded, films summed.

The reference's only multi-worker mode gives worker `taskNum` of `taskCount` a sub-window of the sample
extent and lets the parent COPY each worker's region (lib/dartray_web/render_manager.dart:100-141,
lib/dartray/dartray.dart:1009-1023), which drops filter contributions that cross tile borders.  Here
every rank renders interleaved 1024-pixel blocks (`drt_render_shard`) into its own full-size film and
the films are SUMMED once (NCCL all-reduce over NVLink; gloo in the CPU tests), so the N-GPU image is
the 1-GPU image up to float64 summation order.  There is no exchange during traversal or shading.
"""
from __future__ import annotations

import numpy as np

BLOCK_PIXELS = 1024  # must match render_api.cu


def shard_pixel_count(total_pixels: int, shard: int, n_shards: int, block: int = BLOCK_PIXELS) -> int:
    """Pixels of the sample window that `drt_render_shard(shard, n_shards)` renders."""
    if n_shards <= 1:
        return total_pixels
    n_blocks = (total_pixels + block - 1) // block
    owned = (n_blocks - shard + n_shards - 1) // n_shards if n_blocks > shard else 0
    mine = owned * block
    if owned and (n_blocks - 1) % n_shards == shard:
        mine -= n_blocks * block - total_pixels
    return mine


def shard_pixel_indices(total_pixels: int, shard: int, n_shards: int, block: int = BLOCK_PIXELS) -> np.ndarray:
    """Row-major pixel indices of a shard, in the order the renderer visits them."""
    k = np.arange(shard_pixel_count(total_pixels, shard, n_shards, block), dtype=np.int64)
    if n_shards <= 1:
        return k
    return ((k // block) * n_shards + shard) * block + (k % block)


class _DevicePointer:
    """Exposes a raw device allocation to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


def film_tensor(ctx):
    """The context's film accumulators [height * width * (X, Y, Z, weight)] as a float64 torch tensor
    aliasing device memory."""
    import torch
    ptr, n = ctx.film_device()
    return torch.as_tensor(_DevicePointer(ptr, n), device=torch.device("cuda", ctx.device))


def sum_films(film, group=None):
    """Sum a film tensor over all ranks in place (NCCL on GPU tensors, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(film, op=dist.ReduceOp.SUM, group=group)
    return film


def render_sharded(ctx, rank: int, world: int, group=None):
    """Render this rank's shard and sum the films; afterwards every rank's `ctx.film_read()` returns the
    whole image.

    The sum is in place: after it every rank's film holds the contributions of ALL ranks, so a second sharded render into
    the same film would count the earlier passes `world` times.  Progressive multi-pass use therefore has to call
    `ctx.film_clear()` between passes (and accumulate the passes outside); this is checked."""
    import time

    import torch
    if world > 1 and getattr(ctx, "_film_is_summed", False):
        raise RuntimeError("render_sharded: the film already holds an all-reduced image; call ctx.film_clear() before the next pass")
    t0 = time.perf_counter()
    ctx.render_shard(rank, world)  # blocking: returns when this rank's film is complete
    t1 = time.perf_counter()
    if world > 1:
        film = film_tensor(ctx)
        sum_films(film, group)
        torch.cuda.synchronize(ctx.device)
        ctx._film_is_summed = True
    return {"render_s": t1 - t0, "film_sum_s": time.perf_counter() - t1}
