"""Multi-GPU plumbing for the one place the path shards: samples per pixel.

Rank r of G renders the absolute sample range [r*spp/G, (r+1)*spp/G) of EVERY pixel with the scene and BVH
replicated (each rank builds locally; the build is deterministic). The RNG is keyed by absolute sample index, so
the union of the ranks' samples is exactly the single-GPU sample set. The per-rank accumulators hold SUMS, so the
only collective is one reduce(SUM) to rank 0 — NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple


def shard_samples(total_spp: int, rank: int, world: int) -> Tuple[int, int]:
    """(sample_offset, samples) of `rank`; ranges tile [0, total_spp) exactly, sizes differ by at most 1."""
    if world < 1 or not (0 <= rank < world) or total_spp < 0:
        raise ValueError("bad shard request")
    lo = rank * total_spp // world
    hi = (rank + 1) * total_spp // world
    return lo, hi - lo


class DeviceArray:
    """Minimal __cuda_array_interface__ holder so torch can view the library's accumulator without a copy."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {
            "shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None,
        }


def accumulator_tensor(ctx):
    """torch.float32 view (no copy) of ctx's device accumulator (W*H*3 sums)."""
    import torch

    ptr, n = ctx.accum_device_ptr()
    return torch.as_tensor(DeviceArray(ptr, n), device=f"cuda:{ctx.device}")


def reduce_accumulators(tensor, total_spp: int, dst: int = 0):
    """The single collective of the path: reduce(SUM) of the per-rank sums to `dst`; returns the mean image on dst."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return None
    return tensor / float(total_spp)
