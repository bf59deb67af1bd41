"""How one frame is split over ranks — a thin view of the library's own policy
(rt_gpu_shard_* in csrc/rt_multi.cu; pure host arithmetic, usable without a device).

  * sample split: every rank renders all pixels for a range of sample indices, in multiples of the reference's
    8-sample jitter batch (raytracer.c:641-697);
  * chunk split: the reference's 32x32 chunks (raytracer.c:619-637) dealt round-robin — used when the sample
    batches do not divide evenly over the ranks (spp < 8 x ranks, or a remainder).
The per-rank f32 accumulators are then summed on rank 0 (fused reduce + resolve over NVLink), which denoises."""
from __future__ import annotations

import ctypes as C

from ._ffi import SPLIT_AUTO, SPLIT_CHUNKS, SPLIT_SAMPLES, gpu_lib  # noqa: F401

JITTER_BATCH = 8


def sample_range(rank: int, world: int, samples: int) -> tuple[int, int]:
    """[begin, end) of the samples rank `rank` renders under the sample split; ranks beyond the number of
    8-sample batches get an empty range."""
    if world < 1 or not 0 <= rank < world or samples < 0:
        raise ValueError("bad rank/world/samples")
    lo, hi = C.c_int32(), C.c_int32()
    gpu_lib().rt_gpu_shard_samples(rank, world, samples, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def split_mode(samples: int, world: int, requested: int = SPLIT_AUTO) -> int:
    return int(gpu_lib().rt_gpu_shard_mode(samples, world, requested))


def owned_chunks(rank: int, world: int, width: int, height: int) -> int:
    return int(gpu_lib().rt_gpu_shard_chunks(rank, world, width, height))
