"""How one frame is split over ranks: by sample range, in multiples of the reference's
8-sample jitter batch (raytracer.c:641-697), every rank rendering all pixels.  The per-rank f32
accumulators are then summed to rank 0 (one NCCL reduce), which resolves and denoises."""
from __future__ import annotations

JITTER_BATCH = 8


def sample_range(rank: int, world: int, samples: int) -> tuple[int, int]:
    """[begin, end) of the samples rank `rank` renders.  Batches of 8 are dealt round the ranks as
    evenly as possible; ranks beyond the number of batches get an empty range."""
    if world < 1 or not 0 <= rank < world or samples < 0:
        raise ValueError("bad rank/world/samples")
    batches = (samples + JITTER_BATCH - 1) // JITTER_BATCH
    lo = (batches * rank) // world
    hi = (batches * (rank + 1)) // world
    return min(lo * JITTER_BATCH, samples), min(hi * JITTER_BATCH, samples)
