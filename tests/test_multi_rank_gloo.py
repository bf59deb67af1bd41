"""The N>1 path on CPU: world_size-2 gloo.  Each rank renders its sample range (the oracle stands in
for the device kernel, which is what the GPU tests prove equal to it), the accumulators are reduced
to rank 0 exactly as bench.py does with NCCL, rank 0 resolves; the result must equal the
single-process frame up to f32 summation order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_ffi
from helpers import load
from raytracing_c_b200.sharding import sample_range


def test_sample_ranges_tile_the_frame():
    for world in (1, 2, 3, 4, 8):
        for spp in (1, 8, 16, 24, 100, 1024):
            spans = [sample_range(r, world, spp) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == spp
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo % 8 == 0 for lo, _ in spans)
    assert sample_range(0, 8, 1024) == (0, 128) and sample_range(7, 8, 1024) == (896, 1024)
    with pytest.raises(ValueError):
        sample_range(2, 2, 16)


def _worker(rank, world, port, w, h, spp, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    loaded = load("spheres.glb")
    lo, hi = sample_range(rank, world, spp)
    part = oracle_ffi.render(loaded, w, h, spp, 8, n_threads=2, sample_begin=lo, sample_end=hi)["accum"]
    t = torch.from_numpy(part.copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, t.numpy())
    loaded.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_split_reduces_to_the_single_rank_frame(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    w, h, spp = 40, 30, 16
    out = str(tmp_path / "accum.npy")
    mp.spawn(_worker, args=(2, port, w, h, spp, out), nprocs=2, join=True)
    got = np.load(out)
    loaded = load("spheres.glb")
    try:
        full = oracle_ffi.render(loaded, w, h, spp, 8, n_threads=2)
    finally:
        loaded.close()
    np.testing.assert_allclose(got, full["accum"], rtol=2e-6, atol=1e-6)
    resolved = oracle_ffi.resolve(got, spp)
    # an f32 reassociation may move a value across a u8 truncation boundary: at most one code, rarely
    assert np.abs(resolved.astype(int) - full["pixels"].astype(int)).max() <= 1
    assert (resolved == full["pixels"]).mean() > 0.99
