"""The N>1 path on CPU: world_size-2 gloo.  Each rank renders the share the LIBRARY's policy gives it
(rt_gpu_shard_samples / rt_gpu_shard_chunks: pure host arithmetic in libraytracer_gpu.so, usable without a device);
the oracle stands in for the device kernels (which the GPU tests prove equal to it); the accumulators are summed on
rank 0; rank 0 resolves.  Sample split: equal to the single-process frame up to f32 summation order.  Chunk split:
BIT-identical, because every pixel is summed on one rank and the cross-rank sum only adds zeros.
The device side of the same machinery (shards on one GPU, N GPUs in one process, the C host with --gpus) is
tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_ffi
from helpers import load
from raytracing_c_b200.sharding import SPLIT_CHUNKS, SPLIT_SAMPLES, owned_chunks, sample_range, split_mode


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


def chunk_mask(rank, world, w, h):
    """pixels of the 32x32 chunks whose row-major id is congruent to rank (raytracer.c:619-637 dealt round-robin)"""
    cx = (w + 31) // 32
    ys, xs = np.mgrid[0:h, 0:w]
    return (((ys // 32) * cx + xs // 32) % world == rank).astype(np.uint8)


def _worker(rank, world, port, w, h, spp, out_path, chunks=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    loaded = load("spheres.glb")
    if chunks:
        mask = chunk_mask(rank, world, w, h)
        assert mask.sum() <= owned_chunks(rank, world, w, h) * 1024
        part = oracle_ffi.render(loaded, w, h, spp, 8, n_threads=2, pixel_mask=mask)["accum"]
    else:
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


def test_two_rank_chunk_split_is_bit_identical_to_the_single_rank_frame(tmp_path):
    """spp < 8 x ranks: the policy deals 32x32 chunks instead of sample ranges, and the reduced frame is the SAME bits."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    w, h, spp = 70, 40, 4                      # 3 x 2 chunks, ragged edges; 4 spp over 2 ranks cannot be cut at a batch of 8
    assert split_mode(spp, 2) == SPLIT_CHUNKS and split_mode(1024, 8) == SPLIT_SAMPLES and split_mode(24, 2) == SPLIT_CHUNKS
    assert owned_chunks(0, 2, w, h) + owned_chunks(1, 2, w, h) == 6
    out = str(tmp_path / "accum.npy")
    mp.spawn(_worker, args=(2, port, w, h, spp, out, True), nprocs=2, join=True)
    got = np.load(out)
    loaded = load("spheres.glb")
    try:
        full = oracle_ffi.render(loaded, w, h, spp, 8, n_threads=2)
    finally:
        loaded.close()
    assert np.array_equal(got, full["accum"])
    assert np.array_equal(oracle_ffi.resolve(got, spp), full["pixels"])
