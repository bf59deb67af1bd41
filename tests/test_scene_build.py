"""scene_init (product host builder, raytracing_c_b200/host/scene_build.c) against the oracle's literal
restatement of reference scene.c: byte-identical nodes, SoA positions and AoS records."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_ffi
from helpers import MODELS, load, scene_buffers
from raytracing_c_b200._ffi import Scene, Triangle, TriangleSlice, host_lib


def build_both(tris: np.ndarray):
    """tris: (n, 3, 3) float32 positions; normals/uvs derived deterministically."""
    n = len(tris)
    arr = (Triangle * max(n, 1))()
    for i in range(n):
        for v in range(3):
            arr[i].positions[v].x, arr[i].positions[v].y, arr[i].positions[v].z = map(float, tris[i, v])
            arr[i].normals[v].x, arr[i].normals[v].y, arr[i].normals[v].z = 0.0, 1.0, 0.0
            arr[i].tex_coords[v].x, arr[i].tex_coords[v].y = float(v == 1) * (1 + i % 3), float(v == 2) + 0.25 * (i % 2)
        arr[i].shader.data = 0x1000 + 16 * i
    sl = TriangleSlice(C.cast(arr, C.POINTER(Triangle)), n)
    a, b = Scene(), Scene()
    host_lib().scene_init(C.byref(a), sl)
    oracle_ffi.lib().oracle_scene_init(C.byref(b), sl)
    out = (scene_buffers(a), scene_buffers(b), int(a.bvh.depth), int(a.bvh.nodes.len), int(a.triangles.len),
           np.frombuffer(C.string_at(a.triangles.aos, a.triangles.len * 112), dtype=np.uint64).reshape(-1, 14)[:, 12].copy())
    host_lib().scene_destroy(C.byref(a))
    oracle_ffi.lib().oracle_scene_destroy(C.byref(b))
    return out


@pytest.mark.parametrize("name", ["quad.obj", "fov_test.obj", "tower.obj", "spheres.glb", "sheen.glb", "helmet.glb"])
def test_shipped_models_byte_identical(name):
    a, b = load(name), load(name, builder=oracle_ffi.lib().oracle_scene_init)
    try:
        assert scene_buffers(a.scene) == scene_buffers(b.scene)
        assert a.scene.bvh.depth == b.scene.bvh.depth and a.scene.bvh.last_row_offset == b.scene.bvh.last_row_offset
    finally:
        a.close(); b.close()


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 63, 64, 65, 72, 511, 513, 1000])
def test_random_soups_and_size_formulas(n):
    """Sizes around the 8^k boundaries, incl. the reference's two failure cases (n<=8 and 64+8)."""
    rng = np.random.default_rng(n)
    base = rng.uniform(-5, 5, (n, 1, 3))
    tris = (base + rng.uniform(-0.5, 0.5, (n, 3, 3))).astype(np.float32)
    a, b, depth, nodes, slots, shader = build_both(tris)
    assert a == b
    groups = (n + 7) // 8
    want_depth = 1
    while 8 ** want_depth < groups:
        want_depth += 1
    assert depth == want_depth and nodes == sum(8 ** k for k in range(depth)) and slots == 8 ** depth * 8
    assert sorted(int(s) for s in shader if s) == [0x1000 + 16 * i for i in range(n)], "every triangle lands in exactly one slot"


def test_centroid_ties_are_resolved_like_a_stable_sort():
    """Many identical centroid keys on every axis: membership of each leaf must not depend on the
    sort algorithm (the product builder uses qsort with a composite key, the oracle a merge sort)."""
    rng = np.random.default_rng(42)
    grid = rng.integers(0, 3, (400, 1, 3)).astype(np.float32)
    shape = np.array([[0, 0, 0], [0.25, 0, 0], [0, 0.25, 0]], dtype=np.float32)
    a, b, *_ = build_both(grid + shape)
    assert a == b


def test_degenerate_uvs_and_zero_area():
    tris = np.zeros((12, 3, 3), dtype=np.float32)
    tris[:, 1, 0] = 1
    tris[:, 2, 1] = np.arange(12) % 2          # every other triangle has zero area -> NaN normal, same on both sides
    a, b, *_ = build_both(tris)
    assert a == b


def test_scene_cache_round_trip_and_rejections(tmp_path):
    """Scene cache (reference scene.c:18-76, host/scene_cache.c): same container, material indices instead of the
    reference's raw pointers.  A loaded scene has the same bytes, re-bound shaders, and renders the same radiance."""
    import ctypes as C
    import struct
    from raytracing_c_b200._ffi import Scene, host_lib
    from helpers import scene_buffers
    host = host_lib()
    loaded = load("helmet.glb")
    try:
        path = str(tmp_path / "helmet.scene")
        assert host.scene_save_file(path.encode(), C.byref(loaded.scene), loaded.model.materials, loaded.model.n_materials)
        n_nodes, n_slots = loaded.scene.bvh.nodes.len, loaded.scene.triangles.len
        assert os.path.getsize(path) == 96 + n_nodes * 192 + n_slots * (36 + 112)         # header aligned to 32 B (scene.c:13-16)
        back = Scene()
        assert host.scene_load_file(path.encode(), C.byref(back), loaded.model.materials, loaded.model.n_materials, loaded.shader_proc)
        assert scene_buffers(back) == scene_buffers(loaded.scene)
        aos_a = np.frombuffer(C.string_at(loaded.scene.triangles.aos, n_slots * 112), dtype=np.uint64).reshape(n_slots, 14)
        aos_b = np.frombuffer(C.string_at(back.triangles.aos, n_slots * 112), dtype=np.uint64).reshape(n_slots, 14)
        assert np.array_equal(aos_a[:, 12:], aos_b[:, 12:]), "Shader {data, proc} pointers must be re-bound identically"
        back.background = loaded.scene.background
        want = oracle_ffi.render(loaded, 48, 32, 2, n_threads=2)["accum"]
        keep = loaded.scene
        loaded.scene = back
        got = oracle_ffi.render(loaded, 48, 32, 2, n_threads=2)["accum"]
        loaded.scene = keep
        assert np.array_equal(got, want)
        host.scene_destroy(C.byref(back))
        # rejections: truncated, padded, version 0 (the reference's raw-pointer files), hostile counts
        data = open(path, "rb").read()
        for name, blob in (("cut", data[:-7]), ("pad", data + b"\0" * 32), ("v0", struct.pack("<i", 0) + data[4:]),
                           ("nodes", data[:4] + struct.pack("<i", 2 ** 30) + data[8:])):
            p = str(tmp_path / f"{name}.scene")
            open(p, "wb").write(blob)
            assert not host.scene_load_file(p.encode(), C.byref(Scene()), loaded.model.materials, loaded.model.n_materials, loaded.shader_proc), name
        assert b"version 0" in host.rt_host_last_error() or b"disagree" in host.rt_host_last_error()
    finally:
        loaded.close()
