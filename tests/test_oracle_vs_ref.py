"""Live cross-check: the oracle against the reference's own sources compiled here.

Needs oracle/_ref/libref.so (built from /root/reference in place by `make -C oracle ref`); skipped
where neither the library nor the reference tree exists.  Covers what the golden vectors cannot
carry across machines: REFERENCE seed mode (one sequential RNG stream from 0, single thread,
_mm256_rsqrt_ps primary rays — raytracer.c:596-720 exactly as written).
"""
import numpy as np
import pytest

import oracle_ffi
import ref_ffi
from helpers import MODELS, CAMERAS, scene_buffers
from raytracing_c_b200 import driver

pytestmark = pytest.mark.skipif(not ref_ffi.available(), reason="oracle/_ref/libref.so not buildable here")


def load_pair(name, **override):
    r = ref_ffi.lib()
    cam = driver.look_at(**CAMERAS[name]) if name in CAMERAS else None
    a = driver.load_scene(f"{MODELS}/{name}", shader_proc=r.ref_shader_proc(), background_proc=r.ref_background_proc(),
                          builder=ref_ffi.scene_builder(), camera=cam)
    b = driver.load_scene(f"{MODELS}/{name}", shader_proc=oracle_ffi.shader_proc(), background_proc=oracle_ffi.background_proc(),
                          builder=oracle_ffi.lib().oracle_scene_init, camera=cam)
    for loaded in (a, b):
        for i in range(loaded.model.n_materials):
            for k, v in override.items():
                setattr(loaded.model.materials[i], k, v)
    return a, b


def test_struct_sizes_match_reference_headers():
    r = ref_ffi.lib()
    assert (r.ref_sizeof_triangle(), r.ref_sizeof_triangle_aos(), r.ref_sizeof_bvh_node()) == (112, 112, 192)
    assert (r.ref_sizeof_scene(), r.ref_sizeof_context()) == (208, 80)


@pytest.mark.parametrize("name", ["spheres.glb", "sheen.glb", "tower.obj", "helmet.glb"])
def test_scene_init_bytes(name):
    a, b = load_pair(name)
    try:
        assert scene_buffers(a.scene) == scene_buffers(b.scene)
    finally:
        a.close(); b.close()


@pytest.mark.parametrize("name,w,h,spp,override", [
    ("spheres.glb", 96, 64, 4, {}),
    ("helmet.glb", 64, 48, 3, {}),
    ("sheen.glb", 48, 48, 4, {"sheen": 1.0, "sheen_tint": 1.0}),
    ("tower.obj", 48, 32, 2, {}),
])
def test_reference_mode_image_is_identical(name, w, h, spp, override):
    """render_thread_proc as shipped (T=1) vs the oracle in reference seed mode: same u8 image."""
    a, b = load_pair(name, **override)
    try:
        want = ref_ffi.render_reference_mode(a, w, h, spp)
        got = oracle_ffi.render(b, w, h, spp, 8, n_threads=1, seed_mode=oracle_ffi.SEED_REFERENCE, approx_rsqrt=True)["pixels"]
        assert want.any()
        assert np.array_equal(got, want)
    finally:
        a.close(); b.close()


def test_per_sample_mode_and_denoiser():
    a, b = load_pair("spheres.glb")
    try:
        want = ref_ffi.cast_rays_per_sample(a, 20, 12, 3, user_seed=99)
        got = oracle_ffi.render(b, 20, 12, 3, 8, n_threads=2, user_seed=99, want_per_sample=True)
        assert np.array_equal(got["per_sample"], want)
        assert np.array_equal(ref_ffi.denoise(got["pixels"]), oracle_ffi.denoise(got["pixels"]))
    finally:
        a.close(); b.close()
