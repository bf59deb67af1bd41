"""Known-answer and property tests of the oracle's building blocks (no GPU)."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_ffi
from helpers import load
from oracle_ffi import Ray
from raytracing_c_b200._ffi import Vec3


def _o():
    o = oracle_ffi.lib()
    o.oracle_rand_u32.restype = C.c_uint32
    o.oracle_rand_u32.argtypes = [C.POINTER(C.c_uint32)]
    o.oracle_rand_f32.restype = C.c_float
    o.oracle_rand_f32.argtypes = [C.POINTER(C.c_uint32)]
    o.oracle_path_seed.restype = C.c_uint32
    o.oracle_path_seed.argtypes = [C.c_uint32] * 3
    return o


def test_rand_u32_known_answers():
    """common.h:15-20 evaluated independently in Python integers; the OUTPUT is fed back as state."""
    def step(state):
        s = (state * 747796405 + 2891336453) & 0xFFFFFFFF
        w = (((s >> ((s >> 28) + 4)) ^ s) * 277803737) & 0xFFFFFFFF
        return (w >> 22) ^ w
    o = _o()
    for seed in (0, 1, 0xDEADBEEF):
        st, py = C.c_uint32(seed), seed
        for _ in range(16):
            py = step(py)
            assert o.oracle_rand_u32(C.byref(st)) == py and st.value == py
    st = C.c_uint32(0)
    vals = [o.oracle_rand_f32(C.byref(st)) for _ in range(2000)]
    assert 0.0 <= min(vals) and max(vals) <= 1.0
    assert abs(np.mean(vals) - 0.5) < 0.03


def test_path_seed_separates_pixels_and_samples():
    o = _o()
    seeds = {o.oracle_path_seed(p, s, 0) for p in range(64) for s in range(64)}
    assert len(seeds) == 64 * 64
    assert o.oracle_path_seed(5, 7, 0) != o.oracle_path_seed(5, 7, 1)


def test_hash12_matches_float32_restatement():
    """raytracer.c:584-594 restated with numpy float32 scalars."""
    f = np.float32

    def fract(v):
        return f(v - np.floor(v))

    def hash12(px, py):
        a, b, c = fract(f(px) * f(0.1031)), fract(f(py) * f(0.1031)), fract(f(px) * f(0.1031))
        k = f(33.33)
        d = f(f(f(a * f(b + k)) + f(b * f(c + k))) + f(c * f(a + k)))
        return fract(f(f(f(a + b) + f(d * f(2.0))) * f(c + d)))
    o = oracle_ffi.lib()
    rng = np.random.default_rng(3)
    for _ in range(300):
        x, y = f(rng.integers(0, 1920) * 50 + rng.integers(0, 1024)), f(rng.integers(0, 1080))
        assert o.oracle_hash12(float(x), float(y)) == hash12(x, y)
        assert 0.0 <= hash12(x, y) < 1.0


@pytest.mark.parametrize("fn,ref,lo,hi", [
    ("sin", math.sin, -7.0, 7.0), ("cos", math.cos, -7.0, 7.0), ("asin", math.asin, -1.0, 1.0)])
def test_rt_math_unary_within_one_ulp(fn, ref, lo, hi):
    o = oracle_ffi.lib()
    g = getattr(o, f"oracle_{fn}f")
    g.restype, g.argtypes = C.c_float, [C.c_float]
    xs = np.random.default_rng(1).uniform(lo, hi, 20000).astype(np.float32)
    for x in xs[:4000]:
        want = np.float32(ref(float(x)))
        assert abs(np.float32(g(float(x))) - want) <= np.spacing(np.abs(want)) + 1e-45


def test_rt_math_pow_and_atan2_within_one_ulp():
    o = oracle_ffi.lib()
    o.oracle_powf.restype, o.oracle_powf.argtypes = C.c_float, [C.c_float, C.c_float]
    o.oracle_atan2f.restype, o.oracle_atan2f.argtypes = C.c_float, [C.c_float, C.c_float]
    rng = np.random.default_rng(2)
    for y in (2.4, 1 / 2.4, 5.0, 2.0):
        for x in np.concatenate([rng.uniform(0, 1, 1500), rng.uniform(0, 1e-3, 300), rng.uniform(1, 4, 300)]).astype(np.float32):
            want = np.float32(math.pow(float(x), float(np.float32(y))))
            assert abs(np.float32(o.oracle_powf(float(x), float(np.float32(y)))) - want) <= np.spacing(want)
    assert o.oracle_powf(-0.5, 5.0) == -0.03125 and o.oracle_powf(0.0, 2.4) == 0.0 and o.oracle_powf(3.0, 0.0) == 1.0
    assert math.isnan(o.oracle_powf(-0.1, 2.4))
    for a, b in rng.uniform(-1, 1, (2000, 2)).astype(np.float32):
        want = np.float32(math.atan2(float(a), float(b)))
        assert abs(np.float32(o.oracle_atan2f(float(a), float(b))) - want) <= np.spacing(np.abs(want))
    assert o.oracle_atan2f(0.0, -1.0) == np.float32(math.pi) and o.oracle_atan2f(0.0, 0.0) == 0.0


def test_rt_math_dense_sweeps_stay_within_0_51_ulp():
    """rt_math.h is the libm of BOTH sides, so its accuracy is part of the parity contract: dense sweeps
    against binary64 libm, error measured in ulps of the binary32 result."""
    o = oracle_ffi.lib()
    rng = np.random.default_rng(11)

    def ulps(got, want64):
        want32 = want64.astype(np.float32)
        return np.abs(got.astype(np.float64) - want64) / np.spacing(np.abs(want32)).astype(np.float64)

    o.oracle_powf_array.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_ssize_t]
    for y in (2.4, 1 / 2.4, 0.37, 7.3, -1.5):
        x = np.concatenate([rng.uniform(0, 1, 200000), rng.uniform(1, 100, 50000), np.exp(rng.uniform(-60, 60, 50000))]).astype(np.float32)
        out = np.zeros_like(x)
        o.oracle_powf_array(x.ctypes.data, np.float32(y), out.ctypes.data, x.size)
        want = np.power(x.astype(np.float64), np.float64(np.float32(y)))
        ok = (want > 1e-37) & (want < 3e38)
        assert ulps(out[ok], want[ok]).max() < 0.51
    # subnormal base, overflow, underflow
    assert o.oracle_powf(1e-41, 0.5) == pytest.approx(math.sqrt(1e-41), rel=1e-6)
    assert o.oracle_powf(1e30, 2.4) == float("inf") and o.oracle_powf(1e-30, 2.4) == 0.0

    o.oracle_asin_array.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t]
    x = np.concatenate([rng.uniform(-1, 1, 300000), [0.0, 1.0, -1.0, 0.5, -0.5, 0.99999994, 1e-20]]).astype(np.float32)
    out = np.zeros_like(x)
    o.oracle_asin_array(x.ctypes.data, out.ctypes.data, x.size)
    nz = x != 0
    assert ulps(out[nz], np.arcsin(x[nz].astype(np.float64))).max() < 0.51 and out[~nz].max() == 0
    assert math.isnan(o.oracle_asinf(1.0000001))

    o.oracle_atan2_array.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ssize_t]
    a, b = rng.uniform(-1, 1, 300000).astype(np.float32), rng.uniform(-1, 1, 300000).astype(np.float32)
    out = np.zeros_like(a)
    o.oracle_atan2_array(a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size)
    assert ulps(out, np.arctan2(a.astype(np.float64), b.astype(np.float64))).max() < 0.51


def test_powf_positive_is_bit_identical_to_powf_on_its_domain():
    """The sRGB decode on the device skips rt_powf's case analysis (rt_shade.cuh decode_srgb)."""
    o = oracle_ffi.lib()
    sig = [C.c_void_p, C.c_float, C.c_void_p, C.c_ssize_t]
    o.oracle_powf_array.argtypes = o.oracle_powf_positive_array.argtypes = sig
    x = np.concatenate([np.random.default_rng(9).uniform(0.05, 1.0, 500000), np.arange(256) / 255.999 / 1.055 + 0.055 / 1.055]).astype(np.float32)
    a, b = np.zeros_like(x), np.zeros_like(x)
    for y in (2.4, 1 / 2.4):
        o.oracle_powf_array(x.ctypes.data, np.float32(y), a.ctypes.data, x.size)
        o.oracle_powf_positive_array(x.ctypes.data, np.float32(y), b.ctypes.data, x.size)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_pow5_and_square_are_what_powf_returns():
    """Schlick terms and the GGX denominator call pow_f32(x, 5) / pow_f32(x, 2) in the reference
    (driver.c:204-214); the device code uses rt_pow5f(x) and x * x."""
    o = oracle_ffi.lib()
    o.oracle_powf_array.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_ssize_t]
    o.oracle_pow5_array.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t]
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-1.5, 1.5, 300000), [0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 3e7, np.inf, -np.inf]]).astype(np.float32)
    a, b = np.zeros_like(x), np.zeros_like(x)
    o.oracle_powf_array(x.ctypes.data, 5.0, a.ctypes.data, x.size)
    o.oracle_pow5_array(x.ctypes.data, b.ctypes.data, x.size)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    o.oracle_powf_array(x.ctypes.data, 2.0, a.ctypes.data, x.size)
    assert np.array_equal(a.view(np.uint32), (x * x).view(np.uint32))


def test_fused_sincos_is_bit_identical_to_the_separate_calls():
    """The kernels call rt_sincosf where the reference calls sin_f32 and cos_f32 (driver.c:119-123,239-240)."""
    o = oracle_ffi.lib()
    sig = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ssize_t]
    o.oracle_sincos_array.argtypes = o.oracle_sincos_fused_array.argtypes = sig
    x = np.random.default_rng(5).uniform(-10, 10, 200000).astype(np.float32)
    s1, c1, s2, c2 = (np.zeros_like(x) for _ in range(4))
    o.oracle_sincos_array(x.ctypes.data, s1.ctypes.data, c1.ctypes.data, x.size)
    o.oracle_sincos_fused_array(x.ctypes.data, s2.ctypes.data, c2.ctypes.data, x.size)
    assert np.array_equal(s1.view(np.uint32), s2.view(np.uint32)) and np.array_equal(c1.view(np.uint32), c2.view(np.uint32))


def test_resolve_matches_numpy_restatement():
    """raytracer.c:700-716: /spp, clamp, sRGB OETF (common.h:90-92), *255.999, truncate."""
    rng = np.random.default_rng(5)
    accum = (rng.uniform(0, 3, (50, 3)) ** 3).astype(np.float32)
    accum[0] = (0, 0.0031308 * 4, 1e9)
    accum[1] = (np.nan, -1.0, 4.0)
    got = oracle_ffi.resolve(accum, 4)
    v = accum * np.float32(1.0 / 4)
    v = np.where(np.isnan(v), np.float32(0), v)
    v = np.clip(v, 0, 1).astype(np.float32)
    enc = np.where(v <= np.float32(0.0031308), np.float32(12.92) * v,
                   np.float32(1.055) * np.power(v.astype(np.float64), 1 / 2.4).astype(np.float32) - np.float32(0.055))
    want = (enc.astype(np.float32) * np.float32(255.999)).astype(np.uint8)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1     # pow rounding may move one code
    assert (got == want).mean() > 0.97
    assert tuple(got[1]) == (0, 0, 255)


def numpy_denoise(img):
    """denoiser.c:47-127 restated with numpy float32, pixel by pixel."""
    f = np.float32
    h, w, _ = img.shape
    out = np.zeros_like(img)
    for y in range(h):
        for x in range(w):
            taps = []
            for oy in (-1, 0, 1):
                for ox in (-1, 0, 1):
                    c = img[min(max(y + oy, 0), h - 1), min(max(x + ox, 0), w - 1)].astype(f) / f(255.999)
                    lum = f(f(f(c[0] * f(0.2126)) + f(c[1] * f(0.7152))) + f(c[2] * f(0.0722)))
                    if oy == 0 and ox == 0:
                        centre = (c, lum)
                    pos = len(taps)
                    for i, t in enumerate(taps):
                        if t[1] > lum:
                            pos = i
                            break
                    taps.insert(pos, (c, lum))
            med = taps[4]
            mean = f(0)
            for t in taps[1:8]:
                mean = f(mean + t[1])
            mean = f(mean / f(7))
            noisy = abs(f(med[1] - mean))
            d = f(abs(f(med[1] - centre[1])) - f(noisy * f(5)))
            d = f(min(max(d, f(0)), f(0.0125)) / f(0.0125))
            res = centre[0] * f(f(1) - d) + med[0] * d
            out[y, x] = (res.astype(f) * f(255.999)).astype(np.uint8)
    return out


def test_denoiser_matches_numpy_restatement():
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (9, 11, 3), dtype=np.uint8)
    img[3:6, 3:8] = 90
    assert np.array_equal(oracle_ffi.denoise(img), numpy_denoise(img))
    flat = np.full((5, 5, 3), 77, dtype=np.uint8)
    assert np.array_equal(oracle_ffi.denoise(flat), flat)       # a flat image is a fixed point


def brute_force_hit(scene, origin, direction):
    """Every padded slot through scalar float32 Möller–Trumbore (raytracer.c:115-152), no BVH."""
    f = np.float32
    n = scene.triangles.len
    arr = [np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (n,)) for p in
           list(scene.triangles.x) + list(scene.triangles.y) + list(scene.triangles.z)]
    p = [np.stack([arr[v], arr[3 + v], arr[6 + v]], axis=1) for v in range(3)]
    with np.errstate(all="ignore"):
        e1, e2 = p[1] - p[0], p[2] - p[0]
        o, d = np.asarray(origin, f), np.asarray(direction, f)

        def cross(a, b):
            return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                             a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1).astype(f)

        def dot(a, b):
            return ((a[..., 0] * b[..., 0]).astype(f) + (a[..., 1] * b[..., 1]).astype(f) + (a[..., 2] * b[..., 2]).astype(f)).astype(f)
        pv = cross(np.broadcast_to(d, e2.shape), e2)
        inv = (f(1) / dot(e1, pv)).astype(f)
        tv = (o - p[0]).astype(f)
        qv = cross(tv, e1)
        u, v, t = inv * dot(tv, pv), inv * dot(np.broadcast_to(d, qv.shape), qv), inv * dot(e2, qv)
        eps = f(0.0001)
        miss = (u < -eps) | (u > f(1) + eps) | (v < -eps) | (u + v > f(1) + eps) | (t < eps)
        t = np.where(miss | ~(t > 0), np.inf, t)
    k = int(np.argmin(t))
    return (k, t[k]) if np.isfinite(t[k]) else (-1, np.inf)


@pytest.mark.parametrize("name", ["fov_test.obj", "sheen.glb"])
def test_traversal_finds_the_brute_force_closest_hit(name):
    """The ordered 8-ary traversal may return a different slot only on exact distance ties."""
    loaded = load(name)
    try:
        o = oracle_ffi.lib()
        rng = np.random.default_rng(8)
        eye = np.array([loaded.scene.camera.view_matrix[r][3] for r in range(3)], dtype=np.float32)
        hits = 0
        for _ in range(300):
            target = rng.uniform(-1.5, 1.5, 3).astype(np.float32)
            d = target - eye
            d = (d / np.linalg.norm(d)).astype(np.float32)
            t = C.c_float()
            slot = o.oracle_trace_ray(C.byref(loaded.scene), Ray(Vec3(*eye), Vec3(*d)), C.byref(t))
            want_slot, want_t = brute_force_hit(loaded.scene, eye, d)
            assert (slot < 0) == (want_slot < 0)
            if slot >= 0:
                hits += 1
                assert t.value == want_t
        assert hits > 30
    finally:
        loaded.close()


def test_axis_aligned_rays_with_zero_direction_components():
    """inv_dir = +-inf and 0*inf = NaN lanes (raytracer.c:198-227) must neither crash nor lose the hit."""
    loaded = load("fov_test.obj")
    try:
        o = oracle_ffi.lib()
        t = C.c_float()
        slot = o.oracle_trace_ray(C.byref(loaded.scene), Ray(Vec3(0.0, 0.0, 3.0), Vec3(0.0, 0.0, -1.0)), C.byref(t))
        want_slot, want_t = brute_force_hit(loaded.scene, (0, 0, 3), (0, 0, -1))
        assert (slot >= 0) == (want_slot >= 0) and (slot < 0 or t.value == want_t)
        slot = o.oracle_trace_ray(C.byref(loaded.scene), Ray(Vec3(0.0, 0.0, 3.0), Vec3(0.0, 1.0, 0.0)), C.byref(t))
        assert slot == -1 and math.isinf(t.value)
    finally:
        loaded.close()
