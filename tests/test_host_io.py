"""Host-side C services (librt_host.so): codecs, loaders, camera.  No GPU."""
import ctypes as C
import io
import json
import os
import struct

import numpy as np
import pytest

from helpers import MODELS
from raytracing_c_b200 import driver
from raytracing_c_b200._ffi import Camera, Image, Model, host_lib

PIL = pytest.importorskip("PIL.Image")


def image_array(im: Image) -> np.ndarray:
    return np.frombuffer(C.string_at(im.pixels.data, im.pixels.len), dtype=np.uint8).reshape(im.height, im.width, im.components)


def test_background_is_deterministic_and_round_trips_through_every_encoder(tmp_path):
    host = host_lib()
    a, b = Image(), Image()
    host.rt_generate_background(C.byref(a), 256, 128)
    host.rt_generate_background(C.byref(b), 256, 128)
    px = image_array(a)
    assert np.array_equal(px, image_array(b)) and px.std() > 10
    for ext in ("png", "qoi", "ppm"):
        path = str(tmp_path / f"bg.{ext}")
        assert host.rt_save_image(path.encode(), C.byref(a))
        assert np.array_equal(np.asarray(PIL.open(path).convert("RGB")), px), ext
    # stored-deflate PNG like the reference's writer: IDAT = rows*(1+3w) + 5 per 64K block + zlib framing
    assert os.path.getsize(str(tmp_path / "bg.png")) > 128 * (1 + 3 * 256)
    back = Image()
    assert host.rt_load_texture(str(tmp_path / "bg.png").encode(), C.byref(back))
    assert np.array_equal(image_array(back), px)
    for im in (a, b, back):
        host.rt_image_free(C.byref(im))


def test_png_decoder_filters_and_rgba(tmp_path):
    rng = np.random.default_rng(0)
    rgba = rng.integers(0, 256, (33, 47, 4), dtype=np.uint8)
    rgba[:, :, :3] = (np.add.outer(np.arange(33), np.arange(47))[:, :, None] * [3, 5, 7]) % 256   # filter-friendly gradients
    p = str(tmp_path / "t.png")
    PIL.fromarray(rgba).save(p, optimize=True)
    im = Image()
    assert host_lib().rt_load_texture(p.encode(), C.byref(im))
    assert im.components == 4 and np.array_equal(image_array(im), rgba)
    host_lib().rt_image_free(C.byref(im))


def test_baseline_jpeg_decoder_tracks_libjpeg():
    data = open(os.path.join(MODELS, "helmet.glb"), "rb").read()
    clen = struct.unpack("<I", data[12:16])[0]
    doc = json.loads(data[20:20 + clen])
    off = 20 + clen + 8
    bv = doc["bufferViews"][doc["images"][0]["bufferView"]]
    jpg = data[off + bv["byteOffset"]: off + bv["byteOffset"] + bv["byteLength"]]
    im = Image()
    assert host_lib().rt_image_decode(jpg, len(jpg), C.byref(im))
    mine = image_array(im).astype(int)
    ref = np.asarray(PIL.open(io.BytesIO(jpg)).convert("RGB")).astype(int)
    assert mine.shape == ref.shape == (2048, 2048, 3)
    diff = np.abs(mine - ref)
    assert diff.mean() < 0.1 and diff.max() <= 4
    host_lib().rt_image_free(C.byref(im))
    assert not host_lib().rt_image_decode(b"not an image", 12, C.byref(im))


@pytest.mark.parametrize("name,tris,mats,images", [
    ("quad.obj", 2, 2, 0), ("fov_test.obj", 72, 2, 0), ("tower.obj", 4320, 1, 0),
    ("spheres.glb", 4800, 6, 0), ("sheen.glb", 1920, 2, 0), ("helmet.glb", 15452, 2, 4)])
def test_model_loaders(name, tris, mats, images):
    m, cam = Model(), Camera()
    host = host_lib()
    host.rt_camera_default(C.byref(cam))
    assert host.rt_load_model_file(os.path.join(MODELS, name).encode(), 0x1234, C.byref(m), C.byref(cam))
    assert (m.triangles.len, m.n_materials, m.n_images) == (tris, mats, images)
    assert all(m.triangles.data[i].shader.proc == 0x1234 for i in range(0, tris, max(1, tris // 7)))
    if name.endswith(".glb"):
        assert cam.view_matrix[3][3] == 1.0 and abs(cam.focal_length - 1 / np.tan(cam.fov / 2)) < 1e-5
    else:       # OBJ leaves the default camera untouched (driver.c:510-587, 765-767)
        assert [cam.view_matrix[r][3] for r in range(3)] == [0.0, 0.0, 3.0]
        assert abs(cam.fov - np.deg2rad(70)) < 1e-6
    if name == "helmet.glb":
        mat = m.materials[0]
        assert mat.texture_albedo and mat.texture_normal and mat.texture_metal_roughness and mat.texture_emission
        assert (mat.roughness, mat.metalness, mat.normal_map_strength) == (1.0, 1.0, 1.0)   # glTF 2.0 defaults
        assert mat.texture_albedo.contents.width == 2048
    if name == "spheres.glb":
        assert abs(m.materials[2].metalness - 0.33070865) < 1e-6 and m.materials[1].metalness == 1.0
    host.rt_model_free(C.byref(m))


def test_missing_file_and_bad_extension():
    m, cam = Model(), Camera()
    assert not host_lib().rt_load_model_file(b"/nonexistent/model.obj", 0, C.byref(m), C.byref(cam))
    assert not host_lib().rt_load_model_file(b"model.stl", 0, C.byref(m), C.byref(cam))


def test_look_at_camera():
    cam = driver.look_at((3.0, 0.0, 0.0), (0.0, 0.0, 0.0))
    m = np.array([[cam.view_matrix[r][c] for c in range(4)] for r in range(4)])
    assert np.allclose(m[:3, 3], [3, 0, 0]) and np.allclose(-m[:3, 2], [-1, 0, 0])     # looks down -Z of the camera
    assert np.allclose(m[:3, :3] @ m[:3, :3].T, np.eye(3), atol=1e-6)


def _try_load(path):
    """rt_load_model_file on a hostile file: must return (False or True) without touching memory it does not own."""
    import ctypes as C
    from raytracing_c_b200._ffi import Camera, Model, host_lib
    host = host_lib()
    model, cam = Model(), Camera()
    ok = host.rt_load_model_file(str(path).encode(), None, C.byref(model), C.byref(cam))
    if ok:
        host.rt_model_free(C.byref(model))
    return bool(ok)


def test_malformed_glb_is_rejected_not_read_out_of_bounds(tmp_path):
    """the GLB chunk length, accessor offsets/counts/strides and image bufferViews all come from the file."""
    import json
    import struct
    good = open(os.path.join(MODELS, "spheres.glb"), "rb").read()
    # (1) JSON chunk length larger than the file
    bad = bytearray(good)
    struct.pack_into("<I", bad, 12, len(good) * 4)
    (tmp_path / "chunk.glb").write_bytes(bad)
    assert not _try_load(tmp_path / "chunk.glb")
    # (2) truncations at many lengths: never a crash, and a load that survives must free cleanly
    for cut in (12, 19, 20, 21, 200, len(good) // 3, len(good) // 2, len(good) - 5):
        (tmp_path / "cut.glb").write_bytes(good[:cut])
        _try_load(tmp_path / "cut.glb")
    # (3) hostile numbers in the JSON: negative offsets, counts, strides, buffer indices
    clen = struct.unpack_from("<I", good, 12)[0]
    doc = json.loads(good[20:20 + clen])
    rest = good[20 + clen:]

    def rebuild(d):
        js = json.dumps(d).encode()
        js += b" " * (-len(js) % 4)
        body = struct.pack("<II", len(js), 0x4E4F534A) + js + rest
        return b"glTF" + struct.pack("<II", 2, 12 + len(body)) + body

    for mutate in (lambda d: d["accessors"][0].update(byteOffset=-4096),
                   lambda d: d["accessors"][0].update(count=-7),
                   lambda d: d["accessors"][0].update(count=2 ** 40),
                   lambda d: d["bufferViews"][0].update(byteOffset=-1),
                   lambda d: d["bufferViews"][0].update(byteOffset=2 ** 62),
                   lambda d: d["bufferViews"][0].update(byteStride=-12),
                   lambda d: d["bufferViews"][0].update(buffer=-3),
                   lambda d: d["bufferViews"][0].update(buffer=99)):
        d = json.loads(json.dumps(doc))
        mutate(d)
        (tmp_path / "hostile.glb").write_bytes(rebuild(d))
        _try_load(tmp_path / "hostile.glb")          # may load (the primitive is skipped) or fail; must not crash
    (tmp_path / "same.glb").write_bytes(rebuild(doc))
    assert _try_load(tmp_path / "same.glb")
