import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from raytracing_c_b200 import driver, gpu_lib
from raytracing_c_b200._ffi import gpu_check
gpu_check(gpu_lib().rt_gpu_init(0))
loaded = driver.load_scene("/root/repo/assets/models/helmet.glb")
w, h, spp = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 8
def r(**o):
    driver.set_options(**o)
    driver.render(loaded, w, h, spp, 8)
    return driver.read_accum(w, h)
full = r()
for world in (2, 8):
    tot = np.zeros_like(full)
    for k in range(world):
        tot += r(pixel_rank=k, pixel_world=world)
    bad = np.argwhere((tot != full).any(axis=-1))
    print("world", world, "bad pixels", len(bad))
    if len(bad):
        ys, xs = bad[:, 0], bad[:, 1]
        print(" y range", ys.min(), ys.max(), "x range", xs.min(), xs.max())
        cid = (ys // 32) * 60 + xs // 32
        print(" chunks", np.unique(cid)[:20], "ratio", (tot[ys[0], xs[0]] / np.maximum(full[ys[0], xs[0]], 1e-9)))
