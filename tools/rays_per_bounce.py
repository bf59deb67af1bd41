"""Rays per bounce of the headline frame (counter differences between renders with max_bounces = 1..8)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raytracing_c_b200 import driver, gpu_lib
from raytracing_c_b200._ffi import gpu_check
gpu = gpu_lib(); gpu_check(gpu.rt_gpu_init(0)); torch.cuda.set_device(0)
loaded = driver.load_scene(os.path.join(ROOT, "assets", "models", sys.argv[1] if len(sys.argv) > 1 else "helmet.glb"))
driver.register_callbacks(loaded)
scene = C.byref(loaded.scene)
W, H, SPP = 1920, 1080, 64
accum = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
prev = None
for B in range(1, 9):
    ctr = torch.zeros(8, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gpu_check(gpu.rt_gpu_render_accum_device(scene, W, H, 0, SPP, B, 0, 0, accum.data_ptr(), None, None, ctr.data_ptr(), None))
    e1.record(); torch.cuda.synchronize()
    c = ctr.cpu().tolist()
    d = [a - b for a, b in zip(c, prev)] if prev else c
    print(f"bounce {B - 1}: rays {d[0]:>10d} nodes {d[1]:>11d} leaves {d[2]:>10d}  nodes/ray {d[1] / max(d[0], 1):5.2f} leaves/ray {d[2] / max(d[0], 1):5.2f}  frame {e0.elapsed_time(e1):.2f} ms")
    prev = c
loaded.close()
