"""Times rt_gpu_scene_upload (helmet.glb: 62 MB host -> device, RGBA8 repack on the device) a few times."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from raytracing_c_b200 import driver, gpu_lib
from raytracing_c_b200._ffi import gpu_check
gpu = gpu_lib(); gpu_check(gpu.rt_gpu_init(0))
loaded = driver.load_scene("assets/models/helmet.glb"); driver.register_callbacks(loaded)
scene = C.byref(loaded.scene)
for i in range(6):
    t0 = time.perf_counter(); gpu_check(gpu.rt_gpu_scene_upload(scene)); torch.cuda.synchronize(); print(f"upload {1e3*(time.perf_counter()-t0):.2f} ms")
