set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench1.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench1.log
tail -3 gpurun_out/r2_bench1.log
for v in default no all; do
  if [ $v = no ]; then export RT_GPU_NO_L2_WINDOW=1; else unset RT_GPU_NO_L2_WINDOW; fi
  if [ $v = all ]; then export RT_GPU_L2_WINDOW=all; else unset RT_GPU_L2_WINDOW; fi
  timeout 300 python tools/profile_render.py --spp 64 --reps 6 --stages > gpurun_out/r2_l2_$v.log 2>&1
  tail -7 gpurun_out/r2_l2_$v.log
done
unset RT_GPU_NO_L2_WINDOW RT_GPU_L2_WINDOW
cd /tmp && timeout 300 /root/repo/raytracing_c_b200/host/rt_driver -W 1920 -H 1080 -S 1024 -V -D /root/repo/assets/models/helmet.glb -O /tmp/helmet.png > /root/repo/gpurun_out/r2_rt_driver_1gpu.log 2>&1; cat /root/repo/gpurun_out/r2_rt_driver_1gpu.log
