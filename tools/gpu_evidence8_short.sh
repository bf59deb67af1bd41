# The shorter 8-GPU run behind profiles/r02_bench_8gpu_{helmet,chunks,tower4k_1024,spheres_16}.log of the final kernels:
#   gpurun --gpus 8 --timeout 900 -- bash tools/gpu_evidence8_short.sh   (≈ 2 min of box time; tools/gpu_evidence.sh 8 is the long form)
set -x
N=8
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
          bench.py --gpus $N --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/r02_bench_${N}gpu_$name.log 2>&1; grep '^{' gpurun_out/r02_bench_${N}gpu_$name.log | cut -c1-160; }
run helmet
run chunks --split chunks
run tower4k_1024 --workload tower4k --spp 1024
run spheres_16 --workload spheres --spp 16
(cd /tmp && timeout 200 /root/repo/raytracing_c_b200/host/rt_driver -W 1920 -H 1080 -S 1024 -V -D --gpus $N /root/repo/assets/models/helmet.glb -O /tmp/helmet_$N.png 2>&1 | tr '\r' '\n' | grep -v "^\[" > /root/repo/gpurun_out/r02_rt_driver_${N}gpu_sample_split.log)
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_pytest_multi_${N}gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_multi_${N}gpu.log
