set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_multi_${N}gpu.log; tail -3 gpurun_out/r2_pytest_multi_${N}gpu.log
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/r2_bench_${N}gpu_$name.log 2>&1; grep '^{' gpurun_out/r2_bench_${N}gpu_$name.log | cut -c1-160; }
run samples
run chunks --split chunks --no-cpu
run fast --fast --no-cpu
cd /tmp
for extra in "" "--split chunks"; do
timeout 300 /root/repo/raytracing_c_b200/host/rt_driver -W 1920 -H 1080 -S 1024 -V -D --gpus $N $extra /root/repo/assets/models/helmet.glb -O /tmp/helmet_$N.png 2>&1 | tr '\r' '\n' | grep -v "^\[" > /root/repo/gpurun_out/r2_rt_driver_${N}gpu_$(echo $extra | tr -d ' -')_final.log; grep -E "GPU init|ms$|GPUs:|Time to" /root/repo/gpurun_out/r2_rt_driver_${N}gpu_$(echo $extra | tr -d ' -')_final.log
done
