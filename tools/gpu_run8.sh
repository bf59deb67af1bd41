set -x
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_multi_${N}gpu.log; tail -3 gpurun_out/r2_pytest_multi_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_${N}gpu.log
grep '^{' gpurun_out/r2_bench_${N}gpu.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --split chunks --no-cpu > gpurun_out/r2_bench_${N}gpu_chunks.log 2>&1
grep '^{' gpurun_out/r2_bench_${N}gpu_chunks.log | cut -c1-200
cd /tmp
for extra in "" "--split chunks"; do
timeout 300 /root/repo/raytracing_c_b200/host/rt_driver -W 1920 -H 1080 -S 1024 -V -D --gpus $N $extra /root/repo/assets/models/helmet.glb -O /tmp/helmet_$N.png 2>&1 | tr '\r' '\n' | grep -v "^\[" > /root/repo/gpurun_out/r2_rt_driver_${N}gpu_$(echo $extra | tr -d ' -').log; cat /root/repo/gpurun_out/r2_rt_driver_${N}gpu_$(echo $extra | tr -d ' -').log
done
