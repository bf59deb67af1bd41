#!/bin/bash
# The GPU runs behind profiles/r02_*: run under `gpurun` (one B200, or `gpurun --gpus 8` with N=8 as first argument).
#   gpurun --timeout 3000 -- 'bash tools/gpu_evidence.sh'            tests, bench, reference arm, ncu launch list + traffic capture
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_evidence.sh 8' multi-GPU tests, 8-GPU bench lines, the C host with --gpus 8
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).  A number printed under ncu is never a bench value.
set -x
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log; tail -3 gpurun_out/r02_pytest.log
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_1gpu.log 2>&1; tail -1 gpurun_out/r02_bench_1gpu.log | cut -c1-200
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --fast > gpurun_out/r02_bench_1gpu_fast.log 2>&1; tail -1 gpurun_out/r02_bench_1gpu_fast.log | cut -c1-200
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.log 2>&1; tail -1 gpurun_out/r02_bench_ref.log | cut -c1-200
  python tools/config_sweep.py > gpurun_out/r02_config_sweep_1gpu.jsonl
  CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-parity"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
  timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum \
      --clock-control none -c 1500 --csv --log-file gpurun_out/r02_metrics_bench.csv $CMD > gpurun_out/r02_ncu_metrics.log 2>&1
  # then, here:  python tools/ncu_traffic.py gpurun_out/r02_metrics_bench.csv profiles/r02_trace_traffic.json "<the command>"
  (cd /tmp && /root/repo/raytracing_c_b200/host/rt_driver -W 1920 -H 1080 -S 1024 -V -D /root/repo/assets/models/helmet.glb -O /tmp/helmet.png 2>&1 | tr '\r' '\n' | grep -v "^\[" > /root/repo/gpurun_out/r02_rt_driver_1gpu.log)
else
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_pytest_multi_${N}gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_multi_${N}gpu.log
  run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
            bench.py --gpus $N --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/r02_bench_${N}gpu_$name.log 2>&1; grep '^{' gpurun_out/r02_bench_${N}gpu_$name.log | cut -c1-160; }
  run helmet
  run chunks --split chunks
  run nccl --reduce nccl
  run fast --fast
  run tower4k_16 --workload tower4k --spp 16
  run tower4k_1024 --workload tower4k --spp 1024
  run spheres_16 --workload spheres --spp 16
  for extra in "" "--split chunks" "--reduce nccl"; do
    tag=$(echo $extra | tr -d ' -'); tag=${tag:-samplesplit}
    (cd /tmp && timeout 300 /root/repo/raytracing_c_b200/host/rt_driver -W 1920 -H 1080 -S 1024 -V -D --gpus $N $extra /root/repo/assets/models/helmet.glb -O /tmp/helmet_$N.png 2>&1 | tr '\r' '\n' | grep -v "^\[" > /root/repo/gpurun_out/r02_rt_driver_${N}gpu_$tag.log)
  done
fi
