set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_final.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_final.log; tail -2 gpurun_out/r2_bench_final.log | cut -c1-200
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-parity"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1; tail -1 gpurun_out/r2_ncu_launches.log | cut -c1-200
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_metrics_bench.csv $CMD > gpurun_out/r2_ncu_metrics.log 2>&1; tail -1 gpurun_out/r2_ncu_metrics.log | cut -c1-200
wc -l gpurun_out/r02_launches_bench.csv gpurun_out/r02_metrics_bench.csv
