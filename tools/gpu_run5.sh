set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log; tail -4 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_final.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_final.log; tail -2 gpurun_out/r2_bench_final.log | cut -c1-300
cd /tmp; /root/repo/raytracing_c_b200/host/rt_driver -W 1920 -H 1080 -S 1024 -V -D /root/repo/assets/models/helmet.glb -O /tmp/helmet.png 2>&1 | tr '\r' '\n' | grep -v "^\[" | tee /root/repo/gpurun_out/r2_rt_driver_1gpu_final.log
