set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_b.log 2>&1; grep '^{' gpurun_out/r2_bench_b.log | cut -c1-220
cd /tmp
for pin in 0 1; do
timeout 300 /root/repo/raytracing_c_b200/host/rt_driver -W 1920 -H 1080 -S 1024 -V -D --pinned $pin /root/repo/assets/models/helmet.glb -O /tmp/helmet.png 2>&1 | tr '\r' '\n' | grep -v "^\[" | tail -8
done
