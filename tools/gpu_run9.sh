set -x
N=${1:-8}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/r2_sweep_${N}gpu_$name.log 2>&1; grep '^{' gpurun_out/r2_sweep_${N}gpu_$name.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$name', d['value'], d['ms_per_step'], d['config']['parallelism'][:40], d['parity']['bit_identical'], d['parity']['max_rel'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['upload_ms'])
"; }
run helmet
run tower4k_16 --workload tower4k --spp 16
run tower4k_1024 --workload tower4k --spp 1024
run spheres_16 --workload spheres --spp 16
