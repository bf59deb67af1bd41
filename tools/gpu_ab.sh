# A/B of kernel variants: tools/gpu_ab.sh <spp> <reps> name1 name2 ...   ("base" = the in-tree library)
spp=$1; reps=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset RT_GPU_LIB; else export RT_GPU_LIB=$PWD/raytracing_c_b200/csrc/variants/libraytracer_gpu_$v.so; fi
  echo "== $v"
  timeout 300 python tools/profile_render.py --spp $spp --reps $reps --stages 2>&1 | tail -7 | tee gpurun_out/ab_$v.log
done
