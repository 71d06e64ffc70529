mkdir -p gpurun_out
for v in 256_2 256_4 512_2 512_4; do
  cp build/variants/lib_$v.so deepcoro_clip_b200/libb200clip.so
  timeout 200 python tools/gpu_xfblock_time.py $v 2>&1 | grep -v Warning | tail -2
done 2>&1 | tee gpurun_out/xfblock_variants.log
timeout 300 python tools/gpu_check_xfblock.py 2>&1 | grep -v Warning | tail -12
