#!/bin/bash
# A/B of the packed-FMA (FFMA2) row kernels: tools/gpu_ab_ffma2.sh  (variants built into 3d-vq-vae-2_b200/build/ab/)
L=3d-vq-vae-2_b200/vqvae/libvqvae3d_b200.so
for v in A B C; do
  cp 3d-vq-vae-2_b200/build/ab/lib_$v.so $L
  for c in stack4_512 stack8_256 up8_256; do
    echo "variant $v: $(python tools/prof_case.py $c --reps 5 --batch 2 2>&1 | tail -1)"
  done
done 2>&1 | tee gpurun_out/ab_ffma2.log
cp 3d-vq-vae-2_b200/build/ab/lib_B.so $L
