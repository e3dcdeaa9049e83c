#!/bin/bash
# A/B of row-kernel builds: tools/gpu_ab_ffma2.sh "<variants>" "<cases>" [batch]   (variants = 3d-vq-vae-2_b200/build/ab/lib_<v>.so)
L=3d-vq-vae-2_b200/vqvae/libvqvae3d_b200.so
cp $L /tmp/lib_keep.so
for v in ${1:-A B C}; do
  cp 3d-vq-vae-2_b200/build/ab/lib_$v.so $L
  for c in ${2:-stack4_512 stack8_256 up8_256}; do
    echo "variant $v: $(python tools/prof_case.py $c --reps 20 --batch ${3:-2} 2>&1 | tail -1)"
  done
done 2>&1 | tee -a gpurun_out/ab_ffma2.log
cp /tmp/lib_keep.so $L
