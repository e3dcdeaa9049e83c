#!/bin/bash
# A/B of library builds on prof_case cases: tools/gpu_ab_case.sh "<variants>" "<cases>" [batch]   (3d-vq-vae-2_b200/build/ab/lib_<v>.so)
L=3d-vq-vae-2_b200/vqvae/libvqvae3d_b200.so
cp $L /tmp/lib_keep.so
for v in $1; do
  cp 3d-vq-vae-2_b200/build/ab/lib_$v.so $L
  for c in $2; do
    echo "variant $v: $(python tools/prof_case.py $c --reps 10 --batch ${3:-8} 2>&1 | tail -1)"
  done
done 2>&1 | tee -a gpurun_out/ab_case.log
cp /tmp/lib_keep.so $L
