#!/bin/bash
# A/B of quantizer builds: tools/gpu_ab_vq.sh "<variants>"   (3d-vq-vae-2_b200/build/ab/lib_<v>.so)
L=3d-vq-vae-2_b200/vqvae/libvqvae3d_b200.so
cp $L /tmp/lib_keep.so
for v in $1; do
  cp 3d-vq-vae-2_b200/build/ab/lib_$v.so $L
  echo "== variant $v"
  python tools/bench_quantizer.py --quick 2>&1 | grep -E "K= 512 D= 32|K=1024 D= 64|K=4096 D=128|K= 512 D=128"
done 2>&1 | tee -a gpurun_out/ab_vq.log
cp /tmp/lib_keep.so $L
