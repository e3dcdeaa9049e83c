#!/bin/bash
# role profile of the quantizer kernels: tools/gpu_vq_dbg.sh "<variants>"  (3d-vq-vae-2_b200/build/ab/lib_<v>.so built with -DVQ3D_VQT_DEBUG ...)
L=3d-vq-vae-2_b200/vqvae/libvqvae3d_b200.so
cp $L /tmp/lib_keep.so
for v in $1; do
  cp 3d-vq-vae-2_b200/build/ab/lib_$v.so $L
  echo "== variant $v"
  timeout 300 python tools/debug_vq.py 2>&1
done 2>&1 | tee gpurun_out/vq_dbg.log
cp /tmp/lib_keep.so $L
