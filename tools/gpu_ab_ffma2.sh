#!/bin/bash
# A/B of row-kernel builds: tools/gpu_ab_ffma2.sh "<variants>" "<cases>" [batch]   (variants = 3d-vq-vae-2_b200/build/ab/lib_<v>.so)
# Build a variant here (no GPU needed) by recompiling one file with a knob and relinking, e.g.
#   cd 3d-vq-vae-2_b200 && mkdir -p build/ab && python build.py &&
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true -DVQ3D_FFMA2=0 \
#        -c csrc/preact_row_kernels.cu -o build/ab/row_A.o &&
#   nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o build/ab/lib_A.so \
#        $(ls build/*.o | grep -v preact_row_kernels.o) build/ab/row_A.o && cp vqvae/libvqvae3d_b200.so build/ab/lib_B.so
# Knobs: VQ3D_FFMA2 (packed FMAs), VQ3D_ROW_PACK_C8, VQ3D_ROW_NZ8 (8 z per thread), VQ3D_ROW_NW2 / VQ3D_UP_NW2 (two output
# rows per thread), VQ3D_FUSED_FFMA2.  Results of the round-1 runs: profiles/r01x_ffma2_ab.txt.
L=3d-vq-vae-2_b200/vqvae/libvqvae3d_b200.so
cp $L /tmp/lib_keep.so
for v in ${1:-A B C}; do
  cp 3d-vq-vae-2_b200/build/ab/lib_$v.so $L
  for c in ${2:-stack4_512 stack8_256 up8_256}; do
    echo "variant $v: $(python tools/prof_case.py $c --reps 20 --batch ${3:-2} 2>&1 | tail -1)"
  done
done 2>&1 | tee -a gpurun_out/ab_ffma2.log
cp /tmp/lib_keep.so $L
