#!/bin/bash
# one quantizer iteration on the GPU box: index-exactness tests, the quick sweep, then the role profile of the debug build
# (3d-vq-vae-2_b200/build/ab/lib_dbg.so = the library built with -DVQ3D_VQT_DEBUG)
mkdir -p gpurun_out
L=3d-vq-vae-2_b200/vqvae/libvqvae3d_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantizer" 2>&1 | tail -5 | tee gpurun_out/vq_iter_tests.log
timeout 300 python tools/bench_quantizer.py --quick 2>&1 | tee gpurun_out/vq_iter_bench.log | tail -30
if [ -f 3d-vq-vae-2_b200/build/ab/lib_dbg.so ]; then
  cp $L /tmp/lib_keep.so; cp 3d-vq-vae-2_b200/build/ab/lib_dbg.so $L
  timeout 300 python tools/debug_vq.py 2>&1 | tee gpurun_out/vq_iter_debug.log
  cp /tmp/lib_keep.so $L
fi
