// vq3d_rt.h -- launch/err helpers shared by every kernel file of libvqvae3d_b200.
//
// The product is built by nvcc for sm_100a only.  Defining VQ3D_EMU (done ONLY by
// tests/emu/build_emu.py) swaps <cuda_runtime.h> for a host-thread SIMT emulator so the
// very same kernel sources can be logic-checked against the oracle on a box without a GPU;
// that build is test infrastructure, is never loaded by the Python package, and reports
// vq3d_is_cuda_build() == 0.
#pragma once

#ifdef VQ3D_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/vqvae3d_b200.h"

namespace vq3d {

inline char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

inline int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return VQ3D_OK;
    return fail(VQ3D_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// One launch helper for both builds.  Opts into >48 KB dynamic shared memory when needed.
template <typename... KArgs, typename... Args>
inline int launch(const char *name, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void *stream,
                  Args... args) {
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) return VQ3D_OK;
#ifdef VQ3D_EMU
    emu::launch(grid, block, smem, [&]() { kernel(static_cast<KArgs>(args)...); });
    return VQ3D_OK;
#else
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel),
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return check_cuda(e, name);
    }
    kernel<<<grid, block, smem, static_cast<cudaStream_t>(stream)>>>(static_cast<KArgs>(args)...);
    return check_cuda(cudaGetLastError(), name);
#endif
}

// dynamic shared memory, spelled so that the emulator can provide it too
#ifdef VQ3D_EMU
#define VQ3D_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::dyn_smem())
#else
#define VQ3D_DYN_SMEM(type, name) \
    extern __shared__ __align__(16) unsigned char name##_raw_[]; \
    type *name = reinterpret_cast<type *>(name##_raw_)
#endif

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float ld_scalar(const float *p, float dflt) { return p ? __ldg(p) : dflt; }

// ELU(alpha=1), F.elu of layers.py:117 / model.py:120.  exp(x)-1 like ATen's kernel.
__device__ __forceinline__ float elu1(float v) { return v > 0.0f ? v : (__expf(v) - 1.0f); }

// Packed fp32 FMA (sm_100 `fma.rn.f32x2` -> SASS FFMA2 with a scalar-broadcast multiplicand): two independent
// round-to-nearest FMAs in ONE issue slot, bit-identical to two __fmaf_rn.  The thin-channel SIMT kernels are
// instruction-issue bound, so pairing two output channels per instruction frees issue slots.
#ifndef VQ3D_FFMA2
#define VQ3D_FFMA2 1
#endif
__device__ __forceinline__ float2 ffma2_bcast(float2 w, float x, float2 acc) {
#if defined(VQ3D_EMU) || !VQ3D_FFMA2
    acc.x = __fmaf_rn(w.x, x, acc.x);
    acc.y = __fmaf_rn(w.y, x, acc.y);
    return acc;
#else
    unsigned long long rw, rx, ra, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(rw) : "f"(w.x), "f"(w.y));
    asm("mov.b64 %0, {%1, %1};" : "=l"(rx) : "f"(x));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(acc.x), "f"(acc.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rw), "l"(rx), "l"(ra));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
    return r;
#endif
}

template <bool PACKED>
__device__ __forceinline__ float2 ffma2_bcast_if(float2 w, float x, float2 acc) {
    if constexpr (PACKED) return ffma2_bcast(w, x, acc);
    acc.x = __fmaf_rn(w.x, x, acc.x);
    acc.y = __fmaf_rn(w.y, x, acc.y);
    return acc;
}

__host__ __device__ __forceinline__ int wrap(int i, int n) {  // circular index, |i| < 2n
    return i < 0 ? i + n : (i >= n ? i - n : i);
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace vq3d
