// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal host-thread SIMT emulator: lets g++ compile the product's kernel sources
// (3d-vq-vae-2_b200/csrc/*.cu with -DVQ3D_EMU) so that their indexing/fusion logic can be
// checked against the oracle on a machine without a GPU.  One OS thread per CUDA thread,
// one block at a time, std::barrier for __syncthreads and warp collectives.  It is slow
// (tiny shapes only), cannot run tcgen05/TMA code, and is never loaded by the product
// package (which dlopens only the nvcc-built library and refuses to run without CUDA).
//
// Rules the kernels follow so that this works: no thread returns before the last
// __syncthreads / warp collective of the kernel; warp collectives are called by all lanes.
#pragma once

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }

typedef int cudaError_t;
typedef void *cudaStream_t;
constexpr cudaError_t cudaSuccess = 0;
inline const char *cudaGetErrorString(cudaError_t) { return "emulator"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

namespace emu {

struct WarpState {
    int n;
    uint64_t scratch[32];
    std::barrier<> bar;
    explicit WarpState(int n_) : n(n_), bar(n_) { memset(scratch, 0, sizeof(scratch)); }
};
struct BlockState {
    std::barrier<> bar;
    explicit BlockState(int n) : bar(n) {}
};
struct Ctx {
    dim3 tid, bid;
    int lane = 0;
    WarpState *w = nullptr;
    BlockState *b = nullptr;
};
inline thread_local Ctx ctx;
inline dim3 g_blockDim, g_gridDim;
inline unsigned char *g_dyn = nullptr;
inline unsigned char *dyn_smem() { return g_dyn; }

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &fn) {
    g_blockDim = block;
    g_gridDim = grid;
    const int nthreads = int(block.x * block.y * block.z);
    std::vector<unsigned char> dyn(smem + 256);
    g_dyn = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(dyn.data()) + 127) & ~uintptr_t(127));
    BlockState bs(nthreads);
    std::vector<std::unique_ptr<WarpState>> ws;
    for (int w = 0; w * 32 < nthreads; ++w) ws.emplace_back(new WarpState(std::min(32, nthreads - w * 32)));
    auto worker = [&](int t) {
        ctx.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        ctx.lane = t % 32;
        ctx.w = ws[t / 32].get();
        ctx.b = &bs;
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    ctx.bid = dim3(bx, by, bz);
                    fn();
                    bs.bar.arrive_and_wait();  // next block reuses the static __shared__ storage
                }
    };
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto &t : th) t.join();
    g_dyn = nullptr;
}

template <typename T>
inline T shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    WarpState *w = ctx.w;
    memcpy(&w->scratch[ctx.lane], &v, sizeof(T));
    w->bar.arrive_and_wait();
    if (src < 0 || src >= w->n) src = ctx.lane;
    T r;
    memcpy(&r, &w->scratch[src], sizeof(T));
    w->bar.arrive_and_wait();
    return r;
}
}  // namespace emu

#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
constexpr int warpSize = 32;

inline void __syncthreads() { emu::ctx.b->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::ctx.w->bar.arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl(v, src); }
// width < 32: the warp is split into segments of `width` lanes, src is relative to the caller's segment
template <typename T> inline T __shfl_sync(unsigned, T v, int src, int width) {
    return emu::shfl(v, (emu::ctx.lane / width) * width + (src % width));
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl(v, emu::ctx.lane ^ m); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) { return emu::shfl(v, emu::ctx.lane + d); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, int d) { return emu::shfl(v, emu::ctx.lane - d); }
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::WarpState *w = emu::ctx.w;
    w->scratch[emu::ctx.lane] = pred ? 1 : 0;
    w->bar.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < w->n; ++i) r |= unsigned(w->scratch[i] & 1) << i;
    w->bar.arrive_and_wait();
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }

// atomics
inline float atomicAdd(float *p, float v) {
    uint32_t *u = reinterpret_cast<uint32_t *>(p), old = __atomic_load_n(u, __ATOMIC_RELAXED), nw;
    float f;
    do { memcpy(&f, &old, 4); f += v; memcpy(&nw, &f, 4); } while (!__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    memcpy(&f, &old, 4);
    return f;
}
inline double atomicAdd(double *p, double v) {
    uint64_t *u = reinterpret_cast<uint64_t *>(p), old = __atomic_load_n(u, __ATOMIC_RELAXED), nw;
    double f;
    do { memcpy(&f, &old, 8); f += v; memcpy(&nw, &f, 8); } while (!__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    memcpy(&f, &old, 8);
    return f;
}
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T emu_atomic_minmax(T *p, T v, bool want_min) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while ((want_min ? v < old : v > old) && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
    return old;
}
inline int atomicMin(int *p, int v) { return emu_atomic_minmax(p, v, true); }
inline int atomicMax(int *p, int v) { return emu_atomic_minmax(p, v, false); }
inline unsigned atomicMin(unsigned *p, unsigned v) { return emu_atomic_minmax(p, v, true); }
inline unsigned atomicMax(unsigned *p, unsigned v) { return emu_atomic_minmax(p, v, false); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// math / intrinsics (compile the emulator with -ffp-contract=off so that a*b+c is never fused)
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __expf(float a) { return expf(a); }
inline float __fdividef(float a, float b) { return a / b; }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline void __stcs(T *p, T v) { *p = v; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
using std::max;
using std::min;
