// tc_microbench.cu -- single-SM micro-measurements that the kernel designs in DESIGN.md lean on (round 2):
//   ldtm   tcgen05.ld throughput (32x32b .x16/.x32/.x64/.x128, 4..16 warps, 1 or 2 loads in flight)  -> TMEM read-out ceiling of the quantizer
//   umma   cost of one tcgen05.mma M128 x N x K16 (bf16, SS) vs N and vs the A-operand layout
//          (no-swizzle K-major with far-apart K chunks = the conv kernels' shifted-window layout, dense no-swizzle, SWIZZLE_128B)
//   hmma   legacy mma.sync.m16n8k16 bf16 throughput per SM (what a register-operand implicit GEMM could reach on thin channels)
//   lds    shared-memory read bandwidth seen by LDS.128 (the A-operand fetch competes with it)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench/tc_microbench tools/microbench/tc_microbench.cu
// Run (B200): tools/microbench/tc_microbench > profiles/r02_tc_microbench.txt
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../3d-vq-vae-2_b200/csrc/tc_common.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t *slot, int warp) {
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t base, int warp) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}

// ---- tcgen05.ld ---------------------------------------------------------------------------------------------------------
template <int N> __device__ __forceinline__ uint32_t ldtm_x(uint32_t taddr);
template <> __device__ __forceinline__ uint32_t ldtm_x<16>(uint32_t taddr) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) a ^= r[i];
    return a;
}
#define R8(b) "%" #b
template <> __device__ __forceinline__ uint32_t ldtm_x<32>(uint32_t taddr) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                   "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) a ^= r[i];
    return a;
}
template <> __device__ __forceinline__ uint32_t ldtm_x<64>(uint32_t taddr) {
    uint32_t r[64];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                   "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                   "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
                   "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
                   "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]) : "r"(taddr) : "memory");
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) a ^= r[i];
    return a;
}
// 16x256b.x8: the other common accumulator read shape (128 bits x 2 per lane row pair), 32 registers
template <> __device__ __forceinline__ uint32_t ldtm_x<256>(uint32_t taddr) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                   "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) a ^= r[i];
    return a;
}

// N = registers per lane per load for the 32x32b shapes (16/32/64); 256 = the 16x256b.x8 shape (32 registers)
template <int N, int INFLIGHT>
__global__ void ldtm_kernel(int iters, unsigned long long *cycles, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = tmem_alloc_all(&slot, warp);
    constexpr int NCOL = N == 256 ? 64 : N;          // columns one load covers
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < INFLIGHT; ++j) {
            const uint32_t col = (uint32_t)(((warp >> 2) * INFLIGHT + j + i) * NCOL) & 511u & ~(uint32_t)(NCOL - 1);
            acc ^= ldtm_x<N>(base + lane_sel + col);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) *sink = acc;
    tmem_free_all(base, warp);
}

// ---- tcgen05.mma cost -----------------------------------------------------------------------------------------------------
// layout 0: K-major no-swizzle, K chunks LBO = 48 KB apart... (clamped to the buffer), rows shifted per "tap" like the conv kernels
// layout 1: K-major no-swizzle, dense (LBO = 2048: 128 rows x 16 B per K chunk)
// layout 2: K-major SWIZZLE_128B (128-byte rows, SBO = 1024), K16 steps advance the start address by 32 B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           ((uint64_t)((saddr >> 7) & 7) << 49) | (2ull << 61);
}

template <int LAYOUT>
__global__ void __launch_bounds__(128, 1)
umma_kernel(int N, int iters, int reps, unsigned long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbarrier_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t base = tmem_alloc_all(&slot, warp);
    const uint32_t sA = (s_u32(smem) + 1023u) & ~1023u, sB = sA + 128 * 1024;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lbo_b = (uint32_t)N * 16;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        uint32_t ph = 0;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
                for (int r = 0; r < reps; ++r) {
#pragma unroll
                    for (int j = 0; j < 27; ++j) {      // compile-time operand addresses: the issuing thread spends ~2 instructions per MMA
                        uint64_t ad;
                        if (LAYOUT == 0)      ad = umma_desc(sA + (uint32_t)((j * 37 + 3) * 16), 60000u & ~15u, 128);   // shifted windows, far K chunks
                        else if (LAYOUT == 1) ad = umma_desc(sA + (uint32_t)((j & 7) * 4096), 2048, 128);
                        else                  ad = umma_desc_sw128(sA + (uint32_t)((j & 3) * 32) + (uint32_t)(((j >> 2) & 7) * 16384));
                        const uint64_t bd = umma_desc(sB + (uint32_t)(j * 512), lbo_b, 128);
                        umma_f16(base, ad, bd, idesc, (j > 0 || r > 0) ? 1u : 0u);
                    }
                }
                umma_commit_to(&bar);
            }
            __syncwarp();
            mbarrier_wait(&bar, ph); ph ^= 1u;
        }
        t1 = clock64();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
    tmem_free_all(base, warp);
}

// ---- legacy warp-level MMA -----------------------------------------------------------------------------------------------
__global__ void hmma_kernel(int iters, unsigned long long *cycles, float *sink) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    if (s == 1.2345f) *sink = s;
}

// ---- LDS.128 bandwidth ------------------------------------------------------------------------------------------------------
__global__ void lds_kernel(int iters, unsigned long long *cycles, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float4 *s = reinterpret_cast<float4 *>(smem);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 v = s[(threadIdx.x + j * 256 + it * 32) & 4095];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
    if (acc.x + acc.y + acc.z + acc.w == 1.2345f) *sink = acc.x;
}

template <int N, int INF>
static void run_ldtm(int warps, unsigned long long *dcy, uint32_t *dsink) {
    const int iters = 2000;
    ldtm_kernel<N, INF><<<1, warps * 32>>>(iters, dcy, dsink);
    CK(cudaDeviceSynchronize());
    ldtm_kernel<N, INF><<<1, warps * 32>>>(iters, dcy, dsink);
    CK(cudaDeviceSynchronize());
    unsigned long long cy;
    CK(cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost));
    const int regs = N == 256 ? 32 : N;
    const double bytes = (double)warps * 32 * regs * 4 * INF * iters;
    printf("ldtm shape=%-10s warps=%2d inflight=%d : %8.1f cycles/load/warp  %7.1f B/clk/SM\n", N == 256 ? "16x256b.x8" : (N == 16 ? "32x32b.x16" : N == 32 ? "32x32b.x32" : "32x32b.x64"),
           warps, INF, (double)cy / (iters * INF), bytes / (double)cy);
}

int main() {
    unsigned long long *dcy; uint32_t *dsink; float *fsink;
    CK(cudaMalloc(&dcy, 8)); CK(cudaMalloc(&dsink, 4)); CK(cudaMalloc(&fsink, 4));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("# %s, %d SMs, clock %d kHz (cycles below are SM clocks; single CTA on one SM)\n", prop.name, prop.multiProcessorCount, prop.clockRate);

    printf("## tcgen05.ld throughput (one CTA; warps beyond 4 share the four lane quarters)\n");
    for (int w : {4, 8, 16}) { run_ldtm<16, 1>(w, dcy, dsink); run_ldtm<16, 2>(w, dcy, dsink); run_ldtm<32, 1>(w, dcy, dsink); run_ldtm<32, 2>(w, dcy, dsink);
                               run_ldtm<64, 1>(w, dcy, dsink); run_ldtm<64, 2>(w, dcy, dsink); run_ldtm<256, 1>(w, dcy, dsink); run_ldtm<256, 2>(w, dcy, dsink); }

    printf("## tcgen05.mma M128 x N x K16 bf16 (SS): cycles per MMA, 27 MMAs per commit (one conv M-block) and 270 per commit\n");
    const char *lname[3] = {"no-swizzle, shifted windows, far K chunks", "no-swizzle, dense", "SWIZZLE_128B"};
    void (*uk[3])(int, int, int, unsigned long long *) = {umma_kernel<0>, umma_kernel<1>, umma_kernel<2>};
    for (int layout = 0; layout < 3; ++layout) {
        CK(cudaFuncSetAttribute(uk[layout], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int N : {16, 32, 48, 64, 128, 256})
            for (int reps : {1, 10}) {
                const int iters = reps == 1 ? 400 : 40;
                uk[layout]<<<1, 128, 200 * 1024>>>(N, iters, reps, dcy);
                CK(cudaDeviceSynchronize());
                uk[layout]<<<1, 128, 200 * 1024>>>(N, iters, reps, dcy);
                CK(cudaDeviceSynchronize());
                unsigned long long cy; CK(cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost));
                printf("umma layout=%d (%s) N=%3d per_commit=%3d : %7.1f cycles/MMA  (math floor 128*N/256 = %d)\n", layout, lname[layout], N, 27 * reps,
                       (double)cy / ((double)iters * reps * 27), 128 * N / 256);
            }
    }

    printf("## mma.sync.m16n8k16 bf16 -> fp32 (HMMA), 8 independent accumulators per warp\n");
    for (int w : {4, 8, 16, 32}) {
        const int iters = 4000;
        hmma_kernel<<<1, w * 32>>>(iters, dcy, fsink); CK(cudaDeviceSynchronize());
        hmma_kernel<<<1, w * 32>>>(iters, dcy, fsink); CK(cudaDeviceSynchronize());
        unsigned long long cy; CK(cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost));
        const double macs = (double)w * iters * 8 * 16 * 8 * 16;
        printf("hmma warps=%2d : %7.2f cycles/HMMA/warp  %8.1f MAC/clk/SM (dense bf16 tcgen05 peak is ~4096)\n", w, (double)cy / (iters * 8.0), macs / (double)cy);
    }

    printf("## LDS.128 bandwidth, 256 threads\n");
    CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    lds_kernel<<<1, 256, 64 * 1024>>>(2000, dcy, fsink); CK(cudaDeviceSynchronize());
    lds_kernel<<<1, 256, 64 * 1024>>>(2000, dcy, fsink); CK(cudaDeviceSynchronize());
    { unsigned long long cy; CK(cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost)); printf("lds.128: %7.1f B/clk/SM\n", 256.0 * 16 * 8 * 2000 / (double)cy); }
    return 0;
}
