// tc_kernels.cu -- implicit-GEMM 3-D convolution on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Used for the GEMM-shaped layers of the path: C_in*k^3 >= 64 reduction length with >= 16
// output channels (the 72/64/128/256-channel blocks of the Full model, vqvae/layers.py:134-171
// at the sizes of SURVEY.md 8a).  D[voxel, co] = sum_k A[voxel, k] * W[co, k], k = (ci, kh, kw, kz):
//
//   A  (M = 128 output voxels per CTA, K-major bf16): gathered on the fly by the CTA's warps from
//      the fp32 activation (wrapped coordinates = circular padding, zero padding = zero fill),
//      with the Fixup input transform ELU(x+a)+b applied, packed to bf16 and stored in the UMMA
//      canonical no-swizzle K-major layout (core matrix = 8 rows x 16 B, LBO = K-direction stride,
//      SBO = 8-row-group stride).  These activations are <= a few MB and L2 resident; the gather
//      is coordinate arithmetic, which TMA's zero-fill out-of-bounds mode cannot express for wrap.
//   B  (N = C_out padded to 16, K-major bf16): the reference's (C_out, C_in, k, k, k) fp32 weight
//      is already K-major; rows are converted to bf16 while being staged.
//   D  fp32 accumulators in TMEM (N columns x 128 lanes); one elected thread issues
//      tcgen05.mma.cta_group::1.kind::f16 (M128 x N x K16), tcgen05.commit signals an mbarrier per
//      smem stage so the gather of chunk i+1 overlaps the MMAs of chunk i.
//   epilogue: tcgen05.ld 32x32b (lane = voxel) -> *scale + b + bias[co] + residual (+ELU) -> fp32
//      planar store, coalesced because consecutive lanes are consecutive voxels.
//
// Operands are rounded to bf16 (fp32 accumulate): results match the fp32 path to ~1e-2 relative
// (north_star tolerance for BF16 paths); the quantizer stays fp32 and index-exact on its inputs.
#include "vq3d_rt.h"

#ifndef VQ3D_EMU
#include <cuda_bf16.h>

namespace vq3d {

constexpr int kTcBK = 64;            // K elements per smem stage
constexpr int kTcStages = 2;
constexpr int kTcM = 128;

struct TcParams {
    int B, H, W, Z, C1, C2, Cout, k, stride, pad, circ, pre_act, post_act;
    int Ho, Wo, Zo, Npad, Ktot, G;
    int nsplit, cps;                 // split-K: grid.y splits of cps K-chunks each (1 = no split)
    float *ws;                       // split-K partial sums [nsplit][Npad][tiles * 128] fp32 (plain stores; reduced in a fixed
                                     // order by conv3d_tc_reduce_kernel, so the result is bit-reproducible)
    const float *x1, *x2, *w, *bias, *pre_a, *pre_b, *post_scale, *post_b, *residual;
    float *y;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();   // never hang the GPU: a lost arrival becomes a launch error instead
}

// K-major, no swizzle: bits [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

// threads = 128 * G: thread (row r = tid % 128, group g = tid / 128); a stage holds K = 64 = 8 chunks
// of 8 elements (16 B); group g packs chunks g, g+G, ...
template <int KS>       // kernel size as a compile-time constant: the (ci, kh, kw, kz) decode of the gather folds to shifts / constant divisions
__global__ void __launch_bounds__(512, 2)      // two CTAs per SM: the gather is latency-bound
conv3d_tc_kernel(TcParams p) {
    VQ3D_DYN_SMEM(unsigned char, smem_raw);
    __shared__ __align__(8) uint64_t mma_done[kTcStages];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = tid & (kTcM - 1), g = tid >> 7, G = p.G;
    const int Npad = p.Npad, Ktot = p.Ktot;
    constexpr int k = KS, kk = k * k, k3 = kk * k;
    const int Cin = p.C1 + p.C2;
    // stage layout: A [8 k-chunks][16 row-groups][128 B] = 16 KB, then B [8][Npad/8][128 B]
    const uint32_t a_bytes = kTcM * kTcBK * 2, b_bytes = (uint32_t)Npad * kTcBK * 2;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t a_lbo = (kTcM / 8) * 128, b_lbo = (uint32_t)(Npad / 8) * 128, sbo = 128;
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    unsigned char *smem = smem_raw + (smem_base - smem_u32(smem_raw));

    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)Npad) tmem_cols <<= 1;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    // this thread's output voxel
    const int64_t S = (int64_t)p.H * p.W * p.Z, So = (int64_t)p.Ho * p.Wo * p.Zo;
    const int64_t total = (int64_t)p.B * So;
    const int64_t v = (int64_t)blockIdx.x * kTcM + r;
    const bool row_ok = v < total;
    int b = 0, oh = 0, ow = 0, oz = 0;
    if (row_ok) {
        b = (int)(v / So);
        int rem = (int)(v - (int64_t)b * So);
        oh = rem / (p.Wo * p.Zo);
        rem -= oh * p.Wo * p.Zo;
        ow = rem / p.Zo;
        oz = rem - ow * p.Zo;
    }
    const int ih0 = oh * p.stride - p.pad, iw0 = ow * p.stride - p.pad, iz0 = oz * p.stride - p.pad;
    // k = 4 / 2: this thread's input z of every kz tap, wrapped (circular) or -1 (zero padding, outside): fixed for the whole kernel
    int zoff[KS == 4 || KS == 2 ? KS : 1];
    if constexpr (KS == 4 || KS == 2) {
#pragma unroll
        for (int jz = 0; jz < KS; ++jz) {
            int iz = iz0 + jz;
            if (p.circ) iz = wrap(iz, p.Z); else if (iz < 0 || iz >= p.Z) iz = -1;
            zoff[jz] = iz;
        }
    }
    const float pa = ld_scalar(p.pre_a, 0.f), pb = ld_scalar(p.pre_b, 0.f);
    // instruction descriptor: D fp32 (1<<4), A/B bf16 (1<<7, 1<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Npad >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

    const int nchunks_all = (Ktot + kTcBK - 1) / kTcBK;
    const int chunk0 = (int)blockIdx.y * p.cps;
    const int nchunks = min(p.cps, nchunks_all - chunk0);          // this CTA's share of the K loop (>= 1)
    for (int it = 0; it < nchunks; ++it) {
        const int s = it & 1;
        if (it >= kTcStages) mbar_wait(&mma_done[s], (uint32_t)(((it >> 1) - 1) & 1));   // MMAs that read this stage are done
        unsigned char *sa = smem + (size_t)s * stage_bytes;
        unsigned char *sb = sa + a_bytes;
        const int kbase = (chunk0 + it) * kTcBK;
        // ---- B (first item of this thread): the weight loads are issued before the gather and packed after it, so that
        // both wait for the same memory round trip ----
        float wv0[8];
        const bool w_has = tid < Npad * 8;
        {
            const int n = tid >> 3, k0 = kbase + (tid & 7) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) wv0[e] = (w_has && n < p.Cout && k0 + e < Ktot) ? __ldg(p.w + (size_t)n * Ktot + k0 + e) : 0.0f;
        }
        // ---- A: gather + transform + bf16 pack, one 16-byte chunk (8 k's) at a time ----------
        if constexpr (KS == 4 || KS == 2) {
            // a 16-byte chunk is 2 x 4 (k = 4: two kw, all four kz) or 2 x 2 x 2 (k = 2: one input channel) taps with a common
            // (ci, kh[, kw]) prefix: one plane base per chunk, 32-bit offsets from the row bases and the thread's pre-wrapped z
            // instead of the per-element decode + wrap + 64-bit address of the generic path (the gather was instruction-bound:
            // 69 % issue active, tensor pipe < 1 %, profiles/r01x_down16_128.summary.txt).  Keeping two chunks' loads in flight was
            // measured too: same time (1.73 vs 1.74 ms at 16->16 @128x128x32, batch 16) for 86 registers, not kept
            auto addrs = [&](int c, const float *&base, int (&off)[8]) {
                const int kidx = kbase + c * 8;
                const int ci = kidx / k3, t = kidx - ci * k3;
                const int kh = t / kk, kw = (t - kh * kk) / k;
#pragma unroll
                for (int e = 0; e < 8; ++e) off[e] = -1;
                base = p.x1;
                if (row_ok && ci < Cin) {
                    base = ci < p.C1 ? p.x1 + ((size_t)b * p.C1 + ci) * S : p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S;
                    constexpr int NH = KS == 4 ? 1 : 2, NW = 2;               // kh values, kw values of a chunk
#pragma unroll
                    for (int jh = 0; jh < NH; ++jh) {
                        int ih = ih0 + kh + jh;
                        bool okh = true;
                        if (p.circ) ih = wrap(ih, p.H); else okh = ih >= 0 && ih < p.H;
#pragma unroll
                        for (int jw = 0; jw < NW; ++jw) {
                            int iw = iw0 + kw + jw;
                            bool okw = okh;
                            if (p.circ) iw = wrap(iw, p.W); else okw = okh && iw >= 0 && iw < p.W;
                            const int rowo = (ih * p.W + iw) * p.Z;
#pragma unroll
                            for (int jz = 0; jz < KS; ++jz)
                                if (okw && zoff[jz] >= 0) off[(jh * NW + jw) * KS + jz] = rowo + zoff[jz];
                        }
                    }
                }
            };
            auto finish = [&](int c, const int (&off)[8], float (&vals)[8]) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (off[e] >= 0) vals[e] = p.pre_act ? elu1(vals[e] + pa) + pb : vals[e] + pb;
                uint4 pk;
                pk.x = pack_bf16(vals[0], vals[1]); pk.y = pack_bf16(vals[2], vals[3]);
                pk.z = pack_bf16(vals[4], vals[5]); pk.w = pack_bf16(vals[6], vals[7]);
                *reinterpret_cast<uint4 *>(sa + (size_t)c * a_lbo + (size_t)(r >> 3) * sbo + (size_t)(r & 7) * 16) = pk;
            };
            for (int c = g; c < 8; c += G) {
                const float *b0;
                int o0[8];
                float v0[8];
                addrs(c, b0, o0);
#pragma unroll
                for (int e = 0; e < 8; ++e) v0[e] = o0[e] >= 0 ? __ldg(b0 + o0[e]) : 0.0f;
                finish(c, o0, v0);
            }
        } else {
        for (int c = g; c < 8; c += G) {
            int kidx = kbase + c * 8;
            int ci = kidx / k3, t = kidx - ci * k3;
            int kh = t / kk; t -= kh * kk;
            int kw = t / k, kz = t - kw * k;
            // addresses first, then the eight loads back to back (one memory round trip per chunk), then the transform
            const float *src[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                src[e] = nullptr;
                if (row_ok && ci < Cin) {
                    int ih = ih0 + kh, iw = iw0 + kw, iz = iz0 + kz;
                    bool ok = true;
                    if (p.circ) { ih = wrap(ih, p.H); iw = wrap(iw, p.W); iz = wrap(iz, p.Z); }
                    else ok = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W && iz >= 0 && iz < p.Z;
                    if (ok) {
                        const float *base = ci < p.C1 ? p.x1 + ((size_t)b * p.C1 + ci) * S : p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S;
                        src[e] = base + ((size_t)ih * p.W + iw) * p.Z + iz;
                    }
                }
                if (++kz == k) { kz = 0; if (++kw == k) { kw = 0; if (++kh == k) { kh = 0; ++ci; } } }
            }
            float vals[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) vals[e] = src[e] ? __ldg(src[e]) : 0.0f;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (src[e]) vals[e] = p.pre_act ? elu1(vals[e] + pa) + pb : vals[e] + pb;
            uint4 pk;
            pk.x = pack_bf16(vals[0], vals[1]); pk.y = pack_bf16(vals[2], vals[3]);
            pk.z = pack_bf16(vals[4], vals[5]); pk.w = pack_bf16(vals[6], vals[7]);
            *reinterpret_cast<uint4 *>(sa + (size_t)c * a_lbo + (size_t)(r >> 3) * sbo + (size_t)(r & 7) * 16) = pk;
        }
        }
        // ---- B: weight rows, fp32 -> bf16 -------------------------------------------------------
        for (int i = tid; i < Npad * 8; i += blockDim.x) {
            const int n = i >> 3, c = i & 7;
            const int k0 = kbase + c * 8;
            float wv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) wv[e] = i == tid ? wv0[e] : ((n < p.Cout && k0 + e < Ktot) ? __ldg(p.w + (size_t)n * Ktot + k0 + e) : 0.0f);
            uint4 pk;
            pk.x = pack_bf16(wv[0], wv[1]); pk.y = pack_bf16(wv[2], wv[3]);
            pk.z = pack_bf16(wv[4], wv[5]); pk.w = pack_bf16(wv[6], wv[7]);
            *reinterpret_cast<uint4 *>(sb + (size_t)c * b_lbo + (size_t)(n >> 3) * sbo + (size_t)(n & 7) * 16) = pk;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (tensor core)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes, b_addr = a_addr + a_bytes;
#pragma unroll
            for (int j = 0; j < kTcBK / 16; ++j) {
                const uint64_t adesc = make_desc(a_addr + (uint32_t)(2 * j) * a_lbo, a_lbo, sbo);
                const uint64_t bdesc = make_desc(b_addr + (uint32_t)(2 * j) * b_lbo, b_lbo, sbo);
                umma_bf16(tmem_d, adesc, bdesc, idesc, (it > 0 || j > 0) ? 1u : 0u);
            }
            umma_commit(&mma_done[s]);       // arrives when every MMA issued so far has completed
        }
    }
    // all MMAs complete in order: waiting for the last commit covers everything
    {
        const int last = nchunks - 1;
        mbar_wait(&mma_done[last & 1], (uint32_t)((last >> 1) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., columns split across the G warp groups ----
    {
        const int q = warp & 3, wg = warp >> 2;
        const int row = q * 32 + lane;
        const int64_t vr = (int64_t)blockIdx.x * kTcM + row;
        const bool ok = vr < total;
        int bb = 0; int64_t rem = 0;
        if (ok) { bb = (int)(vr / So); rem = vr - (int64_t)bb * So; }
        const float sc = ld_scalar(p.post_scale, 1.f), sbias = ld_scalar(p.post_b, 0.f);
        for (int c0 = wg * 8; c0 < Npad; c0 += G * 8) {
            uint32_t acc[8];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ok) {
                if (p.nsplit > 1) {        // partial sums of this K slice -> its own workspace slab, [column][voxel] (coalesced)
                    const size_t rows = (size_t)gridDim.x * kTcM;
                    float *wcol = p.ws + ((size_t)blockIdx.y * Npad + c0) * rows + vr;
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (c0 + e < p.Cout) wcol[(size_t)e * rows] = __uint_as_float(acc[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int co = c0 + e;
                        if (co < p.Cout) {
                            const size_t o = ((size_t)bb * p.Cout + co) * So + rem;
                            float yv = __fmaf_rn(__uint_as_float(acc[e]), sc, sbias);
                            if (p.bias) yv += __ldg(p.bias + co);
                            if (p.residual) yv += __ldg(p.residual + o);
                            if (p.post_act) yv = elu1(yv);
                            p.y[o] = yv;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

// split-K reduction + output transform: thread = (voxel, output channel); the nsplit partial sums are added in slice
// order (deterministic), then *scale + b + bias + residual (+ELU) exactly like the unsplit epilogue
__global__ void __launch_bounds__(256)
conv3d_tc_reduce_kernel(TcParams p, int64_t rows) {
    const int64_t So = (int64_t)p.Ho * p.Wo * p.Zo, total = (int64_t)p.B * So;
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int co = blockIdx.y;
    if (v >= total) return;
    const float *src = p.ws + (size_t)co * rows + v;
    const size_t slab = (size_t)p.Npad * rows;
    float acc = 0.0f;
    int s = 0;
    for (; s + 4 <= p.nsplit; s += 4) {
        const float a0 = __ldcg(src + (size_t)s * slab), a1 = __ldcg(src + (size_t)(s + 1) * slab),
                    a2 = __ldcg(src + (size_t)(s + 2) * slab), a3 = __ldcg(src + (size_t)(s + 3) * slab);
        acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, a0), a1), a2), a3);
    }
    for (; s < p.nsplit; ++s) acc = __fadd_rn(acc, __ldcg(src + (size_t)s * slab));
    const int b = (int)(v / So);
    const int64_t rem = v - (int64_t)b * So;
    const size_t o = ((size_t)b * p.Cout + co) * So + rem;
    float yv = __fmaf_rn(acc, ld_scalar(p.post_scale, 1.f), ld_scalar(p.post_b, 0.f));
    if (p.bias) yv += __ldg(p.bias + co);
    if (p.residual) yv += __ldg(p.residual + o);
    if (p.post_act) yv = elu1(yv);
    p.y[o] = yv;
}

}  // namespace vq3d
#endif  // !VQ3D_EMU

using namespace vq3d;

#ifndef VQ3D_EMU
// split-K plan: few M tiles (tiny spatial extent, long K = C_in*k^3) leave most SMs idle and make the K loop
// a serial latency chain; slice K across grid.y and reduce through an fp32 workspace
static void tc_split_plan(const vq3d_conv_desc *d, int64_t *tiles_out, int *nsplit, int *cps) {
    const int Ho = (d->H + 2 * d->pad - d->k) / d->stride + 1, Wo = (d->W + 2 * d->pad - d->k) / d->stride + 1,
              Zo = (d->Z + 2 * d->pad - d->k) / d->stride + 1;
    const int64_t total = (int64_t)d->B * Ho * Wo * Zo;
    const int64_t tiles = ceil_div(total, kTcM);
    const int nchunks = (int)ceil_div((int64_t)(d->C1 + d->C2) * d->k * d->k * d->k, kTcBK);
    int ns = 1;
    if (tiles * 2 <= kNumSMs && nchunks >= 4) {
        ns = (int)((2 * kNumSMs) / tiles);
        if (ns > nchunks) ns = nchunks;
    }
    int c = (int)ceil_div(nchunks, ns);
    ns = (int)ceil_div(nchunks, c);          // no empty slices
    *tiles_out = tiles; *nsplit = ns; *cps = c;
}
#endif

extern "C" size_t vq3d_conv3d_tc_workspace(const vq3d_conv_desc *d) {
#ifdef VQ3D_EMU
    (void)d;
    return 0;
#else
    if (!d || d->B < 1 || d->H < 1 || d->W < 1 || d->Z < 1 || d->k < 1 || d->stride < 1 || d->Cout < 1) return 0;
    int64_t tiles; int ns, cps;
    tc_split_plan(d, &tiles, &ns, &cps);
    if (ns <= 1) return 0;
    const int Npad = (d->Cout + 15) & ~15;
    return (size_t)ns * Npad * tiles * kTcM * 4;
#endif
}

extern "C" int vq3d_conv3d_tc(const vq3d_conv_desc *d, void *ws, size_t ws_bytes, void *stream) {
#ifdef VQ3D_EMU
    (void)d; (void)ws; (void)ws_bytes; (void)stream;
    return fail(VQ3D_ERR_UNSUPPORTED, "conv3d_tc: tensor-core kernels cannot run in the host emulator");
#else
    if (!d) return fail(VQ3D_ERR_INVALID, "conv3d_tc: null descriptor");
    if (!d->x1 || !d->w || !d->y) return fail(VQ3D_ERR_INVALID, "conv3d_tc: null x1/w/y");
    if (d->C2 > 0 && !d->x2) return fail(VQ3D_ERR_INVALID, "conv3d_tc: C2 > 0 but x2 is NULL");
    if (d->B < 1 || d->H < 1 || d->W < 1 || d->Z < 1 || d->C1 < 1 || d->C2 < 0 || d->Cout < 1) return fail(VQ3D_ERR_INVALID, "conv3d_tc: bad sizes");
    if (d->k < 1 || d->k > 4 || (d->stride != 1 && d->stride != 2) || d->pad < 0 || d->pad >= d->k) return fail(VQ3D_ERR_INVALID, "conv3d_tc: unsupported geometry");
    if (d->pad_circular && (d->pad > d->H || d->pad > d->W || d->pad > d->Z)) return fail(VQ3D_ERR_INVALID, "conv3d_tc: circular padding larger than the input");
    const int Cin = d->C1 + d->C2;
    const int64_t Ktot = (int64_t)Cin * d->k * d->k * d->k;
    if (d->Cout > 256) return fail(VQ3D_ERR_UNSUPPORTED, "conv3d_tc: Cout > 256");
    if (Ktot < 32 || d->Cout < 8) return fail(VQ3D_ERR_UNSUPPORTED, "conv3d_tc: not GEMM-shaped (K=%lld, N=%d)", (long long)Ktot, d->Cout);
    TcParams p;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z; p.C1 = d->C1; p.C2 = d->C2; p.Cout = d->Cout;
    p.k = d->k; p.stride = d->stride; p.pad = d->pad; p.circ = d->pad_circular; p.pre_act = d->pre_act; p.post_act = d->post_act;
    p.Ho = (d->H + 2 * d->pad - d->k) / d->stride + 1;
    p.Wo = (d->W + 2 * d->pad - d->k) / d->stride + 1;
    p.Zo = (d->Z + 2 * d->pad - d->k) / d->stride + 1;
    if (p.Ho < 1 || p.Wo < 1 || p.Zo < 1) return fail(VQ3D_ERR_INVALID, "conv3d_tc: empty output");
    p.Npad = (d->Cout + 15) & ~15;
    p.Ktot = (int)Ktot;
    p.x1 = d->x1; p.x2 = d->x2; p.w = d->w; p.bias = d->bias; p.pre_a = d->pre_a; p.pre_b = d->pre_b;
    p.post_scale = d->post_scale; p.post_b = d->post_b; p.residual = d->residual; p.y = d->y;
    int64_t tiles;
    tc_split_plan(d, &tiles, &p.nsplit, &p.cps);
    p.ws = nullptr;
    if (p.nsplit > 1) {
        const size_t need = vq3d_conv3d_tc_workspace(d);
        if (!ws || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 15) != 0) {       // no workspace: run unsplit
            p.nsplit = 1;
            p.cps = (int)ceil_div(Ktot, kTcBK);
        } else {
            p.ws = static_cast<float *>(ws);
        }
    }
    p.G = tiles * p.nsplit >= 2 * kNumSMs ? 2 : 4;      // few CTAs: put more gather threads on each
    const size_t smem = (size_t)kTcStages * ((size_t)kTcM * kTcBK * 2 + (size_t)p.Npad * kTcBK * 2) + 256;
    const dim3 grid((unsigned)tiles, (unsigned)p.nsplit), block((unsigned)(kTcM * p.G));
    int rc;
    switch (p.k) {
        case 1: rc = launch("conv3d_tc", conv3d_tc_kernel<1>, grid, block, smem, stream, p); break;
        case 2: rc = launch("conv3d_tc", conv3d_tc_kernel<2>, grid, block, smem, stream, p); break;
        case 3: rc = launch("conv3d_tc", conv3d_tc_kernel<3>, grid, block, smem, stream, p); break;
        default: rc = launch("conv3d_tc", conv3d_tc_kernel<4>, grid, block, smem, stream, p); break;
    }
    if (rc || p.nsplit <= 1) return rc;
    const int64_t total = (int64_t)p.B * p.Ho * p.Wo * p.Zo;
    return launch("conv3d_tc_reduce", conv3d_tc_reduce_kernel, dim3((unsigned)ceil_div(total, 256), (unsigned)p.Cout), dim3(256), 0, stream,
                  p, (int64_t)tiles * kTcM);
#endif
}
