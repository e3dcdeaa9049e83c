// preact_tc_kernels.cu -- runs of 'same' PreActFixupResBlocks (vqvae/layers.py:102-216; the
// nn.Sequential stacks of layers.py:492-494,566-569) with the k3 convolution on the 5th-gen tensor
// cores, up to 24 blocks per launch: persistent warp-specialised CTAs (one per SM) + grid barrier.
//
//   t1 = ELU(conv1x1(ELU(x+b1a)+b1b)+b2a)+b2b           (pointwise)
//   t2 = ELU(conv3x3x3_circular(t1)+b3a)+b3b            (tensor cores)
//   y  = conv1x1(t2)*scale + b4 + x                     (pointwise)
//
// The pointwise halves of two CONSECUTIVE blocks are fused: the epilogue that produces y_i also
// produces t1_{i+1} from it (no halo needed for 1x1 convolutions), rounds it to bf16 and writes it
// to a global workspace in the layout the tensor core wants: [chunk of 8 channels][h][w][z+1 of Z+2]
// [16 B], the depth rows padded with their circular halo.  Block i+1 then needs no arithmetic to
// build its A operand: a haloed (th+2, tw+2, Z+2) box is (th+2)(tw+2) row copies per chunk, issued as
// cp.async.bulk (UBLKCP) straight into the UMMA canonical K-major no-swizzle layout
// ([chunk][box voxel L][16 B], L = (h*IW + w)*IZ + z).
//
//   producer warp   waits for a free A buffer, issues the row copies of the next tile (mbarrier tx)
//   3 MMA lanes     the 27 taps of conv2 = 27 accumulating tcgen05.mma (M128 x N=CBP x K16 per 16
//                   branch channels) per M-block of 128 CONSECUTIVE box indices: the A operand of tap
//                   (kh,kw,kz) is the same smem array shifted by ((kh-1)*IW + (kw-1))*IZ + (kz-1)
//                   rows -- a descriptor start-address change, no im2col.  Rows that fall on halo
//                   positions compute garbage that is never stored.  fp32 accumulators in TMEM.
//   consumer warps  tcgen05.ld (lane = voxel) -> ELU -> conv3 -> *scale + b4 + x -> store y (fp32,
//                   the reference's planar layout, in place) -> ELU -> conv1 of the next block -> ELU ->
//                   bf16 -> t1 workspace.
// A buffers and TMEM accumulators are double buffered, so copies, MMAs and the SIMT epilogue of
// neighbouring tiles overlap inside the single resident CTA (kernels that use tcgen05 get one CTA
// per SM from this driver).  Between blocks every CTA passes a grid barrier (cooperative launch).
#include "vq3d_rt.h"

#ifndef VQ3D_EMU
#include <cuda_bf16.h>
#include <cstdlib>

namespace vq3d {

constexpr int kTcsMaxBlocks = 24;          // blocks per launch (kernel parameter space)
constexpr int kTcsAuxWarps = 4;            // warp 0 producer, warps 1..3 MMA issuers
constexpr int kTcsMmaWarps = 3;

struct TcsBlock {
    const float *w1, *w2, *w3;
    const float *b1a, *b1b, *b2a, *b2b, *b3a, *b3b, *b4, *scale;
};

struct TcsParams {
    int B, H, W, Z;
    int th, tw, nth, ntw;
    int IH, IW, IZ, NL;
    int L0, NMB, NLA;
    int ntiles, nblocks;
    uint32_t tmem_cols;
    unsigned int *sync;                    // grid barrier counter (zeroed by the host before the launch)
    unsigned long long *trace;             // debug timeline or NULL
    const float *x;                        // input of block 0 (may alias y)
    float *y;                              // output of every block (updated in place from block 1 on)
    uint4 *t1[2];                          // bf16 t1 ping-pong: [B][NCH][H][W][Z+2] 16-byte units
    TcsBlock blk[kTcsMaxBlocks];
};

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarrier_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbarrier_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = s_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();   // a lost arrival becomes a launch error, never a hung GPU
}

__device__ __forceinline__ void mbarrier_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(s_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbarrier_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}

// global -> shared bulk copy (UBLKCP); completion is signalled as tx bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}

// K-major, no swizzle: [0,14) addr>>4, [16,30) LBO>>4 (K-direction core-matrix stride), [32,46) SBO>>4
// (8-row-group stride), [46,48) version = 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit_to(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ int pmodi(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}

__device__ __forceinline__ void grid_barrier(unsigned int *ctr, unsigned int target) {
    asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy global stores -> later bulk-copy (async proxy) reads
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned int v = 0;
        uint32_t spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            __nanosleep(20);
        } while (++spins < (1u << 26));
        if (v < target) __trap();
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
}

constexpr int rup(int v, int m) { return (v + m - 1) / m * m; }

// debug timeline (VQ3D_TC_TRACE=1): CTA 0 stamps %globaltimer for blocks 1 and 2 into 9 slots each
__device__ __forceinline__ void tc_trace(unsigned long long *tr, int blk, int ev) {
    if (tr != nullptr && blockIdx.x == 0 && (blk == 1 || blk == 2)) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        tr[(blk - 1) * 9 + ev] = t;
    }
}

template <int C, int CB>
struct TcsCfg {
    static constexpr int CBP = rup(CB, 16);            // MMA N and K extent
    static constexpr int NK = CBP / 16;
    static constexpr int NCH = CBP / 8;                // 8-channel (16 B) chunks per voxel
    static constexpr int CB4 = rup(CB, 4), C4 = rup(C, 4);
    static constexpr int OC = C <= 32 ? C : 24;        // conv3 output channels per register pass
    static constexpr uint32_t W2TAP = (uint32_t)CBP * CBP * 2;   // bytes of one tap's B operand
    static constexpr uint32_t W2BYTES = 27u * W2TAP;
    static constexpr uint32_t LBO_B = (uint32_t)CBP * 16;
    static constexpr size_t w_floats = (size_t)C * CB4 + (size_t)CB * C4;
    static size_t smem_bytes(int NLA) { return 128 + W2BYTES + 2 * (size_t)NCH * NLA * 16 + w_floats * 4; }
};

// pointwise first half of a block from the (fp32) values of one voxel, streamed channel by channel
template <int C, int CB>
struct T1Acc {
    static constexpr int CBP = TcsCfg<C, CB>::CBP, CB4 = TcsCfg<C, CB>::CB4, NCH = TcsCfg<C, CB>::NCH;
    float acc[CBP];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int c = 0; c < CBP; ++c) acc[c] = 0.0f;
    }
    // v = value of input channel ci (before ELU); sw1 = [C][CB4] transposed conv1 weights
    __device__ __forceinline__ void add(int ci, float v, const float *sw1, float b1a, float b1b) {
        const float a = elu1(v + b1a) + b1b;
        const float4 *wr = reinterpret_cast<const float4 *>(sw1 + ci * CB4);
#pragma unroll
        for (int j = 0; j < CB4 / 4; ++j) {
            const float4 w = wr[j];
            acc[4 * j + 0] = __fmaf_rn(w.x, a, acc[4 * j + 0]);
            if (4 * j + 1 < CB) acc[4 * j + 1] = __fmaf_rn(w.y, a, acc[4 * j + 1]);
            if (4 * j + 2 < CB) acc[4 * j + 2] = __fmaf_rn(w.z, a, acc[4 * j + 2]);
            if (4 * j + 3 < CB) acc[4 * j + 3] = __fmaf_rn(w.w, a, acc[4 * j + 3]);
        }
    }
    // ELU, bf16, store into the z-padded chunk-planar workspace (+ the circular depth halo copies)
    __device__ __forceinline__ void store(uint4 *t1, int b, int oh, int ow, int oz, int H, int W, int Z, float b2a, float b2b) {
#pragma unroll
        for (int c = 0; c < CBP; ++c) acc[c] = c < CB ? elu1(acc[c] + b2a) + b2b : 0.0f;
#pragma unroll
        for (int kc = 0; kc < NCH; ++kc) {
            uint4 pk;
            pk.x = bf16x2(acc[8 * kc + 0], acc[8 * kc + 1]); pk.y = bf16x2(acc[8 * kc + 2], acc[8 * kc + 3]);
            pk.z = bf16x2(acc[8 * kc + 4], acc[8 * kc + 5]); pk.w = bf16x2(acc[8 * kc + 6], acc[8 * kc + 7]);
            uint4 *row = t1 + ((((size_t)b * NCH + kc) * H + oh) * W + ow) * (size_t)(Z + 2);
            row[oz + 1] = pk;
            if (oz == 0) row[Z + 1] = pk;
            if (oz == Z - 1) row[0] = pk;
        }
    }
};

template <int C, int CB, int NCW>
__global__ void __launch_bounds__((kTcsAuxWarps + NCW) * 32, 1)
preact_tc_kernel(const __grid_constant__ TcsParams p) {
    using Cfg = TcsCfg<C, CB>;
    constexpr int CBP = Cfg::CBP, NK = Cfg::NK, NCH = Cfg::NCH, CB4 = Cfg::CB4, C4 = Cfg::C4, OC = Cfg::OC;
    constexpr int NT = (kTcsAuxWarps + NCW) * 32;
    constexpr int NG = NCW / 4;
    static_assert(NCW % 4 == 0 && NCW >= 4, "consumer warps come in groups of 4 (TMEM lane quarters)");
    VQ3D_DYN_SMEM(unsigned char, smem_raw);
    __shared__ __align__(8) uint64_t bar_full[2], bar_sa_empty[2], bar_tm_full[2], bar_tm_empty[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (s_u32(smem_raw) + 127u) & ~127u;
    unsigned char *smem = smem_raw + (base - s_u32(smem_raw));
    unsigned char *sW2 = smem;                                        // 27 x [NCH][CBP rows][16 B]
    const uint32_t lbo_a = (uint32_t)p.NLA * 16;
    const uint32_t sa_bytes = (uint32_t)NCH * lbo_a;                  // one A buffer: [NCH][NLA][16 B]
    float *sw3 = reinterpret_cast<float *>(smem + Cfg::W2BYTES + 2 * (size_t)sa_bytes);   // [CB][C4]  this block
    float *sw1n = sw3 + CB * C4;                                      // [C][CB4]  NEXT block's conv1
    const uint32_t sW2_addr = base, sA_addr = base + Cfg::W2BYTES;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < 2; ++s) {
            mbarrier_init(&bar_full[s], 1);
            mbarrier_init(&bar_sa_empty[s], kTcsMmaWarps);
            mbarrier_init(&bar_tm_full[s], kTcsMmaWarps);
            mbarrier_init(&bar_tm_empty[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    // D fp32 (1<<4), A/B bf16 (1<<7, 1<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CBP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    const int H = p.H, W = p.W, Z = p.Z;
    const int64_t S = (int64_t)H * W * Z;
    const int IW = p.IW, IZ = p.IZ, IWZ = p.IW * p.IZ;

    // ---- prologue: t1 of block 0 from x (pointwise, grid-stride over voxels) -----------------------
    {
        const TcsBlock &b0 = p.blk[0];
        for (int i = tid; i < C * CB4; i += NT) {
            const int cb = i % CB4, ci = i / CB4;
            sw1n[i] = cb < CB ? __ldg(b0.w1 + cb * C + ci) : 0.0f;
        }
        const float b1a = ld_scalar(b0.b1a, 0.f), b1b = ld_scalar(b0.b1b, 0.f), b2a = ld_scalar(b0.b2a, 0.f), b2b = ld_scalar(b0.b2b, 0.f);
        __syncthreads();
        const int64_t total = (int64_t)p.B * S;
        for (int64_t v = (int64_t)blockIdx.x * NT + tid; v < total; v += (int64_t)gridDim.x * NT) {
            const int b = (int)(v / S);
            const int64_t r = v - (int64_t)b * S;
            const int oz = (int)(r % Z);
            const int64_t hw = r / Z;
            const int ow = (int)(hw % W), oh = (int)(hw / W);
            const float *px = p.x + (size_t)b * C * S + r;
            T1Acc<C, CB> t;
            t.clear();
#pragma unroll(C <= 32 ? C : 8)
            for (int ci = 0; ci < C; ++ci) t.add(ci, __ldcg(px + (size_t)ci * S), sw1n, b1a, b1b);
            t.store(p.t1[0], b, oh, ow, oz, H, W, Z, b2a, b2b);
        }
        grid_barrier(p.sync, gridDim.x);
    }

    // tiles of this CTA (the same sequence in every block and for every role)
    const int my_tiles = p.ntiles > (int)blockIdx.x ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    uint32_t kt = 0;          // tiles processed by this CTA so far: buffer = kt & 1, use count = kt >> 1

    for (int blk = 0; blk < p.nblocks; ++blk) {
        const TcsBlock &bp = p.blk[blk];
        const bool has_next = blk + 1 < p.nblocks;
        if (tid == 0) tc_trace(p.trace, blk, 0);
        // ---- weights of this block (conv2 as the B operand, conv3) and of the next block's conv1 -----
        for (int i = tid; i < CB * C4; i += NT) {
            const int c = i % C4, cb = i / C4;
            sw3[i] = c < C ? __ldg(bp.w3 + c * CB + cb) : 0.0f;
        }
        if (has_next) {
            const float *w1n = p.blk[blk + 1].w1;
            for (int i = tid; i < C * CB4; i += NT) {
                const int cb = i % CB4, ci = i / CB4;
                sw1n[i] = cb < CB ? __ldg(w1n + cb * C + ci) : 0.0f;
            }
        }
        for (int i = tid; i < 27 * NCH * CBP; i += NT) {
            const int n = i % CBP, kc = (i / CBP) % NCH, t = i / (CBP * NCH);
            float wv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ci = kc * 8 + e;
                wv[e] = (n < CB && ci < CB) ? __ldg(bp.w2 + ((size_t)n * CB + ci) * 27 + t) : 0.0f;
            }
            uint4 pk;
            pk.x = bf16x2(wv[0], wv[1]); pk.y = bf16x2(wv[2], wv[3]); pk.z = bf16x2(wv[4], wv[5]); pk.w = bf16x2(wv[6], wv[7]);
            *reinterpret_cast<uint4 *>(sW2 + (size_t)t * Cfg::W2TAP + (size_t)kc * Cfg::LBO_B + (size_t)n * 16) = pk;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // sW2 written by the generic proxy, read by the tensor core
        __syncthreads();
        if (tid == 0) tc_trace(p.trace, blk, 1);

        const uint4 *t1src = p.t1[blk & 1];
        uint4 *t1dst = p.t1[(blk + 1) & 1];
        const float *resid = blk == 0 ? p.x : p.y;

        if (warp == 0) {
            // ================= producer: haloed box rows -> A buffer (bulk copies) =======================
            const int nrows = p.IH * IW;
            asm volatile("fence.proxy.async;" ::: "memory");     // t1 was written with generic-proxy stores (other CTAs, before the barrier)
            for (int i = 0; i < my_tiles; ++i) {
                const uint32_t k = kt + (uint32_t)i, s = k & 1u, u = k >> 1;
                int t = (int)blockIdx.x + i * (int)gridDim.x;
                const int twi = t % p.ntw; t /= p.ntw;
                const int thi = t % p.nth; t /= p.nth;
                const int b = t, oh0 = thi * p.th, ow0 = twi * p.tw;
                mbarrier_wait(&bar_sa_empty[s], (u & 1u) ^ 1u);
                if (lane == 0) mbarrier_arrive_expect_tx(&bar_full[s], (uint32_t)p.NL * NCH * 16u);
                __syncwarp();
                const uint32_t row_bytes = (uint32_t)IZ * 16u;
                for (int r = lane; r < nrows * NCH; r += 32) {
                    const int kc = r / nrows, rr = r - kc * nrows;
                    const int lh = rr / IW, lw = rr - lh * IW;
                    const int gh = pmodi(oh0 - 1 + lh, H), gw = pmodi(ow0 - 1 + lw, W);
                    const uint4 *src = t1src + ((((size_t)b * NCH + kc) * H + gh) * W + gw) * (size_t)(Z + 2);
                    bulk_g2s(sA_addr + s * sa_bytes + (uint32_t)kc * lbo_a + (uint32_t)rr * row_bytes, src, row_bytes, &bar_full[s]);
                }
                if (i == 0 && lane == 0) tc_trace(p.trace, blk, 2);
            }
        } else if (warp < kTcsAuxWarps) {
            // ================= MMA issuers: lane 0 of warps 1..3, M-blocks round-robin ====================
            if (lane == 0) {
                for (int i = 0; i < my_tiles; ++i) {
                    const uint32_t k = kt + (uint32_t)i, s = k & 1u, u = k >> 1;
                    mbarrier_wait(&bar_full[s], u & 1u);
                    mbarrier_wait(&bar_tm_empty[s], (u & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (i == 0 && warp == 1) tc_trace(p.trace, blk, 3);
                    const uint32_t a_base = sA_addr + s * sa_bytes;
                    for (int mb = warp - 1; mb < p.NMB; mb += kTcsMmaWarps) {
                        const uint32_t row0 = (uint32_t)(p.L0 + mb * 128);
                        const uint32_t d_addr = tmem_d + (uint32_t)((s * p.NMB + mb) * CBP);
#pragma unroll
                        for (int tp = 0; tp < 27; ++tp) {
                            const int kh = tp / 9, kw = (tp / 3) % 3, kz = tp % 3;
                            const uint32_t arow = row0 + (uint32_t)((kh - 1) * IWZ + (kw - 1) * IZ + (kz - 1));
#pragma unroll
                            for (int ks = 0; ks < NK; ++ks) {
                                const uint64_t adesc = umma_desc(a_base + arow * 16u + (uint32_t)(2 * ks) * lbo_a, lbo_a, 128);
                                const uint64_t bdesc = umma_desc(sW2_addr + (uint32_t)tp * Cfg::W2TAP + (uint32_t)(2 * ks) * Cfg::LBO_B, Cfg::LBO_B, 128);
                                umma_f16(d_addr, adesc, bdesc, idesc, (tp > 0 || ks > 0) ? 1u : 0u);
                            }
                        }
                    }
                    umma_commit_to(&bar_tm_full[s]);       // accumulators of this thread's M-blocks are complete
                    umma_commit_to(&bar_sa_empty[s]);      // ... and its reads of the A buffer are done
                    if (i == 0 && warp == 1) tc_trace(p.trace, blk, 4);
                }
            }
        } else {
            // ================= consumers: TMEM -> conv3 + residual -> y, conv1 of the next block -> t1 =======
            const int cw = warp - kTcsAuxWarps, g = cw >> 2, q = warp & 3;
            const float b3a = ld_scalar(bp.b3a, 0.f), b3b = ld_scalar(bp.b3b, 0.f), b4 = ld_scalar(bp.b4, 0.f), sc = ld_scalar(bp.scale, 1.f);
            float n1a = 0.f, n1b = 0.f, n2a = 0.f, n2b = 0.f;
            if (has_next) {
                const TcsBlock &nb = p.blk[blk + 1];
                n1a = ld_scalar(nb.b1a, 0.f); n1b = ld_scalar(nb.b1b, 0.f); n2a = ld_scalar(nb.b2a, 0.f); n2b = ld_scalar(nb.b2b, 0.f);
            }
            for (int i = 0; i < my_tiles; ++i) {
                const uint32_t k = kt + (uint32_t)i, s = k & 1u, u = k >> 1;
                int t = (int)blockIdx.x + i * (int)gridDim.x;
                const int twi = t % p.ntw; t /= p.ntw;
                const int thi = t % p.nth; t /= p.nth;
                const int b = t, oh0 = thi * p.th, ow0 = twi * p.tw;
                bool waited = false;
                for (int mb = g; mb < p.NMB; mb += NG) {
                    const int L = p.L0 + mb * 128 + q * 32 + lane;
                    const int lz = L % IZ, r = L / IZ;
                    const int lw = r % IW, lh = r / IW;
                    const int oh = oh0 + lh - 1, ow = ow0 + lw - 1, oz = lz - 1;
                    const bool valid = lh >= 1 && lh <= p.th && lw >= 1 && lw <= p.tw && lz >= 1 && lz <= Z && oh < H && ow < W;
                    const size_t off = valid ? ((size_t)oh * W + ow) * Z + oz : 0;
                    const float *px = resid + (size_t)b * C * S + off;
                    float xr[OC];
                    if (C <= 32 && valid) {       // residual loads in flight while the MMAs finish
#pragma unroll
                        for (int j = 0; j < OC; ++j) xr[j] = __ldcg(px + (size_t)j * S);
                    }
                    if (!waited) {
                        mbarrier_wait(&bar_tm_full[s], u & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        waited = true;
                        if (i == 0 && tid == kTcsAuxWarps * 32) tc_trace(p.trace, blk, 5);
                    }
                    float t2[CBP];
#pragma unroll
                    for (int ks = 0; ks < NK; ++ks)
                        tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)((s * p.NMB + mb) * CBP + ks * 16), t2 + ks * 16);
                    if (valid) {
#pragma unroll
                        for (int cb = 0; cb < CB; ++cb) t2[cb] = elu1(t2[cb] + b3a) + b3b;
                        float *py = p.y + (size_t)b * C * S + off;
                        T1Acc<C, CB> tn;
                        tn.clear();
#pragma unroll
                        for (int c0 = 0; c0 < C; c0 += OC) {
                            float out[OC];
                            if (C > 32) {
#pragma unroll
                                for (int j = 0; j < OC; ++j)
                                    if (c0 + j < C) xr[j] = __ldcg(px + (size_t)(c0 + j) * S);
                            }
#pragma unroll
                            for (int j = 0; j < OC; ++j) out[j] = 0.0f;
#pragma unroll
                            for (int cb = 0; cb < CB; ++cb) {
                                const float *wr = sw3 + cb * C4 + c0;
#pragma unroll
                                for (int j = 0; j < OC; ++j)
                                    if (c0 + j < C) out[j] = __fmaf_rn(wr[j], t2[cb], out[j]);
                            }
#pragma unroll
                            for (int j = 0; j < OC; ++j) {
                                if (c0 + j < C) {
                                    const float yv = __fmaf_rn(out[j], sc, b4) + xr[j];
                                    py[(size_t)(c0 + j) * S] = yv;
                                    if (has_next) tn.add(c0 + j, yv, sw1n, n1a, n1b);
                                }
                            }
                        }
                        if (has_next) tn.store(t1dst, b, oh, ow, oz, H, W, Z, n2a, n2b);
                    }
                }
                if (!waited) mbarrier_wait(&bar_tm_full[s], u & 1u);    // keep the phase bookkeeping in step
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbarrier_arrive(&bar_tm_empty[s]);
                if (i == 0 && tid == kTcsAuxWarps * 32) tc_trace(p.trace, blk, 6);
            }
        }
        kt += (uint32_t)my_tiles;
        if (tid == kTcsAuxWarps * 32) tc_trace(p.trace, blk, 7);
        if (has_next) grid_barrier(p.sync, (unsigned int)(blk + 2) * gridDim.x);
        if (tid == 0) tc_trace(p.trace, blk, 8);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
struct TcsPlan {
    int th, tw, IH, IW, IZ, NL, L0, NMB, NLA, ntiles;
    uint32_t tmem_cols;
    size_t smem;
    double cost;
};

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
    }
    return n;
}

template <int C, int CB>
static bool plan_tile(const vq3d_preact_desc *d, TcsPlan &best) {
    using Cfg = TcsCfg<C, CB>;
    const int cand[] = {1, 2, 4, 8, 16, 32, 64};
    int forced[2] = {0, 0};
    if (const char *e = getenv("VQ3D_TC_TILE")) sscanf(e, "%d,%d", &forced[0], &forced[1]);
    bool found = false;
    const int nsm = sm_count();
    const size_t smem_cap = 227 * 1024 - 1024;
    for (int a : cand) for (int bq : cand) {
        int th = a, tw = bq;
        if (forced[0] > 0) { th = forced[0]; tw = forced[1]; }
        if (th > d->H) { if (a != cand[0] && th / 2 >= d->H) continue; th = d->H; }
        if (tw > d->W) { if (bq != cand[0] && tw / 2 >= d->W) continue; tw = d->W; }
        TcsPlan pl;
        pl.th = th; pl.tw = tw;
        pl.IH = th + 2; pl.IW = tw + 2; pl.IZ = d->Z + 2;
        pl.NL = pl.IH * pl.IW * pl.IZ;
        pl.L0 = (pl.IW + 1) * pl.IZ + 1;
        const int Lend = (th * pl.IW + tw) * pl.IZ + d->Z;
        pl.NMB = (Lend - pl.L0 + 1 + 127) / 128;
        bool ok = 2 * pl.NMB * Cfg::CBP <= 512;
        pl.NLA = (pl.L0 + pl.NMB * 128 + pl.L0 + 7) & ~7;
        if (pl.NLA < pl.NL) pl.NLA = (pl.NL + 7) & ~7;
        ok = ok && (size_t)pl.NLA * 16 <= 0x3fffu * 16;                   // LBO field
        pl.smem = Cfg::smem_bytes(pl.NLA);
        ok = ok && pl.smem <= smem_cap;
        if (ok) {
            uint32_t cols = 32;
            while (cols < (uint32_t)(2 * pl.NMB * Cfg::CBP)) cols <<= 1;
            pl.tmem_cols = cols;
            pl.ntiles = d->B * (int)ceil_div(d->H, th) * (int)ceil_div(d->W, tw);
            // cycles per tile on one SM: SIMT epilogue (128 lanes/cycle), MMA issue, L2 -> smem copies; they overlap
            const double simt = ((double)pl.NMB * 128 * 40 + (double)th * tw * d->Z * (2.0 * C * CB + 14.0 * C + 16.0 * CB + 80.0)) / 128.0;
            const double mma = (double)pl.NMB * 27 * Cfg::NK * (Cfg::CBP / 2 > 24 ? Cfg::CBP / 2 : 24) / kTcsMmaWarps;
            const double load = (double)pl.NL * Cfg::NCH * 16 / 48.0 + (double)pl.IH * pl.IW * Cfg::NCH * 4.0;
            double tile = simt > mma ? simt : mma;
            if (load > tile) tile = load;
            pl.cost = (double)ceil_div(pl.ntiles, nsm) * (tile + 400.0) + 2500.0 + load + mma;
            if (!found || pl.cost < best.cost) { best = pl; found = true; }
        }
        if (forced[0] > 0) return found;
    }
    return found;
}

template <int C, int CB>
static size_t t1_units(const vq3d_preact_desc *d) {       // 16-byte units of ONE t1 buffer
    return (size_t)d->B * TcsCfg<C, CB>::NCH * d->H * d->W * (size_t)(d->Z + 2);
}

template <int C, int CB>
static size_t ws_bytes(const vq3d_preact_desc *d) { return 256 + 2 * t1_units<C, CB>(d) * 16; }

template <int C, int CB, int NCW>
static int launch_tcs(const vq3d_preact_desc *blocks, int n, void *ws, size_t ws_size, void *stream) {
    constexpr int NT = (kTcsAuxWarps + NCW) * 32;
    const vq3d_preact_desc *d = &blocks[0];
    if (ws_size < ws_bytes<C, CB>(d)) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: workspace too small (%zu < %zu bytes)", ws_size, ws_bytes<C, CB>(d));
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: workspace must be 256-byte aligned");
    TcsPlan pl;
    if (!plan_tile<C, CB>(d, pl)) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: no tile fits shared memory / TMEM");
    auto kernel = preact_tc_kernel<C, CB, NCW>;
    cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(attr)");
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, NT, pl.smem);
    if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(occupancy)");
    if (occ < 1) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: kernel does not fit on an SM");
    // one CTA per SM: the kernel may hold the SM's whole TMEM, a second resident CTA could spin in tcgen05.alloc forever
    int grid = sm_count();
    if (grid > pl.ntiles) grid = pl.ntiles;
    size_t smem = pl.smem;
    if (occ > 1 && smem < 120 * 1024) {
        smem = 120 * 1024;
        e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(attr)");
    }
    if (getenv("VQ3D_TC_DEBUG"))
        fprintf(stderr, "preact_stack_tc<%d,%d>: %dx%dx%d n=%d tile %dx%dx%d NL=%d NMB=%d ntiles=%d occ=%d grid=%d smem=%zu tmem=%u\n", C, CB,
                d->H, d->W, d->Z, n, pl.th, pl.tw, d->Z, pl.NL, pl.NMB, pl.ntiles, occ, grid, smem, pl.tmem_cols);

    TcsParams p;
    memset(&p, 0, sizeof(p));
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z;
    p.th = pl.th; p.tw = pl.tw;
    p.nth = (int)ceil_div(d->H, pl.th); p.ntw = (int)ceil_div(d->W, pl.tw);
    p.IH = pl.IH; p.IW = pl.IW; p.IZ = pl.IZ; p.NL = pl.NL; p.L0 = pl.L0; p.NMB = pl.NMB; p.NLA = pl.NLA;
    p.ntiles = pl.ntiles; p.tmem_cols = pl.tmem_cols;
    p.sync = reinterpret_cast<unsigned int *>(ws);
    const bool trace = getenv("VQ3D_TC_TRACE") != nullptr;
    p.trace = trace ? reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(ws) + 64) : nullptr;
    p.t1[0] = reinterpret_cast<uint4 *>(static_cast<unsigned char *>(ws) + 256);
    p.t1[1] = p.t1[0] + t1_units<C, CB>(d);
    p.y = blocks[n - 1].y;
    for (int i0 = 0; i0 < n; i0 += kTcsMaxBlocks) {
        const int nb = n - i0 < kTcsMaxBlocks ? n - i0 : kTcsMaxBlocks;
        p.nblocks = nb;
        p.x = i0 == 0 ? blocks[0].x : p.y;
        for (int j = 0; j < nb; ++j) {
            const vq3d_preact_desc &s = blocks[i0 + j];
            TcsBlock &t = p.blk[j];
            t.w1 = s.w1; t.w2 = s.w2; t.w3 = s.w3;
            t.b1a = s.b1a; t.b1b = s.b1b; t.b2a = s.b2a; t.b2b = s.b2b; t.b3a = s.b3a; t.b3b = s.b3b; t.b4 = s.b4; t.scale = s.scale;
        }
        e = cudaMemsetAsync(p.sync, 0, sizeof(unsigned int), static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(memset)");
        void *args[] = {&p};
        e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(kernel), dim3((unsigned)grid), dim3(NT), args, smem, static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(cooperative launch)");
        if (trace) {
            unsigned long long h[18];
            cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
            cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
            const char *names[9] = {"block start", "weights staged", "copies issued", "A full (MMA)", "MMAs committed", "TMEM full (consumer)",
                                    "tile 0 drained", "all tiles drained", "grid barrier passed"};
            for (int b = 0; b < 2; ++b)
                for (int ev = 0; ev < 9; ++ev)
                    fprintf(stderr, "  trace blk %d %-22s +%6.2f us\n", b + 1, names[ev], (double)(long long)(h[b * 9 + ev] - h[b * 9]) * 1e-3);
        }
    }
    return VQ3D_OK;
}

struct TcsEntry {
    int c, cb;
    int (*fn)(const vq3d_preact_desc *, int, void *, size_t, void *);
    size_t (*ws)(const vq3d_preact_desc *);
};

#define VQ3D_TCS(C, CB, NCW) {C, CB, launch_tcs<C, CB, NCW>, ws_bytes<C, CB>}
static const TcsEntry kTcs[] = {
    VQ3D_TCS(8, 4, 12), VQ3D_TCS(16, 8, 12), VQ3D_TCS(18, 9, 12), VQ3D_TCS(32, 16, 12), VQ3D_TCS(64, 32, 8), VQ3D_TCS(72, 36, 8),
};

static const TcsEntry *find_tcs(const vq3d_preact_desc *d) {
    for (const TcsEntry &e : kTcs)
        if (e.c == d->Cin && e.cb == d->Cb && d->Cout == d->Cin) return &e;
    return nullptr;
}

}  // namespace vq3d
#endif  // !VQ3D_EMU

using namespace vq3d;

extern "C" size_t vq3d_preact_stack_tc_workspace(const vq3d_preact_desc *first_block) {
#ifdef VQ3D_EMU
    (void)first_block;
    return 0;
#else
    if (!first_block || first_block->B < 1 || first_block->H < 1 || first_block->W < 1 || first_block->Z < 1) return 0;
    const TcsEntry *e = find_tcs(first_block);
    return e ? e->ws(first_block) : 0;
#endif
}

extern "C" int vq3d_preact_stack_tc(const vq3d_preact_desc *blocks, int n, void *ws, size_t ws_size, void *stream) {
#ifdef VQ3D_EMU
    (void)blocks; (void)n; (void)ws; (void)ws_size; (void)stream;
    return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: tensor-core kernels cannot run in the host emulator");
#else
    if (!blocks || n < 1) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: bad arguments");
    const vq3d_preact_desc *d = &blocks[0];
    if (!d->x || !blocks[n - 1].y) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: null input/output");
    if (d->B < 1 || d->H < 1 || d->W < 1 || d->Z < 1) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: bad sizes");
    for (int i = 0; i < n; ++i) {
        const vq3d_preact_desc &s = blocks[i];
        if (!s.w1 || !s.w2 || !s.w3) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: null weights");
        if (s.mode != 0 || s.wskip || s.Cin != d->Cin || s.Cb != d->Cb || s.Cout != d->Cin)
            return fail(VQ3D_ERR_INVALID, "preact_stack_tc: blocks must be equal-shape 'same' blocks without skip");
    }
    if ((int64_t)d->B * d->Cin * d->H * d->W * d->Z > ((int64_t)1 << 40)) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: tensor too large");
    const TcsEntry *e = find_tcs(d);
    if (!e) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: no tensor-core instantiation for C=%d Cb=%d", d->Cin, d->Cb);
    if (!ws) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: null workspace (see vq3d_preact_stack_tc_workspace)");
    return e->fn(blocks, n, ws, ws_size, stream);
#endif
}
