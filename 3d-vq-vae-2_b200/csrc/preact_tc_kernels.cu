// preact_tc_kernels.cu -- runs of 'same' PreActFixupResBlocks (vqvae/layers.py:102-216; the
// nn.Sequential stacks of layers.py:492-494,566-569) with the k3 convolution on the 5th-gen tensor
// cores, several blocks per launch (persistent CTAs + grid barrier).
//
//   o = conv1x1(ELU(x+b1a)+b1b); o = conv3x3x3_circular(ELU(o+b2a)+b2b); o = conv1x1(ELU(o+b3a)+b3b)
//   y = o*scale + b4 + x
//
// Per CTA and tile of (th, tw, tz) output voxels:
//   stage A  SIMT: conv1 + ELU on the haloed box (wrapped coordinates = circular padding), rounded to
//            bf16 and stored voxel-major in the UMMA canonical K-major no-swizzle layout
//            ([channel chunk of 8][box voxel L][16 B]); L = (h*IW + w)*IZ + z is the linear box index.
//   stage B  the 27 taps of conv2 are 27 accumulating tcgen05.mma (M128 x N=CBP x K16 per 16 branch
//            channels): for an M-block of 128 CONSECUTIVE box indices, the A operand of tap
//            (kh,kw,kz) is the same smem array shifted by ((kh-1)*IW + (kw-1))*IZ + (kz-1) rows --
//            a descriptor start-address change, no im2col.  Rows that fall on halo positions compute
//            garbage that is never stored (~20-25% of the rows).  fp32 accumulators live in TMEM.
//   stage C  tcgen05.ld (lane = voxel), ELU, conv3 in registers, *scale + b4 + x, coalesced store.
// Weights of conv2 are staged once per block as the B operand (27 x [CBP x CBP] bf16).
//
// Between consecutive blocks of the stack every CTA passes a grid barrier (cooperative launch, so
// all CTAs are co-resident); activations ping-pong between two global buffers that stay in L2 for
// every tensor of the model below the 256x256x64 level.
#include "vq3d_rt.h"

#ifndef VQ3D_EMU
#include <cuda_bf16.h>
#include <cstdlib>

namespace vq3d {

constexpr int kTcsThreads = 256;
constexpr int kTcsMaxBlocks = 24;          // blocks per launch (kernel parameter space)
constexpr int kTcsMaxMB = 32;              // M-blocks per tile (512 TMEM columns / 16)

struct TcsBlock {
    const float *w1, *w2, *w3, *ws;
    const float *b1a, *b1b, *b2a, *b2b, *b3a, *b3b, *b4, *scale, *b1c, *b1d;
};

struct TcsParams {
    int B, H, W, Z;
    int th, tw, tz, nth, ntw, ntz;
    int IH, IW, IZ, NL;
    int L0, NMB, NLA;
    int ntiles, nblocks;
    uint32_t tmem_cols;
    unsigned int *sync;                    // grid barrier counter (zeroed by the host before the launch)
    const float *x;                        // input of block 0
    float *buf[2];                         // block i writes buf[i & 1]
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
    }
    __syncthreads();
}

constexpr int rup(int v, int m) { return (v + m - 1) / m * m; }

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
    static size_t smem_bytes(int NLA) { return 128 + W2BYTES + (size_t)NCH * NLA * 16 + w_floats * 4; }
};

template <int C, int CB, int MINB>
__global__ void __launch_bounds__(kTcsThreads, MINB)
preact_tc_kernel(const __grid_constant__ TcsParams p) {
    using Cfg = TcsCfg<C, CB>;
    constexpr int CBP = Cfg::CBP, NK = Cfg::NK, NCH = Cfg::NCH, CB4 = Cfg::CB4, C4 = Cfg::C4, OC = Cfg::OC;
    VQ3D_DYN_SMEM(unsigned char, smem_raw);
    __shared__ __align__(8) uint64_t mbar[kTcsMaxMB];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (s_u32(smem_raw) + 127u) & ~127u;
    unsigned char *smem = smem_raw + (base - s_u32(smem_raw));
    unsigned char *sW2 = smem;                                       // 27 x [NCH][CBP rows][16 B]
    unsigned char *sA = sW2 + Cfg::W2BYTES;                          // [NCH][NLA][16 B]
    const uint32_t lbo_a = (uint32_t)p.NLA * 16;
    float *sw1 = reinterpret_cast<float *>(sA + (size_t)NCH * lbo_a);   // [C][CB4]
    float *sw3 = sw1 + C * CB4;                                      // [CB][C4]
    const uint32_t sW2_addr = base, sA_addr = base + Cfg::W2BYTES;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < kTcsMaxMB; ++i) mbarrier_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    // D fp32 (1<<4), A/B bf16 (1<<7, 1<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CBP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    const int64_t S = (int64_t)p.H * p.W * p.Z;
    const int IW = p.IW, IZ = p.IZ, IWZ = p.IW * p.IZ;
    uint32_t phase = 0;                      // parity of the mbarriers: flips once per processed tile

    for (int blk = 0; blk < p.nblocks; ++blk) {
        const TcsBlock &bp = p.blk[blk];
        const float *src = blk == 0 ? p.x : p.buf[(blk - 1) & 1];
        float *dst = p.buf[blk & 1];
        // ---- weights of this block -> shared ----------------------------------------------------
        for (int i = tid; i < C * CB4; i += kTcsThreads) {
            const int cb = i % CB4, ci = i / CB4;
            sw1[i] = cb < CB ? __ldg(bp.w1 + cb * C + ci) : 0.0f;
        }
        for (int i = tid; i < CB * C4; i += kTcsThreads) {
            const int c = i % C4, cb = i / C4;
            sw3[i] = c < C ? __ldg(bp.w3 + c * CB + cb) : 0.0f;
        }
        for (int i = tid; i < 27 * NCH * CBP; i += kTcsThreads) {
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
        const float b1a = ld_scalar(bp.b1a, 0.f), b1b = ld_scalar(bp.b1b, 0.f), b2a = ld_scalar(bp.b2a, 0.f), b2b = ld_scalar(bp.b2b, 0.f);
        const float b3a = ld_scalar(bp.b3a, 0.f), b3b = ld_scalar(bp.b3b, 0.f), b4 = ld_scalar(bp.b4, 0.f), sc = ld_scalar(bp.scale, 1.f);
        __syncthreads();

        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            int t = tile;
            const int tzi = t % p.ntz; t /= p.ntz;
            const int twi = t % p.ntw; t /= p.ntw;
            const int thi = t % p.nth; t /= p.nth;
            const int b = t;
            const int oh0 = thi * p.th, ow0 = twi * p.tw, oz0 = tzi * p.tz;
            const float *xb = src + (size_t)b * C * S;

            // ---- stage A: t1 = ELU(conv1(ELU(x+b1a)+b1b)+b2a)+b2b on the haloed box -> bf16 A operand
            for (int i = tid; i < p.NL; i += kTcsThreads) {
                const int lz = i % IZ, r = i / IZ;
                const int lw = r % IW, lh = r / IW;
                const int gh = pmodi(oh0 - 1 + lh, p.H), gw = pmodi(ow0 - 1 + lw, p.W), gz = pmodi(oz0 - 1 + lz, p.Z);
                const float *px = xb + ((size_t)gh * p.W + gw) * p.Z + gz;
                float acc[CBP];
#pragma unroll
                for (int c = 0; c < CBP; ++c) acc[c] = 0.0f;
#pragma unroll(C <= 32 ? C : 8)
                for (int ci = 0; ci < C; ++ci) {
                    const float v = elu1(__ldcg(px + (size_t)ci * S) + b1a) + b1b;
                    const float4 *wr = reinterpret_cast<const float4 *>(sw1 + ci * CB4);
#pragma unroll
                    for (int j = 0; j < CB4 / 4; ++j) {
                        const float4 w = wr[j];
                        acc[4 * j + 0] = __fmaf_rn(w.x, v, acc[4 * j + 0]);
                        if (4 * j + 1 < CB) acc[4 * j + 1] = __fmaf_rn(w.y, v, acc[4 * j + 1]);
                        if (4 * j + 2 < CB) acc[4 * j + 2] = __fmaf_rn(w.z, v, acc[4 * j + 2]);
                        if (4 * j + 3 < CB) acc[4 * j + 3] = __fmaf_rn(w.w, v, acc[4 * j + 3]);
                    }
                }
#pragma unroll
                for (int c = 0; c < CBP; ++c) acc[c] = c < CB ? elu1(acc[c] + b2a) + b2b : 0.0f;
#pragma unroll
                for (int kc = 0; kc < NCH; ++kc) {
                    uint4 pk;
                    pk.x = bf16x2(acc[8 * kc + 0], acc[8 * kc + 1]); pk.y = bf16x2(acc[8 * kc + 2], acc[8 * kc + 3]);
                    pk.z = bf16x2(acc[8 * kc + 4], acc[8 * kc + 5]); pk.w = bf16x2(acc[8 * kc + 6], acc[8 * kc + 7]);
                    *reinterpret_cast<uint4 *>(sA + (size_t)kc * lbo_a + (size_t)i * 16) = pk;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();

            // ---- stage B: lane 0 of warp w issues the 27*NK MMAs of M-blocks w, w+8, ... ---------------
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int mb = warp; mb < p.NMB; mb += kTcsThreads / 32) {
                    const uint32_t row0 = (uint32_t)(p.L0 + mb * 128);
                    const uint32_t d_addr = tmem_d + (uint32_t)(mb * CBP);
#pragma unroll
                    for (int tp = 0; tp < 27; ++tp) {
                        const int kh = tp / 9, kw = (tp / 3) % 3, kz = tp % 3;
                        const uint32_t arow = row0 + (uint32_t)((kh - 1) * IWZ + (kw - 1) * IZ + (kz - 1));
#pragma unroll
                        for (int ks = 0; ks < NK; ++ks) {
                            const uint64_t adesc = umma_desc(sA_addr + arow * 16u + (uint32_t)(2 * ks) * lbo_a, lbo_a, 128);
                            const uint64_t bdesc = umma_desc(sW2_addr + (uint32_t)tp * Cfg::W2TAP + (uint32_t)(2 * ks) * Cfg::LBO_B, Cfg::LBO_B, 128);
                            umma_f16(d_addr, adesc, bdesc, idesc, (tp > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit_to(&mbar[mb]);
                }
            }
            __syncwarp();

            // ---- stage C: warps (q = warp%4 -> TMEM lanes 32q..), group g = warp/4 takes M-blocks g, g+2, ...
            {
                const int q = warp & 3, g = warp >> 2;
                for (int mb = g; mb < p.NMB; mb += 2) {
                    mbarrier_wait(&mbar[mb], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int L = p.L0 + mb * 128 + q * 32 + lane;
                    const int lz = L % IZ, r = L / IZ;
                    const int lw = r % IW, lh = r / IW;
                    const int oh = oh0 + lh - 1, ow = ow0 + lw - 1, oz = oz0 + lz - 1;
                    const bool valid = lh >= 1 && lh <= p.th && lw >= 1 && lw <= p.tw && lz >= 1 && lz <= p.tz &&
                                       oh < p.H && ow < p.W && oz < p.Z;
                    float t2[CBP];
#pragma unroll
                    for (int ks = 0; ks < NK; ++ks)
                        tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * CBP + ks * 16), t2 + ks * 16);
                    if (valid) {
#pragma unroll
                        for (int cb = 0; cb < CB; ++cb) t2[cb] = elu1(t2[cb] + b3a) + b3b;
                        const size_t off = ((size_t)oh * p.W + ow) * p.Z + oz;
                        const float *px = xb + off;
                        float *py = dst + (size_t)b * C * S + off;
#pragma unroll
                        for (int c0 = 0; c0 < C; c0 += OC) {
                            float out[OC];
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
                            for (int j = 0; j < OC; ++j)
                                if (c0 + j < C) py[(size_t)(c0 + j) * S] = __fmaf_rn(out[j], sc, b4) + __ldcg(px + (size_t)(c0 + j) * S);
                        }
                    }
                }
            }
            phase ^= 1u;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();     // TMEM drained and every MMA retired before the next tile reuses sA / TMEM
        }
        if (blk + 1 < p.nblocks) grid_barrier(p.sync, (unsigned int)(blk + 1) * gridDim.x);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
struct TcsPlan {
    int th, tw, tz, IH, IW, IZ, NL, L0, NMB, NLA, ntiles, occ;
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
static bool plan_tile(const vq3d_preact_desc *d, int minb, TcsPlan &best) {
    using Cfg = TcsCfg<C, CB>;
    const int cand[] = {1, 2, 4, 8, 16, 32, 64};
    int forced[3] = {0, 0, 0};
    if (const char *e = getenv("VQ3D_TC_TILE")) sscanf(e, "%d,%d,%d", &forced[0], &forced[1], &forced[2]);
    bool found = false;
    const int nsm = sm_count();
    const size_t smem_cap = 227 * 1024 - 2048;
    for (int a : cand) for (int bq : cand) for (int c : cand) {
        int th = a, tw = bq, tz = c;
        if (forced[0] > 0) { th = forced[0]; tw = forced[1]; tz = forced[2]; }
        if (th > d->H) { if (a != cand[0] && th / 2 >= d->H) continue; th = d->H; }
        if (tw > d->W) { if (bq != cand[0] && tw / 2 >= d->W) continue; tw = d->W; }
        if (tz > d->Z) { if (c != cand[0] && tz / 2 >= d->Z) continue; tz = d->Z; }
        TcsPlan pl;
        pl.th = th; pl.tw = tw; pl.tz = tz;
        pl.IH = th + 2; pl.IW = tw + 2; pl.IZ = tz + 2;
        pl.NL = pl.IH * pl.IW * pl.IZ;
        pl.L0 = (pl.IW + 1) * pl.IZ + 1;
        const int Lend = (th * pl.IW + tw) * pl.IZ + tz;
        pl.NMB = (Lend - pl.L0 + 1 + 127) / 128;
        if (pl.NMB > kTcsMaxMB || pl.NMB * Cfg::CBP > 512) continue;
        pl.NLA = (pl.L0 + pl.NMB * 128 + pl.L0 + 7) & ~7;
        if (pl.NLA < pl.NL) pl.NLA = (pl.NL + 7) & ~7;
        if ((size_t)pl.NLA * 16 > 0x3fffu * 16) continue;                 // LBO field
        pl.smem = Cfg::smem_bytes(pl.NLA);
        if (pl.smem > smem_cap) continue;
        uint32_t cols = 32;
        while (cols < (uint32_t)(pl.NMB * Cfg::CBP)) cols <<= 1;
        pl.tmem_cols = cols;
        int occ = (int)((size_t)(227 * 1024) / (pl.smem + 1024));
        if (occ > (int)(512 / cols)) occ = (int)(512 / cols);
        if (occ > minb) occ = minb;
        if (occ < 1) continue;
        pl.occ = occ;
        pl.ntiles = d->B * (int)ceil_div(d->H, th) * (int)ceil_div(d->W, tw) * (int)ceil_div(d->Z, tz);
        // cycles per tile on one SM (128 lanes/cycle), plus a fixed latency per tile
        const double stage_a = (double)pl.NL * (C * CB + 8.0 * C + 8.0 * CB + 40.0) / 128.0;
        const double stage_c = (double)th * tw * tz * (C * CB + 6.0 * C + 8.0 * CB + 60.0) / 128.0 + pl.NMB * 128.0 * 30.0 / 128.0;
        const double mma = (double)pl.NMB * 27 * Cfg::NK * (Cfg::CBP / 2 > 16 ? Cfg::CBP / 2 : 16) / (pl.NMB < 8 ? pl.NMB : 8);
        const double work = stage_a + stage_c + mma;
        const double waves_sm = (double)ceil_div(pl.ntiles, nsm), waves_cta = (double)ceil_div(pl.ntiles, (int64_t)nsm * occ);
        pl.cost = waves_sm * work / (occ > 1 ? 1.0 : 0.8) + waves_cta * 2500.0;
        if (!found || pl.cost < best.cost) { best = pl; found = true; }
        if (forced[0] > 0) return found;
    }
    return found;
}

template <int C, int CB, int MINB>
static int launch_tcs(const vq3d_preact_desc *blocks, int n, float *tmp, unsigned int *sync_ws, void *stream) {
    using Cfg = TcsCfg<C, CB>;
    const vq3d_preact_desc *d = &blocks[0];
    TcsPlan pl;
    if (!plan_tile<C, CB>(d, MINB, pl)) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: no tile fits shared memory / TMEM");
    auto kernel = preact_tc_kernel<C, CB, MINB>;
    cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(attr)");
    e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(carveout)");
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kTcsThreads, pl.smem);
    if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(occupancy)");
    if (occ > (int)(512 / pl.tmem_cols)) occ = (int)(512 / pl.tmem_cols);   // never more resident CTAs than TMEM allows
    if (occ < 1) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: kernel does not fit on an SM");
    // TMEM: a resident CTA beyond 512/cols would spin in tcgen05.alloc; keep the smem request large enough to exclude it
    size_t smem = pl.smem;
    const size_t min_smem = (size_t)(227 * 1024) / (512 / pl.tmem_cols + 1) + 1;
    if (smem < min_smem) smem = min_smem;
    if (smem != pl.smem) {
        e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(attr)");
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kTcsThreads, smem);
        if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(occupancy)");
        if (occ < 1) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: kernel does not fit on an SM");
    }
    int grid = occ * sm_count();
    if (grid > pl.ntiles) grid = pl.ntiles;
    if (getenv("VQ3D_TC_DEBUG")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, reinterpret_cast<const void *>(kernel));
        int occ0 = -1, occ1 = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, kernel, kTcsThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, kernel, kTcsThreads, smem);
        fprintf(stderr, "  regs=%d static_smem=%zu maxdyn=%d carveout=%d occ(0 smem)=%d occ(smem)=%d\n", fa.numRegs, fa.sharedSizeBytes,
                fa.maxDynamicSharedSizeBytes, fa.preferredShmemCarveout, occ0, occ1);
    }
    if (getenv("VQ3D_TC_DEBUG"))
        fprintf(stderr, "preact_stack_tc<%d,%d>: %dx%dx%d n=%d tile %dx%dx%d NL=%d NMB=%d ntiles=%d occ=%d grid=%d smem=%zu tmem=%u\n", C, CB,
                d->H, d->W, d->Z, n, pl.th, pl.tw, pl.tz, pl.NL, pl.NMB, pl.ntiles, occ, grid, smem, pl.tmem_cols);

    TcsParams p;
    memset(&p, 0, sizeof(p));
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z;
    p.th = pl.th; p.tw = pl.tw; p.tz = pl.tz;
    p.nth = (int)ceil_div(d->H, pl.th); p.ntw = (int)ceil_div(d->W, pl.tw); p.ntz = (int)ceil_div(d->Z, pl.tz);
    p.IH = pl.IH; p.IW = pl.IW; p.IZ = pl.IZ; p.NL = pl.NL; p.L0 = pl.L0; p.NMB = pl.NMB; p.NLA = pl.NLA;
    p.ntiles = pl.ntiles; p.tmem_cols = pl.tmem_cols; p.sync = sync_ws;
    float *out = blocks[n - 1].y;
    const float *src = blocks[0].x;
    for (int i0 = 0; i0 < n; i0 += kTcsMaxBlocks) {
        const int nb = n - i0 < kTcsMaxBlocks ? n - i0 : kTcsMaxBlocks;
        p.nblocks = nb;
        p.x = src;
        // global block i must land in out when (n-1-i) is even; inside the chunk block j writes buf[j & 1]
        float *even = ((n - 1 - i0) % 2 == 0) ? out : tmp;
        float *odd = even == out ? tmp : out;
        p.buf[0] = even; p.buf[1] = odd;
        for (int j = 0; j < nb; ++j) {
            const vq3d_preact_desc &s = blocks[i0 + j];
            TcsBlock &t = p.blk[j];
            t.w1 = s.w1; t.w2 = s.w2; t.w3 = s.w3; t.ws = s.wskip;
            t.b1a = s.b1a; t.b1b = s.b1b; t.b2a = s.b2a; t.b2b = s.b2b; t.b3a = s.b3a; t.b3b = s.b3b;
            t.b4 = s.b4; t.scale = s.scale; t.b1c = s.b1c; t.b1d = s.b1d;
        }
        if (nb > 1) {
            e = cudaMemsetAsync(sync_ws, 0, sizeof(unsigned int), static_cast<cudaStream_t>(stream));
            if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(memset)");
            void *args[] = {&p};
            e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(kernel), dim3((unsigned)grid), dim3(kTcsThreads), args, smem,
                                            static_cast<cudaStream_t>(stream));
            if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(cooperative launch)");
        } else {
            kernel<<<dim3((unsigned)grid), dim3(kTcsThreads), smem, static_cast<cudaStream_t>(stream)>>>(p);
            e = cudaGetLastError();
            if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(launch)");
        }
        src = p.buf[(nb - 1) & 1];
    }
    return VQ3D_OK;
}

struct TcsEntry {
    int c, cb;
    int (*fn)(const vq3d_preact_desc *, int, float *, unsigned int *, void *);
};

static const TcsEntry kTcs[] = {
    {8, 4, launch_tcs<8, 4, 2>},     {16, 8, launch_tcs<16, 8, 2>},   {18, 9, launch_tcs<18, 9, 2>},
    {32, 16, launch_tcs<32, 16, 2>}, {64, 32, launch_tcs<64, 32, 1>}, {72, 36, launch_tcs<72, 36, 1>},
};

}  // namespace vq3d
#endif  // !VQ3D_EMU

using namespace vq3d;

extern "C" int vq3d_preact_stack_tc(const vq3d_preact_desc *blocks, int n, float *tmp, uint32_t *sync_ws, void *stream) {
#ifdef VQ3D_EMU
    (void)blocks; (void)n; (void)tmp; (void)sync_ws; (void)stream;
    return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: tensor-core kernels cannot run in the host emulator");
#else
    if (!blocks || n < 1 || !sync_ws) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: bad arguments");
    if (n > 1 && !tmp) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: n > 1 needs a scratch activation buffer");
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
    for (const TcsEntry &e : kTcs)
        if (e.c == d->Cin && e.cb == d->Cb) return e.fn(blocks, n, tmp, sync_ws, stream);
    return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: no tensor-core instantiation for C=%d Cb=%d", d->Cin, d->Cb);
#endif
}
