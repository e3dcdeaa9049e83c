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
//   warp 0 (1 lane) the 27 taps of conv2 = 27 accumulating tcgen05.mma (M128 x N=CBP x K16 per 16
//                   branch channels) per M-block of 128 CONSECUTIVE box indices: the A operand of tap
//                   (kh,kw,kz) is the same smem array shifted by ((kh-1)*IW + (kw-1))*IZ + (kz-1)
//                   rows -- a descriptor start-address change, no im2col.  Rows that fall on halo
//                   positions compute garbage that is never stored.  fp32 accumulators in TMEM.
//   warps 1..3      producers: wait for a free A buffer, issue the row copies of the next tile
//   consumer warps  in groups of 4 (one per TMEM lane quarter), one M-block at a time per group:
//                   tcgen05.ld D2 (lane = voxel) -> ELU -> bf16 -> smem -> tcgen05.mma with W3 (conv3)
//                   -> tcgen05.ld -> *scale + b4 + x -> store y (fp32, the reference's planar layout,
//                   in place) -> ELU -> bf16 -> smem -> tcgen05.mma with the NEXT block's W1 (conv1)
//                   -> tcgen05.ld -> ELU -> bf16 -> t1 workspace.  The two pointwise MMAs are issued
//                   by the group's own leader lane; SIMT only does ELU / convert / the residual add.
// Weights are converted once per call by a small prep kernel into bf16 B-operand images; a block's
// images (conv2, conv3, next conv1) are one contiguous bulk copy into shared memory.
// A buffers and TMEM accumulators are double buffered where they fit, so copies, MMAs and the
// epilogue of neighbouring tiles overlap inside the single resident CTA (kernels that use tcgen05
// get one CTA per SM from this driver).  Between blocks every CTA passes a grid barrier
// (cooperative launch).
#include "vq3d_rt.h"

#ifndef VQ3D_EMU
#include <cuda.h>            // CUtensorMap (the encoder itself is fetched through cudaGetDriverEntryPoint: no libcuda link)
#include <cuda_bf16.h>
#include <cstdlib>

namespace vq3d {

constexpr int kTcsMaxBlocks = 24;          // blocks per launch (kernel parameter space)
constexpr int kTcsAuxWarps = 4;            // warp 0 MMA issuer, warps 1..3 producers
constexpr int kTcsProdWarps = 3;
constexpr int kTcsMaxMB = 16;              // M-blocks per tile
#ifndef VQ3D_TCS_PACE
#define VQ3D_TCS_PACE 0
#endif
constexpr bool kTcsPace = VQ3D_TCS_PACE != 0;   // limit the conv2 batches queued in the tensor pipe (measured: no gain)

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
    int t1_ready;                          // 1: t1[0] was written by an earlier kernel and the residual of block 0 is already in y
                                           //    (the 'up' block: vq3d_preact_up_tc); no prologue, no leading grid barrier
    uint32_t tmem_cols;
    unsigned int *sync;                    // grid barrier counter (zeroed by the host before the launch)
    unsigned long long *trace;             // debug timeline or NULL
    const float *x;                        // input of block 0 (may alias y)
    float *y;                              // output of every block (updated in place from block 1 on)
    uint4 *t1[2];                          // bf16 t1 ping-pong: [B][NCH][H][W][Z+2] 16-byte units
    const unsigned char *wimg;             // per block [W1 | W2 | W3] bf16 B-operand images (+ one trailing W1 slot)
    // 5-D tensor maps over the two t1 buffers (dims, fastest first: 8 bf16 | Z+2 | W | H | B*NCH; box 8 x (Z+2) x (tw+2) x (th+2) x 1):
    // the whole haloed box of an interior tile is ONE cp.async.bulk.tensor per 8-channel chunk, landing in shared memory in
    // exactly the [h][w][z][16 B] order the shifted-window MMAs read.  Tiles whose halo wraps around the (circular) volume
    // edge keep the row-by-row bulk copies.
    int use_tma;
    alignas(64) CUtensorMap tmap[2];
    TcsBlock blk[kTcsMaxBlocks];
};

#include "tc_common.cuh"

// TMA tiled load of a 5-D box (UTMALDG): coordinates fastest dimension first; completion as tx bytes on `bar`
__device__ __forceinline__ void tma_load_5d(uint32_t dst_smem, const CUtensorMap *tmap, int c0, int c1, int c2, int c3, int c4, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(s_u32(bar)) : "memory");
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

// debug timeline (VQ3D_TC_TRACE=1): CTA 0 stamps %globaltimer of block 1 into 16 slots
__device__ __forceinline__ void tc_trace(unsigned long long *tr, int blk, int ev) {
    if (tr != nullptr && blockIdx.x == 0 && blk == 1) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        tr[ev] = t;
    }
}

template <int C, int CB>
struct TcsCfg {
    static constexpr int CBP = rup(CB, 16);            // branch channels padded: N and K of conv2, K of conv3, N of conv1
    static constexpr int CP = rup(C, 16);              // block channels padded: N of conv3, K of conv1
    static constexpr int NK2 = CBP / 16, NK1 = CP / 16;
    static constexpr int NCH = CBP / 8, CPCH = CP / 8; // 8-channel (16 B) chunks per voxel
    static constexpr int CB4 = rup(CB, 4);
    static constexpr uint32_t W2TAP = (uint32_t)CBP * CBP * 2;   // bytes of one tap's B operand
    // C_b = 9 (the 18-channel stack, the 18 -> 9 -> 8 'up' block): the ninth channel would double the reduction length of
    // every tap (K 9 -> 16).  Instead its three depth neighbours ride in the spare slots of the second 16-byte chunk
    // ([c8(z), c8(z-1), c8(z+1), 0 x 5]) and two taps share one K16 MMA through the descriptor's K-chunk stride (LBO):
    //   MMA a of (kh, kw):  channels 0..7 at depth tap -1  |  channels 0..7 at depth tap 0     (LBO = one row = 16 B)
    //   MMA b of (kh, kw):  channels 0..7 at depth tap +1  |  channel 8 at all three depth taps (LBO = chunk stride - 16 B)
    // 18 MMAs instead of 27 for the same products (an M128 x N16 x K16 MMA costs ~43 clk whatever it multiplies: its A operand
    // is fetched from shared memory at 128 B/clk, profiles/r02_tc_microbench.txt).
    static constexpr bool K9 = CB == 9;
    static constexpr int NMMA2 = K9 ? 18 : 27 * (CBP / 16);      // conv2 MMAs per M-block
    static constexpr uint32_t W2BYTES = K9 ? 18u * 512u : 27u * W2TAP;
    static constexpr uint32_t WPBYTES = (uint32_t)CP * CBP * 2;  // bytes of a pointwise (conv1 or conv3) B operand
    static constexpr uint32_t WIMG = W2BYTES + 2 * WPBYTES;      // per block image [W1 | W2 | W3]
    static constexpr uint32_t LBO_B2 = (uint32_t)CBP * 16;       // conv2 and conv1 images: CBP rows per k-chunk
    static constexpr uint32_t LBO_B3 = (uint32_t)CP * 16;        // conv3 image: CP rows per k-chunk
    static constexpr uint32_t SAP = (uint32_t)(CP > CBP ? CP : CBP) * 256;   // per consumer group staging: 128 rows x max(CP, CBP) bf16 (A3 aliases A1)
    static size_t smem_bytes(int NLA, int nbuf, int ng) {
        return 128 + WIMG + (size_t)nbuf * NCH * NLA * 16 + (size_t)ng * SAP;
    }
};

// ELU(v + a) + b for 16 accumulator columns, packed to two 16-byte bf16 chunks; columns >= nreal become 0
__device__ __forceinline__ void elu_pack16(const float *v, float a, float b, int nreal, uint4 &lo, uint4 &hi) {
    float t[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = j < nreal ? elu_bl(v[j] + a) + b : 0.0f;
    lo.x = bf16x2(t[0], t[1]); lo.y = bf16x2(t[2], t[3]); lo.z = bf16x2(t[4], t[5]); lo.w = bf16x2(t[6], t[7]);
    hi.x = bf16x2(t[8], t[9]); hi.y = bf16x2(t[10], t[11]); hi.z = bf16x2(t[12], t[13]); hi.w = bf16x2(t[14], t[15]);
}

// C_b = 9: channel 8 of voxel z goes into slot 0 of its own second chunk, slot 1 of voxel z+1's and slot 2 of voxel z-1's
// (circular in depth; three disjoint 2-byte stores, slots 3..7 stay zero from the launch's memset).  row1 = the (b, h, w) row of
// chunk 1 ([Z+2] 16-byte units, unit u = depth u - 1).
__device__ __forceinline__ void store_c8(uint4 *row1, int oz, int Z, float v8) {
    const unsigned short h = __bfloat16_as_ushort(__float2bfloat16_rn(v8));
    unsigned short *r = reinterpret_cast<unsigned short *>(row1);
    const int zn = oz + 1 == Z ? 0 : oz + 1, zp = oz == 0 ? Z - 1 : oz - 1;
    r[(oz + 1) * 8 + 0] = h;
    r[(zn + 1) * 8 + 1] = h;
    r[(zp + 1) * 8 + 2] = h;
}

// weights -> bf16 B-operand images (K-major, no swizzle: [k-chunk of 8][row n][16 B]), once per call
template <int C, int CB>
__global__ void __launch_bounds__(256)
tcs_prep_kernel(const __grid_constant__ TcsParams p, unsigned char *wimg) {
    using Cfg = TcsCfg<C, CB>;
    constexpr int CBP = Cfg::CBP, CP = Cfg::CP, NCH = Cfg::NCH, CPCH = Cfg::CPCH;
    const int blk = blockIdx.y;                         // 0..nblocks (the last one is the zero W1 slot only)
    unsigned char *img = wimg + (size_t)blk * Cfg::WIMG;
    const bool real = blk < p.nblocks;
    const TcsBlock &bp = p.blk[real ? blk : 0];
    const int n1 = CPCH * CBP, n2 = (Cfg::K9 ? 18 * 2 : 27 * NCH) * CBP, n3 = NCH * CP;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + (real ? n2 + n3 : 0); i += gridDim.x * blockDim.x) {
        float wv[8];
        unsigned char *dst;
        if (i < n1) {                                   // W1: rows n = cb, K = c
            const int n = i % CBP, kc = i / CBP;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = kc * 8 + e;
                wv[e] = (real && bp.w1 != nullptr && n < CB && c < C) ? __ldg(bp.w1 + n * C + c) : 0.0f;
            }
            dst = img + (size_t)kc * Cfg::LBO_B2 + (size_t)n * 16;
        } else if (i < n1 + n2) {                       // W2: per tap, rows n = co, K = ci
            const int j = i - n1;
            if (Cfg::K9) {                              // per MMA m = 2 * (kh * 3 + kw) + {a, b}: [k-chunk 2][n 16][16 B]
                const int n = j % CBP, kc = (j / CBP) % 2, m = j / (CBP * 2);
                const int hw = m >> 1;
                const float *wr = bp.w2 + (size_t)n * CB * 27;
#pragma unroll
                for (int e = 0; e < 8; ++e) wv[e] = 0.0f;
                if (n < CB) {
                    if ((m & 1) == 0 || kc == 0) {      // channels 0..7 of depth tap -1 / 0 (MMA a) or +1 (MMA b, first chunk)
                        const int kz = (m & 1) ? 2 : kc;
#pragma unroll
                        for (int e = 0; e < 8; ++e) wv[e] = __ldg(wr + e * 27 + hw * 3 + kz);
                    } else {                            // channel 8: slots = depth taps 0, -1, +1
                        wv[0] = __ldg(wr + 8 * 27 + hw * 3 + 1);
                        wv[1] = __ldg(wr + 8 * 27 + hw * 3 + 0);
                        wv[2] = __ldg(wr + 8 * 27 + hw * 3 + 2);
                    }
                }
                dst = img + Cfg::WPBYTES + (size_t)m * 512 + (size_t)kc * 256 + (size_t)n * 16;
            } else {
            const int n = j % CBP, kc = (j / CBP) % NCH, t = j / (CBP * NCH);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ci = kc * 8 + e;
                wv[e] = (n < CB && ci < CB) ? __ldg(bp.w2 + ((size_t)n * CB + ci) * 27 + t) : 0.0f;
            }
            dst = img + Cfg::WPBYTES + (size_t)t * Cfg::W2TAP + (size_t)kc * Cfg::LBO_B2 + (size_t)n * 16;
            }
        } else {                                        // W3: rows n = c, K = cb
            const int j = i - n1 - n2;
            const int n = j % CP, kc = j / CP;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int cb = kc * 8 + e;
                wv[e] = (n < C && cb < CB) ? __ldg(bp.w3 + n * CB + cb) : 0.0f;
            }
            dst = img + Cfg::WPBYTES + Cfg::W2BYTES + (size_t)kc * Cfg::LBO_B3 + (size_t)n * 16;
        }
        uint4 pk;
        pk.x = bf16x2(wv[0], wv[1]); pk.y = bf16x2(wv[2], wv[3]); pk.z = bf16x2(wv[4], wv[5]); pk.w = bf16x2(wv[6], wv[7]);
        *reinterpret_cast<uint4 *>(dst) = pk;
    }
}

template <int C, int CB, int NCW, int NBUF>
__global__ void __launch_bounds__((kTcsAuxWarps + NCW) * 32, 1)
preact_tc_kernel(const __grid_constant__ TcsParams p) {
    using Cfg = TcsCfg<C, CB>;
    constexpr int CBP = Cfg::CBP, CP = Cfg::CP, NK2 = Cfg::NK2, NK1 = Cfg::NK1, NCH = Cfg::NCH, CB4 = Cfg::CB4;
    constexpr int NT = (kTcsAuxWarps + NCW) * 32;
    constexpr int NG = NCW / 4;
    static_assert(NCW % 4 == 0 && NCW >= 4, "consumer warps come in groups of 4 (TMEM lane quarters)");
    static_assert(NBUF == 1 || NBUF == 2, "one or two A / accumulator buffers");
    VQ3D_DYN_SMEM(unsigned char, smem_raw);
    __shared__ __align__(8) uint64_t bar_full[2], bar_sa_empty[2], bar_tm_empty[2], bar_w, bar_g[NG], bar_mb[2][kTcsMaxMB];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (s_u32(smem_raw) + 127u) & ~127u;
    unsigned char *smem = smem_raw + (base - s_u32(smem_raw));
    // [W2 | W3 | W1next] (one contiguous image copy) | A buffers | per-group staging
    const uint32_t sW2_addr = base, sW3_addr = base + Cfg::W2BYTES, sW1_addr = sW3_addr + Cfg::WPBYTES;
    const uint32_t lbo_a = (uint32_t)p.NLA * 16;
    const uint32_t sa_bytes = (uint32_t)NCH * lbo_a;                  // one A buffer: [NCH][NLA][16 B]
    const uint32_t sA_addr = base + Cfg::WIMG;
    unsigned char *sA = smem + Cfg::WIMG;
    const uint32_t sAp_addr = sA_addr + NBUF * sa_bytes;
    unsigned char *sAp = sA + (size_t)NBUF * sa_bytes;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < 2; ++s) {
            mbarrier_init(&bar_full[s], kTcsProdWarps);
            mbarrier_init(&bar_sa_empty[s], 1);
            for (int m = 0; m < kTcsMaxMB; ++m) mbarrier_init(&bar_mb[s][m], 1);
            mbarrier_init(&bar_tm_empty[s], NCW);
        }
        mbarrier_init(&bar_w, 1);
        for (int g = 0; g < NG; ++g) mbarrier_init(&bar_g[g], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    const uint32_t tmem_grp = tmem_d + (uint32_t)(NBUF * p.NMB * CBP);    // per group: D3 (CP columns) then D1 (CBP columns)
    // D fp32 (1<<4), A/B bf16 (1<<7, 1<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc_b = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CBP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // N = CBP
    const uint32_t idesc_c = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);    // N = CP

    const int H = p.H, W = p.W, Z = p.Z;
    const int64_t S = (int64_t)H * W * Z;
    const int IW = p.IW, IZ = p.IZ, IWZ = p.IW * p.IZ;

    // ---- prologue: t1 of block 0 from x (pointwise SIMT, grid-stride over voxels; fp32 weights in the A area) ----
    if (!p.t1_ready) {
        const TcsBlock &b0 = p.blk[0];
        float *sw1 = reinterpret_cast<float *>(sA);          // [C][CB4]
        for (int i = tid; i < C * CB4; i += NT) {
            const int cb = i % CB4, ci = i / CB4;
            sw1[i] = cb < CB ? __ldg(b0.w1 + cb * C + ci) : 0.0f;
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
            float acc[CBP];
#pragma unroll
            for (int c = 0; c < CBP; ++c) acc[c] = 0.0f;
#pragma unroll(C <= 32 ? C : 8)
            for (int ci = 0; ci < C; ++ci) {
                const float a = elu1(__ldcg(px + (size_t)ci * S) + b1a) + b1b;
                const float4 *wr = reinterpret_cast<const float4 *>(sw1 + ci * CB4);
#pragma unroll
                for (int j = 0; j < CB4 / 4; ++j) {
                    const float4 w = wr[j];
                    acc[4 * j + 0] = __fmaf_rn(w.x, a, acc[4 * j + 0]);
                    acc[4 * j + 1] = __fmaf_rn(w.y, a, acc[4 * j + 1]);
                    acc[4 * j + 2] = __fmaf_rn(w.z, a, acc[4 * j + 2]);
                    acc[4 * j + 3] = __fmaf_rn(w.w, a, acc[4 * j + 3]);
                }
            }
#pragma unroll
            for (int ks = 0; ks < NK2; ++ks) {
                uint4 lo, hi;
                elu_pack16(acc + 16 * ks, b2a, b2b, CB - 16 * ks, lo, hi);
                uint4 *r0 = p.t1[0] + ((((size_t)b * NCH + 2 * ks) * H + oh) * W + ow) * (size_t)(Z + 2);
                uint4 *r1 = r0 + (size_t)H * W * (Z + 2);
                r0[oz + 1] = lo;
                if (oz == 0) r0[Z + 1] = lo;
                if (oz == Z - 1) r0[0] = lo;
                if (Cfg::K9) {
                    store_c8(r1, oz, Z, elu_bl(acc[8] + b2a) + b2b);
                } else {
                    r1[oz + 1] = hi;
                    if (oz == 0) r1[Z + 1] = hi;
                    if (oz == Z - 1) r1[0] = hi;
                }
            }
        }
        grid_barrier(p.sync, gridDim.x);
    }

    // tiles of this CTA (the same sequence in every block and for every role)
    const int my_tiles = p.ntiles > (int)blockIdx.x ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    uint32_t kt = 0;          // tiles processed by this CTA so far: buffer = kt % NBUF, use count = kt / NBUF
    uint32_t gph = 0;         // parity of this consumer group's pointwise-MMA barrier

    for (int blk = 0; blk < p.nblocks; ++blk) {
        const TcsBlock &bp = p.blk[blk];
        const bool has_next = blk + 1 < p.nblocks;
        if (tid == 0) tc_trace(p.trace, blk, 0);
        const uint4 *t1src = p.t1[blk & 1];
        uint4 *t1dst = p.t1[(blk + 1) & 1];
        const float *resid = (blk == 0 && !p.t1_ready) ? p.x : p.y;

        if (warp == 0) {
            // ================= weights (one image copy) + MMA issue for conv2 (warp 0, one elected lane) ==========
            if (elect_one()) {
                const unsigned char *src = p.wimg + (size_t)blk * Cfg::WIMG + Cfg::WPBYTES;    // [W2 | W3 | W1 of blk+1]
                mbarrier_arrive_expect_tx(&bar_w, Cfg::WIMG);
                for (uint32_t o = 0; o < Cfg::WIMG; o += 32768u) {
                    const uint32_t n = Cfg::WIMG - o < 32768u ? Cfg::WIMG - o : 32768u;
                    bulk_g2s(sW2_addr + o, src + o, n, &bar_w);
                }
            }
            __syncwarp();
            mbarrier_wait(&bar_w, (uint32_t)blk & 1u);
            if (lane == 0) tc_trace(p.trace, blk, 1);
            // conv2: one batch of 27*NK2 MMAs per M-block, each committed to its own barrier so the consumers
            // can start on M-block 0 while the rest of the tile is still in the pipe.  The tensor pipe
            // executes in issue order and the consumer groups issue their pointwise MMAs themselves, so
            // the queue is kept short: a batch is only issued once the batch two before it has completed.
            uint64_t *pace_bar[2] = {nullptr, nullptr};      // barriers of the two most recently issued batches
            uint32_t pace_par[2] = {0, 0};
            for (int i = 0; i < my_tiles; ++i) {
                const uint32_t k = kt + (uint32_t)i, s = k % NBUF, u = k / NBUF;
                mbarrier_wait(&bar_full[s], u & 1u);
                mbarrier_wait(&bar_tm_empty[s], (u & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (i == 0 && lane == 0) tc_trace(p.trace, blk, 3);
                const uint32_t a_base = sA_addr + s * sa_bytes;
                for (int mb = 0; mb < p.NMB; ++mb) {
                    if (kTcsPace && pace_bar[0] != nullptr) mbarrier_wait(pace_bar[0], pace_par[0]);     // at most two batches queued in the pipe
                    pace_bar[0] = pace_bar[1]; pace_par[0] = pace_par[1];
                    pace_bar[1] = &bar_mb[s][mb]; pace_par[1] = u & 1u;
                    if (elect_one()) {
                        const uint32_t row0 = (uint32_t)(p.L0 + mb * 128);
                        const uint32_t d_addr = tmem_d + (uint32_t)((s * p.NMB + mb) * CBP);
                        if constexpr (Cfg::K9) {
#pragma unroll
                            for (int hw = 0; hw < 9; ++hw) {
                                const int kh = hw / 3, kw = hw % 3;
                                const uint32_t arow = row0 + (uint32_t)((kh - 1) * IWZ + (kw - 1) * IZ);       // depth tap 0
                                // a: channels 0..7 at depth taps -1 | 0 (the second K chunk starts one row further)
                                umma_f16(d_addr, umma_desc(a_base + (arow - 1u) * 16u, 16, 128),
                                         umma_desc(sW2_addr + (uint32_t)(2 * hw) * 512u, 256, 128), idesc_b, hw > 0 ? 1u : 0u);
                                // b: channels 0..7 at depth tap +1 | channel 8's three depth taps (chunk 1 of the row itself)
                                umma_f16(d_addr, umma_desc(a_base + (arow + 1u) * 16u, lbo_a - 16u, 128),
                                         umma_desc(sW2_addr + (uint32_t)(2 * hw + 1) * 512u, 256, 128), idesc_b, 1u);
                            }
                        } else {
#pragma unroll
                        for (int tp = 0; tp < 27; ++tp) {
                            const int kh = tp / 9, kw = (tp / 3) % 3, kz = tp % 3;
                            const uint32_t arow = row0 + (uint32_t)((kh - 1) * IWZ + (kw - 1) * IZ + (kz - 1));
#pragma unroll
                            for (int ks = 0; ks < NK2; ++ks) {
                                const uint64_t adesc = umma_desc(a_base + arow * 16u + (uint32_t)(2 * ks) * lbo_a, lbo_a, 128);
                                const uint64_t bdesc = umma_desc(sW2_addr + (uint32_t)tp * Cfg::W2TAP + (uint32_t)(2 * ks) * Cfg::LBO_B2, Cfg::LBO_B2, 128);
                                umma_f16(d_addr, adesc, bdesc, idesc_b, (tp > 0 || ks > 0) ? 1u : 0u);
                            }
                        }
                        }
                        umma_commit_to(&bar_mb[s][mb]);                      // this M-block's accumulator is complete
                        if (mb + 1 == p.NMB) umma_commit_to(&bar_sa_empty[s]);  // ... and the A buffer has been read
                    }
                    __syncwarp();
                }
                if (i == 0 && lane == 0) tc_trace(p.trace, blk, 4);
            }
        } else if (warp < kTcsAuxWarps) {
            // ================= producers: haloed box rows -> A buffer (bulk copies), rows split over 3 warps ===
            const int nrows = p.IH * IW;
            const uint32_t row_bytes = (uint32_t)IZ * 16u;
            const int pw = warp - 1;
            const int ncopies = nrows * NCH;
            const int mine = ncopies > pw * 32 + lane ? (ncopies - 1 - (pw * 32 + lane)) / (kTcsProdWarps * 32) + 1 : 0;
            uint32_t wbytes = (uint32_t)mine * row_bytes;       // bytes this warp will deliver
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wbytes += __shfl_xor_sync(0xffffffffu, wbytes, o);
            // TMA path: chunk kc of an interior tile is one box copy, issued by lane 0 of producer warp kc % 3
            const int my_boxes = NCH > pw ? (NCH - 1 - pw) / kTcsProdWarps + 1 : 0;
            const uint32_t box_bytes = (uint32_t)p.NL * 16u;
            const CUtensorMap *tmap = &p.tmap[blk & 1];
            for (int i = 0; i < my_tiles; ++i) {
                const uint32_t k = kt + (uint32_t)i, s = k % NBUF, u = k / NBUF;
                int t = (int)blockIdx.x + i * (int)gridDim.x;
                const int twi = t % p.ntw; t /= p.ntw;
                const int thi = t % p.nth; t /= p.nth;
                const int b = t, oh0 = thi * p.th, ow0 = twi * p.tw;
                mbarrier_wait(&bar_sa_empty[s], (u & 1u) ^ 1u);
                if (p.use_tma && oh0 >= 1 && oh0 + p.th + 1 <= H && ow0 >= 1 && ow0 + p.tw + 1 <= W) {
                    if (lane == 0) {
                        mbarrier_arrive_expect_tx(&bar_full[s], (uint32_t)my_boxes * box_bytes);
                        for (int kc = pw; kc < NCH; kc += kTcsProdWarps)
                            tma_load_5d(sA_addr + s * sa_bytes + (uint32_t)kc * lbo_a, tmap, 0, 0, ow0 - 1, oh0 - 1, b * NCH + kc, &bar_full[s]);
                    }
                    __syncwarp();
                    if (i == 0 && tid == 32) tc_trace(p.trace, blk, 2);
                    continue;
                }
                if (lane == 0) mbarrier_arrive_expect_tx(&bar_full[s], wbytes);
                __syncwarp();
                for (int r = pw * 32 + lane; r < ncopies; r += kTcsProdWarps * 32) {
                    const int kc = r / nrows, rr = r - kc * nrows;
                    const int lh = rr / IW, lw = rr - lh * IW;
                    const int gh = pmodi(oh0 - 1 + lh, H), gw = pmodi(ow0 - 1 + lw, W);
                    const uint4 *src = t1src + ((((size_t)b * NCH + kc) * H + gh) * W + gw) * (size_t)(Z + 2);
                    bulk_g2s(sA_addr + s * sa_bytes + (uint32_t)kc * lbo_a + (uint32_t)rr * row_bytes, src, row_bytes, &bar_full[s]);
                }
                if (i == 0 && tid == 32) tc_trace(p.trace, blk, 2);
            }
        } else {
            // ================= consumers ==============================================================
            const int cw = warp - kTcsAuxWarps, g = cw >> 2, q = warp & 3;
            const int row = q * 32 + lane;
            const float b3a = ld_scalar(bp.b3a, 0.f), b3b = ld_scalar(bp.b3b, 0.f), b4 = ld_scalar(bp.b4, 0.f), sc = ld_scalar(bp.scale, 1.f);
            float n1a = 0.f, n1b = 0.f, n2a = 0.f, n2b = 0.f;
            if (has_next) {
                const TcsBlock &nb = p.blk[blk + 1];
                n1a = ld_scalar(nb.b1a, 0.f); n1b = ld_scalar(nb.b1b, 0.f); n2a = ld_scalar(nb.b2a, 0.f); n2b = ld_scalar(nb.b2b, 0.f);
            }
            unsigned char *my_ap = sAp + (size_t)g * Cfg::SAP + (size_t)row * 16;      // this row in the group's staging tile
            const uint32_t ap_addr = sAp_addr + (uint32_t)g * Cfg::SAP;
            const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
            const uint32_t d3_addr = tmem_grp + (uint32_t)(g * (CP + CBP)), d1_addr = d3_addr + CP;
            const int last_mb = p.NMB > g ? g + ((p.NMB - 1 - g) / NG) * NG : -1;
            const float inv_iz = 1.0f / (float)IZ, inv_iw = 1.0f / (float)IW;
            mbarrier_wait(&bar_w, (uint32_t)blk & 1u);          // the group leaders issue MMAs that read W3 / W1
            for (int i = 0; i < my_tiles; ++i) {
                const uint32_t k = kt + (uint32_t)i, s = k % NBUF, u = k / NBUF;
                int t = (int)blockIdx.x + i * (int)gridDim.x;
                const int twi = t % p.ntw; t /= p.ntw;
                const int thi = t % p.nth; t /= p.nth;
                const int b = t, oh0 = thi * p.th, ow0 = twi * p.tw;
                for (int mb = g; mb < p.NMB; mb += NG) {
                    const int L = p.L0 + mb * 128 + row;
                    // box index -> (lh, lw, lz) by reciprocal multiplication (exact: L < 2^15, divisors <= 258; the runtime-divisor
                    // integer divisions were 7 % of the consumers' samples)
                    const int r = (int)(((float)L + 0.5f) * inv_iz), lz = L - r * IZ;
                    const int lh = (int)(((float)r + 0.5f) * inv_iw), lw = r - lh * IW;
                    const int oh = oh0 + lh - 1, ow = ow0 + lw - 1, oz = lz - 1;
                    const bool valid = lh >= 1 && lh <= p.th && lw >= 1 && lw <= p.tw && lz >= 1 && lz <= Z && oh < H && ow < W;
                    const size_t off = valid ? (size_t)b * C * S + ((size_t)oh * W + ow) * Z + oz : 0;
                    const float *px = resid + off;
                    float xr[C <= 32 ? C : 16];
                    if (C <= 32 && valid) {       // residual loads in flight while the MMAs finish
                        const float *pc = px;
#pragma unroll
                        for (int j = 0; j < (C <= 32 ? C : 16); ++j) { xr[j] = __ldcg(pc); pc += S; }
                    }
                    mbarrier_wait(&bar_mb[s][mb], u & 1u);
                    if (i == 0 && mb == 0 && tid == kTcsAuxWarps * 32) tc_trace(p.trace, blk, 5);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // ---- phase 1: t2 = ELU(D2 + b3a) + b3b -> bf16 A operand of conv3 -------------------------
                    float v[16];
#pragma unroll
                    for (int ks = 0; ks < NK2; ++ks) {
                        tmem_ld16(tmem_d + lane_sel + (uint32_t)((s * p.NMB + mb) * CBP + ks * 16), v);
                        uint4 lo, hi;
                        elu_pack16(v, b3a, b3b, CB - 16 * ks, lo, hi);
                        *reinterpret_cast<uint4 *>(my_ap + (size_t)(2 * ks) * 2048) = lo;
                        *reinterpret_cast<uint4 *>(my_ap + (size_t)(2 * ks + 1) * 2048) = hi;
                    }
                    if (mb == last_mb) {          // this warp has read its last accumulator rows of the tile
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbarrier_arrive(&bar_tm_empty[s]);
                    }
                    const bool tr0 = i == 0 && mb == 0 && tid == kTcsAuxWarps * 32;
                    if (tr0) tc_trace(p.trace, blk, 9);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
                    if (q == 0 && elect_one()) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                        for (int ks = 0; ks < NK2; ++ks)
                            umma_f16(d3_addr, umma_desc(ap_addr + (uint32_t)(2 * ks) * 2048u, 2048, 128),
                                     umma_desc(sW3_addr + (uint32_t)(2 * ks) * Cfg::LBO_B3, Cfg::LBO_B3, 128), idesc_c, ks > 0 ? 1u : 0u);
                        umma_commit_to(&bar_g[g]);
                    }
                    mbarrier_wait(&bar_g[g], gph); gph ^= 1u;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (tr0) tc_trace(p.trace, blk, 10);
                    // ---- phase 2: y = D3*scale + b4 + x -> global; ELU(y + b1a') + b1b' -> bf16 A operand of conv1' ----
                    float *py = p.y + off;
#pragma unroll
                    for (int kc2 = 0; kc2 < NK1; ++kc2) {
                        constexpr int XB = C <= 32 ? 0 : 1;      // C > 32: residual loaded chunk by chunk
                        if (XB && valid) {
                            const float *pc = px + (size_t)(16 * kc2) * S;
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (16 * kc2 + j < C) { xr[j] = __ldcg(pc); pc += S; }
                        }
                        tmem_ld16(d3_addr + lane_sel + (uint32_t)(kc2 * 16), v);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (16 * kc2 + j < C) v[j] = __fmaf_rn(v[j], sc, b4);
                        if (valid) {
                            float *pc = py + (size_t)(16 * kc2) * S;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (16 * kc2 + j < C) {
                                    v[j] += xr[XB ? j : 16 * kc2 + j];
                                    *pc = v[j];
                                    pc += S;
                                }
                            }
                        }
                        if (has_next) {
                            uint4 lo, hi;
                            elu_pack16(v, n1a, n1b, C - 16 * kc2, lo, hi);
                            *reinterpret_cast<uint4 *>(my_ap + (size_t)(2 * kc2) * 2048) = lo;
                            *reinterpret_cast<uint4 *>(my_ap + (size_t)(2 * kc2 + 1) * 2048) = hi;
                        }
                    }
                    if (tr0) tc_trace(p.trace, blk, 11);
                    if (has_next) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
                        if (q == 0 && elect_one()) {
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                            for (int ks = 0; ks < NK1; ++ks)
                                umma_f16(d1_addr, umma_desc(ap_addr + (uint32_t)(2 * ks) * 2048u, 2048, 128),
                                         umma_desc(sW1_addr + (uint32_t)(2 * ks) * Cfg::LBO_B2, Cfg::LBO_B2, 128), idesc_b, ks > 0 ? 1u : 0u);
                            umma_commit_to(&bar_g[g]);
                        }
                        mbarrier_wait(&bar_g[g], gph); gph ^= 1u;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (tr0) tc_trace(p.trace, blk, 12);
                        // ---- phase 3: t1' = ELU(D1 + b2a') + b2b' -> bf16 -> workspace (with the circular depth halo) ----
#pragma unroll
                        for (int ks = 0; ks < NK2; ++ks) {
                            tmem_ld16(d1_addr + lane_sel + (uint32_t)(ks * 16), v);
                            uint4 lo, hi;
                            elu_pack16(v, n2a, n2b, CB - 16 * ks, lo, hi);
                            if (valid) {
                                uint4 *r0 = t1dst + ((((size_t)b * NCH + 2 * ks) * H + oh) * W + ow) * (size_t)(Z + 2);
                                uint4 *r1 = r0 + (size_t)H * W * (Z + 2);
                                r0[oz + 1] = lo;
                                if (oz == 0) r0[Z + 1] = lo;
                                if (oz == Z - 1) r0[0] = lo;
                                if (Cfg::K9) {
                                    store_c8(r1, oz, Z, elu_bl(v[8] + n2a) + n2b);
                                } else {
                                    r1[oz + 1] = hi;
                                    if (oz == 0) r1[Z + 1] = hi;
                                    if (oz == Z - 1) r1[0] = hi;
                                }
                            }
                        }
                        if (tr0) tc_trace(p.trace, blk, 13);
                    }
                }
                if (last_mb < 0) {                  // a group without M-blocks in this tile still keeps the phases in step
                    mbarrier_wait(&bar_mb[s][p.NMB - 1], u & 1u);
                    __syncwarp();
                    if (lane == 0) mbarrier_arrive(&bar_tm_empty[s]);
                }
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

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library links cudart statically and not libcuda)
typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encoder() {
    static TmapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TmapEncodeFn>(ptr);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// tensor maps of the two t1 buffers for a (th, tw) tiling; false = not available (the kernel then copies row by row)
static bool make_t1_tmaps(TcsParams &p, int NCH) {
    if (getenv("VQ3D_TC_NO_TMA")) return false;
    TmapEncodeFn enc = tmap_encoder();
    if (!enc || p.Z + 2 > 256 || p.tw + 2 > 256 || p.th + 2 > 256) return false;
    const cuuint64_t dims[5] = {8, (cuuint64_t)(p.Z + 2), (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B * NCH};
    const cuuint64_t strides[4] = {16, 16ull * (p.Z + 2), 16ull * (p.Z + 2) * p.W, 16ull * (p.Z + 2) * p.W * p.H};
    const cuuint32_t box[5] = {8, (cuuint32_t)(p.Z + 2), (cuuint32_t)(p.tw + 2), (cuuint32_t)(p.th + 2), 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 2; ++i)
        if (enc(&p.tmap[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.t1[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    return true;
}

template <int C, int CB, int NCW, int NBUF>
static bool plan_tile(const vq3d_preact_desc *d, TcsPlan &best) {
    using Cfg = TcsCfg<C, CB>;
    constexpr int NG = NCW / 4;
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
        if (pl.NMB > kTcsMaxMB) { if (forced[0] > 0) return false; continue; }
        const int cols_needed = NBUF * pl.NMB * Cfg::CBP + NG * (Cfg::CP + Cfg::CBP);
        bool ok = cols_needed <= 512;
        pl.NLA = (pl.L0 + pl.NMB * 128 + pl.L0 + 7) & ~7;
        if (pl.NLA < pl.NL) pl.NLA = (pl.NL + 7) & ~7;
        ok = ok && (size_t)pl.NLA * 16 <= 0x3fffu * 16;                   // LBO field
        pl.smem = Cfg::smem_bytes(pl.NLA, NBUF, NG);
        if (pl.smem < 128 + (size_t)Cfg::WIMG + (size_t)C * Cfg::CB4 * 4) ok = false;   // prologue scratch lives in the A area
        ok = ok && pl.smem <= smem_cap;
        if (ok) {
            uint32_t cols = 32;
            while (cols < (uint32_t)cols_needed) cols <<= 1;
            pl.tmem_cols = cols;
            pl.ntiles = d->B * (int)ceil_div(d->H, th) * (int)ceil_div(d->W, tw);
            // cycles per tile on one SM; epilogue (SIMT), tensor pipe (smem operand fetch bound) and copies overlap
            const double rows = (double)pl.NMB * 128;
            const double simt = rows * (14.0 * (C + 2 * CB) + 6.0 * C + 150.0) / 128.0 * (12.0 / NCW) + (double)ceil_div(pl.NMB, NG) * 1500.0;
            const double mma = rows * 27 * Cfg::NK2 * (32.0 + Cfg::CBP * 2.0) / 128.0 + rows * (Cfg::NK2 + Cfg::NK1) * 48.0 / 128.0;
            const double load = (double)pl.NL * Cfg::NCH * 16 / 64.0 + (double)pl.IH * pl.IW * Cfg::NCH * 60.0 / kTcsProdWarps;
            double tile = simt > mma ? simt : mma;
            if (load > tile) tile = load;
            if (NBUF == 1) tile = simt + mma + load;
            pl.cost = (double)ceil_div(pl.ntiles, nsm) * (tile + 300.0) + (NBUF == 1 ? 0.0 : load + mma + 3000.0);
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

// [256 B: barrier counter, trace] [t1 x 2] [weight images of up to kTcsMaxBlocks blocks + one W1 slot]
template <int C, int CB>
static size_t ws_bytes(const vq3d_preact_desc *d) {
    return 256 + 2 * t1_units<C, CB>(d) * 16 + (size_t)(kTcsMaxBlocks + 1) * TcsCfg<C, CB>::WIMG;
}

// C_b = 9: slots 3..7 of every second-chunk unit must read as zero (they meet zero weights, but 0 x NaN would poison the sum)
template <int C, int CB>
static cudaError_t zero_c8_planes(uint4 *t1, const vq3d_preact_desc *d, cudaStream_t st) {
    if (!TcsCfg<C, CB>::K9) return cudaSuccess;
    const size_t plane = (size_t)d->H * d->W * (size_t)(d->Z + 2) * 16;
    return cudaMemset2DAsync(reinterpret_cast<unsigned char *>(t1) + plane, 2 * plane, 0, plane, (size_t)d->B, st);
}

template <int C, int CB, int NCW, int NBUF>
static int launch_tcs(const vq3d_preact_desc *blocks, int n, void *ws, size_t ws_size, void *stream) {
    using Cfg = TcsCfg<C, CB>;
    constexpr int NT = (kTcsAuxWarps + NCW) * 32;
    const vq3d_preact_desc *d = &blocks[0];
    if (ws_size < ws_bytes<C, CB>(d)) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: workspace too small (%zu < %zu bytes)", ws_size, ws_bytes<C, CB>(d));
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(VQ3D_ERR_INVALID, "preact_stack_tc: workspace must be 256-byte aligned");
    TcsPlan pl;
    if (!plan_tile<C, CB, NCW, NBUF>(d, pl)) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: no tile fits shared memory / TMEM");
    auto kernel = preact_tc_kernel<C, CB, NCW, NBUF>;
    size_t smem = pl.smem < 120 * 1024 ? 120 * 1024 : pl.smem;     // > half an SM: never two resident CTAs (they would fight over TMEM)
    cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(attr)");
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, NT, smem);
    if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(occupancy)");
    if (occ < 1) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_tc: kernel does not fit on an SM");
    int grid = sm_count();
    if (grid > pl.ntiles) grid = pl.ntiles;
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
    unsigned char *wsb = static_cast<unsigned char *>(ws);
    p.sync = reinterpret_cast<unsigned int *>(wsb);
    const bool trace = getenv("VQ3D_TC_TRACE") != nullptr;
    p.trace = trace ? reinterpret_cast<unsigned long long *>(wsb + 64) : nullptr;
    p.t1[0] = reinterpret_cast<uint4 *>(wsb + 256);
    p.t1[1] = p.t1[0] + t1_units<C, CB>(d);
    unsigned char *wimg = reinterpret_cast<unsigned char *>(p.t1[1] + t1_units<C, CB>(d));
    p.wimg = wimg;
    p.y = blocks[n - 1].y;
    p.use_tma = make_t1_tmaps(p, Cfg::NCH) ? 1 : 0;
    if (getenv("VQ3D_TC_DEBUG")) fprintf(stderr, "preact_stack_tc<%d,%d>: tensor-map TMA box loads %s\n", C, CB, p.use_tma ? "on" : "off (row copies)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    e = zero_c8_planes<C, CB>(p.t1[0], d, st);
    if (e == cudaSuccess) e = zero_c8_planes<C, CB>(p.t1[1], d, st);
    if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(memset chunk 1)");
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
        e = cudaMemsetAsync(p.sync, 0, sizeof(unsigned int), st);
        if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(memset)");
        const int items = Cfg::CPCH * Cfg::CBP + 27 * Cfg::NCH * Cfg::CBP + Cfg::NCH * Cfg::CP;
        tcs_prep_kernel<C, CB><<<dim3((unsigned)ceil_div(items, 256), (unsigned)(nb + 1)), 256, 0, st>>>(p, wimg);
        e = cudaGetLastError();
        if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(prep)");
        void *args[] = {&p};
        e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(kernel), dim3((unsigned)grid), dim3(NT), args, smem, st);
        if (e != cudaSuccess) return check_cuda(e, "preact_stack_tc(cooperative launch)");
        if (trace) {
            unsigned long long h[16];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
            const char *names[14] = {"block start", "weights landed", "first copies issued", "A full (MMA)", "MMAs committed", "TMEM full (consumer)",
                                     "tile 0 drained", "all tiles drained", "grid barrier passed", "  mb0 phase 1 staged", "  mb0 conv3 MMA done",
                                     "  mb0 phase 2 staged", "  mb0 conv1 MMA done", "  mb0 phase 3 stored"};
            const int order[14] = {0, 1, 2, 3, 4, 5, 9, 10, 11, 12, 13, 6, 7, 8};
            for (int k = 0; k < 14; ++k)
                fprintf(stderr, "  trace %-24s +%6.2f us\n", names[order[k]], (double)(long long)(h[order[k]] - h[0]) * 1e-3);
        }
    }
    return VQ3D_OK;
}

struct TcsEntry {
    int c, cb;
    int (*fn)(const vq3d_preact_desc *, int, void *, size_t, void *);
    size_t (*ws)(const vq3d_preact_desc *);
};

#define VQ3D_TCS(C, CB, NCW, NBUF) {C, CB, launch_tcs<C, CB, NCW, NBUF>, ws_bytes<C, CB>}
static const TcsEntry kTcs[] = {
    VQ3D_TCS(8, 4, 16, 2), VQ3D_TCS(16, 8, 24, 2), VQ3D_TCS(18, 9, 20, 2), VQ3D_TCS(32, 16, 16, 2), VQ3D_TCS(64, 32, 8, 2), VQ3D_TCS(72, 36, 8, 1),
};

static const TcsEntry *find_tcs(const vq3d_preact_desc *d) {
    for (const TcsEntry &e : kTcs)
        if (e.c == d->Cin && e.cb == d->Cb && d->Cout == d->Cin) return &e;
    return nullptr;
}

// =====================================================================================================================
// 'up' blocks (mode 2, layers.py:124-132,176-195,591-597) on the same tensor-core kernel.
//
//   t1_lo = ELU(conv1x1(ELU(x+b1a)+b1b) + b2a) + b2b                  low resolution, fp32        (up_lo_kernel, SIMT)
//   s_lo  = conv1x1(x + b1c; w_skip) + b1d                            low resolution, fp32        (the same kernel)
//   t1_hi = bf16(trilinear_x2(t1_lo))   -> the stack kernel's haloed workspace layout             (up_expand_kernel, SIMT)
//   y     = trilinear_x2(s_lo)          (the 1x1 skip convolution commutes with the interpolation, whose weights sum to 1)
//   y    += conv1x1(ELU(conv3x3x3_circular(t1_hi) + b3a) + b3b) * scale + b4     preact_tc_kernel<Cout, Cb>, one block,
//                                                                                t1_ready: no prologue, residual = y
// The low-resolution work is 1/8 of the voxels; the expansion is write bound (32 B of bf16 t1 + 4 Cout B per output voxel).
template <int T>
__global__ void __launch_bounds__(T)
up_lo_kernel(const float *__restrict__ x, const float *__restrict__ w1, const float *__restrict__ wskip, const float *b1a_p, const float *b1b_p,
             const float *b2a_p, const float *b2b_p, const float *b1c_p, const float *b1d_p, float *__restrict__ t1_lo, float *__restrict__ s_lo,
             int64_t total, int64_t S, int Cin, int Cb, int Cout) {
    VQ3D_DYN_SMEM(float, sm);
    const int Cb4 = (Cb + 3) & ~3, Co4 = (Cout + 3) & ~3;
    float *sw1 = sm;                                   // [Cin][Cb4]
    float *swk = sw1 + (size_t)Cin * Cb4;              // [Cin][Co4]
    float *sx = swk + (size_t)Cin * Co4;               // [Cin][T] activated input
    float *sr = sx + (size_t)Cin * T;                  // [Cin][T] raw input + b1c
    const int tid = threadIdx.x;
    for (int i = tid; i < Cin * Cb4; i += T) { const int cb = i % Cb4, ci = i / Cb4; sw1[i] = cb < Cb ? __ldg(w1 + (size_t)cb * Cin + ci) : 0.0f; }
    for (int i = tid; i < Cin * Co4; i += T) { const int co = i % Co4, ci = i / Co4; swk[i] = co < Cout ? __ldg(wskip + (size_t)co * Cin + ci) : 0.0f; }
    const float b1a = ld_scalar(b1a_p, 0.f), b1b = ld_scalar(b1b_p, 0.f), b2a = ld_scalar(b2a_p, 0.f), b2b = ld_scalar(b2b_p, 0.f);
    const float b1c = ld_scalar(b1c_p, 0.f), b1d = ld_scalar(b1d_p, 0.f);
    __syncthreads();
    for (int64_t v = (int64_t)blockIdx.x * T + tid; v < total; v += (int64_t)gridDim.x * T) {
        const int64_t b = v / S, r = v - b * S;
        const float *px = x + (size_t)b * Cin * S + r;
        for (int c0 = 0; c0 < Cin; c0 += 6) {             // six independent loads in flight per pass
            float xv[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) xv[j] = c0 + j < Cin ? __ldcs(px + (size_t)(c0 + j) * S) : 0.0f;
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (c0 + j < Cin) {
                    sx[(c0 + j) * T + tid] = elu1(xv[j] + b1a) + b1b;
                    sr[(c0 + j) * T + tid] = xv[j] + b1c;
                }
        }
        float *pt = t1_lo + (size_t)b * Cb * S + r;
        for (int c0 = 0; c0 < Cb; c0 += 4) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int ci = 0; ci < Cin; ++ci) {
                const float a = sx[ci * T + tid];
                const float4 w = *reinterpret_cast<const float4 *>(sw1 + ci * Cb4 + c0);
                a0 = __fmaf_rn(w.x, a, a0); a1 = __fmaf_rn(w.y, a, a1); a2 = __fmaf_rn(w.z, a, a2); a3 = __fmaf_rn(w.w, a, a3);
            }
            const float o[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < Cb) pt[(size_t)(c0 + j) * S] = elu1(o[j] + b2a) + b2b;
        }
        float *ps = s_lo + (size_t)b * Cout * S + r;
        for (int c0 = 0; c0 < Cout; c0 += 4) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int ci = 0; ci < Cin; ++ci) {
                const float a = sr[ci * T + tid];
                const float4 w = *reinterpret_cast<const float4 *>(swk + ci * Co4 + c0);
                a0 = __fmaf_rn(w.x, a, a0); a1 = __fmaf_rn(w.y, a, a1); a2 = __fmaf_rn(w.z, a, a2); a3 = __fmaf_rn(w.w, a, a3);
            }
            const float o[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < Cout) ps[(size_t)(c0 + j) * S] = o[j] + b1d;
        }
    }
}

// nn.Upsample(scale_factor=2, 'trilinear', align_corners=False) (layers.py:594): out[2i] = .25 in[max(i-1,0)] + .75 in[i],
// out[2i+1] = .75 in[i] + .25 in[min(i+1,n-1)] per axis.  One work item = one low-resolution depth index of one output (h2, w2)
// column = two output voxels of 8 channels; a CTA owns a compact 4 x 8 patch of output columns so that the 12 low-resolution
// values an item reads per channel are L1 hits for all but the first toucher (a row-major item order made every load an L2
// access: 13.7 GB of L2 traffic for the 18 -> 9 -> 8 block at batch 8).  blockIdx.y selects the 8-channel job: the bf16 t1 chunks
// first (written in the tensor-core operand layout, depth halo included), then the fp32 skip chunks (written into y).
constexpr int kUpPH = 4, kUpPW = 8;
__global__ void __launch_bounds__(256)
up_expand_kernel(const float *__restrict__ t1_lo, const float *__restrict__ s_lo, uint4 *__restrict__ t1_hi, float *__restrict__ y,
                 int B, int Cb, int Cout, int H, int W, int Z, int NCH, int k9) {
    const int H2 = 2 * H, W2 = 2 * W, Z2 = 2 * Z;
    const int nph = (H2 + kUpPH - 1) / kUpPH, npw = (W2 + kUpPW - 1) / kUpPW;
    int t = blockIdx.x;
    const int pwi = t % npw; t /= npw;
    const int phi = t % nph;
    const int b = t / nph;
    const int job = blockIdx.y;
    const bool is_t1 = job < NCH;
    const float *src = is_t1 ? t1_lo : s_lo;
    const int Cs = is_t1 ? Cb : Cout, c0 = (is_t1 ? job : job - NCH) * 8;
    const int nc = min(8, Cs - c0);                     // real channels of this job (uniform over the CTA)
    const size_t S = (size_t)H * W * Z;
    for (int item = threadIdx.x; item < kUpPH * kUpPW * Z; item += blockDim.x) {
        const int zl = item % Z, col = item / Z;
        const int h2 = phi * kUpPH + col / kUpPW, w2 = pwi * kUpPW + col % kUpPW;
        if (h2 >= H2 || w2 >= W2) continue;
        int h0, h1, w0, w1;
        float fh0, fh1, fw0, fw1;
        if (h2 & 1) { h0 = h2 >> 1; h1 = min(h0 + 1, H - 1); fh0 = 0.75f; fh1 = 0.25f; } else { h1 = h2 >> 1; h0 = max(h1 - 1, 0); fh0 = 0.25f; fh1 = 0.75f; }
        if (w2 & 1) { w0 = w2 >> 1; w1 = min(w0 + 1, W - 1); fw0 = 0.75f; fw1 = 0.25f; } else { w1 = w2 >> 1; w0 = max(w1 - 1, 0); fw0 = 0.25f; fw1 = 0.75f; }
        const int zm = max(zl - 1, 0), zp = min(zl + 1, Z - 1);
        // twelve source pointers (4 (h, w) corners x 3 depth taps), advanced by one channel volume per channel: two integer
        // instructions per load instead of the six a recomputed 64-bit address costs
        const int o00 = (h0 * W + w0) * Z, o01 = (h0 * W + w1) * Z, o10 = (h1 * W + w0) * Z, o11 = (h1 * W + w1) * Z;
        const float f00 = fh0 * fw0, f01 = fh0 * fw1, f10 = fh1 * fw0, f11 = fh1 * fw1;
        float e[8], o[8];                               // even / odd output depth
        const float *pc = src + ((size_t)b * Cs + c0) * S;
        const float *q00m = pc + o00 + zm, *q01m = pc + o01 + zm, *q10m = pc + o10 + zm, *q11m = pc + o11 + zm;
        const float *q00c = pc + o00 + zl, *q01c = pc + o01 + zl, *q10c = pc + o10 + zl, *q11c = pc + o11 + zl;
        const float *q00p = pc + o00 + zp, *q01p = pc + o01 + zp, *q10p = pc + o10 + zp, *q11p = pc + o11 + zp;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            e[j] = 0.0f; o[j] = 0.0f;
            if (j < nc) {
                const float vm = f00 * __ldg(q00m) + f01 * __ldg(q01m) + f10 * __ldg(q10m) + f11 * __ldg(q11m);
                const float vc = f00 * __ldg(q00c) + f01 * __ldg(q01c) + f10 * __ldg(q10c) + f11 * __ldg(q11c);
                const float vp = f00 * __ldg(q00p) + f01 * __ldg(q01p) + f10 * __ldg(q10p) + f11 * __ldg(q11p);
                e[j] = 0.25f * vm + 0.75f * vc;
                o[j] = 0.75f * vc + 0.25f * vp;
                q00m += S; q01m += S; q10m += S; q11m += S; q00c += S; q01c += S; q10c += S; q11c += S; q00p += S; q01p += S; q10p += S; q11p += S;
            }
        }
        if (is_t1 && k9 && job == 1) {                 // C_b = 9: the ninth channel and its depth neighbours (TcsCfg::K9)
            uint4 *row = t1_hi + ((((size_t)b * NCH + 1) * H2 + h2) * W2 + w2) * (size_t)(Z2 + 2);
            store_c8(row, 2 * zl, Z2, e[0]);
            store_c8(row, 2 * zl + 1, Z2, o[0]);
        } else if (is_t1) {
            uint4 ue, uo;
            ue.x = bf16x2(e[0], e[1]); ue.y = bf16x2(e[2], e[3]); ue.z = bf16x2(e[4], e[5]); ue.w = bf16x2(e[6], e[7]);
            uo.x = bf16x2(o[0], o[1]); uo.y = bf16x2(o[2], o[3]); uo.z = bf16x2(o[4], o[5]); uo.w = bf16x2(o[6], o[7]);
            uint4 *row = t1_hi + ((((size_t)b * NCH + job) * H2 + h2) * W2 + w2) * (size_t)(Z2 + 2);
            row[1 + 2 * zl] = ue;
            row[2 + 2 * zl] = uo;
            if (zl == 0) row[Z2 + 1] = ue;              // circular depth halo (the padding is applied AFTER the upsampling)
            if (zl == Z - 1) row[0] = uo;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nc)
                    __stcs(reinterpret_cast<float2 *>(y + ((((size_t)b * Cout + c0 + j) * H2 + h2) * W2 + w2) * (size_t)Z2 + 2 * zl), make_float2(e[j], o[j]));
        }
    }
}

static vq3d_preact_desc up_hi_desc(const vq3d_preact_desc *d) {      // the one-block 'same' problem the tensor-core kernel sees
    vq3d_preact_desc h = *d;
    h.H = 2 * d->H; h.W = 2 * d->W; h.Z = 2 * d->Z;
    h.Cin = d->Cout;
    return h;
}

// [256 B: barrier counter, trace] [t1_hi] [weight images: 2 slots] [t1_lo fp32] [s_lo fp32]
template <int C, int CB>
static size_t up_ws_bytes(const vq3d_preact_desc *d) {
    const vq3d_preact_desc h = up_hi_desc(d);
    const size_t lo = (size_t)d->B * d->H * d->W * d->Z;
    return 256 + t1_units<C, CB>(&h) * 16 + 2 * (size_t)TcsCfg<C, CB>::WIMG + 256 + lo * (size_t)(d->Cb + d->Cout) * 4;
}

template <int C, int CB, int NCW, int NBUF>
static int launch_up_tc(const vq3d_preact_desc *d, void *ws, size_t ws_size, void *stream) {
    using Cfg = TcsCfg<C, CB>;
    constexpr int NT = (kTcsAuxWarps + NCW) * 32;
    if (ws_size < up_ws_bytes<C, CB>(d)) return fail(VQ3D_ERR_INVALID, "preact_up_tc: workspace too small (%zu < %zu bytes)", ws_size, up_ws_bytes<C, CB>(d));
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(VQ3D_ERR_INVALID, "preact_up_tc: workspace must be 256-byte aligned");
    const vq3d_preact_desc h = up_hi_desc(d);
    TcsPlan pl;
    if (!plan_tile<C, CB, NCW, NBUF>(&h, pl)) return fail(VQ3D_ERR_UNSUPPORTED, "preact_up_tc: no tile fits shared memory / TMEM");
    auto kernel = preact_tc_kernel<C, CB, NCW, NBUF>;
    size_t smem = pl.smem < 120 * 1024 ? 120 * 1024 : pl.smem;
    cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "preact_up_tc(attr)");
    int grid = sm_count();
    if (grid > pl.ntiles) grid = pl.ntiles;
    if (getenv("VQ3D_TC_DEBUG"))
        fprintf(stderr, "preact_up_tc<%d,%d>: lo %dx%dx%d Cin=%d tile %dx%dx%d NL=%d NMB=%d ntiles=%d grid=%d smem=%zu tmem=%u\n", C, CB, d->H, d->W, d->Z,
                d->Cin, pl.th, pl.tw, h.Z, pl.NL, pl.NMB, pl.ntiles, grid, smem, pl.tmem_cols);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char *wsb = static_cast<unsigned char *>(ws);
    const size_t t1u = t1_units<C, CB>(&h);
    uint4 *t1_hi = reinterpret_cast<uint4 *>(wsb + 256);
    unsigned char *wimg = reinterpret_cast<unsigned char *>(t1_hi + t1u);
    const size_t lo = (size_t)d->B * d->H * d->W * d->Z;
    float *t1_lo = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(wimg + 2 * (size_t)Cfg::WIMG) + 255) & ~(uintptr_t)255);
    float *s_lo = t1_lo + lo * d->Cb;

    e = zero_c8_planes<C, CB>(t1_hi, &h, st);
    if (e != cudaSuccess) return check_cuda(e, "preact_up_tc(memset chunk 1)");
    // low-resolution pointwise stage
    {
        constexpr int T = 128;
        const int Cb4 = (d->Cb + 3) & ~3, Co4 = (d->Cout + 3) & ~3;
        const size_t sm = ((size_t)d->Cin * (Cb4 + Co4) + 2 * (size_t)d->Cin * T) * 4;
        if (sm > 200 * 1024) return fail(VQ3D_ERR_UNSUPPORTED, "preact_up_tc: Cin = %d too wide for the low-resolution stage", d->Cin);
        int64_t blocks = ceil_div((int64_t)lo, T);
        if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
        int rc = launch("up_lo", up_lo_kernel<T>, dim3((unsigned)blocks), dim3(T), sm, stream, d->x, d->w1, d->wskip, d->b1a, d->b1b, d->b2a, d->b2b,
                        d->b1c, d->b1d, t1_lo, s_lo, (int64_t)lo, (int64_t)d->H * d->W * d->Z, (int)d->Cin, (int)d->Cb, (int)d->Cout);
        if (rc != VQ3D_OK) return rc;
    }
    // expansion: bf16 t1 in the tensor-core layout + the interpolated skip path into y
    {
        const unsigned jobs = (unsigned)(Cfg::NCH + (d->Cout + 7) / 8);
        const unsigned patches = (unsigned)(d->B * ceil_div(h.H, kUpPH) * ceil_div(h.W, kUpPW));
        int rc = launch("up_expand", up_expand_kernel, dim3(patches, jobs), dim3(256), 0, stream, (const float *)t1_lo, (const float *)s_lo,
                        t1_hi, d->y, (int)d->B, (int)d->Cb, (int)d->Cout, (int)d->H, (int)d->W, (int)d->Z, (int)Cfg::NCH, (int)Cfg::K9);
        if (rc != VQ3D_OK) return rc;
    }
    TcsParams p;
    memset(&p, 0, sizeof(p));
    p.B = h.B; p.H = h.H; p.W = h.W; p.Z = h.Z;
    p.th = pl.th; p.tw = pl.tw;
    p.nth = (int)ceil_div(h.H, pl.th); p.ntw = (int)ceil_div(h.W, pl.tw);
    p.IH = pl.IH; p.IW = pl.IW; p.IZ = pl.IZ; p.NL = pl.NL; p.L0 = pl.L0; p.NMB = pl.NMB; p.NLA = pl.NLA;
    p.ntiles = pl.ntiles; p.tmem_cols = pl.tmem_cols;
    p.nblocks = 1;
    p.t1_ready = 1;
    p.sync = reinterpret_cast<unsigned int *>(wsb);
    p.trace = nullptr;
    p.t1[0] = t1_hi; p.t1[1] = t1_hi;
    p.wimg = wimg;
    p.use_tma = make_t1_tmaps(p, Cfg::NCH) ? 1 : 0;
    p.x = d->y; p.y = d->y;
    TcsBlock &t = p.blk[0];
    t.w1 = nullptr; t.w2 = d->w2; t.w3 = d->w3;
    t.b3a = d->b3a; t.b3b = d->b3b; t.b4 = d->b4; t.scale = d->scale;
    e = cudaMemsetAsync(p.sync, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return check_cuda(e, "preact_up_tc(memset)");
    const int items = Cfg::CPCH * Cfg::CBP + 27 * Cfg::NCH * Cfg::CBP + Cfg::NCH * Cfg::CP;
    tcs_prep_kernel<C, CB><<<dim3((unsigned)ceil_div(items, 256), 2u), 256, 0, st>>>(p, wimg);
    e = cudaGetLastError();
    if (e != cudaSuccess) return check_cuda(e, "preact_up_tc(prep)");
    void *args[] = {&p};
    e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(kernel), dim3((unsigned)grid), dim3(NT), args, smem, st);
    if (e != cudaSuccess) return check_cuda(e, "preact_up_tc(cooperative launch)");
    return VQ3D_OK;
}

struct UpEntry {
    int cout, cb;
    int (*fn)(const vq3d_preact_desc *, void *, size_t, void *);
    size_t (*ws)(const vq3d_preact_desc *);
};
#define VQ3D_UP(C, CB, NCW, NBUF) {C, CB, launch_up_tc<C, CB, NCW, NBUF>, up_ws_bytes<C, CB>}
// (Cout, Cb) of the Full / downscaled models' wide 'up' blocks: 18->9->8, 32->16->16, 16->8->8, 72->36->32
static const UpEntry kUp[] = {VQ3D_UP(8, 9, 20, 2), VQ3D_UP(16, 16, 20, 2), VQ3D_UP(8, 8, 20, 2), VQ3D_UP(32, 36, 8, 1)};

static const UpEntry *find_up(const vq3d_preact_desc *d) {
    if (d->mode != 2 || !d->wskip) return nullptr;
    for (const UpEntry &e : kUp)
        if (e.cout == d->Cout && e.cb == d->Cb) return &e;
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

extern "C" size_t vq3d_preact_up_tc_workspace(const vq3d_preact_desc *d) {
#ifdef VQ3D_EMU
    (void)d;
    return 0;
#else
    if (!d || d->B < 1 || d->H < 1 || d->W < 1 || d->Z < 1 || d->Cin < 1) return 0;
    const UpEntry *e = find_up(d);
    return e ? e->ws(d) : 0;
#endif
}

extern "C" int vq3d_preact_up_tc(const vq3d_preact_desc *d, void *ws, size_t ws_size, void *stream) {
#ifdef VQ3D_EMU
    (void)d; (void)ws; (void)ws_size; (void)stream;
    return fail(VQ3D_ERR_UNSUPPORTED, "preact_up_tc: tensor-core kernels cannot run in the host emulator");
#else
    if (!d) return fail(VQ3D_ERR_INVALID, "preact_up_tc: null descriptor");
    if (d->mode != 2) return fail(VQ3D_ERR_INVALID, "preact_up_tc: mode must be 2 (up)");
    if (!d->x || !d->y || !d->w1 || !d->w2 || !d->w3 || !d->wskip) return fail(VQ3D_ERR_INVALID, "preact_up_tc: null tensor");
    if (d->B < 1 || d->H < 1 || d->W < 1 || d->Z < 1 || d->Cin < 1) return fail(VQ3D_ERR_INVALID, "preact_up_tc: bad sizes");
    if (d->out_w || d->pre_w) return fail(VQ3D_ERR_UNSUPPORTED, "preact_up_tc: fused leading / trailing 1x1 convolutions are not supported here");
    if ((int64_t)d->B * d->Cout * d->H * d->W * d->Z * 8 > ((int64_t)1 << 40) || (int64_t)d->H * d->W * d->Z >= ((int64_t)1 << 31))
        return fail(VQ3D_ERR_INVALID, "preact_up_tc: tensor too large");
    const UpEntry *e = find_up(d);
    if (!e) return fail(VQ3D_ERR_UNSUPPORTED, "preact_up_tc: no tensor-core instantiation for Cout=%d Cb=%d", d->Cout, d->Cb);
    if (!ws) return fail(VQ3D_ERR_INVALID, "preact_up_tc: null workspace (see vq3d_preact_up_tc_workspace)");
    return e->fn(d, ws, ws_size, stream);
#endif
}
