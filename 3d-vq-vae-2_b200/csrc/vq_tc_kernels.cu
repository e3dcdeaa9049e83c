// vq_tc_kernels.cu -- nearest-codeword search of the EMA quantizer (vqvae/layers.py:700-703) for the
// large problems of the quantizer sweep (>= 32k latent vectors, embedding_dim 32/64/128, any K):
// a tensor-core CANDIDATE pass followed by the exact fp32 re-rank, so that the indices stay
// bit-identical to the reference's cdist(direct form) + argmin(first minimum).
//
//   score[n,k] = ||e_k||^2 - 2 x_n.e_k   (= d2[n,k] - ||x_n||^2: same argmin)
// is ONE accumulator of an augmented GEMM on tcgen05: A row = [x_n | 1 1 1 0...], B row =
// [-2 e_k | c_hi c_lo c_lo2 0...] with c = ||e_k||^2 split into three bf16 pieces.  x and e are split
// into bf16 hi + lo parts and three MMAs (hi.hi, hi.lo, lo.hi; fp32 accumulation in TMEM) give the dot
// product to ~2^-15 relative, so the approximate and exact distances differ by less than
// margin = 2^-12 (||x||^2 + max_k ||e_k||^2).  Every code whose score is within `margin` of the row
// minimum is kept as a candidate (normally one or two), and only those are re-evaluated with the exact
// direct-difference fp32 arithmetic of ATen's cdist (4-wide two-rounding body + fma tail), sqrt, and the
// lowest-index tie rule.  The candidate list holds 8 entries; when it fills up (degenerate codebooks, massive
// ties) it is folded into a running exact best, so the worst case degrades to the exact scan, never to an
// approximate answer.
//
// Per CTA (one per SM, persistent): warp 0 streams codebook tiles (bf16 hi|lo images, prepared once per
// call) through a 2-stage shared-memory ring with bulk copies; warp 1 issues the MMAs; 2-4 groups of 4
// warps own 128 latent vectors each (thread = vector = TMEM lane): they stage their A operand (planar
// (B, D, S) global layout -> K-major bf16 hi/lo), sweep each accumulator tile twice out of TMEM (row
// minimum, then candidates of the 16-column groups that can contain one), and finish with the exact
// re-rank, codeword gather, straight-through value, squared error and (training) EMA statistics.
#include "vq3d_rt.h"

#ifndef VQ3D_EMU
#include <cuda_bf16.h>
#include <cstdlib>

namespace vq3d {

#include "tc_common.cuh"

constexpr int kVqtStages = 2;
constexpr int kVqtMaxCand = 8;

template <int D>
struct VqtCfg {
    static constexpr int DA = D + 16;                      // augmented reduction length
    static constexpr int KC = DA / 8;                      // 16-byte chunks per row
    static constexpr int KS = DA / 16;                     // K16 steps
    static constexpr int NT = D <= 64 ? 128 : 64;          // codes per tile (MMA N)
    static constexpr int NG = D <= 32 ? 4 : (D <= 64 ? 3 : 2);         // groups of 128 latent vectors per CTA (smem / TMEM budget)
    static constexpr int THREADS = (4 + 4 * NG) * 32;      // warp 0 producer, warp 1 MMA, warps 4.. the groups
    static constexpr uint32_t IMG = (uint32_t)NT * DA * 2; // one codebook tile image (hi or lo)
    static constexpr uint32_t STAGE = 2 * IMG;             // hi | lo
    static constexpr uint32_t AIMG = 128u * DA * 2;        // one group's A image (hi or lo)
    static constexpr uint32_t LBO_A = 128 * 16, LBO_B = (uint32_t)NT * 16;
    static constexpr size_t smem = 128 + (size_t)kVqtStages * STAGE + (size_t)NG * 2 * AIMG + (size_t)NG * 128 * kVqtMaxCand * 2;
};

struct VqtParams {
    const float *x, *embed;
    int64_t B, S;
    int K, Kpad;
    const unsigned char *wimg;       // per tile [hi image | lo image]
    const float *cmax;               // max_k ||e_k||^2
    float *quant;
    int64_t *idx;
    double *sqerr;
    float *counts, *dw;
    uint32_t tmem_cols;
};

// codebook -> bf16 hi/lo B-operand images of -2e with the ||e||^2 columns, padded to whole tiles
template <int D>
__global__ void __launch_bounds__(256)
vqt_prep_kernel(const float *__restrict__ embed, int K, int Kpad, unsigned char *wimg, float *cmax) {
    using Cfg = VqtCfg<D>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Kpad * Cfg::KC) return;
    const int k = i / Cfg::KC, kc = i % Cfg::KC;
    const int t = k / Cfg::NT, n = k % Cfg::NT;
    float v[8];
    if (kc * 8 < D) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = k < K ? -2.0f * __ldg(embed + (size_t)k * D + kc * 8 + e) : 0.0f;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.0f;
        if (kc * 8 == D) {
            float c = 1e30f;                              // padding rows can never be candidates
            if (k < K) {
                c = 0.0f;
                for (int d = 0; d < D; ++d) { const float ev = __ldg(embed + (size_t)k * D + d); c = __fmaf_rn(ev, ev, c); }
                atomicMax(reinterpret_cast<int *>(cmax), __float_as_int(c));      // c >= 0: int order = float order
            }
            const float c0 = __bfloat162float(__float2bfloat16_rn(c));
            const float c1 = __bfloat162float(__float2bfloat16_rn(c - c0));
            v[0] = c0; v[1] = c1; v[2] = (c - c0) - c1;
        }
    }
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        hi[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
        lo[e] = v[e] - hi[e];
    }
    uint4 ph, pl;
    ph.x = bf16x2(hi[0], hi[1]); ph.y = bf16x2(hi[2], hi[3]); ph.z = bf16x2(hi[4], hi[5]); ph.w = bf16x2(hi[6], hi[7]);
    pl.x = bf16x2(lo[0], lo[1]); pl.y = bf16x2(lo[2], lo[3]); pl.z = bf16x2(lo[4], lo[5]); pl.w = bf16x2(lo[6], lo[7]);
    unsigned char *base = wimg + (size_t)t * Cfg::STAGE + (size_t)kc * Cfg::LBO_B + (size_t)n * 16;
    *reinterpret_cast<uint4 *>(base) = ph;
    *reinterpret_cast<uint4 *>(base + Cfg::IMG) = pl;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// exact squared distance in the reference's summation order (see oracle/vq_oracle.c); x read with a stride
template <int D>
__device__ __forceinline__ float vqt_ref_dist2(const float *__restrict__ xs, int64_t xstride, const float *__restrict__ e) {
    constexpr int NV = (D / 4) * 4;
    float agg = 0.0f;
#pragma unroll 8
    for (int d = 0; d < NV; ++d) {
        const float diff = __fsub_rn(__ldg(xs + (size_t)d * xstride), __ldg(e + d));
        agg = __fadd_rn(agg, __fmul_rn(diff, diff));
    }
#pragma unroll
    for (int d = NV; d < D; ++d) {
        const float diff = __fsub_rn(__ldg(xs + (size_t)d * xstride), __ldg(e + d));
        agg = __fmaf_rn(diff, diff, agg);
    }
    return agg;
}

template <int D>
__global__ void __launch_bounds__(VqtCfg<D>::THREADS, 1)
vq_tc_kernel(const __grid_constant__ VqtParams p) {
    using Cfg = VqtCfg<D>;
    constexpr int NT = Cfg::NT, KS = Cfg::KS, NG = Cfg::NG, NJ = NT / 16;
    VQ3D_DYN_SMEM(unsigned char, smem_raw);
    __shared__ __align__(8) uint64_t e_full[kVqtStages], e_empty[kVqtStages], a_full[NG], d_full[NG], d_empty[NG];
    __shared__ uint32_t tmem_slot;
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (s_u32(smem_raw) + 127u) & ~127u;
    unsigned char *smem = smem_raw + (base - s_u32(smem_raw));
    const uint32_t ring_addr = base, a_addr = base + kVqtStages * Cfg::STAGE;
    unsigned char *sA = smem + (size_t)kVqtStages * Cfg::STAGE;
    unsigned short *s_cand = reinterpret_cast<unsigned short *>(sA + (size_t)NG * 2 * Cfg::AIMG);    // [kVqtMaxCand][NG*128]

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < kVqtStages; ++s) { mbarrier_init(&e_full[s], 1); mbarrier_init(&e_empty[s], 1); }
        for (int g = 0; g < NG; ++g) { mbarrier_init(&a_full[g], 4); mbarrier_init(&d_full[g], 1); mbarrier_init(&d_empty[g], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    const int64_t N = p.B * p.S;
    const int64_t nsuper = (N + NG * 128 - 1) / (NG * 128);
    const int my_super = nsuper > (int64_t)blockIdx.x ? (int)((nsuper - 1 - blockIdx.x) / gridDim.x + 1) : 0;
    const int ntiles = p.Kpad / NT;
    double err_acc = 0.0;

    if (warp == 0) {
        // ===== producer: codebook tiles through the ring (the same tile sequence for every super-tile) =====
        if (elect_one()) {
            uint32_t cnt = 0;
            for (int i = 0; i < my_super; ++i)
                for (int t = 0; t < ntiles; ++t, ++cnt) {
                    const uint32_t slot = cnt % kVqtStages, u = cnt / kVqtStages;
                    mbarrier_wait(&e_empty[slot], (u & 1u) ^ 1u);
                    mbarrier_arrive_expect_tx(&e_full[slot], Cfg::STAGE);
                    const unsigned char *src = p.wimg + (size_t)t * Cfg::STAGE;
                    for (uint32_t o = 0; o < Cfg::STAGE; o += 16384u) {
                        const uint32_t nb = Cfg::STAGE - o < 16384u ? Cfg::STAGE - o : 16384u;
                        bulk_g2s(ring_addr + slot * Cfg::STAGE + o, src + o, nb, &e_full[slot]);
                    }
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        uint32_t cnt = 0, du = 0;
        for (int i = 0; i < my_super; ++i) {
            for (int t = 0; t < ntiles; ++t, ++cnt, ++du) {
                const uint32_t slot = cnt % kVqtStages, u = cnt / kVqtStages;
                mbarrier_wait(&e_full[slot], u & 1u);
                const uint32_t e_hi = ring_addr + slot * Cfg::STAGE, e_lo = e_hi + Cfg::IMG;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (t == 0) mbarrier_wait(&a_full[g], (uint32_t)i & 1u);
                    mbarrier_wait(&d_empty[g], (du & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t ahi = a_addr + (uint32_t)g * 2 * Cfg::AIMG, alo = ahi + Cfg::AIMG;
                        const uint32_t d_addr = tmem_d + (uint32_t)(g * NT);
#pragma unroll
                        for (int sp = 0; sp < 3; ++sp) {
                            const uint32_t aa = sp == 2 ? alo : ahi, bb = sp == 1 ? e_lo : e_hi;
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks)
                                umma_f16(d_addr, umma_desc(aa + (uint32_t)(2 * ks) * Cfg::LBO_A, Cfg::LBO_A, 128),
                                         umma_desc(bb + (uint32_t)(2 * ks) * Cfg::LBO_B, Cfg::LBO_B, 128), idesc, (sp > 0 || ks > 0) ? 1u : 0u);
                        }
                        umma_commit_to(&d_full[g]);
                        if (g == NG - 1) umma_commit_to(&e_empty[slot]);      // every group's MMAs on this tile have been issued
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ===== vector groups =====
        const int g = (warp - 4) >> 2, q = warp & 3, row = q * 32 + lane;
        unsigned char *a_hi = sA + (size_t)g * 2 * Cfg::AIMG + (size_t)row * 16, *a_lo = a_hi + Cfg::AIMG;
        unsigned short *my_cand = s_cand + g * 128 + row;                // stride NG*128 between entries
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const float cmax = __ldg(p.cmax);
        const bool want_stats = p.counts != nullptr;
        uint32_t du = 0;
        for (int i = 0; i < my_super; ++i) {
            const int64_t sup = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
            const int64_t v = (sup * NG + g) * 128 + row;
            const bool active = v < N;
            const int64_t b = active ? v / p.S : 0, s = active ? v - b * p.S : 0;
            const float *xs = p.x + (size_t)b * D * p.S + s;
            // ---- stage the A operand: [x | 1 1 1 0 ...] as bf16 hi / lo, K-major ----
            float xx = 0.0f;
#pragma unroll
            for (int kc = 0; kc < D / 8; ++kc) {
                float hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xv = active ? __ldg(xs + (size_t)(kc * 8 + e) * p.S) : 0.0f;
                    xx = __fmaf_rn(xv, xv, xx);
                    hi[e] = __bfloat162float(__float2bfloat16_rn(xv));
                    lo[e] = xv - hi[e];
                }
                uint4 ph, pl;
                ph.x = bf16x2(hi[0], hi[1]); ph.y = bf16x2(hi[2], hi[3]); ph.z = bf16x2(hi[4], hi[5]); ph.w = bf16x2(hi[6], hi[7]);
                pl.x = bf16x2(lo[0], lo[1]); pl.y = bf16x2(lo[2], lo[3]); pl.z = bf16x2(lo[4], lo[5]); pl.w = bf16x2(lo[6], lo[7]);
                *reinterpret_cast<uint4 *>(a_hi + (size_t)kc * Cfg::LBO_A) = ph;
                *reinterpret_cast<uint4 *>(a_lo + (size_t)kc * Cfg::LBO_A) = pl;
            }
            {
                uint4 one, zero;
                one.x = bf16x2(1.0f, 1.0f); one.y = bf16x2(1.0f, 0.0f); one.z = 0u; one.w = 0u;
                zero.x = zero.y = zero.z = zero.w = 0u;
                *reinterpret_cast<uint4 *>(a_hi + (size_t)(D / 8) * Cfg::LBO_A) = one;
                *reinterpret_cast<uint4 *>(a_hi + (size_t)(D / 8 + 1) * Cfg::LBO_A) = zero;
                *reinterpret_cast<uint4 *>(a_lo + (size_t)(D / 8) * Cfg::LBO_A) = zero;
                *reinterpret_cast<uint4 *>(a_lo + (size_t)(D / 8 + 1) * Cfg::LBO_A) = zero;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbarrier_arrive(&a_full[g]);
            const float margin = 2.44140625e-4f * (xx + cmax) + 1e-30f;        // 2^-12
            float m_run = __int_as_float(0x7f800000);
            int nc = 0;
            float best_r = __int_as_float(0x7f800000);       // exact (sqrt distance, index) of the best candidate folded in so far
            int best_k = 0x7fffffff;
            // exact re-rank of the collected candidates; lexicographic (sqrt(d2), k) minimum = the reference's argmin
            auto flush = [&]() {
                for (int c = 0; c < nc; ++c) {
                    const int k = my_cand[c * (NG * 128)];
                    if (k < p.K) {
                        const float r = __fsqrt_rn(vqt_ref_dist2<D>(xs, p.S, p.embed + (size_t)k * D));
                        if (r < best_r || (r == best_r && k < best_k)) { best_r = r; best_k = k; }
                    }
                }
                nc = 0;
            };
            // ---- sweep the accumulator tiles ----
            for (int t = 0; t < ntiles; ++t, ++du) {
                mbarrier_wait(&d_full[g], du & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_addr = tmem_d + lane_sel + (uint32_t)(g * NT);
                float lmin[NJ];
                const float m_old = m_run;
#pragma unroll
                for (int j2 = 0; j2 < NJ / 2; ++j2) {
                    float sc[32];
                    tmem_ld32(d_addr + (uint32_t)(j2 * 32), sc);
                    float m0 = sc[0], m1 = sc[16];
#pragma unroll
                    for (int e = 1; e < 16; ++e) { m0 = fminf(m0, sc[e]); m1 = fminf(m1, sc[16 + e]); }
                    lmin[2 * j2] = m0; lmin[2 * j2 + 1] = m1;
                    m_run = fminf(m_run, fminf(m0, m1));
                }
                const float thr = m_run + margin;
                if (thr + margin < m_old) nc = 0;            // every earlier candidate is now further than margin from the minimum
                // 16-column groups that can hold a candidate of this row; visit the union over the warp (tcgen05.ld is
                // warp-wide) in ONE rolled loop so that the push / flush code exists once (instruction cache)
                uint32_t gmask = 0;
#pragma unroll
                for (int j = 0; j < NJ; ++j) gmask |= (lmin[j] <= thr ? 1u : 0u) << j;
                uint32_t um = __reduce_or_sync(0xffffffffu, gmask);
#pragma unroll 1
                while (um) {
                    const int j = __ffs(um) - 1;
                    um &= um - 1;
                    float sc[16];
                    tmem_ld16(d_addr + (uint32_t)(j * 16), sc);
                    uint32_t hits = 0;
#pragma unroll
                    for (int e = 0; e < 16; ++e) hits |= (sc[e] <= thr ? 1u : 0u) << e;
                    if (!((gmask >> j) & 1u)) hits = 0;
#pragma unroll 1
                    while (hits) {
                        const int e = __ffs(hits) - 1;
                        hits &= hits - 1;
                        if (nc == kVqtMaxCand) { if (active) flush(); else nc = 0; }
                        my_cand[nc * (NG * 128)] = (unsigned short)(t * NT + j * 16 + e);
                        ++nc;
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbarrier_arrive(&d_empty[g]);
            }
            // ---- exact re-rank, gather, straight-through value, loss partial, statistics ----
            if (active) {
                flush();
                if (best_k == 0x7fffffff) {          // nothing finite collected (NaN input): exact full scan, first minimum
                    for (int k = 0; k < p.K; ++k) {
                        const float r = __fsqrt_rn(vqt_ref_dist2<D>(xs, p.S, p.embed + (size_t)k * D));
                        if (r < best_r || best_k == 0x7fffffff) { best_r = r; best_k = k; }
                    }
                }
                const float *e = p.embed + (size_t)best_k * D;
                float err = 0.0f;
#pragma unroll 8
                for (int d = 0; d < D; ++d) {
                    const float qv = __ldg(e + d);
                    const float xv = __ldg(xs + (size_t)d * p.S);
                    const float df = qv - xv;
                    err = __fmaf_rn(df, df, err);
                    p.quant[((size_t)b * D + d) * p.S + s] = __fadd_rn(xv, __fsub_rn(qv, xv));    // layers.py:720, two roundings
                    if (want_stats) atomicAdd(&p.dw[(size_t)best_k * D + d], xv);
                }
                if (want_stats) atomicAdd(&p.counts[best_k], 1.0f);
                p.idx[(size_t)b * p.S + s] = best_k;
                err_acc += (double)err;
            }
        }
    }
    // ---- squared-error total of the CTA ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err_acc += __shfl_xor_sync(0xffffffffu, err_acc, o);
    if (lane == 0) red[warp] = err_acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0 && p.sqerr != nullptr) {
        double tot = 0.0;
        for (int w = 0; w < Cfg::THREADS / 32; ++w) tot += red[w];
        atomicAdd(p.sqerr, tot);
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
    }
}

template <int D>
static size_t vqt_ws_bytes(int K) {
    using Cfg = VqtCfg<D>;
    const int Kpad = (K + Cfg::NT - 1) / Cfg::NT * Cfg::NT;
    return 256 + (size_t)(Kpad / Cfg::NT) * Cfg::STAGE;
}

template <int D>
static int launch_vqt(const float *x, const float *embed, int64_t B, int64_t S, int K, float *quant, int64_t *idx, double *sqerr,
                      float *counts, float *dw, void *ws, size_t ws_size, void *stream) {
    using Cfg = VqtCfg<D>;
    if (ws_size < vqt_ws_bytes<D>(K) || (reinterpret_cast<uintptr_t>(ws) & 255) != 0)
        return fail(VQ3D_ERR_INVALID, "vq_assign_tc: workspace too small or not 256-byte aligned");
    if (K > 65535) return fail(VQ3D_ERR_UNSUPPORTED, "vq_assign_tc: K > 65535");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VqtParams p;
    p.x = x; p.embed = embed; p.B = B; p.S = S; p.K = K;
    p.Kpad = (K + Cfg::NT - 1) / Cfg::NT * Cfg::NT;
    unsigned char *wsb = static_cast<unsigned char *>(ws);
    p.cmax = reinterpret_cast<float *>(wsb);
    p.wimg = wsb + 256;
    p.quant = quant; p.idx = idx; p.sqerr = sqerr; p.counts = counts; p.dw = dw;
    uint32_t cols = 32;
    while (cols < (uint32_t)(Cfg::NG * Cfg::NT)) cols <<= 1;
    p.tmem_cols = cols;
    cudaError_t e = cudaMemsetAsync(wsb, 0, 256, st);
    if (e != cudaSuccess) return check_cuda(e, "vq_assign_tc(memset)");
    vqt_prep_kernel<D><<<(unsigned)ceil_div((int64_t)p.Kpad * Cfg::KC, 256), 256, 0, st>>>(embed, K, p.Kpad, wsb + 256, reinterpret_cast<float *>(wsb));
    e = cudaGetLastError();
    if (e != cudaSuccess) return check_cuda(e, "vq_assign_tc(prep)");
    auto kernel = vq_tc_kernel<D>;
    size_t smem = Cfg::smem < 120 * 1024 ? 120 * 1024 : Cfg::smem;       // one CTA per SM (TMEM)
    e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "vq_assign_tc(attr)");
    const int64_t nsuper = ceil_div(B * S, Cfg::NG * 128);
    int grid = kNumSMs;
    if (grid > nsuper) grid = (int)nsuper;
    kernel<<<dim3((unsigned)grid), dim3(Cfg::THREADS), smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "vq_assign_tc");
}

}  // namespace vq3d
#endif  // !VQ3D_EMU

using namespace vq3d;

extern "C" size_t vq3d_vq_assign_tc_workspace(int D, int K) {
#ifdef VQ3D_EMU
    (void)D; (void)K;
    return 0;
#else
    if (K < 1 || K > 65535) return 0;
    switch (D) {
        case 32: return vqt_ws_bytes<32>(K);
        case 64: return vqt_ws_bytes<64>(K);
        case 128: return vqt_ws_bytes<128>(K);
        default: return 0;
    }
#endif
}

extern "C" int vq3d_vq_assign_tc(const float *x, const float *embed, int64_t B, int D, int64_t S, int K, float *quant, int64_t *idx,
                                 double *sqerr, float *counts, float *dw, void *ws, size_t ws_bytes, void *stream) {
#ifdef VQ3D_EMU
    (void)x; (void)embed; (void)B; (void)D; (void)S; (void)K; (void)quant; (void)idx; (void)sqerr; (void)counts; (void)dw; (void)ws; (void)ws_bytes; (void)stream;
    return fail(VQ3D_ERR_UNSUPPORTED, "vq_assign_tc: tensor-core kernels cannot run in the host emulator");
#else
    if (!x || !embed || !quant || !idx || !ws) return fail(VQ3D_ERR_INVALID, "vq_assign_tc: null pointer");
    if (B < 1 || S < 1 || K < 1) return fail(VQ3D_ERR_INVALID, "vq_assign_tc: bad sizes");
    if ((counts == nullptr) != (dw == nullptr)) return fail(VQ3D_ERR_INVALID, "vq_assign_tc: counts and dw must both be given or both NULL");
    switch (D) {
        case 32: return launch_vqt<32>(x, embed, B, S, K, quant, idx, sqerr, counts, dw, ws, ws_bytes, stream);
        case 64: return launch_vqt<64>(x, embed, B, S, K, quant, idx, sqerr, counts, dw, ws, ws_bytes, stream);
        case 128: return launch_vqt<128>(x, embed, B, S, K, quant, idx, sqerr, counts, dw, ws, ws_bytes, stream);
        default: return fail(VQ3D_ERR_UNSUPPORTED, "vq_assign_tc: embedding_dim %d (32, 64 and 128 are instantiated)", D);
    }
#endif
}
