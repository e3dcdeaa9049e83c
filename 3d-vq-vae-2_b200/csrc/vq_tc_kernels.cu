// vq_tc_kernels.cu -- nearest-codeword search of the EMA quantizer (vqvae/layers.py:700-703) for the
// large problems of the quantizer sweep (>= 32k latent vectors, embedding_dim 32/64/128, any K):
// a tensor-core CANDIDATE pass followed by the exact fp32 re-rank, so that the indices stay
// bit-identical to the reference's cdist(direct form) + argmin(first minimum).
//
//   score[n,k] = ||e_k||^2 - 2 x_n.e_k   (= d2[n,k] - ||x_n||^2: same argmin)
// is ONE accumulator of an augmented GEMM on tcgen05: A row = [x_n | 1 1 1 0...], B row =
// [-2 e_k | c_hi c_lo c_lo2 0...] with c = ||e_k||^2 split into three bf16 pieces (fp32 accumulation in TMEM).
// For embedding_dim 64 and 128 x and -2e are single bf16 values: the score is within H = 2^-6 ||x|| ||e_k|| (+ the
// accumulation term) of the reference's squared distance.  For embedding_dim 32 both are bf16 hi + lo pairs and the score
// is hi.hi + hi.lo + lo.hi (7 MMAs instead of 3): H = 6 * 2^-16 ||x|| ||e_k||.  Every code whose score can still be the
// reference's minimum given H is kept as a candidate (with hi/lo operands: one, the argmin itself, for 99.75 % of the
// vectors of the randn sweep; with single bf16 two or more for ~25 %), and only rows with several candidates are
// re-evaluated with the exact direct-difference fp32 arithmetic of ATen's cdist (4-wide two-rounding body + fma tail),
// sqrt, and the lowest-index tie rule.  The candidate lists are bounded; when one fills up (degenerate codebooks, massive
// ties) the row falls back to the exact scan of the whole codebook, never to an approximate answer.
//
// Per CTA (one per SM, persistent): warp 0 streams codebook tiles (bf16 images, prepared once per call) through a
// shared-memory ring with bulk copies (resident when the codebook fits); warp 1 issues the MMAs; sweep warps read the
// score tiles out of TMEM (two threads per latent vector, half of the columns each); load/epilogue warps stage the A
// operand (planar (B, D, S) global layout -> K-major bf16) and finish with the merge of the two half rows, the exact
// re-rank, codeword gather, straight-through value, squared error and (training) EMA statistics.
#include "vq3d_rt.h"

#ifndef VQ3D_EMU
#include <cuda_bf16.h>
#include <cstdlib>

namespace vq3d {

#include "tc_common.cuh"

// VQ3D_VQT_DEBUG builds accumulate, per role, the clocks spent inside the named waits (debug counters 8..15 of the
// workspace header: who waits for whom) -- tools/debug_vq.py prints them
#ifdef VQ3D_VQT_DEBUG
#define VQT_TIMED_WAIT(slot, call) do { const long long t0_ = clock64(); call; if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long *>(const_cast<unsigned *>(p.dbg)) + (slot), (unsigned long long)(clock64() - t0_)); } while (0)
#else
#define VQT_TIMED_WAIT(slot, call) call
#endif

constexpr int kVqtMaxCand = 8;       // per half-row list (two column halves per row -> up to 16 candidates per latent vector); 7 for embedding_dim 64 (shared memory)
constexpr int kVqtGsCand = 4;        // the same list of the group-store sweep, which only takes the elements of a displaced live group

// Warp roles of one persistent CTA (one per SM; a "super-tile" is NG groups x 128 latent vectors):
//   warp 0        codebook producer: bf16 B-operand tiles -> shared-memory ring by bulk copies (the whole codebook stays
//                 resident when it fits the ring, so it is fetched from L2 once per CTA, not once per super-tile)
//   warp 1        MMA issuer (one elected lane)
//   warps 4..19   sweep warps, 8 per group: thread = (latent vector = TMEM lane, column half of the score tile).  They read
//                 the score tiles out of TMEM and keep the running row minimum and the list of columns within `margin`
//                 of it.  Two warps per TMEM lane quadrant and group = four sweep warps per SM sub-partition (tcgen05.ld
//                 delivers 370-460 B/clk per SM with 16 warps, tools/microbench: TMEM reads are not this kernel's floor).
//                 embedding_dim 32 uses the group-store sweep (VqtCfg::GS) instead of the candidate scan.
//   warps 20..27  load/epilogue warps, 4 per group: stage the A operand of super-tile i+2 while the sweep warps work on i
//                 and i+1, then merge super-tile i's two half-row lists, resolve the candidates (exact fp32 re-rank only
//                 when there is more than one), gather the codeword, write the straight-through value, index, loss
//                 partial, statistics
// so global-memory latency (x loads, codeword gather, stores) never sits between two MMAs of the same accumulator.
template <int D>
struct VqtCfg {
    // embedding_dim 32 runs the hi/lo-split, group-store variant of the kernel (the shared memory allows it there):
    //  HL  the operands are bf16 hi + lo pairs, rows [xh | xl | 1 1 1 0..] and [(-2e)h | (-2e)l | c0 c1 c2 0..], and the score is
    //      xh.eh + xh.el + xl.eh + c (7 K16 MMAs instead of 3): the error bound of a score drops from 2^-6 to 6 * 2^-16 of
    //      ||x|| ||e||, so that candidates beyond the argmin itself (and with them the exact re-rank) become rare;
    //  GS  group-store sweep, see below.
    static constexpr bool HL = D == 32;
    static constexpr int DA = (HL ? 2 * D : D) + 16;       // augmented reduction length
    static constexpr int KC = DA / 8;                      // 16-byte chunks per row
    static constexpr int KS = HL ? 3 * (D / 16) + 1 : DA / 16;          // K16 MMA steps per score tile
    static constexpr int NT = 128;                         // codes per tile (MMA N)
    static constexpr int NG = 2;                           // groups of 128 latent vectors per super-tile (2 x 2 x 128 TMEM columns)
    static constexpr int NABUF = D <= 64 ? 2 : 1;          // A-operand buffers per group
    // group-store sweep (embedding_dim 32, where the shared memory allows it): the sweep threads do not scan a 16-column
    // group for candidates; a group whose minimum is within margin of the running bound is stored whole (its 16 raw
    // scores, predicated stores, no divergent branch) as the thread's "live group", and the load/epilogue warps pick the
    // candidates out of it once the final bound is known.  Only when a still-live group is displaced (two groups within
    // margin of each other: a few per cent of the rows) its elements go through the old candidate list.
    static constexpr bool GS = D == 32;
    static constexpr int LC = GS ? kVqtGsCand : (D == 64 ? kVqtMaxCand - 1 : kVqtMaxCand);      // list entries per (row, half)
    static constexpr int NSTAGE = D <= 64 ? 4 : 2;                    // codebook ring stages (shared-memory budget)
    static constexpr int NRB = GS ? 1 : 2;                              // result buffers per group (sweep of i+1 while i is merged)
    static constexpr int SWEEP_WARP0 = 4, LE_WARP0 = 4 + 8 * NG;
    static constexpr int THREADS = (LE_WARP0 + 4 * NG) * 32;
    static constexpr uint32_t STAGE = (uint32_t)NT * DA * 2;           // one codebook tile image
    static constexpr uint32_t AIMG = 128u * DA * 2;                    // one group's A image
    static constexpr uint32_t LBO_A = 128 * 16, LBO_B = (uint32_t)NT * 16;
    // per (group, result buffer): ||x||^2 [128]; per (group, result buffer, half): row minimum [128], count [128],
    // columns [LC][128] (u16), score lower bound [LC][128]
    // per (group, A buffer parity): the row's two margin inputs [128] (float2); per (group, result buffer, half): row minimum ...
    static constexpr size_t RES = (size_t)NG * 2 * 128 * 8 + (size_t)NG * NRB * 2 * 128 * (4 + 4 + 6 * LC);
    // group-store sweep, per (group, result buffer, half): live group minimum [128], its first column [128], its 16 raw
    // scores [4 quads][128] x 16 bytes (a quad per 128-bit store: conflict-free)
    static constexpr size_t LIVE = GS ? (size_t)NG * NRB * 2 * 128 * (4 + 4 + 64) : 0;
    static constexpr size_t QUEUE = (size_t)4 * NG * 32 * 8;          // per load/epilogue warp: 32 (code | lane, distance) pairs
    static constexpr size_t BEST = (size_t)NG * 2 * 128 * 4;           // resolved code per row, handed to the sweep warps for the gather
    static constexpr size_t TNORM = (HL ? 1 : 3) * 128 * 4;           // per-tile bounds (K <= 16 384: 128 tiles): ||e||; single bf16 also ||wh||, ||wl||
    static constexpr size_t smem = 128 + (size_t)NSTAGE * STAGE + (size_t)NG * NABUF * AIMG + RES + QUEUE + TNORM + BEST + LIVE;
    static constexpr size_t SMEM_LIMIT = 227 * 1024 - 1024;           // minus the static part (barriers, reduction scratch)
};

struct VqtParams {
    const float *x, *embed;
    int64_t B, S;
    int K, Kpad;
    const unsigned char *wimg;       // per tile: bf16 image of [-2e | c0 c1 c2 0...], codes sorted by ascending norm
    const int *perm;                 // column -> code index (the norm-sorted order)
    const float *tnorm;              // per tile: largest ||e_k|| in it
    const float *tbound;             // per tile (single-bf16 operands): largest ||bf16(-2e_k)|| and largest ||-2e_k - bf16(-2e_k)||
    const unsigned *dbg;             // debug counters (VQ3D_VQT_DEBUG builds)
    int perm_in_smem;                // the column -> code table fits behind the other shared-memory regions (u16 entries)
    float *quant;
    int64_t *idx;
    double *sqerr;
    float *counts, *dw;
    uint32_t tmem_cols;
};

// ||e_k||^2 in fp32 (one fma chain per code)
__global__ void __launch_bounds__(256)
vqt_norm_kernel(const float *__restrict__ embed, int K, int D, float *__restrict__ cnorm) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float c = 0.0f;
    for (int d = 0; d < D; ++d) { const float ev = __ldg(embed + (size_t)k * D + d); c = __fmaf_rn(ev, ev, c); }
    cnorm[k] = c;
}

// columns are the codes in ascending-norm order (ties: ascending index), so that the codes of one tile have similar norms
// and the per-tile error bound of the tensor-core scores is tight whatever outliers the codebook holds
__global__ void __launch_bounds__(256)
vqt_rank_kernel(const float *__restrict__ cnorm, int K, int Kpad, int *__restrict__ perm) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kpad) return;
    if (k >= K) { perm[k] = k; return; }                 // padding columns stay at the end
    const float c = cnorm[k];
    const float ck = c == c ? c : __int_as_float(0x7f800000);
    int rank = 0;
    for (int j = 0; j < K; ++j) {
        const float cj0 = __ldg(cnorm + j);
        const float cj = cj0 == cj0 ? cj0 : __int_as_float(0x7f800000);
        rank += (cj < ck || (cj == ck && j < k)) ? 1 : 0;
    }
    perm[rank] = k;
}

// codebook -> bf16 B-operand images of -2e with the ||e||^2 columns (three bf16 pieces: exact to 2^-24), padded to whole tiles
template <int D>
__global__ void __launch_bounds__(256)
vqt_prep_kernel(const float *__restrict__ embed, const float *__restrict__ cnorm, const int *__restrict__ perm, int K, int Kpad,
                unsigned char *wimg, float *tnorm) {
    using Cfg = VqtCfg<D>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Kpad * Cfg::KC) return;
    const int col = i / Cfg::KC, kc = i % Cfg::KC;
    const int t = col / Cfg::NT, n = col % Cfg::NT;
    const int k = col < K ? perm[col] : -1;
    constexpr int DX = Cfg::HL ? 2 * D : D;          // operand columns before the ||e||^2 pieces
    float v[8];
    if (kc * 8 < DX) {
        // hi/lo split: chunks [0, D/8) hold bf16(-2e), chunks [D/8, 2D/8) hold bf16(-2e - hi)
        const int d0 = (kc * 8) % D;
        const bool lo = kc * 8 >= D;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float w = k >= 0 ? -2.0f * __ldg(embed + (size_t)k * D + d0 + e) : 0.0f;
            v[e] = lo ? w - __bfloat162float(__float2bfloat16_rn(w)) : w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.0f;
        if (kc * 8 == DX) {
            const float c = k >= 0 ? cnorm[k] : 1e30f;     // padding columns can never be candidates
            const float c0 = __bfloat162float(__float2bfloat16_rn(c));
            const float c1 = __bfloat162float(__float2bfloat16_rn(c - c0));
            v[0] = c0; v[1] = c1; v[2] = (c - c0) - c1;
            // the last real column of a tile carries the tile's largest norm
            if (k >= 0 && (n == Cfg::NT - 1 || col == K - 1)) tnorm[t] = sqrtf(c);
        }
    }
    uint4 ph;
    ph.x = bf16x2(v[0], v[1]); ph.y = bf16x2(v[2], v[3]); ph.z = bf16x2(v[4], v[5]); ph.w = bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4 *>(wimg + (size_t)t * Cfg::STAGE + (size_t)kc * Cfg::LBO_B + (size_t)n * 16) = ph;
}

// per tile: largest ||bf16(-2e_k)|| and largest ||-2e_k - bf16(-2e_k)|| over its codes (inflated by 1e-4 against the
// roundings of the norm itself); tb is zeroed before the launch, positive floats order like their bit patterns
template <int D>
__global__ void __launch_bounds__(256)
vqt_tile_bounds_kernel(const float *__restrict__ embed, const int *__restrict__ perm, int K, int NT, float *tb) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= K) return;
    const int k = perm[col];
    float h2 = 0.0f, l2 = 0.0f;
    for (int d = 0; d < D; ++d) {
        const float w = -2.0f * __ldg(embed + (size_t)k * D + d);
        const float wh = __bfloat162float(__float2bfloat16_rn(w)), wl = w - wh;
        h2 = __fmaf_rn(wh, wh, h2);
        l2 = __fmaf_rn(wl, wl, l2);
    }
    atomicMax(reinterpret_cast<int *>(tb) + (col / NT) * 2, __float_as_int(sqrtf(h2) * 1.0001f));
    atomicMax(reinterpret_cast<int *>(tb) + (col / NT) * 2 + 1, __float_as_int(sqrtf(l2) * 1.0001f));
}

// tcgen05.ld is asynchronous: the destination registers are valid after tcgen05.wait::ld.  The wait names them as
// read-write operands so that the compiler cannot schedule a use above it.
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// |tensor-core score - exact score| for a code k of tile t, H_t.  With it, true_k >= score_k - H_t(k) and
// true_j <= score_j + H_t(j): the reference's argmin k satisfies score_k <= U + H_t(k), U = min_j (score_j + H_t(j)).
//  * single-bf16 operands (embedding_dim 64, 128): x = xh + xl, w = -2e = wh + wl exactly (xh, wh the bf16 roundings), and
//    the MMA computes xh.wh, so the dot product is off by xl.wh + xh.wl + xl.wl, at most
//        ||xl|| max_t ||wh|| + ||xh|| max_t ||wl|| + ||xl|| max_t ||wl||                      (Cauchy-Schwarz)
//    with the ACTUAL residual norms of this latent vector (r0 = ||xh||, r1 = ||xl||, computed when the A operand is staged)
//    and of the tile's codewords (vqt_tile_bounds_kernel).  For random data that is ~0.8 * 2^-7 ||x|| ||e||, half of the
//    worst case 2^-6 ||x|| ||e|| that a bound from the unit roundoff alone has to assume -- and still a worst-case bound;
//  * hi/lo-split operands (embedding_dim 32, r0 = ||x||, r1 = ||x||^2): with u = 2^-8 the unit roundoff of bf16,
//    |x - xh| <= u |x| and |x - xh - xl| <= u^2 |x| (the same for w), so the three dropped terms xl.wl, (x - xh - xl).w and
//    x.(w - wh - wl) change x.w by at most 3 u^2 sum|x_i||w_i| <= 6 * 2^-16 ||x|| ||e_k|| = 0.92e-4 ||x|| ||e_k||;
//  * the ||e||^2 columns are exact to 2^-24, and the fp32 accumulation (three or seven steps) / the reference's own roundings
//    stay below 2^-17 (||x||^2 + ||e_k||^2).
template <bool HL>
__device__ __forceinline__ float vqt_half_margin(float r0, float r1, float tn, float wh, float wl) {
    if (HL) return 1.0e-4f * r0 * tn + 1.0e-5f * (r1 + tn * tn) + 1e-30f;
    const float xn = r0 + r1;                    // >= ||x||
    return r1 * wh + r0 * wl + r1 * wl + 0.8e-5f * (xn * xn + tn * tn) + 1e-30f;
}

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// exact squared distance in the reference's summation order (see oracle/vq_oracle.c); x read with a stride.  Loads are
// issued 16 dimensions (x) + 4 float4 (codeword) at a time so that one evaluation costs D/16 memory round trips.
template <int D>
__device__ __forceinline__ float vqt_ref_dist2(const float *__restrict__ xs, int64_t xstride, const float *__restrict__ e) {
    static_assert(D % 32 == 0, "embedding_dim");
    float agg = 0.0f;
#pragma unroll 1
    for (int ch = 0; ch < D / 32; ++ch) {
        // 32 dimensions per memory round trip (an L2 round trip costs ~1 us under this kernel's load: they, not the
        // arithmetic, set the pace of the load/epilogue warps)
        float xv[32];
        float4 ev[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ev[j] = __ldg(reinterpret_cast<const float4 *>(e) + ch * 8 + j);
#pragma unroll
        for (int j = 0; j < 32; ++j) xv[j] = __ldg(xs + (size_t)(ch * 32 + j) * xstride);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float q4[4] = {ev[j].x, ev[j].y, ev[j].z, ev[j].w};
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const float diff = __fsub_rn(xv[j * 4 + l], q4[l]);
                agg = __fadd_rn(agg, __fmul_rn(diff, diff));
            }
        }
    }
    return agg;
}

// per-thread state of a sweep: running minimum over this thread's columns and the columns that were within `margin`
// of it when they were seen.  Instead of its score an entry keeps the minimum of its 16-column group (<= its score, and
// equal to it for the group's smallest element, i.e. nearly always): an entry whose `seen` value is above (final
// minimum + margin) cannot be a candidate; the others are kept (a superset).
struct VqtSweep {
    float m_run, m_push, u_run, hm;      // running minimum, its value at the last push, running U, this tile's half margin
    int nc;
    bool ovf;
    unsigned short *cand;      // [VqtCfg::LC] entries, stride 128
    float *seen;               // [VqtCfg::LC] entries, stride 128
};

// drop the entries that cannot be within margin of the final minimum any more; returns the new count
__device__ __noinline__ int vqt_compact(float thr, int nc, unsigned short *cand, float *seen) {
    int w = 0;
    for (int c = 0; c < nc; ++c) {
        const float sc = seen[c * 128];
        if (sc <= thr) {
            if (w != c) { seen[w * 128] = sc; cand[w * 128] = cand[c * 128]; }
            ++w;
        }
    }
    return w;
}

// one 16-column group of scores (registers): row minimum, then the candidates -- only lanes whose group minimum is
// within margin of their running minimum enter the element scan
template <int LCAP>
__device__ __forceinline__ void vqt_group(VqtSweep &s, const uint32_t (&r)[16], int colbase) {
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
    const float g = fmin3(fmin3(fmin3(v[0], v[1], v[2]), fmin3(v[3], v[4], v[5]), fmin3(v[6], v[7], v[8])),
                          fmin3(fmin3(v[9], v[10], v[11]), fmin3(v[12], v[13], v[14]), v[15]), __int_as_float(0x7f800000));
    s.m_run = fminf(s.m_run, g);
    s.u_run = fminf(s.u_run, g + s.hm);
    const float thr = s.u_run + s.hm;
    if (g <= thr) {
        if (thr < s.m_push) s.nc = 0;       // every listed score is >= the minimum at its push > U + this tile's half margin >= U + its own: stale
        // branch-free miss mask: thr - v is negative exactly when v > thr; its sign bit is shifted into one of four
        // independent chains (one FMA-pipe and one ALU-pipe instruction per column; both pipes issue every other cycle)
        uint32_t c4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int e = 0; e < 16; ++e)
            c4[e >> 2] = __funnelshift_l(__float_as_uint(__fsub_rn(thr, v[e])), c4[e >> 2], 1);
        uint32_t hits = ~((c4[0] << 12) | (c4[1] << 8) | (c4[2] << 4) | c4[3]) & 0xffffu;      // column e is bit 15 - e
#pragma unroll 1
        while (hits) {
            const int e = __clz(hits) - 16;
            hits &= ~(0x8000u >> e);
            if (s.nc == LCAP) s.nc = vqt_compact(thr, s.nc, s.cand, s.seen);
            if (s.nc == LCAP) { s.ovf = true; break; }    // degenerate codebook / massive ties: exact scan in the epilogue
            s.cand[s.nc * 128] = (unsigned short)(colbase + e);
            s.seen[s.nc * 128] = g;
            ++s.nc;
        }
        s.m_push = s.m_run;
    }
}

// ---- group-store sweep (VqtCfg::GS) ----
// per-thread state: running bound U = min (score + half margin), the live group (minimum, first column; its 16 raw scores
// sit in shared memory) and the count of the spill list (-1: overflow)
struct VqtSweepGs {
    float u_run, hm, g_live;
    int col_live, nc;
    uint32_t raw_addr;         // shared-memory address of this thread's quad 0 (quads are 2048 bytes apart)
    uint32_t smem_base;        // CTA's aligned dynamic shared memory (the rare spill path derives the list addresses from it)
};

__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

// a live group is displaced while its minimum is still within margin: its elements within `thr` go to the list
// (rare: two 16-column groups of the same half row within margin of each other).  All addresses are 32-bit shared-memory
// addresses derived here, so that the hot loop carries none of them: raw_off/cand_off/seen_off are the regions' offsets
// from the CTA's shared-memory base.
template <int LC>
__device__ __noinline__ int vqt_spill(float thr, int col_live, int nc, uint32_t raw_addr, uint32_t smem_base, uint32_t raw_off,
                                      uint32_t cand_off, uint32_t seen_off) {
    if (nc < 0) return nc;
    const uint32_t idx = (raw_addr - smem_base - raw_off) >> 4;            // slot * 512 + row
    const uint32_t ent = (idx >> 9) * (uint32_t)(LC * 128) + (idx & 127u);
    const uint32_t cand = smem_base + cand_off + ent * 2u, seen = smem_base + seen_off + ent * 4u;
#pragma unroll 1
    for (int e = 0; e < 16; ++e) {
        const float v = lds_f32(raw_addr + (uint32_t)(e >> 2) * 2048u + (uint32_t)(e & 3) * 4u);
        if (v <= thr) {
            if (nc == LC) {          // drop the entries that cannot be within margin of the final bound any more
                int w = 0;
                for (int c = 0; c < LC; ++c) {
                    const float sc = lds_f32(seen + (uint32_t)c * 512u);
                    if (sc <= thr) {
                        if (w != c) { sts_f32(seen + (uint32_t)w * 512u, sc); sts_u16(cand + (uint32_t)w * 256u, lds_u16(cand + (uint32_t)c * 256u)); }
                        ++w;
                    }
                }
                nc = w;
            }
            if (nc == LC) return -1;         // degenerate codebook / massive ties: exact scan in the epilogue
            sts_u16(cand + (uint32_t)nc * 256u, (uint32_t)(col_live + e));
            sts_f32(seen + (uint32_t)nc * 512u, v);
            ++nc;
        }
    }
    return nc;
}

// one 16-column group: minimum (8 three-input minima), bound update, and -- predicated, no branch on the usual path -- the
// group becomes the thread's live group when its minimum is within margin of the bound
template <int LC>
__device__ __forceinline__ void vqt_group_gs(VqtSweepGs &s, const uint32_t (&r)[16], int colbase, uint32_t raw_off, uint32_t cand_off, uint32_t seen_off) {
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
    const float g = fmin3(fmin3(fmin3(v[0], v[1], v[2]), fmin3(v[3], v[4], v[5]), fmin3(v[6], v[7], v[8])),
                          fmin3(fmin3(v[9], v[10], v[11]), fmin3(v[12], v[13], v[14]), v[15]), __int_as_float(0x7f800000));
    s.u_run = fminf(s.u_run, g + s.hm);
    const float thr = s.u_run + s.hm;
    const bool push = g <= thr;
    if (push && s.g_live <= thr) s.nc = vqt_spill<LC>(thr, s.col_live, s.nc, s.raw_addr, s.smem_base, raw_off, cand_off, seen_off);
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %17, 0;\n\t"
        "@q st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n\t"
        "@q st.shared.v4.b32 [%0 + 2048], {%5, %6, %7, %8};\n\t"
        "@q st.shared.v4.b32 [%0 + 4096], {%9, %10, %11, %12};\n\t"
        "@q st.shared.v4.b32 [%0 + 6144], {%13, %14, %15, %16};\n\t}"
        :: "r"(s.raw_addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"((uint32_t)push)
        : "memory");
    s.g_live = push ? g : s.g_live;
    s.col_live = push ? colbase : s.col_live;
}

// codeword gather, straight-through value, squared error and EMA statistics of one latent vector for the dimensions
// [dim0, dim0 + DH); the thread with dim0 == 0 also writes the index and the code count.  Returns the squared-error partial.
template <int D, int DH>
__device__ __forceinline__ float vqt_gather(const VqtParams &p, int64_t b, int64_t s, int best_k, int dim0, bool want_stats, bool dw_vec) {
    const float *xs = p.x + ((size_t)b * D + dim0) * p.S + s;
    float *qs = p.quant + ((size_t)b * D + dim0) * p.S + s;
    const float *e = p.embed + (size_t)best_k * D + dim0;
    float err = 0.0f;
#pragma unroll 1
    for (int ch = 0; ch < DH / 16; ++ch) {
        float4 ev[4];
        float xv[16];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) ev[jj] = __ldg(reinterpret_cast<const float4 *>(e) + ch * 4 + jj);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) xv[jj] = __ldg(xs + (size_t)(ch * 16 + jj) * p.S);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float qv[4] = {ev[jj].x, ev[jj].y, ev[jj].z, ev[jj].w};
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const float df = qv[l] - xv[jj * 4 + l];
                err = __fmaf_rn(df, df, err);
                qs[(size_t)(ch * 16 + jj * 4 + l) * p.S] = __fadd_rn(xv[jj * 4 + l], __fsub_rn(qv[l], xv[jj * 4 + l]));    // layers.py:720, two roundings
            }
            if (want_stats) {
                float *dst = p.dw + (size_t)best_k * D + dim0 + ch * 16 + jj * 4;
                if (dw_vec) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(xv[jj * 4]), "f"(xv[jj * 4 + 1]), "f"(xv[jj * 4 + 2]), "f"(xv[jj * 4 + 3]) : "memory");
                else { atomicAdd(dst, xv[jj * 4]); atomicAdd(dst + 1, xv[jj * 4 + 1]); atomicAdd(dst + 2, xv[jj * 4 + 2]); atomicAdd(dst + 3, xv[jj * 4 + 3]); }
            }
        }
    }
    if (dim0 == 0) {
        if (want_stats) atomicAdd(&p.counts[best_k], 1.0f);
        p.idx[(size_t)b * p.S + s] = best_k;
    }
    return err;
}

template <int D>
__global__ void __launch_bounds__(VqtCfg<D>::THREADS, 1)
vq_tc_kernel(const __grid_constant__ VqtParams p) {
    using Cfg = VqtCfg<D>;
    constexpr int NT = Cfg::NT, KS = Cfg::KS, NG = Cfg::NG, NABUF = Cfg::NABUF, NSTAGE = Cfg::NSTAGE;
    VQ3D_DYN_SMEM(unsigned char, smem_raw);
    __shared__ __align__(8) uint64_t e_full[NSTAGE], e_empty[NSTAGE], a_full[NG][NABUF], a_empty[NG][NABUF],
        d_full[NG][2], d_empty[NG][2], r_full[NG][2], r_empty[NG][2], b_full[NG][2], b_empty[NG][2];
    __shared__ uint32_t tmem_slot;
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (s_u32(smem_raw) + 127u) & ~127u;
    unsigned char *smem = smem_raw + (base - s_u32(smem_raw));
    const uint32_t ring_addr = base, a_addr = base + NSTAGE * Cfg::STAGE;
    unsigned char *sA = smem + (size_t)NSTAGE * Cfg::STAGE;
    float2 *s_xx = reinterpret_cast<float2 *>(sA + (size_t)NG * NABUF * Cfg::AIMG);    // [NG][2][128]: the row's margin inputs (r0, r1)
    float *s_min = reinterpret_cast<float *>(s_xx + NG * 2 * 128);                      // [NG][2][half][128]  (running U of the half)
    constexpr int NRB = Cfg::NRB;
    int *s_nc = reinterpret_cast<int *>(s_min + NG * NRB * 2 * 128);                    // [NG][NRB][half][128]  (-1: overflow)
    constexpr int LC = Cfg::LC;
    float *s_seen = reinterpret_cast<float *>(s_nc + NG * NRB * 2 * 128);               // [NG][NRB][half][LC][128]
    unsigned short *s_cand = reinterpret_cast<unsigned short *>(s_seen + NG * NRB * 2 * LC * 128);   // same shape
    uint32_t *s_queue = reinterpret_cast<uint32_t *>(s_cand + NG * NRB * 2 * LC * 128);              // [4*NG warps][64] + distances
    // per-tile norms and (when it fits) the column -> code table live in shared memory: the epilogue warps look them up on
    // their critical path, where every global-memory round trip costs ~1 us
    float *s_tnorm = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(s_queue) + Cfg::QUEUE);
    const float *s_twh = s_tnorm + (Cfg::HL ? 0 : 128), *s_twl = s_tnorm + (Cfg::HL ? 0 : 256);    // single bf16: per-tile ||wh||, ||wl||
    int *s_best = reinterpret_cast<int *>(s_tnorm + (Cfg::HL ? 1 : 3) * 128);           // [NG][2][128]  (-1: row past the end)
    // group-store sweep: live group minimum / first column / raw scores per (group, result buffer, half)
    float *s_glive = reinterpret_cast<float *>(s_best + NG * 2 * 128);                  // [NG][2][half][128]
    int *s_collive = reinterpret_cast<int *>(s_glive + (Cfg::GS ? NG * NRB * 2 * 128 : 0));
    float *s_raw = reinterpret_cast<float *>(s_collive + (Cfg::GS ? NG * NRB * 2 * 128 : 0));   // [NG][NRB][half][4 quads][128][4]
    unsigned short *s_perm = reinterpret_cast<unsigned short *>(s_raw + (Cfg::GS ? NG * NRB * 2 * 4 * 128 * 4 : 0));
    for (int i = threadIdx.x; i < p.Kpad / NT; i += Cfg::THREADS) {
        s_tnorm[i] = __ldg(p.tnorm + i);
        if constexpr (!Cfg::HL) { s_tnorm[128 + i] = __ldg(p.tbound + 2 * i); s_tnorm[256 + i] = __ldg(p.tbound + 2 * i + 1); }
    }
    if (p.perm_in_smem)
        for (int i = threadIdx.x; i < p.Kpad; i += Cfg::THREADS) s_perm[i] = (unsigned short)__ldg(p.perm + i);
    auto code_of = [&](int col) -> int { return p.perm_in_smem ? (int)s_perm[col] : __ldg(p.perm + col); };

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < NSTAGE; ++s) { mbarrier_init(&e_full[s], 1); mbarrier_init(&e_empty[s], 1); }
        for (int g = 0; g < NG; ++g) {
            for (int b = 0; b < NABUF; ++b) { mbarrier_init(&a_full[g][b], 4); mbarrier_init(&a_empty[g][b], 9); }
            for (int b = 0; b < 2; ++b) {
                mbarrier_init(&d_full[g][b], 1); mbarrier_init(&d_empty[g][b], 8);
                mbarrier_init(&r_full[g][b], 8); mbarrier_init(&r_empty[g][b], 4);
                mbarrier_init(&b_full[g][b], 4); mbarrier_init(&b_empty[g][b], 8);
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

#ifdef VQ3D_VQT_DEBUG
    const long long t_kernel0 = clock64();
#endif
    const int64_t N = p.B * p.S;
    const int64_t nsuper = (N + NG * 128 - 1) / (NG * 128);
    const int my_super = nsuper > (int64_t)blockIdx.x ? (int)((nsuper - 1 - blockIdx.x) / gridDim.x + 1) : 0;
    const int ntiles = p.Kpad / NT;
    const bool resident = ntiles <= NSTAGE;
    // who gathers: with few codebook tiles the load/epilogue warps pace the kernel, so the sweep warps (two threads per row,
    // idle between super-tiles) take the gather; with many tiles the sweep is the critical path and the epilogue keeps it
    const bool gather_in_sweep = ntiles <= 8;
    const bool want_stats = p.counts != nullptr;
#ifdef VQ3D_VQT_SCALAR_RED
    const bool dw_vec = false;
#else
    const bool dw_vec = (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0;
#endif
    double err_acc = 0.0;

    if (warp == 0) {
        // ===== producer: codebook tiles (once when resident, else the same tile sequence for every super-tile) =====
        if (elect_one() && my_super > 0) {
            uint32_t cnt = 0;
            const int rounds = resident ? 1 : my_super;
            for (int i = 0; i < rounds; ++i)
                for (int t = 0; t < ntiles; ++t, ++cnt) {
                    const uint32_t slot = cnt % NSTAGE, u = cnt / NSTAGE;
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
        // ===== MMA issuer: score tile (g, t) -> accumulator g*2 + (unit & 1) =====
        uint32_t cnt = 0, u = 0;
        for (int i = 0; i < my_super; ++i) {
            const int ab = i % NABUF;
            const uint32_t apar = (uint32_t)(i / NABUF) & 1u;
            for (int t = 0; t < ntiles; ++t, ++cnt, ++u) {
                const uint32_t slot = resident ? (uint32_t)t : cnt % NSTAGE;
                if (!resident || i == 0) mbarrier_wait(&e_full[slot], (cnt / NSTAGE) & 1u);
                const uint32_t e_img = ring_addr + slot * Cfg::STAGE;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (t == 0) VQT_TIMED_WAIT(4, mbarrier_wait(&a_full[g][ab], apar));
                    VQT_TIMED_WAIT(5, mbarrier_wait(&d_empty[g][u & 1u], ((u >> 1) & 1u) ^ 1u));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t a_img = a_addr + (uint32_t)(g * NABUF + ab) * Cfg::AIMG;
                        const uint32_t d_addr = tmem_d + (uint32_t)((g * 2 + (int)(u & 1u)) * NT);
                        if constexpr (Cfg::HL) {
                            // xh.eh, xh.el, xl.eh (D/16 K16 steps each), then the ||e||^2 columns
                            constexpr int KD = D / 16;
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) {
                                const int part = ks / KD, st = ks % KD;
                                const int ca = ks == KS - 1 ? 2 * (D / 8) : (part == 2 ? D / 8 : 0) + 2 * st;
                                const int cb = ks == KS - 1 ? 2 * (D / 8) : (part == 1 ? D / 8 : 0) + 2 * st;
                                umma_f16(d_addr, umma_desc(a_img + (uint32_t)ca * Cfg::LBO_A, Cfg::LBO_A, 128),
                                         umma_desc(e_img + (uint32_t)cb * Cfg::LBO_B, Cfg::LBO_B, 128), idesc, ks > 0 ? 1u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks)
                                umma_f16(d_addr, umma_desc(a_img + (uint32_t)(2 * ks) * Cfg::LBO_A, Cfg::LBO_A, 128),
                                         umma_desc(e_img + (uint32_t)(2 * ks) * Cfg::LBO_B, Cfg::LBO_B, 128), idesc, ks > 0 ? 1u : 0u);
                        }
                        umma_commit_to(&d_full[g][u & 1u]);
                        if (t == ntiles - 1) umma_commit_to(&a_empty[g][ab]);          // this super-tile's A image has been consumed
                        if (!resident && g == NG - 1) umma_commit_to(&e_empty[slot]);    // every group's MMAs on this tile have been issued
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= Cfg::SWEEP_WARP0 && warp < Cfg::LE_WARP0) {
        // ===== sweep warps =====
        const int g = ((warp - Cfg::SWEEP_WARP0) >> 2) & 1, half = (warp - Cfg::SWEEP_WARP0) >> 3, q = warp & 3, row = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        constexpr int HC = NT / 2;           // columns per half
        VqtSweep sw;
        VqtSweepGs gs;
        uint32_t u = 0;
        // offsets of the group-store regions from the CTA's shared-memory base (compile-time but for the base)
        const uint32_t raw_off = s_u32(s_raw) - base, cand_off = s_u32(s_cand) - base, seen_off = s_u32(s_seen) - base;
        // the two per-thread address bases of the hot loop, made opaque so that they live in registers instead of being
        // re-derived from the thread index in every iteration
        uint32_t tm_base = tmem_d + lane_sel + (uint32_t)(g * 2 * NT + half * HC);
        uint32_t raw_row = base + raw_off + (uint32_t)((g * NRB * 2 + half) * 512 + row) * 16u;
        asm volatile("" : "+r"(tm_base), "+r"(raw_row));
        // ---- gather / straight-through / loss partial / statistics of super-tile j for this thread's half of the dimensions
        // (the code index comes from the load/epilogue warps through s_best) ----
        auto gather = [&](int j) {
            const int rb = j & 1;
            mbarrier_wait(&b_full[g][rb], ((uint32_t)j >> 1) & 1u);
            const int best_k = s_best[(g * 2 + rb) * 128 + row];
            __syncwarp();
            if (lane == 0) mbarrier_arrive(&b_empty[g][rb]);
            if (best_k < 0) return;
            const int64_t sup = (int64_t)blockIdx.x + (int64_t)j * gridDim.x;
            const int64_t v = (sup * NG + g) * 128 + row;
            const int64_t b = p.B == 1 ? 0 : v / p.S, s = v - b * p.S;
            err_acc += (double)vqt_gather<D, D / 2>(p, b, s, best_k, half * (D / 2), want_stats, dw_vec);
        };
        for (int i = 0; i < my_super; ++i) {
            const int rb = i & 1, ab = i % NABUF, rs = i % NRB;       // s_xx buffer, A buffer, result buffer
            if (gather_in_sweep && i >= 2) VQT_TIMED_WAIT(16, gather(i - 2));
            VQT_TIMED_WAIT(6, mbarrier_wait(&r_empty[g][rs], ((uint32_t)(i / NRB) & 1u) ^ 1u));
            VQT_TIMED_WAIT(7, mbarrier_wait(&a_full[g][ab], (uint32_t)(i / NABUF) & 1u));
            const float2 rr = s_xx[(g * 2 + rb) * 128 + row];       // the row's margin inputs
            __syncwarp();
            if (lane == 0) mbarrier_arrive(&a_empty[g][ab]);
            const int slot_res = (g * NRB + rs) * 2 + half;
            if constexpr (Cfg::GS) {
                gs.u_run = __int_as_float(0x7f800000);
                gs.g_live = __int_as_float(0x7f800000);
                gs.col_live = 0;
                gs.nc = 0;
                gs.smem_base = base;
                gs.raw_addr = raw_row + (uint32_t)rs * (2u * 512u * 16u);
            } else {
                sw.u_run = __int_as_float(0x7f800000);
                sw.m_run = __int_as_float(0x7f800000);
                sw.m_push = __int_as_float(0x7f800000);
                sw.nc = 0;
                sw.ovf = false;
                sw.cand = s_cand + (size_t)slot_res * LC * 128 + row;
                sw.seen = s_seen + (size_t)slot_res * LC * 128 + row;
            }
            for (int t = 0; t < ntiles; ++t, ++u) {
                VQT_TIMED_WAIT(8, mbarrier_wait(&d_full[g][u & 1u], (u >> 1) & 1u));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_addr = tm_base + (u & 1u) * (uint32_t)NT;
                const int col0 = t * NT + half * HC;
                if constexpr (Cfg::GS) gs.hm = vqt_half_margin<Cfg::HL>(rr.x, rr.y, s_tnorm[t], s_twh[t], s_twl[t]);
                else sw.hm = vqt_half_margin<Cfg::HL>(rr.x, rr.y, s_tnorm[t], s_twh[t], s_twl[t]);
                // ping-pong over the 16-column groups: the next group's tcgen05.ld is in flight while this one is scanned
                uint32_t ra[16], rb16[16];
#ifdef VQ3D_VQT_DEBUG
                const long long t_tile0 = clock64();
#endif
                tmem_ld16_async(d_addr, ra);
                tmem_wait_ld16(ra);
#pragma unroll 1
                for (int c = 0; c < HC / 16; c += 2) {
                    tmem_ld16_async(d_addr + (uint32_t)((c + 1) * 16), rb16);
                    if constexpr (Cfg::GS) vqt_group_gs<LC>(gs, ra, col0 + c * 16, raw_off, cand_off, seen_off);
                    else vqt_group<LC>(sw, ra, col0 + c * 16);
                    tmem_wait_ld16(rb16);
                    if (c + 2 < HC / 16) tmem_ld16_async(d_addr + (uint32_t)((c + 2) * 16), ra);
                    if constexpr (Cfg::GS) vqt_group_gs<LC>(gs, rb16, col0 + (c + 1) * 16, raw_off, cand_off, seen_off);
                    else vqt_group<LC>(sw, rb16, col0 + (c + 1) * 16);
                    tmem_wait_ld16(ra);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbarrier_arrive(&d_empty[g][u & 1u]);
#ifdef VQ3D_VQT_DEBUG
                if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(const_cast<unsigned *>(p.dbg)) + 17, (unsigned long long)(clock64() - t_tile0));
#endif
            }
            if constexpr (Cfg::GS) {
                s_min[slot_res * 128 + row] = gs.u_run;
                s_nc[slot_res * 128 + row] = gs.nc;
                s_glive[slot_res * 128 + row] = gs.g_live;
                s_collive[slot_res * 128 + row] = gs.col_live;
            } else {
                s_min[slot_res * 128 + row] = sw.u_run;
                s_nc[slot_res * 128 + row] = sw.ovf ? -1 : sw.nc;
            }
            __syncwarp();
            if (lane == 0) mbarrier_arrive(&r_full[g][rs]);
        }
        if (gather_in_sweep)
            for (int j = my_super >= 2 ? my_super - 2 : 0; j < my_super; ++j) gather(j);
    } else if (warp >= Cfg::LE_WARP0) {
        // ===== load / epilogue warps =====
        const int g = (warp - Cfg::LE_WARP0) >> 2, q = warp & 3, row = q * 32 + lane;
        const bool want_stats = p.counts != nullptr;
        #ifdef VQ3D_VQT_SCALAR_RED
        const bool dw_vec = false;
#else
        const bool dw_vec = (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0;
#endif
        auto locate = [&](int i, int64_t &b, int64_t &s) -> bool {
            const int64_t sup = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
            const int64_t v = (sup * NG + g) * 128 + row;
            const bool active = v < N;
            if (p.B == 1) { b = 0; s = active ? v : 0; return active; }       // the usual case: no 64-bit division
            b = active ? v / p.S : 0;
            s = active ? v - b * p.S : 0;
            return active;
        };
        // ---- stage the A operand of super-tile i: [x | 1 1 1 0 ...] as bf16, K-major; ||x||^2 for the margin ----
        auto stage = [&](int i) {
            const int ab = i % NABUF, rb = i & 1;
            int64_t b, s;
            const bool active = locate(i, b, s);
            const float *xs = p.x + (size_t)b * D * p.S + s;
            VQT_TIMED_WAIT(9, mbarrier_wait(&a_empty[g][ab], ((uint32_t)(i / NABUF) & 1u) ^ 1u));
            unsigned char *a_img = sA + (size_t)(g * NABUF + ab) * Cfg::AIMG + (size_t)row * 16;
            float xx = 0.0f, xl2 = 0.0f;
            const float *px = xs;
#pragma unroll 4
            for (int kc = 0; kc < D / 8; ++kc) {
                float xv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e, px += p.S) xv[e] = active ? __ldg(px) : 0.0f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if constexpr (Cfg::HL) {
                        xx = __fmaf_rn(xv[e], xv[e], xx);
                    } else {        // ||xh||^2 and ||xl||^2: the actual rounding residual of this vector
                        const float xh = __bfloat162float(__float2bfloat16_rn(xv[e])), xl = xv[e] - xh;
                        xx = __fmaf_rn(xh, xh, xx);
                        xl2 = __fmaf_rn(xl, xl, xl2);
                    }
                }
                uint4 ph;
                ph.x = bf16x2(xv[0], xv[1]); ph.y = bf16x2(xv[2], xv[3]); ph.z = bf16x2(xv[4], xv[5]); ph.w = bf16x2(xv[6], xv[7]);
                *reinterpret_cast<uint4 *>(a_img + (size_t)kc * Cfg::LBO_A) = ph;
                if constexpr (Cfg::HL) {       // lo parts: x - bf16(x), exact in fp32
                    float xl[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) xl[e] = xv[e] - __bfloat162float(__float2bfloat16_rn(xv[e]));
                    uint4 pl;
                    pl.x = bf16x2(xl[0], xl[1]); pl.y = bf16x2(xl[2], xl[3]); pl.z = bf16x2(xl[4], xl[5]); pl.w = bf16x2(xl[6], xl[7]);
                    *reinterpret_cast<uint4 *>(a_img + (size_t)(D / 8 + kc) * Cfg::LBO_A) = pl;
                }
            }
            uint4 one, zero;
            one.x = bf16x2(1.0f, 1.0f); one.y = bf16x2(1.0f, 0.0f); one.z = 0u; one.w = 0u;
            zero.x = zero.y = zero.z = zero.w = 0u;
            constexpr int KX = (Cfg::HL ? 2 * D : D) / 8;
            *reinterpret_cast<uint4 *>(a_img + (size_t)KX * Cfg::LBO_A) = one;
            *reinterpret_cast<uint4 *>(a_img + (size_t)(KX + 1) * Cfg::LBO_A) = zero;
            // margin inputs: hi/lo operands (||x||, ||x||^2); single bf16 (||xh||, ||xl||), inflated against their own roundings
            s_xx[(g * 2 + rb) * 128 + row] = Cfg::HL ? make_float2(sqrtf(xx), xx) : make_float2(sqrtf(xx) * 1.0001f, sqrtf(xl2) * 1.0001f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbarrier_arrive(&a_full[g][ab]);
        };
        if (my_super > 0) stage(0);
        if (my_super > 1) stage(1);
        for (int i = 0; i < my_super; ++i) {
            const int rb = i & 1, rs = i % NRB;
            int64_t b, s;
            const bool active = locate(i, b, s);
            VQT_TIMED_WAIT(10, mbarrier_wait(&r_full[g][rs], (uint32_t)(i / NRB) & 1u));
#ifdef VQ3D_VQT_DEBUG
            long long t_ph = clock64();
            auto phase = [&](int slot) { const long long t1 = clock64(); if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(const_cast<unsigned *>(p.dbg)) + slot, (unsigned long long)(t1 - t_ph)); t_ph = t1; };
#else
            auto phase = [&](int) {};
#endif
            // merge the two half-row lists: final minimum, then the entries that can still be within margin of it
            const int res0 = (g * NRB + rs) * 2;
            const float u_fin = fminf(s_min[res0 * 128 + row], s_min[(res0 + 1) * 128 + row]);
            const float2 rr = s_xx[(g * 2 + rb) * 128 + row];
            // entry (h, c) can still hold the reference's argmin: its score lower bound is within its tile's half margin of U
            auto kept = [&](int h, int c) -> bool {
                const int col = (int)s_cand[((size_t)(res0 + h) * LC + c) * 128 + row];
                return s_seen[((size_t)(res0 + h) * LC + c) * 128 + row] <=
                       u_fin + vqt_half_margin<Cfg::HL>(rr.x, rr.y, s_tnorm[col / NT], s_twh[col / NT], s_twl[col / NT]);
            };
            // group-store sweep: the elements of half h's live group within thr_live_h of U are candidates (-inf: the group went stale)
            float thr_live0 = __int_as_float(0xff800000), thr_live1 = thr_live0;
            int col_live0 = 0, col_live1 = 0;
            // the first four candidate columns stay in registers (the re-rank takes them from there; a fifth and later one,
            // rare, is looked up again in enumeration order)
            int nc = 0, kc0 = 0x7fffffff, kc1 = 0, kc2 = 0, kc3 = 0;
            auto add = [&](int col) {
                kc3 = nc == 3 ? col : kc3;
                kc2 = nc == 2 ? col : kc2;
                kc1 = nc == 1 ? col : kc1;
                kc0 = nc == 0 ? col : kc0;
                ++nc;
            };
            bool ovf = false;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int nh = s_nc[(res0 + h) * 128 + row];
                ovf |= nh < 0;
                for (int c = 0; c < nh; ++c)
                    if (kept(h, c)) add((int)s_cand[((size_t)(res0 + h) * LC + c) * 128 + row]);
                if constexpr (Cfg::GS) {
                    const int cl = s_collive[(res0 + h) * 128 + row];
                    const float thr = u_fin + vqt_half_margin<Cfg::HL>(rr.x, rr.y, s_tnorm[cl / NT], s_twh[cl / NT], s_twl[cl / NT]);
                    if (s_glive[(res0 + h) * 128 + row] <= thr) {
                        (h ? thr_live1 : thr_live0) = thr;
                        (h ? col_live1 : col_live0) = cl;
                        const float4 *rw = reinterpret_cast<const float4 *>(s_raw + (size_t)(res0 + h) * 4 * 512 + row * 4);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const float4 v4 = rw[qd * 128];
                            if (v4.x <= thr) add(cl + qd * 4);
                            if (v4.y <= thr) add(cl + qd * 4 + 1);
                            if (v4.z <= thr) add(cl + qd * 4 + 2);
                            if (v4.w <= thr) add(cl + qd * 4 + 3);
                        }
                    }
                }
            }
            // candidate enumeration for the re-rank: position -> (is a candidate, column); PER positions per half
            constexpr int PER = LC + (Cfg::GS ? 16 : 0);
            auto cand_at = [&](int pos, int &col) -> bool {
                const int h = pos / PER, j = pos - h * PER;
                if (Cfg::GS && j >= LC) {
                    const int e = j - LC;
                    col = (h ? col_live1 : col_live0) + e;
                    return s_raw[((size_t)(res0 + h) * 4 + (e >> 2)) * 512 + row * 4 + (e & 3)] <= (h ? thr_live1 : thr_live0);
                }
                if (j >= s_nc[(res0 + h) * 128 + row]) return false;
                col = (int)s_cand[((size_t)(res0 + h) * LC + j) * 128 + row];
                return kept(h, j);
            };
            // n-th candidate in enumeration order (n >= 4: the ones that did not fit the registers)
            auto nth_cand = [&](int n) -> int {
                int seen_n = 0;
#pragma unroll 1
                for (int pos = 0; pos < 2 * PER; ++pos) {
                    int c = 0;
                    if (cand_at(pos, c)) {
                        if (seen_n == n) return c;
                        ++seen_n;
                    }
                }
                return p.Kpad;         // cannot happen (the count came from the same tests)
            };
            const int k_one = kc0;
            int best_k = 0x7fffffff;
            bool need_scan = false;
            if (active) {
#ifdef VQ3D_VQT_DEBUG
                {
                    unsigned *dbg = const_cast<unsigned *>(p.dbg);
                    atomicAdd(dbg + (ovf ? 3 : (nc == 1 ? 0 : (nc >= 2 ? 1 : 2))), 1u);
                    if (nc >= 2) atomicAdd(dbg + 4, (unsigned)nc);
                }
#endif
                if (!ovf && nc == 1 && k_one < p.K) best_k = code_of(k_one);      // alone within the error bound: it IS the reference's argmin
            }
            phase(12);       // merge
            // exact re-rank of the vectors with several candidates, densely packed over the warp: the (vector, code) pairs
            // go through a 32-entry queue, every lane evaluates one pair per round with the reference's arithmetic, the
            // owners keep the lexicographic (sqrt(d2), k) minimum = the reference's argmin (first minimum)
            {
                const bool requester = active && !ovf && best_k == 0x7fffffff && nc >= 1;
                uint32_t *q_key = s_queue + (size_t)(warp - Cfg::LE_WARP0) * 64;
                float *q_r = reinterpret_cast<float *>(q_key + 32);
                int remaining = requester ? nc : 0, cursor = 0;
                float best_r = __int_as_float(0x7f800000);
#pragma unroll 1
                while (__any_sync(0xffffffffu, remaining > 0)) {
                    int incl = remaining;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    const int excl = incl - remaining;
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    int take = 32 - excl;
                    take = take < 0 ? 0 : (take > remaining ? remaining : take);
                    for (int j = 0; j < take; ++j) {
                        // next candidate of this row
                        const int col = cursor == 0 ? kc0 : (cursor == 1 ? kc1 : (cursor == 2 ? kc2 : (cursor == 3 ? kc3 : nth_cand(cursor))));
                        q_key[excl + j] = (uint32_t)(col < p.K ? code_of(col) : 0xffff) | ((uint32_t)lane << 16);
                        ++cursor;
                    }
                    __syncwarp();
                    phase(18);       // re-rank: queueing
                    const int nq = total < 32 ? total : 32;
#pragma unroll 1
                    for (int pp = lane; pp < 32; pp += 32) {
                        const uint32_t key = pp < nq ? q_key[pp] : 0u;
                        const int owner = (int)(key >> 16), k = (int)(key & 0xffffu);
                        const int64_t ob = __shfl_sync(0xffffffffu, b, owner), os = __shfl_sync(0xffffffffu, s, owner);
                        if (pp < nq)
                            q_r[pp] = k < p.K ? __fsqrt_rn(vqt_ref_dist2<D>(p.x + (size_t)ob * D * p.S + os, p.S, p.embed + (size_t)k * D))
                                              : __int_as_float(0x7fc00000);
                    }
                    __syncwarp();
                    for (int j = 0; j < take; ++j) {
                        const float r = q_r[excl + j];
                        const int k = (int)(q_key[excl + j] & 0xffffu);
                        if (r < best_r || (r == best_r && k < best_k)) { best_r = r; best_k = k; }
                    }
                    remaining -= take;
                    __syncwarp();
                    phase(19);       // re-rank: evaluation of a round
#ifdef VQ3D_VQT_DEBUG
                    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(const_cast<unsigned *>(p.dbg)) + 20, 1ull);
#endif
                }
            }
            phase(13);       // re-rank
            if (active) need_scan = best_k == 0x7fffffff;        // list overflow or nothing finite (NaN input)
            __syncwarp();
            if (lane == 0) mbarrier_arrive(&r_empty[g][rs]);
            // exact scan of the whole codebook for the (rare) vectors without a usable list: the warp splits the codes,
            // every lane keeps its first minimum, then a lexicographic (distance, index) reduction
            __syncwarp();
            unsigned scan_mask = __ballot_sync(0xffffffffu, need_scan);
#pragma unroll 1
            while (scan_mask) {
                const int src = __ffs(scan_mask) - 1;
                scan_mask &= scan_mask - 1;
                const int64_t sb = __shfl_sync(0xffffffffu, b, src), ss = __shfl_sync(0xffffffffu, s, src);
                const float *sxs = p.x + (size_t)sb * D * p.S + ss;
                float r_best = __int_as_float(0x7f800000);
                int k_best = 0x7fffffff;
                for (int k = lane; k < p.K; k += 32) {
                    const float r = __fsqrt_rn(vqt_ref_dist2<D>(sxs, p.S, p.embed + (size_t)k * D));
                    if (r < r_best || k_best == 0x7fffffff) { r_best = r; k_best = k; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float r2 = __shfl_xor_sync(0xffffffffu, r_best, o);
                    const int k2 = __shfl_xor_sync(0xffffffffu, k_best, o);
                    const bool take = k2 != 0x7fffffff && (k_best == 0x7fffffff || r2 < r_best || (r2 == r_best && k2 < k_best) ||
                                                           (r_best != r_best && r2 != r2 && k2 < k_best));
                    if (take) { r_best = r2; k_best = k2; }
                }
                if (lane == src) best_k = k_best;
            }
            // the resolved code goes to the sweep warps, which gather the codeword and write the outputs in their idle time
            // between two super-tiles (two threads per row there, half of the dimensions each)
            if (gather_in_sweep) {
                VQT_TIMED_WAIT(9, mbarrier_wait(&b_empty[g][rb], (((uint32_t)i >> 1) & 1u) ^ 1u));
                s_best[(g * 2 + rb) * 128 + row] = active ? best_k : -1;
                __syncwarp();
                if (lane == 0) mbarrier_arrive(&b_full[g][rb]);
            } else if (active) {
                err_acc += (double)vqt_gather<D, D>(p, b, s, best_k, 0, want_stats, dw_vec);
            }
            phase(14);       // fallback scan + hand-over
            if (i + 2 < my_super) stage(i + 2);
            phase(15);       // stage (incl. its a_empty wait)
        }
    }
#ifdef VQ3D_VQT_DEBUG
    if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long *>(const_cast<unsigned *>(p.dbg)) + 11, (unsigned long long)(clock64() - t_kernel0));
#endif
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

// workspace: [0, 256) header (debug counters), ||e||^2 [Kpad], column -> code [Kpad], per-tile largest norm [ntiles], then
// the tile images (256-byte aligned)
static size_t vqt_ws_tables(int Kpad, int ntiles) { return (256 + (size_t)Kpad * 8 + (size_t)ntiles * 12 + 255) / 256 * 256; }

template <int D>
static size_t vqt_ws_bytes(int K) {
    using Cfg = VqtCfg<D>;
    const int Kpad = (K + Cfg::NT - 1) / Cfg::NT * Cfg::NT;
    return vqt_ws_tables(Kpad, Kpad / Cfg::NT) + (size_t)(Kpad / Cfg::NT) * Cfg::STAGE;
}

template <int D>
static int launch_vqt(const float *x, const float *embed, int64_t B, int64_t S, int K, float *quant, int64_t *idx, double *sqerr,
                      float *counts, float *dw, void *ws, size_t ws_size, void *stream) {
    using Cfg = VqtCfg<D>;
    if (ws_size < vqt_ws_bytes<D>(K) || (reinterpret_cast<uintptr_t>(ws) & 255) != 0)
        return fail(VQ3D_ERR_INVALID, "vq_assign_tc: workspace too small or not 256-byte aligned");
    if (K > 16384) return fail(VQ3D_ERR_UNSUPPORTED, "vq_assign_tc: K > 16384");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VqtParams p;
    p.x = x; p.embed = embed; p.B = B; p.S = S; p.K = K;
    p.Kpad = (K + Cfg::NT - 1) / Cfg::NT * Cfg::NT;
    const int ntiles = p.Kpad / Cfg::NT;
    unsigned char *wsb = static_cast<unsigned char *>(ws);
    float *cnorm = reinterpret_cast<float *>(wsb + 256);
    int *perm = reinterpret_cast<int *>(cnorm + p.Kpad);
    float *tnorm = reinterpret_cast<float *>(perm + p.Kpad);
    unsigned char *wimg = wsb + vqt_ws_tables(p.Kpad, ntiles);
    p.dbg = reinterpret_cast<const unsigned *>(wsb + 16);
    float *tbound = tnorm + ntiles;
    p.perm = perm; p.tnorm = tnorm; p.tbound = tbound; p.wimg = wimg;
    p.quant = quant; p.idx = idx; p.sqerr = sqerr; p.counts = counts; p.dw = dw;
    p.tmem_cols = 512;                                                   // NG groups x 2 accumulators x NT columns
    cudaError_t e = cudaMemsetAsync(wsb, 0, 256, st);
    if (e == cudaSuccess && sqerr != nullptr) e = cudaMemsetAsync(sqerr, 0, sizeof(double), st);
    if (e != cudaSuccess) return check_cuda(e, "vq_assign_tc(memset)");
    vqt_norm_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, st>>>(embed, K, D, cnorm);
    vqt_rank_kernel<<<(unsigned)ceil_div(p.Kpad, 256), 256, 0, st>>>(cnorm, K, p.Kpad, perm);
    vqt_prep_kernel<D><<<(unsigned)ceil_div((int64_t)p.Kpad * Cfg::KC, 256), 256, 0, st>>>(embed, cnorm, perm, K, p.Kpad, wimg, tnorm);
    if constexpr (!Cfg::HL) {
        e = cudaMemsetAsync(tbound, 0, (size_t)ntiles * 8, st);
        if (e != cudaSuccess) return check_cuda(e, "vq_assign_tc(memset)");
        vqt_tile_bounds_kernel<D><<<(unsigned)ceil_div(K, 256), 256, 0, st>>>(embed, perm, K, Cfg::NT, tbound);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return check_cuda(e, "vq_assign_tc(prep)");
    auto kernel = vq_tc_kernel<D>;
    static_assert(Cfg::NG * 2 * Cfg::NT == 512, "TMEM budget");
    size_t smem = Cfg::smem;                                             // > 113 KB: one CTA per SM (it owns all of TMEM)
    p.perm_in_smem = smem + (size_t)p.Kpad * 2 <= Cfg::SMEM_LIMIT ? 1 : 0;
    if (p.perm_in_smem) smem += (size_t)p.Kpad * 2;
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
    if (K < 1 || K > 16384) return 0;
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
