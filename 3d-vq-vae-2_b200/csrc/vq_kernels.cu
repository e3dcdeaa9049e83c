// vq_kernels.cu -- the EMA vector quantizer of vqvae/layers.py:602-728 as sm_100a kernels.
//
// Data layout in HBM: latents stay in the reference's (B, D, S) planar layout (S = H*W*Z
// innermost), so one thread per latent vector reads D fully coalesced planes and no
// (B,D,H,W,Z)->(B,H,W,Z,D) transpose copy (layers.py:690-693) is ever materialised.  The
// codebook (K*D fp32, <= 64 KB at the model's sizes) is staged in shared memory and read
// as warp-wide broadcasts.  The N x K distance matrix (layers.py:701) and the N x K
// one-hot (layers.py:638) never exist.
//
// Index exactness: distances are accumulated in the exact fp32 order of ATen's cdist
// ('donot_use_mm_for_euclid_dist'): two roundings (mul, add) for d < 4*floor(D/4), fma for
// the D%4 tail; the reference then compares sqrt(d2), and argmin keeps the FIRST minimum.
// sqrt can merge distinct d2 into one float, so the scan tracks the equivalence class of
// sqrt(best) and only moves the index when the class changes (see scan_update).
#include "vq3d_rt.h"

namespace vq3d {

constexpr int kVqThreads = 256;
constexpr int kVqCodebookSmemBytes = 32 * 1024;   // codebook chunk staged per pass
constexpr int kVqStatsSmemBytes = 72 * 1024;      // per-CTA [K] + [K,D] accumulators if they fit

struct Best {
    float d2;    // smallest squared distance so far
    float root;  // sqrtf(d2) of the class the index belongs to
    int idx;     // first index whose sqrt(d2) equals `root`
};

__device__ __forceinline__ void scan_update(Best &b, float d2, int k) {
    if (d2 < b.d2) {                       // strict: ties in d2 keep the lower index
        const float r = __fsqrt_rn(d2);
        if (r != b.root) {                 // a strictly smaller sqrt: new class, k is its first member
            b.root = r;
            b.idx = k;
        }                                  // else sqrt rounds to the same float: reference sees a tie, lower index stays
        b.d2 = d2;
    }
}

template <int D>
__device__ __forceinline__ float ref_dist2(const float (&x)[D], const float *__restrict__ e) {
    constexpr int NV = (D / 4) * 4;
    float agg = 0.0f;
#pragma unroll
    for (int d = 0; d < NV; ++d) {
        const float diff = __fsub_rn(x[d], e[d]);
        agg = __fadd_rn(agg, __fmul_rn(diff, diff));
    }
#pragma unroll
    for (int d = NV; d < D; ++d) {
        const float diff = __fsub_rn(x[d], e[d]);
        agg = __fmaf_rn(diff, diff, agg);
    }
    return agg;
}

__device__ __forceinline__ float ref_dist2_dyn(const float *__restrict__ xs, int xstride, const float *__restrict__ e, int D) {
    const int nv = (D / 4) * 4;
    float agg = 0.0f;
    int d = 0;
    for (; d < nv; ++d) {
        const float diff = __fsub_rn(xs[d * xstride], e[d]);
        agg = __fadd_rn(agg, __fmul_rn(diff, diff));
    }
    for (; d < D; ++d) {
        const float diff = __fsub_rn(xs[d * xstride], e[d]);
        agg = __fmaf_rn(diff, diff, agg);
    }
    return agg;
}

__device__ __forceinline__ double block_sum_double(double v, double *red /* [32] shared */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = (lane < (int)((blockDim.x + 31) >> 5)) ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}

// DT > 0: compile-time embedding dim, latent vector in registers.
// DT == 0: run-time D, latent vector staged in shared memory (column per thread).
template <int DT>
__global__ void __launch_bounds__(kVqThreads)
vq_assign_kernel(const float *__restrict__ x, const float *__restrict__ embed, int64_t S, int Drt, int K, int kchunk,
                 int64_t tiles_per_batch, int64_t total_tiles, int smem_stats,
                 float *__restrict__ quant, int64_t *__restrict__ idx_out, double *__restrict__ sqerr,
                 float *__restrict__ counts, float *__restrict__ dw) {
    const int D = DT > 0 ? DT : Drt;
    VQ3D_DYN_SMEM(float, smem);
    __shared__ double red[32];
    float *s_code = smem;                                   // [kchunk * D]
    float *s_x = s_code + (size_t)kchunk * D;               // DT == 0: [D * kVqThreads]
    float *s_stats = s_x + (DT > 0 ? 0 : (size_t)D * kVqThreads);   // [K + K*D] if smem_stats
    const bool want_stats = counts != nullptr;
    const int tid = threadIdx.x;

    if (want_stats && smem_stats) {
        for (int i = tid; i < K * (D + 1); i += kVqThreads) s_stats[i] = 0.0f;
    }
    double err_acc = 0.0;

    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int64_t b = tile / tiles_per_batch;
        const int64_t s = (tile - b * tiles_per_batch) * kVqThreads + tid;
        const bool active = s < S;
        const float *xb = x + (size_t)b * D * S;
        float xr[DT > 0 ? DT : 1];
        if (DT > 0) {
#pragma unroll
            for (int d = 0; d < (DT > 0 ? DT : 1); ++d) xr[d] = active ? xb[(size_t)d * S + s] : 0.0f;
        } else {
            for (int d = 0; d < D; ++d) s_x[d * kVqThreads + tid] = active ? xb[(size_t)d * S + s] : 0.0f;
        }
        Best best;
        best.d2 = __int_as_float(0x7f800000);   // +inf
        best.root = best.d2;
        best.idx = 0;
        for (int k0 = 0; k0 < K; k0 += kchunk) {
            const int kc = min(kchunk, K - k0);
            __syncthreads();                      // previous chunk fully consumed
            for (int i = tid; i < kc * D; i += kVqThreads) s_code[i] = embed[(size_t)k0 * D + i];
            __syncthreads();
            for (int k = 0; k < kc; ++k) {
                float d2;
                if (DT > 0) d2 = ref_dist2<(DT > 0 ? DT : 1)>(xr, s_code + k * D);
                else d2 = ref_dist2_dyn(s_x + tid, kVqThreads, s_code + k * D, D);
                scan_update(best, d2, k0 + k);
            }
        }
        // gather + straight-through value + loss partial + EMA statistics
        if (active) {
            const float *e = embed + (size_t)best.idx * D;
            float err = 0.0f;
            for (int d = 0; d < D; ++d) {
                const float q = __ldg(e + d);
                const float xv = DT > 0 ? xr[DT > 0 ? d : 0] : s_x[d * kVqThreads + tid];
                const float df = q - xv;
                err = __fmaf_rn(df, df, err);
                // the reference returns inputs + (quantized - inputs).detach() (layers.py:720):
                // two fp32 roundings, not q itself -- reproduce them so the decoder sees the same bits
                quant[((size_t)b * D + d) * S + s] = __fadd_rn(xv, __fsub_rn(q, xv));
                if (want_stats) {
                    if (smem_stats) atomicAdd(&s_stats[K + best.idx * D + d], xv);
                    else atomicAdd(&dw[(size_t)best.idx * D + d], xv);
                }
            }
            if (want_stats) {
                if (smem_stats) atomicAdd(&s_stats[best.idx], 1.0f);
                else atomicAdd(&counts[best.idx], 1.0f);
            }
            idx_out[(size_t)b * S + s] = best.idx;
            err_acc += (double)err;
        }
    }
    const double tot = block_sum_double(err_acc, red);
    if (tid == 0 && sqerr != nullptr) atomicAdd(sqerr, tot);
    if (want_stats && smem_stats) {
        __syncthreads();
        for (int i = tid; i < K * (D + 1); i += kVqThreads) {
            const float v = s_stats[i];
            if (v != 0.0f) atomicAdd(i < K ? &counts[i] : &dw[i - K], v);
        }
    }
}

__global__ void vq_loss_kernel(const double *__restrict__ sqerr, float cc, double inv_numel, float *__restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *loss = __fmul_rn(cc, (float)(*sqerr * inv_numel));
}

// single CTA: K <= a few thousand, K*D <= ~1M
__global__ void __launch_bounds__(1024)
vq_ema_update_kernel(const float *__restrict__ counts, const float *__restrict__ dw, int K, int D, float decay, float omd,
                     float alpha, float k_alpha, float *__restrict__ cluster_size, float *__restrict__ embed_avg,
                     float *__restrict__ embed) {
    __shared__ double red[32];
    __shared__ float s_n;
    double part = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float cs = __fmaf_rn(counts[k], omd, __fmul_rn(cluster_size[k], decay));
        cluster_size[k] = cs;
        part += (double)cs;
    }
    const double tot = block_sum_double(part, red);
    if (threadIdx.x == 0) s_n = (float)tot;
    __syncthreads();
    const float n = s_n;
    const float denom = __fadd_rn(n, k_alpha);
    for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
        const int k = i / D;
        const float ea = __fmaf_rn(dw[i], omd, __fmul_rn(embed_avg[i], decay));
        embed_avg[i] = ea;
        const float smoothed = __fmul_rn(n, (cluster_size[k] + alpha) / denom);
        embed[i] = ea / smoothed;
    }
}

// grid (chunks, D, B): double sums of x and x^2 per channel
__global__ void __launch_bounds__(256)
vq_init_sums_kernel(const float *__restrict__ x, int D, int64_t S, double *__restrict__ scratch) {
    __shared__ double red[32];
    const int d = blockIdx.y;
    const float *p = x + ((size_t)blockIdx.z * D + d) * S;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)p[s];
        s1 += v;
        s2 += v * v;
    }
    const double t1 = block_sum_double(s1, red);
    const double t2 = block_sum_double(s2, red);
    if (threadIdx.x == 0) {
        atomicAdd(&scratch[d], t1);
        atomicAdd(&scratch[D + d], t2);
    }
}

__global__ void vq_init_finish_kernel(const double *__restrict__ scratch, int D, double n, float *__restrict__ meanstd) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D) {
        const double mean = scratch[d] / n;
        double var = (scratch[D + d] - scratch[d] * mean) / (n - 1.0);   // unbiased (torch.std default)
        if (var < 0.0) var = 0.0;
        meanstd[d] = (float)mean;
        meanstd[D + d] = (float)sqrt(var);
    }
}

__global__ void vq_init_apply_kernel(const float *__restrict__ meanstd, int K, int D, float add_cluster,
                                     float *__restrict__ embed, float *__restrict__ embed_avg,
                                     float *__restrict__ cluster_size, int64_t *__restrict__ first_pass) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K * D) {
        const int d = i % D;
        const float e = __fadd_rn(__fmul_rn(embed[i], meanstd[D + d]), meanstd[d]);   // embed.mul_(std).add_(mean)
        embed[i] = e;
        embed_avg[i] = e;
    }
    if (i < K) cluster_size[i] = __fadd_rn(cluster_size[i], add_cluster);
    if (i == 0) *first_pass = 0;
}

__global__ void vq_embed_code_kernel(const int64_t *__restrict__ idx, const float *__restrict__ embed, int64_t n, int D,
                                     int K, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * D) {
        const int64_t row = i / D;
        const int d = (int)(i - row * D);
        int64_t k = idx[row];
        k = k < 0 ? 0 : (k >= K ? K - 1 : k);
        out[i] = embed[k * D + d];
    }
}

__global__ void vq_backward_kernel(const float *__restrict__ gq, const float *__restrict__ gloss, const float *__restrict__ x,
                                   const float *__restrict__ q, int64_t numel, float coef, float *__restrict__ gx) {
    const float c = coef * __ldg(gloss);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = gq ? gq[i] : 0.0f;
        gx[i] = __fmaf_rn(c, x[i] - q[i], g);
    }
}

template <int DT>
static int launch_assign(const float *x, const float *embed, int64_t B, int D, int64_t S, int K, float *quant,
                         int64_t *idx, double *sqerr, float *counts, float *dw, void *stream) {
    int kchunk = kVqCodebookSmemBytes / (4 * D);
    if (kchunk < 1) kchunk = 1;
    if (kchunk > K) kchunk = K;
    const int64_t tiles_per_batch = ceil_div(S, kVqThreads);
    const int64_t total_tiles = tiles_per_batch * B;
    const size_t stats_bytes = (size_t)K * (D + 1) * 4;
    const int smem_stats = (counts != nullptr && stats_bytes <= (size_t)kVqStatsSmemBytes) ? 1 : 0;
    size_t smem = (size_t)kchunk * D * 4 + (DT > 0 ? 0 : (size_t)D * kVqThreads * 4) + (smem_stats ? stats_bytes : 0);
    int64_t grid = total_tiles < (int64_t)kNumSMs * 4 ? total_tiles : (int64_t)kNumSMs * 4;
    return launch("vq_assign", vq_assign_kernel<DT>, dim3((unsigned)grid), dim3(kVqThreads), smem, stream, x, embed, S, D, K,
                  kchunk, tiles_per_batch, total_tiles, smem_stats, quant, idx, sqerr, counts, dw);
}

}  // namespace vq3d

using namespace vq3d;

extern "C" int vq3d_vq_assign(const float *x, const float *embed, int64_t B, int D, int64_t S, int K, float *quant,
                              int64_t *idx, double *sqerr, float *counts, float *dw, void *stream) {
    if (!x || !embed || !quant || !idx) return fail(VQ3D_ERR_INVALID, "vq_assign: null pointer");
    if (B < 0 || S < 0 || D < 1 || K < 1 || D > 4096) return fail(VQ3D_ERR_INVALID, "vq_assign: bad sizes B=%lld D=%d S=%lld K=%d", (long long)B, D, (long long)S, K);
    if ((counts == nullptr) != (dw == nullptr)) return fail(VQ3D_ERR_INVALID, "vq_assign: counts and dw must both be given or both NULL");
    if (sqerr != nullptr) {      // the accumulator is zeroed here (a memset node on the stream), not by the caller
        const int rc = check_cuda(cudaMemsetAsync(sqerr, 0, sizeof(double), static_cast<cudaStream_t>(stream)), "vq_assign(memset)");
        if (rc != VQ3D_OK) return rc;
    }
    if (B == 0 || S == 0) return VQ3D_OK;
    switch (D) {
        case 1: return launch_assign<1>(x, embed, B, D, S, K, quant, idx, sqerr, counts, dw, stream);
        case 2: return launch_assign<2>(x, embed, B, D, S, K, quant, idx, sqerr, counts, dw, stream);
        case 4: return launch_assign<4>(x, embed, B, D, S, K, quant, idx, sqerr, counts, dw, stream);
        case 8: return launch_assign<8>(x, embed, B, D, S, K, quant, idx, sqerr, counts, dw, stream);
        case 16: return launch_assign<16>(x, embed, B, D, S, K, quant, idx, sqerr, counts, dw, stream);
        case 32: return launch_assign<32>(x, embed, B, D, S, K, quant, idx, sqerr, counts, dw, stream);
        case 64: return launch_assign<64>(x, embed, B, D, S, K, quant, idx, sqerr, counts, dw, stream);
        default:
            if (D > 192) return fail(VQ3D_ERR_UNSUPPORTED, "vq_assign: embedding_dim %d > 192 not supported", D);
            return launch_assign<0>(x, embed, B, D, S, K, quant, idx, sqerr, counts, dw, stream);
    }
}

extern "C" int vq3d_vq_loss(const double *sqerr, double commitment_cost, int64_t numel, float *loss, void *stream) {
    if (!sqerr || !loss || numel <= 0) return fail(VQ3D_ERR_INVALID, "vq_loss: bad arguments");
    return launch("vq_loss", vq_loss_kernel, dim3(1), dim3(32), 0, stream, sqerr, (float)commitment_cost, 1.0 / (double)numel, loss);
}

extern "C" int vq3d_vq_ema_update(const float *counts, const float *dw, int K, int D, double decay, double laplace_alpha,
                                  float *cluster_size, float *embed_avg, float *embed, void *stream) {
    if (!counts || !dw || !cluster_size || !embed_avg || !embed || K < 1 || D < 1) return fail(VQ3D_ERR_INVALID, "vq_ema_update: bad arguments");
    return launch("vq_ema_update", vq_ema_update_kernel, dim3(1), dim3(1024), 0, stream, counts, dw, K, D, (float)decay,
                  (float)(1.0 - decay), (float)laplace_alpha, (float)(K * laplace_alpha), cluster_size, embed_avg, embed);
}

extern "C" int vq3d_vq_init_stats(const float *x, int64_t B, int D, int64_t S, double *scratch, float *meanstd, void *stream) {
    if (!x || !scratch || !meanstd || D < 1 || D > 65535 || B < 1 || B > 65535 || S < 1) return fail(VQ3D_ERR_INVALID, "vq_init_stats: bad arguments");
    int rc = check_cuda(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * D, (cudaStream_t)stream), "vq_init_stats memset");
    if (rc) return rc;
    int64_t chunks = ceil_div(S, 256 * 8);
    if (chunks > 1024) chunks = 1024;
    rc = launch("vq_init_sums", vq_init_sums_kernel, dim3((unsigned)chunks, (unsigned)D, (unsigned)B), dim3(256), 0, stream, x, D, S, scratch);
    if (rc) return rc;
    return launch("vq_init_finish", vq_init_finish_kernel, dim3((unsigned)ceil_div(D, 128)), dim3(128), 0, stream,
                  (const double *)scratch, D, (double)B * (double)S, meanstd);
}

extern "C" int vq3d_vq_init_apply(const float *meanstd, int K, int D, double total_vectors, float *embed, float *embed_avg,
                                  float *cluster_size, int64_t *first_pass, void *stream) {
    if (!meanstd || !embed || !embed_avg || !cluster_size || !first_pass || K < 1 || D < 1) return fail(VQ3D_ERR_INVALID, "vq_init_apply: bad arguments");
    const int n = K * D > K ? K * D : K;
    return launch("vq_init_apply", vq_init_apply_kernel, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, stream, meanstd, K, D,
                  (float)(total_vectors / (double)K), embed, embed_avg, cluster_size, first_pass);
}

extern "C" int vq3d_vq_embed_code(const int64_t *idx, const float *embed, int64_t n, int D, int K, float *out, void *stream) {
    if (!idx || !embed || !out || D < 1 || K < 1 || n < 0) return fail(VQ3D_ERR_INVALID, "vq_embed_code: bad arguments");
    if (n == 0) return VQ3D_OK;
    return launch("vq_embed_code", vq_embed_code_kernel, dim3((unsigned)ceil_div(n * D, 256)), dim3(256), 0, stream, idx, embed, n, D, K, out);
}

extern "C" int vq3d_vq_backward(const float *grad_quant, const float *grad_loss, const float *x, const float *quant,
                                int64_t numel, double commitment_cost, float *grad_x, void *stream) {
    if (!grad_loss || !x || !quant || !grad_x || numel < 0) return fail(VQ3D_ERR_INVALID, "vq_backward: bad arguments");
    if (numel == 0) return VQ3D_OK;
    int64_t blocks = ceil_div(numel, 256 * 4);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    return launch("vq_backward", vq_backward_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, grad_quant, grad_loss, x, quant,
                  numel, (float)(2.0 * commitment_cost / (double)numel), grad_x);
}
