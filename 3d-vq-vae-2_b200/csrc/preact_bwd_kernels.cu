// preact_bwd_kernels.cu -- fused backward of a 'same' PreActFixupResBlock without skip convolution
// (vqvae/layers.py:176-195 under autograd; the blocks of the 50 / 150-deep stacks, layers.py:492-494,566-569):
//
//   forward   t1 = x + b1a        a1 = ELU(t1) + b1b      c1 = W1 a1                      (1x1, C -> Cb)
//             t2 = c1 + b2a       a2 = ELU(t2) + b2b      c2 = conv3x3x3_circular(W2, a2)
//             t3 = c2 + b3a       a3 = ELU(t3) + b3b      c3 = W3 a3                      (1x1, Cb -> C)
//             y  = c3 * scale + b4 + x
//   backward  d b4 = sum gy       d scale = sum gy c3     g3 = gy * scale
//             d W3 = sum g3 (x) a3            ga3 = W3^T g3          d b3b = sum ga3
//             gt3 = ga3 ELU'(t3)              d b3a = sum gt3
//             d W2[co,ci,k] = sum_v gt3[co](v) a2[ci](v + k - 1)     ga2[ci](u) = sum_{co,k} W2[co,ci,k] gt3[co](u - k + 1)
//             d b2b = sum ga2     gt2 = ga2 ELU'(t2)      d b2a = sum gt2
//             d W1 = sum gt2 (x) a1           ga1 = W1^T gt2         d b1b = sum ga1
//             gt1 = ga1 ELU'(t1)              d b1a = sum gt1        gx = gt1 + gy
//
// The training forward saves ONLY the block input x (one fused forward launch per block); the backward recomputes the
// intermediates inside two tiled kernels that exchange gt3 and c1 (Cb channels each) through a workspace:
//
//   kernel A  a2 on the tile + 1 halo (pointwise recompute) -> c2 -> a3, ELU'(t3); with gy: c3, d scale, d b4, ga3, gt3 -> workspace;
//             d W3 as a [C x T] x [T x Cb] product out of the staged tile; d b3a, d b3b; c1 -> workspace
//   kernel B  gt3 on the tile + 1 halo -> ga2 (transposed convolution) -> gt2 (c1 from the workspace) -> ga1 -> gx = gt1 + gy;
//             d W1 from the staged tile; d b2a, d b2b, d b1a, d b1b
//   d W2      the tiled weight-gradient kernel of backward_kernels.cu on (c1 with its pre-activation, gt3)
//
// = 3 launches (+ 1 forward) per block instead of 9 of the composed path.  fp32 SIMT (exact against autograd to summation
// order); C <= 32, Cb <= 16.
#include "vq3d_rt.h"

namespace vq3d {

constexpr int kPbThreads = 256;
constexpr int kPbMaxCb = 16;

struct PbParams {
    int B, H, W, Z, C, Cb;
    int TH, TW, TZ, T;                     // tile of output voxels (T = TH * TW * TZ <= 512)
    int nth, ntw, ntz;
    const float *x, *gy, *w1, *w2, *w3;
    const float *b1a, *b1b, *b2a, *b2b, *b3a, *b3b, *b4, *scale;
    float *c1ws, *g3ws;                    // [B][Cb][S] each
    float *gx, *gw1, *gw3, *gscal;         // gscal[8]: d b1a, b1b, b2a, b2b, b3a, b3b, b4, scale (accumulated)
};

__device__ __forceinline__ float pb_block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;    // valid in thread 0
}

__device__ __forceinline__ int pb_wrap(int i, int n) {       // any i >= -n
    i %= n;
    return i < 0 ? i + n : i;
}

__device__ __forceinline__ void pb_tile(const PbParams &p, int &b, int &h0, int &w0, int &z0) {
    int t = blockIdx.x;
    const int tz = t % p.ntz; t /= p.ntz;
    const int tw = t % p.ntw; t /= p.ntw;
    const int th = t % p.nth;
    b = t / p.nth;
    h0 = th * p.TH; w0 = tw * p.TW; z0 = tz * p.TZ;
}

// pair phase shared by both kernels: out[r * ncol + q] += sum_tv sRow[r][tv] * sCol[q][tv] * mul  (one atomicAdd per pair and CTA)
__device__ __forceinline__ void pb_pairs(const float *sRow, int nrow, const float *sCol, int ncol, int T, int ld, float mul, float *out) {
    for (int pr = threadIdx.x; pr < nrow * ncol; pr += kPbThreads) {
        const float *a = sRow + (pr / ncol) * ld, *c = sCol + (pr % ncol) * ld;
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;      // four independent chains
        int tv = 0;
        for (; tv + 3 < T; tv += 4) {
            s0 = __fmaf_rn(a[tv], c[tv], s0); s1 = __fmaf_rn(a[tv + 1], c[tv + 1], s1);
            s2 = __fmaf_rn(a[tv + 2], c[tv + 2], s2); s3 = __fmaf_rn(a[tv + 3], c[tv + 3], s3);
        }
        for (; tv < T; ++tv) s0 = __fmaf_rn(a[tv], c[tv], s0);
        atomicAdd(out + pr, ((s0 + s1) + (s2 + s3)) * mul);
    }
}

// NZ: depth neighbours per thread in the convolution loops (2 when the tile depth is even: the window values and every weight
// vector are loaded once for both voxels -- 13 shared loads per 54 FMAs at C_b = 9 instead of 13 per 27)
template <int NZ>
__global__ void __launch_bounds__(kPbThreads, 2)
preact_same_bwd_a_kernel(PbParams p) {
    VQ3D_DYN_SMEM(float, sm);
    __shared__ float red[32];
    const int C = p.C, Cb = p.Cb, Cb4 = (Cb + 3) & ~3, T = p.T, LD = p.T + 1;
    const int IH = p.TH + 2, IW = p.TW + 2, IZ = p.TZ + 2, R1 = IH * IW * IZ;
    float *sW1 = sm;                              // [Cb][C]
    float *sW2 = sW1 + Cb * C;                    // [ci][tap][Cb4]  (output channels contiguous)
    float *sW3 = sW2 + Cb * 27 * Cb4;             // [C][Cb]
    float *sA2 = sW3 + C * Cb;                    // [Cb][R1]
    float *sA3 = sA2 + Cb * R1;                   // [Cb][LD]
    float *sG3 = sA3 + Cb * LD;                   // [C][LD]
    const int tid = threadIdx.x;
    for (int i = tid; i < Cb * C; i += kPbThreads) { sW1[i] = __ldg(p.w1 + i); sW3[i] = __ldg(p.w3 + i); }
    for (int i = tid; i < Cb * 27 * Cb4; i += kPbThreads) {
        const int co = i % Cb4, tap = (i / Cb4) % 27, ci = i / (Cb4 * 27);
        sW2[i] = co < Cb ? __ldg(p.w2 + ((size_t)co * Cb + ci) * 27 + tap) : 0.0f;
    }
    const float b1a = ld_scalar(p.b1a, 0.f), b1b = ld_scalar(p.b1b, 0.f), b2a = ld_scalar(p.b2a, 0.f), b2b = ld_scalar(p.b2b, 0.f);
    const float b3a = ld_scalar(p.b3a, 0.f), b3b = ld_scalar(p.b3b, 0.f), sc = ld_scalar(p.scale, 1.f);
    int b, h0, w0, z0;
    pb_tile(p, b, h0, w0, z0);
    const size_t S = (size_t)p.H * p.W * p.Z;
    const float *xb = p.x + (size_t)b * C * S;
    __syncthreads();
    // ---- a2 on the tile + halo 1 (c1 of the tile's own voxels goes to the workspace) ----
    for (int r = tid; r < R1; r += kPbThreads) {
        const int lz = r % IZ, lw = (r / IZ) % IW, lh = r / (IZ * IW);
        const int gh = pb_wrap(h0 + lh - 1, p.H), gw = pb_wrap(w0 + lw - 1, p.W), gz = pb_wrap(z0 + lz - 1, p.Z);
        const size_t off = ((size_t)gh * p.W + gw) * p.Z + gz;
        float c1[kPbMaxCb];
#pragma unroll
        for (int j = 0; j < kPbMaxCb; ++j) c1[j] = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float t1 = __ldg(xb + (size_t)c * S + off) + b1a;
            const float a1 = (t1 > 0.0f ? t1 : __expf(t1) - 1.0f) + b1b;
#pragma unroll
            for (int j = 0; j < kPbMaxCb; ++j)
                if (j < Cb) c1[j] = __fmaf_rn(sW1[j * C + c], a1, c1[j]);
        }
        const bool own = lh >= 1 && lh <= p.TH && lw >= 1 && lw <= p.TW && lz >= 1 && lz <= p.TZ &&
                         h0 + lh - 1 < p.H && w0 + lw - 1 < p.W && z0 + lz - 1 < p.Z;
#pragma unroll
        for (int j = 0; j < kPbMaxCb; ++j)
            if (j < Cb) {
                const float t2 = c1[j] + b2a;
                sA2[j * R1 + r] = (t2 > 0.0f ? t2 : __expf(t2) - 1.0f) + b2b;
                if (own) p.c1ws[((size_t)b * Cb + j) * S + off] = c1[j];
            }
    }
    __syncthreads();
    // ---- the tile's voxels: c2 -> a3, ELU'(t3); c3, ga3, gt3 ----
    float s_scale = 0.0f, s_b4 = 0.0f, s_b3b = 0.0f, s_b3a = 0.0f;
    for (int tq = tid; tq * NZ < T; tq += kPbThreads) {
        const int tv0 = tq * NZ;
        const int dz = tv0 % p.TZ, dw = (tv0 / p.TZ) % p.TW, dh = tv0 / (p.TZ * p.TW);
        const int oh = h0 + dh, ow = w0 + dw;
        float c2[NZ][kPbMaxCb];
#pragma unroll
        for (int v = 0; v < NZ; ++v)
#pragma unroll
            for (int j = 0; j < kPbMaxCb; ++j) c2[v][j] = 0.0f;
        if (oh < p.H && ow < p.W && z0 + dz < p.Z) {
            for (int ci = 0; ci < Cb; ++ci) {
                const float *a = sA2 + ci * R1 + (dh * IW + dw) * IZ + dz;       // window origin = tap (0,0,0) of voxel 0
#pragma unroll
                for (int hw = 0; hw < 9; ++hw) {
                    const float *ar = a + ((hw / 3) * IW + hw % 3) * IZ;
                    float av[NZ + 2];
#pragma unroll
                    for (int t = 0; t < NZ + 2; ++t) av[t] = ar[t];
#pragma unroll
                    for (int kz = 0; kz < 3; ++kz) {
                        const float *wr = sW2 + (ci * 27 + hw * 3 + kz) * Cb4;
#pragma unroll
                        for (int j = 0; j < kPbMaxCb; ++j)
                            if (j < Cb) {
                                const float wv = wr[j];
#pragma unroll
                                for (int v = 0; v < NZ; ++v) c2[v][j] = __fmaf_rn(wv, av[v + kz], c2[v][j]);
                            }
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < NZ; ++v) {
            const int tv = tv0 + v, oz = z0 + dz + v;
            const bool active = oh < p.H && ow < p.W && oz < p.Z;
            float a3[kPbMaxCb], ga3[kPbMaxCb];
#pragma unroll
            for (int j = 0; j < kPbMaxCb; ++j) {
                const float t3 = c2[v][j] + b3a;
                const float e = __expf(t3);
                a3[j] = (active && j < Cb) ? (t3 > 0.0f ? t3 : e - 1.0f) + b3b : 0.0f;
                c2[v][j] = t3 > 0.0f ? 1.0f : e;           // from here on: ELU'(t3)
                ga3[j] = 0.0f;
                if (j < Cb) sA3[j * LD + tv] = a3[j];
            }
            const size_t off = active ? ((size_t)oh * p.W + ow) * p.Z + oz : 0;
            for (int c = 0; c < C; ++c) {
                const float g = active ? __ldg(p.gy + ((size_t)b * C + c) * S + off) : 0.0f;
                float c3 = 0.0f;
#pragma unroll
                for (int j = 0; j < kPbMaxCb; ++j)
                    if (j < Cb) c3 = __fmaf_rn(sW3[c * Cb + j], a3[j], c3);
                s_scale = __fmaf_rn(g, c3, s_scale);
                s_b4 += g;
                const float g3 = g * sc;
                sG3[c * LD + tv] = g3;
#pragma unroll
                for (int j = 0; j < kPbMaxCb; ++j)
                    if (j < Cb) ga3[j] = __fmaf_rn(sW3[c * Cb + j], g3, ga3[j]);
            }
            if (active) {
#pragma unroll
                for (int j = 0; j < kPbMaxCb; ++j)
                    if (j < Cb) {
                        const float gt3 = ga3[j] * c2[v][j];
                        s_b3b += ga3[j];
                        s_b3a += gt3;
                        p.g3ws[((size_t)b * Cb + j) * S + off] = gt3;
                    }
            }
        }
    }
    __syncthreads();
    if (p.gw3) pb_pairs(sG3, C, sA3, Cb, T, LD, 1.0f, p.gw3);                  // d W3 [C][Cb]
    if (p.gscal) {
        const float t0 = pb_block_sum(s_b3a, red), t1 = pb_block_sum(s_b3b, red), t2 = pb_block_sum(s_b4, red), t3 = pb_block_sum(s_scale, red);
        if (tid == 0) {
            atomicAdd(p.gscal + 4, t0); atomicAdd(p.gscal + 5, t1); atomicAdd(p.gscal + 6, t2); atomicAdd(p.gscal + 7, t3);
        }
    }
}

template <int NZ>
__global__ void __launch_bounds__(kPbThreads, 2)
preact_same_bwd_b_kernel(PbParams p) {
    VQ3D_DYN_SMEM(float, sm);
    __shared__ float red[32];
    const int C = p.C, Cb = p.Cb, Cb4 = (Cb + 3) & ~3, T = p.T, LD = p.T + 1;
    const int IH = p.TH + 2, IW = p.TW + 2, IZ = p.TZ + 2, R1 = IH * IW * IZ;
    float *sW1 = sm;                              // [Cb][C]
    float *sW2 = sW1 + Cb * C;                    // [co][tap][Cb4]  (input channels contiguous)
    float *sG = sW2 + Cb * 27 * Cb4;              // [Cb][R1]  gt3 on the tile + halo 1
    float *sG1 = sG + Cb * R1;                    // [Cb][LD]  gt2
    float *sA1 = sG1 + Cb * LD;                   // [C][LD]   a1
    const int tid = threadIdx.x;
    for (int i = tid; i < Cb * C; i += kPbThreads) sW1[i] = __ldg(p.w1 + i);
    for (int i = tid; i < Cb * 27 * Cb4; i += kPbThreads) {
        const int ci = i % Cb4, tap = (i / Cb4) % 27, co = i / (Cb4 * 27);
        sW2[i] = ci < Cb ? __ldg(p.w2 + ((size_t)co * Cb + ci) * 27 + tap) : 0.0f;
    }
    const float b1a = ld_scalar(p.b1a, 0.f), b1b = ld_scalar(p.b1b, 0.f), b2a = ld_scalar(p.b2a, 0.f);
    int b, h0, w0, z0;
    pb_tile(p, b, h0, w0, z0);
    const size_t S = (size_t)p.H * p.W * p.Z;
    for (int i = tid; i < Cb * R1; i += kPbThreads) {
        const int r = i % R1, j = i / R1;
        const int lz = r % IZ, lw = (r / IZ) % IW, lh = r / (IZ * IW);
        const int gh = pb_wrap(h0 + lh - 1, p.H), gw = pb_wrap(w0 + lw - 1, p.W), gz = pb_wrap(z0 + lz - 1, p.Z);
        sG[i] = __ldg(p.g3ws + ((size_t)b * Cb + j) * S + ((size_t)gh * p.W + gw) * p.Z + gz);
    }
    __syncthreads();
    float s_b2b = 0.0f, s_b2a = 0.0f, s_b1b = 0.0f, s_b1a = 0.0f;
    for (int tq = tid; tq * NZ < T; tq += kPbThreads) {
        const int tv0 = tq * NZ;
        const int dz = tv0 % p.TZ, dw = (tv0 / p.TZ) % p.TW, dh = tv0 / (p.TZ * p.TW);
        const int oh = h0 + dh, ow = w0 + dw;
        float ga2[NZ][kPbMaxCb];
#pragma unroll
        for (int v = 0; v < NZ; ++v)
#pragma unroll
            for (int j = 0; j < kPbMaxCb; ++j) ga2[v][j] = 0.0f;
        if (oh < p.H && ow < p.W && z0 + dz < p.Z) {
            for (int co = 0; co < Cb; ++co) {
                // ga2(u) += W2[co][.][k] gt3[co](u - k + 1): box coordinate of u - k + 1 is (d + 1) - (k - 1) = d + 2 - k
                const float *g = sG + co * R1 + ((dh + 2) * IW + (dw + 2)) * IZ + (dz + 2);
#pragma unroll
                for (int hw = 0; hw < 9; ++hw) {
                    const float *gr = g - ((hw / 3) * IW + hw % 3) * IZ;
                    float gv[NZ + 2];                      // gv[t] = gt3 at depth offset t - 2 (t = 0 .. NZ + 1)
#pragma unroll
                    for (int t = 0; t < NZ + 2; ++t) gv[t] = gr[t - 2];
#pragma unroll
                    for (int kz = 0; kz < 3; ++kz) {
                        const float *wr = sW2 + (co * 27 + hw * 3 + kz) * Cb4;
#pragma unroll
                        for (int j = 0; j < kPbMaxCb; ++j)
                            if (j < Cb) {
                                const float wv = wr[j];
#pragma unroll
                                for (int v = 0; v < NZ; ++v) ga2[v][j] = __fmaf_rn(wv, gv[v + 2 - kz], ga2[v][j]);
                            }
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < NZ; ++v) {
            const int tv = tv0 + v, oz = z0 + dz + v;
            const bool active = oh < p.H && ow < p.W && oz < p.Z;
            const size_t off = active ? ((size_t)oh * p.W + ow) * p.Z + oz : 0;
            float gt2[kPbMaxCb];
#pragma unroll
            for (int j = 0; j < kPbMaxCb; ++j) {
                gt2[j] = 0.0f;
                if (j < Cb) {
                    if (active) {
                        const float t2 = __ldg(p.c1ws + ((size_t)b * Cb + j) * S + off) + b2a;
                        gt2[j] = ga2[v][j] * (t2 > 0.0f ? 1.0f : __expf(t2));
                        s_b2b += ga2[v][j];
                        s_b2a += gt2[j];
                    }
                    sG1[j * LD + tv] = gt2[j];
                }
            }
            for (int c = 0; c < C; ++c) {
                float a1 = 0.0f;
                if (active) {
                    const float t1 = __ldg(p.x + ((size_t)b * C + c) * S + off) + b1a;
                    const float e = __expf(t1);
                    a1 = (t1 > 0.0f ? t1 : e - 1.0f) + b1b;
                    float ga1 = 0.0f;
#pragma unroll
                    for (int j = 0; j < kPbMaxCb; ++j)
                        if (j < Cb) ga1 = __fmaf_rn(sW1[j * C + c], gt2[j], ga1);
                    const float gt1 = ga1 * (t1 > 0.0f ? 1.0f : e);
                    s_b1b += ga1;
                    s_b1a += gt1;
                    if (p.gx) p.gx[((size_t)b * C + c) * S + off] = gt1 + __ldg(p.gy + ((size_t)b * C + c) * S + off);
                }
                sA1[c * LD + tv] = a1;
            }
        }
    }
    __syncthreads();
    if (p.gw1) pb_pairs(sG1, Cb, sA1, C, T, LD, 1.0f, p.gw1);                  // d W1 [Cb][C]
    if (p.gscal) {
        const float t0 = pb_block_sum(s_b1a, red), t1 = pb_block_sum(s_b1b, red), t2 = pb_block_sum(s_b2a, red), t3 = pb_block_sum(s_b2b, red);
        if (tid == 0) {
            atomicAdd(p.gscal + 0, t0); atomicAdd(p.gscal + 1, t1); atomicAdd(p.gscal + 2, t2); atomicAdd(p.gscal + 3, t3);
        }
    }
}

static bool pb_plan(const vq3d_preact_desc *d, PbParams &p, size_t &smem_a, size_t &smem_b) {
    const int C = d->Cin, Cb = d->Cb, Cb4 = (Cb + 3) & ~3;
    int T = (C + Cb) <= 27 ? 512 : 256;
    p.TZ = d->Z < 16 ? d->Z : 16;
    p.TW = d->W < 8 ? d->W : 8;
    int th = T / (p.TZ * p.TW);
    if (th < 1) th = 1;
    p.TH = d->H < th ? d->H : th;
    p.T = p.TH * p.TW * p.TZ;
    p.nth = (int)ceil_div(d->H, p.TH); p.ntw = (int)ceil_div(d->W, p.TW); p.ntz = (int)ceil_div(d->Z, p.TZ);
    const size_t R1 = (size_t)(p.TH + 2) * (p.TW + 2) * (p.TZ + 2), LD = (size_t)p.T + 1;
    smem_a = ((size_t)2 * Cb * C + (size_t)Cb * 27 * Cb4 + (size_t)Cb * R1 + (size_t)(Cb + C) * LD) * sizeof(float);
    smem_b = ((size_t)Cb * C + (size_t)Cb * 27 * Cb4 + (size_t)Cb * R1 + (size_t)(Cb + C) * LD) * sizeof(float);
    return smem_a <= 200 * 1024 && smem_b <= 200 * 1024;
}

}  // namespace vq3d

using namespace vq3d;

extern "C" size_t vq3d_preact_same_backward_workspace(const vq3d_preact_desc *d) {
    if (!d || d->B < 1 || d->H < 1 || d->W < 1 || d->Z < 1 || d->Cb < 1 || d->Cb > kPbMaxCb || d->Cin < 1 || d->Cin > 32 || d->Cin != d->Cout ||
        d->mode != 0 || d->wskip)
        return 0;
    PbParams p;
    size_t sa, sb;
    if (!pb_plan(d, p, sa, sb)) return 0;
    return (size_t)2 * d->B * d->Cb * d->H * d->W * d->Z * sizeof(float);
}

extern "C" int vq3d_preact_same_backward(const vq3d_preact_desc *d, const float *gy, void *ws, size_t ws_bytes, float *gx, float *gw1, float *gw2,
                                         float *gw3, float *gscalars, void *stream) {
    if (!d || !gy || !ws) return fail(VQ3D_ERR_INVALID, "preact_same_backward: null descriptor / gy / workspace");
    if (d->mode != 0 || d->wskip || d->Cin != d->Cout) return fail(VQ3D_ERR_UNSUPPORTED, "preact_same_backward: 'same' blocks without skip convolution only");
    if (d->Cb > kPbMaxCb || d->Cin > 32) return fail(VQ3D_ERR_UNSUPPORTED, "preact_same_backward: C = %d, Cb = %d exceed the fused kernels (32 / 16)", d->Cin, d->Cb);
    if (!d->x || !d->w1 || !d->w2 || !d->w3) return fail(VQ3D_ERR_INVALID, "preact_same_backward: null tensor");
    if (d->out_w || d->pre_w) return fail(VQ3D_ERR_UNSUPPORTED, "preact_same_backward: fused leading / trailing 1x1 convolutions are not differentiated here");
    const size_t need = vq3d_preact_same_backward_workspace(d);
    if (need == 0) return fail(VQ3D_ERR_UNSUPPORTED, "preact_same_backward: no tile fits shared memory");
    if (ws_bytes < need) return fail(VQ3D_ERR_INVALID, "preact_same_backward: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    PbParams p;
    memset(&p, 0, sizeof(p));
    size_t sa, sb;
    pb_plan(d, p, sa, sb);
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z; p.C = d->Cin; p.Cb = d->Cb;
    p.x = d->x; p.gy = gy; p.w1 = d->w1; p.w2 = d->w2; p.w3 = d->w3;
    p.b1a = d->b1a; p.b1b = d->b1b; p.b2a = d->b2a; p.b2b = d->b2b; p.b3a = d->b3a; p.b3b = d->b3b; p.b4 = d->b4; p.scale = d->scale;
    const size_t S = (size_t)d->H * d->W * d->Z;
    p.c1ws = static_cast<float *>(ws);
    p.g3ws = p.c1ws + (size_t)d->B * d->Cb * S;
    p.gx = gx; p.gw1 = gw1; p.gw3 = gw3; p.gscal = gscalars;
    const dim3 grid((unsigned)((int64_t)d->B * p.nth * p.ntw * p.ntz));
    const bool pairz = p.TZ % 2 == 0;        // even tile depth: two depth neighbours per thread
    int rc = pairz ? launch("preact_same_bwd_a", preact_same_bwd_a_kernel<2>, grid, dim3(kPbThreads), sa, stream, p)
                   : launch("preact_same_bwd_a", preact_same_bwd_a_kernel<1>, grid, dim3(kPbThreads), sa, stream, p);
    if (rc) return rc;
    rc = pairz ? launch("preact_same_bwd_b", preact_same_bwd_b_kernel<2>, grid, dim3(kPbThreads), sb, stream, p)
               : launch("preact_same_bwd_b", preact_same_bwd_b_kernel<1>, grid, dim3(kPbThreads), sb, stream, p);
    if (rc) return rc;
    if (gw2) {      // d W2: the tiled weight-gradient kernel on (c1 with its pre-activation, gt3)
        vq3d_conv_desc cd;
        memset(&cd, 0, sizeof(cd));
        cd.B = d->B; cd.H = d->H; cd.W = d->W; cd.Z = d->Z; cd.C1 = d->Cb; cd.C2 = 0; cd.Cout = d->Cb; cd.k = 3; cd.stride = 1; cd.pad = 1;
        cd.pad_circular = 1; cd.pre_act = 1; cd.x1 = p.c1ws; cd.w = d->w2; cd.pre_a = d->b2a; cd.pre_b = d->b2b;
        vq3d_conv_bwd g;
        memset(&g, 0, sizeof(g));
        g.gy = p.g3ws; g.gw = gw2; g.skip_input_grads = 1;
        rc = vq3d_conv3d_backward(&cd, &g, stream);
        if (rc) return rc;
    }
    return VQ3D_OK;
}
