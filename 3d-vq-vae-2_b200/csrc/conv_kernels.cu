// conv_kernels.cu -- direct 3-D convolution, trilinear x2 upsampling and the Huber epilogue.
//
// vq3d_conv3d is the shape-generic path: any (C1+C2 -> Cout, k in 1..4, stride 1|2, circular
// or zero padding) with the Fixup input transform ELU(x+a)+b and the output transform
// *scale + b + bias[co] + residual fused in.  It covers every nn.Conv3d call site of
// vqvae/layers.py (134-171, 249-274, 377, 490, 508, 535) including the torch.cat + 1x1
// `proj` (385, 512) without materialising the concatenation.  Layers that dominate time get
// dedicated fused kernels (preact_kernels.cu); this one is what the rest falls back to.
//
// Mapping: one thread = one output voxel x CO_T output channels; blockIdx.y walks C_out in
// chunks of CO_T.  Weights of the chunk are staged transposed in shared memory
// ([ci][tap][CO_T]) so that the inner product reads them as 128-bit warp broadcasts.
#include "vq3d_rt.h"

namespace vq3d {

constexpr int kConvThreads = 128;
constexpr int kConvWeightFloats = 10240;   // 40 KB weight stage

struct ConvParams {
    int B, H, W, Z, C1, C2, Cout, k, stride, pad, circ, pre_act, post_act;
    int Ho, Wo, Zo;
    const float *x1, *x2, *w, *bias, *pre_a, *pre_b, *post_scale, *post_b, *residual;
    float *y;
};

template <int CO_T>
__global__ void __launch_bounds__(kConvThreads)
conv3d_generic_kernel(ConvParams p) {
    __shared__ __align__(16) float s_w[kConvWeightFloats];
    const int Cin = p.C1 + p.C2;
    const int k = p.k, k3 = k * k * k;
    const int64_t S = (int64_t)p.H * p.W * p.Z;
    const int64_t So = (int64_t)p.Ho * p.Wo * p.Zo;
    const int64_t total = (int64_t)p.B * So;
    const int64_t v = (int64_t)blockIdx.x * kConvThreads + threadIdx.x;
    const bool active = v < total;
    const int co0 = blockIdx.y * CO_T;

    int b = 0, oh = 0, ow = 0, oz = 0;
    if (active) {
        b = (int)(v / So);
        int64_t r = v - (int64_t)b * So;
        oh = (int)(r / ((int64_t)p.Wo * p.Zo));
        r -= (int64_t)oh * p.Wo * p.Zo;
        ow = (int)(r / p.Zo);
        oz = (int)(r - (int64_t)ow * p.Zo);
    }
    const float pa = ld_scalar(p.pre_a, 0.0f), pb = ld_scalar(p.pre_b, 0.0f);

    float acc[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[j] = 0.0f;

    int cic = kConvWeightFloats / (k3 * CO_T);
    if (cic > Cin) cic = Cin;
    for (int ci0 = 0; ci0 < Cin; ci0 += cic) {
        const int nci = min(cic, Cin - ci0);
        __syncthreads();
        for (int i = threadIdx.x; i < nci * k3 * CO_T; i += kConvThreads) {
            const int j = i % CO_T;
            const int t = (i / CO_T) % k3;
            const int cl = i / (CO_T * k3);
            const int co = co0 + j;
            s_w[i] = co < p.Cout ? p.w[((size_t)co * Cin + (ci0 + cl)) * k3 + t] : 0.0f;
        }
        __syncthreads();
        if (active) {
            for (int cl = 0; cl < nci; ++cl) {
                const int ci = ci0 + cl;
                const float *src = ci < p.C1 ? p.x1 + ((size_t)b * p.C1 + ci) * S
                                             : p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S;
                const float *wp = s_w + (size_t)cl * k3 * CO_T;
                for (int kh = 0; kh < k; ++kh) {
                    int ih = oh * p.stride - p.pad + kh;
                    if (p.circ) ih = wrap(ih, p.H);
                    else if (ih < 0 || ih >= p.H) continue;
                    for (int kw = 0; kw < k; ++kw) {
                        int iw = ow * p.stride - p.pad + kw;
                        if (p.circ) iw = wrap(iw, p.W);
                        else if (iw < 0 || iw >= p.W) continue;
                        const float *row = src + ((size_t)ih * p.W + iw) * p.Z;
                        for (int kz = 0; kz < k; ++kz) {
                            int iz = oz * p.stride - p.pad + kz;
                            if (p.circ) iz = wrap(iz, p.Z);
                            else if (iz < 0 || iz >= p.Z) continue;
                            float xv = __ldg(row + iz);
                            xv = p.pre_act ? elu1(xv + pa) + pb : xv + pb;
                            const float *wt = wp + ((kh * k + kw) * k + kz) * CO_T;
#pragma unroll
                            for (int j = 0; j < CO_T; ++j) acc[j] = __fmaf_rn(wt[j], xv, acc[j]);
                        }
                    }
                }
            }
        }
    }
    if (active) {
        const float sc = ld_scalar(p.post_scale, 1.0f), sb = ld_scalar(p.post_b, 0.0f);
        const int64_t r = v - (int64_t)b * So;
#pragma unroll
        for (int j = 0; j < CO_T; ++j) {
            const int co = co0 + j;
            if (co < p.Cout) {
                const size_t o = ((size_t)b * p.Cout + co) * So + r;
                float yv = __fmaf_rn(acc[j], sc, sb);
                if (p.bias) yv += __ldg(p.bias + co);
                if (p.residual) yv += __ldg(p.residual + o);
                if (p.post_act) yv = elu1(yv);
                p.y[o] = yv;
            }
        }
    }
}

// k = 3, stride 1, pad 1 ("same") convolutions with few input channels on big tensors (the composed blocks of the training
// path and the forward-convolution form of their input gradients): a CTA stages the pre-transformed input halo tile of a
// 4 x 4 x 32 output box once (the generic kernel re-applies wrap arithmetic and ELU for each of the 27 taps), then every
// thread computes two output voxels x CO_T channels from shared memory.  Same summation order as the generic kernel
// (ci, kh, kw, kz; zero-padded taps add an exact 0), so the results are bit-identical to it.
constexpr int kCtTH = 4, kCtTW = 4, kCtTZ = 32, kCtThreads = 256, kCtMaxCin = 16;

template <int CO_T>
__global__ void __launch_bounds__(kCtThreads)
conv3d_tiled_kernel(ConvParams p, int tilesH, int tilesW, int tilesZ) {
    VQ3D_DYN_SMEM(float, smem);
    constexpr int HH = kCtTH + 2, HW = kCtTW + 2, HZ = kCtTZ + 2, HV = HH * HW * HZ, TV = kCtTH * kCtTW * kCtTZ;
    constexpr int CO_P = (CO_T + 3) & ~3;                      // weight row padded to whole float4s (128-bit broadcast loads)
    const int Cin = p.C1 + p.C2;
    float *su = smem, *s_w = smem + (((size_t)Cin * HV + 3) & ~(size_t)3);      // s_w: [ci][tap][CO_P], 16-byte aligned
    int tile = blockIdx.x;
    const int tz = tile % tilesZ; tile /= tilesZ;
    const int tw = tile % tilesW; tile /= tilesW;
    const int th = tile % tilesH;
    const int b = tile / tilesH;
    const int h0 = th * kCtTH, w0 = tw * kCtTW, z0 = tz * kCtTZ, co0 = blockIdx.y * CO_T;
    const int64_t S = (int64_t)p.H * p.W * p.Z;
    const float pa = ld_scalar(p.pre_a, 0.0f), pb = ld_scalar(p.pre_b, 0.0f);
    for (int i = threadIdx.x; i < Cin * 27 * CO_P; i += kCtThreads) {
        const int j = i % CO_P, t = (i / CO_P) % 27, ci = i / (CO_P * 27), co = co0 + j;
        s_w[i] = (j < CO_T && co < p.Cout) ? p.w[((size_t)co * Cin + ci) * 27 + t] : 0.0f;
    }
    for (int i = threadIdx.x; i < Cin * HV; i += kCtThreads) {
        const int ci = i / HV;
        int r = i - ci * HV;
        const int hh = r / (HW * HZ); r -= hh * HW * HZ;
        const int ww = r / HZ, zz = r - ww * HZ;
        int ih = h0 - 1 + hh, iw = w0 - 1 + ww, iz = z0 - 1 + zz;
        if (p.circ) {
            ih = ih < 0 ? ih + p.H : (ih >= p.H ? ih - p.H : ih);
            iw = iw < 0 ? iw + p.W : (iw >= p.W ? iw - p.W : iw);
            iz = iz < 0 ? iz + p.Z : (iz >= p.Z ? iz - p.Z : iz);
        }
        float u = 0.0f;
        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W && iz >= 0 && iz < p.Z) {
            const float *src = ci < p.C1 ? p.x1 + ((size_t)b * p.C1 + ci) * S : p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S;
            u = __ldg(src + ((size_t)ih * p.W + iw) * p.Z + iz);
            u = p.pre_act ? elu1(u + pa) + pb : u + pb;
        }
        su[i] = u;
    }
    __syncthreads();
    const float sc = ld_scalar(p.post_scale, 1.0f), sb = ld_scalar(p.post_b, 0.0f);
#pragma unroll 1
    for (int v = threadIdx.x; v < TV; v += kCtThreads) {
        const int dh = v / (kCtTW * kCtTZ), dw = (v / kCtTZ) % kCtTW, dz = v % kCtTZ;
        const int oh = h0 + dh, ow = w0 + dw, oz = z0 + dz;
        float acc[CO_T];
#pragma unroll
        for (int j = 0; j < CO_T; ++j) acc[j] = 0.0f;
        for (int ci = 0; ci < Cin; ++ci) {
            const float *ub = su + (size_t)ci * HV + (dh * HW + dw) * HZ + dz;
            const float *wp = s_w + (size_t)ci * 27 * CO_P;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                    for (int kz = 0; kz < 3; ++kz) {
                        const float xv = ub[(kh * HW + kw) * HZ + kz];
                        const float4 *wt = reinterpret_cast<const float4 *>(wp + ((kh * 3 + kw) * 3 + kz) * CO_P);
#pragma unroll
                        for (int j4 = 0; j4 < CO_P / 4; ++j4) {
                            const float4 w4 = wt[j4];
                            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int l = 0; l < 4; ++l)
                                if (j4 * 4 + l < CO_T) acc[j4 * 4 + l] = __fmaf_rn(wv[l], xv, acc[j4 * 4 + l]);
                        }
                    }
        }
        if (oh < p.H && ow < p.W && oz < p.Z) {
            const int64_t r = ((int64_t)oh * p.W + ow) * p.Z + oz;
#pragma unroll
            for (int j = 0; j < CO_T; ++j) {
                const int co = co0 + j;
                if (co < p.Cout) {
                    const size_t o = ((size_t)b * p.Cout + co) * S + r;
                    float yv = __fmaf_rn(acc[j], sc, sb);
                    if (p.bias) yv += __ldg(p.bias + co);
                    if (p.residual) yv += __ldg(p.residual + o);
                    if (p.post_act) yv = elu1(yv);
                    p.y[o] = yv;
                }
            }
        }
    }
}

template <int CO_T>
static int launch_conv_tiled(const ConvParams &p, void *stream) {
    const int tH = (int)ceil_div(p.H, kCtTH), tW = (int)ceil_div(p.W, kCtTW), tZ = (int)ceil_div(p.Z, kCtTZ);
    constexpr int CO_P = (CO_T + 3) & ~3;
    const size_t smem = ((((size_t)(p.C1 + p.C2) * (kCtTH + 2) * (kCtTW + 2) * (kCtTZ + 2) + 3) & ~(size_t)3) + (size_t)(p.C1 + p.C2) * 27 * CO_P) * sizeof(float);
    return launch("conv3d_tiled", conv3d_tiled_kernel<CO_T>, dim3((unsigned)((int64_t)p.B * tH * tW * tZ), (unsigned)ceil_div(p.Cout, CO_T)),
                  dim3(kCtThreads), smem, stream, p, tH, tW, tZ);
}

// 1x1x1 convolutions on big tensors (parse_input 1 -> 4 at 512^3, the cat + proj 18 -> 18 at 128x128x32,
// layers.py:535,385,512): pure streaming.  A thread owns 4 consecutive voxels (float4 along the contiguous
// depth axis) x CO_T output channels; weights of the chunk sit in shared memory ([ci][CO_T], broadcast reads).
constexpr int kPwThreads = 256;
constexpr int kPwMaxCin = 160;

template <int CO_T>
__global__ void __launch_bounds__(kPwThreads)
pointwise_kernel(ConvParams p) {
    __shared__ __align__(16) float s_w[kPwMaxCin * CO_T];
    const int Cin = p.C1 + p.C2;
    const int64_t S = (int64_t)p.H * p.W * p.Z, S4 = S >> 2;
    const int co0 = blockIdx.y * CO_T;
    for (int i = threadIdx.x; i < Cin * CO_T; i += kPwThreads) {
        const int j = i % CO_T, ci = i / CO_T;
        s_w[i] = co0 + j < p.Cout ? p.w[(size_t)(co0 + j) * Cin + ci] : 0.0f;
    }
    const float pa = ld_scalar(p.pre_a, 0.0f), pb = ld_scalar(p.pre_b, 0.0f);
    const float sc = ld_scalar(p.post_scale, 1.0f), sb = ld_scalar(p.post_b, 0.0f);
    __syncthreads();
    const int64_t total = (int64_t)p.B * S4;
    for (int64_t v = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; v < total; v += (int64_t)gridDim.x * kPwThreads) {
        const int b = (int)(v / S4);
        const int64_t r4 = v - (int64_t)b * S4;
        float acc[CO_T][4];
#pragma unroll
        for (int j = 0; j < CO_T; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
        for (int ci = 0; ci < Cin; ++ci) {
            const float *src = ci < p.C1 ? p.x1 + ((size_t)b * p.C1 + ci) * S : p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S;
            float4 xv = __ldg(reinterpret_cast<const float4 *>(src) + r4);
            if (p.pre_act) { xv.x = elu1(xv.x + pa) + pb; xv.y = elu1(xv.y + pa) + pb; xv.z = elu1(xv.z + pa) + pb; xv.w = elu1(xv.w + pa) + pb; }
            else { xv.x += pb; xv.y += pb; xv.z += pb; xv.w += pb; }
            const float *wt = s_w + ci * CO_T;
#pragma unroll
            for (int j = 0; j < CO_T; ++j) {
                const float w = wt[j];
                acc[j][0] = __fmaf_rn(w, xv.x, acc[j][0]); acc[j][1] = __fmaf_rn(w, xv.y, acc[j][1]);
                acc[j][2] = __fmaf_rn(w, xv.z, acc[j][2]); acc[j][3] = __fmaf_rn(w, xv.w, acc[j][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < CO_T; ++j) {
            const int co = co0 + j;
            if (co < p.Cout) {
                const float bias = p.bias ? __ldg(p.bias + co) : 0.0f;
                const size_t o4 = (((size_t)b * p.Cout + co) * S >> 2) + r4;
                float4 yv;
                yv.x = __fmaf_rn(acc[j][0], sc, sb) + bias; yv.y = __fmaf_rn(acc[j][1], sc, sb) + bias;
                yv.z = __fmaf_rn(acc[j][2], sc, sb) + bias; yv.w = __fmaf_rn(acc[j][3], sc, sb) + bias;
                if (p.residual) {
                    const float4 rv = __ldg(reinterpret_cast<const float4 *>(p.residual) + o4);
                    yv.x += rv.x; yv.y += rv.y; yv.z += rv.z; yv.w += rv.w;
                }
                if (p.post_act) { yv.x = elu1(yv.x); yv.y = elu1(yv.y); yv.z = elu1(yv.z); yv.w = elu1(yv.w); }
                reinterpret_cast<float4 *>(p.y)[o4] = yv;
            }
        }
    }
}

template <int CO_T>
static int launch_pointwise(const ConvParams &p, void *stream) {
    const int64_t total4 = (int64_t)p.B * p.H * p.W * p.Z / 4;
    int64_t gx = ceil_div(total4, kPwThreads);
    if (gx > (int64_t)kNumSMs * 16) gx = (int64_t)kNumSMs * 16;
    return launch("pointwise", pointwise_kernel<CO_T>, dim3((unsigned)gx, (unsigned)ceil_div(p.Cout, CO_T)), dim3(kPwThreads), 0, stream, p);
}

// nn.Upsample(scale_factor=2, mode='trilinear', align_corners=False): src = o/2 - 0.25 clamped at 0
__device__ __forceinline__ void up_taps(int o, int n, int &i0, int &i1, float &l1) {
    float src = 0.5f * (float)o - 0.25f;
    if (src < 0.0f) src = 0.0f;
    i0 = (int)src;
    l1 = src - (float)i0;
    i1 = i0 + (i0 < n - 1 ? 1 : 0);
}

__global__ void __launch_bounds__(256)
upsample2x_kernel(const float *__restrict__ x, int64_t BC, int H, int W, int Z, int pre_act, const float *pre_a,
                  const float *pre_b, float *__restrict__ y) {
    const int Ho = 2 * H, Wo = 2 * W, Zo = 2 * Z;
    const int64_t So = (int64_t)Ho * Wo * Zo, S = (int64_t)H * W * Z;
    const int64_t total = BC * So;
    const float pa = ld_scalar(pre_a, 0.0f), pb = ld_scalar(pre_b, 0.0f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bc = i / So;
        int64_t r = i - bc * So;
        const int oh = (int)(r / ((int64_t)Wo * Zo));
        r -= (int64_t)oh * Wo * Zo;
        const int ow = (int)(r / Zo), oz = (int)(r - (int64_t)ow * Zo);
        int h0, h1, w0, w1, z0, z1;
        float lh, lw, lz;
        up_taps(oh, H, h0, h1, lh);
        up_taps(ow, W, w0, w1, lw);
        up_taps(oz, Z, z0, z1, lz);
        const float *src = x + bc * S;
        float vals[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int hh = (t & 4) ? h1 : h0, ww = (t & 2) ? w1 : w0, zz = (t & 1) ? z1 : z0;
            float xv = __ldg(src + ((size_t)hh * W + ww) * Z + zz);
            vals[t] = pre_act ? elu1(xv + pa) + pb : xv + pb;
        }
        const float a00 = vals[0] * (1.0f - lz) + vals[1] * lz, a01 = vals[2] * (1.0f - lz) + vals[3] * lz;
        const float a10 = vals[4] * (1.0f - lz) + vals[5] * lz, a11 = vals[6] * (1.0f - lz) + vals[7] * lz;
        const float b0 = a00 * (1.0f - lw) + a01 * lw, b1 = a10 * (1.0f - lw) + a11 * lw;
        y[i] = b0 * (1.0f - lh) + b1 * lh;
    }
}

// model.py:120-152 fused: ELU, depth mask, cylinder mask, smooth-L1 sum
__global__ void __launch_bounds__(256)
huber_elu_mask_kernel(const float *__restrict__ dec, const float *__restrict__ x, const int *__restrict__ num_valid,
                      const uint8_t *__restrict__ mask_hw, int64_t B, int HW, int Z, double *sum, double *count) {
    __shared__ double red[32];
    const int64_t total = B * (int64_t)HW * Z;
    double acc = 0.0, cnt = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % Z);
        const int64_t r = i / Z;
        const int hw = (int)(r % HW);
        const int b = (int)(r / HW);
        if (mask_hw && !mask_hw[hw]) continue;
        float loc = elu1(dec[i]);
        if (num_valid && z >= num_valid[b]) loc = 0.0f;
        const float d = fabsf(loc - x[i]);
        acc += (double)(d < 1.0f ? 0.5f * d * d : d - 0.5f);
        cnt += 1.0;
    }
    // block reduce (all threads reach here)
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ double red2[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[warp] = acc; red2[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[w]; c += red2[w]; }
        atomicAdd(sum, a);
        atomicAdd(count, c);
    }
}

// The same pass with every statistic the reference logs next to the loss (model.py:143-149: sub_metric_log_dict of the
// unreduced loss and of the reconstruction, nmse and psnr of metrics/evaluate.py:18-24), as partial sums:
//   sums[0..7] = sum loss, count, sum (loc - x)^2, sum x^2, sum loc, sum loc^2, sum loss^2, unused
//   minmax[0..3] = min loc, max loc, min loss, max loss   (caller initialises to +inf, -inf, +inf, -inf)
__device__ __forceinline__ void atomic_min_f32(float *addr, float v) {      // works for any sign: ordered-int trick
    if (v >= 0.0f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float *addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256)
huber_elu_mask_stats_kernel(const float *__restrict__ dec, const float *__restrict__ x, const int *__restrict__ num_valid,
                            const uint8_t *__restrict__ mask_hw, int64_t B, int HW, int Z, double *sums, float *minmax) {
    __shared__ double red[7][8];
    __shared__ float redm[4][8];
    const int64_t total = B * (int64_t)HW * Z;
    double a[7] = {0, 0, 0, 0, 0, 0, 0};
    float lo_loc = __int_as_float(0x7f800000), hi_loc = -lo_loc, lo_l = lo_loc, hi_l = -lo_loc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % Z);
        const int64_t r = i / Z;
        const int hw = (int)(r % HW);
        const int b = (int)(r / HW);
        if (mask_hw && !mask_hw[hw]) continue;
        float loc = elu1(dec[i]);
        if (num_valid && z >= num_valid[b]) loc = 0.0f;
        const float xv = x[i];
        const float df = loc - xv, d = fabsf(df);
        const float l = d < 1.0f ? 0.5f * d * d : d - 0.5f;
        a[0] += (double)l; a[1] += 1.0; a[2] += (double)df * df; a[3] += (double)xv * xv; a[4] += (double)loc; a[5] += (double)loc * loc;
        a[6] += (double)l * l;
        lo_loc = fminf(lo_loc, loc); hi_loc = fmaxf(hi_loc, loc); lo_l = fminf(lo_l, l); hi_l = fmaxf(hi_l, l);
    }
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 7; ++k) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
        lo_loc = fminf(lo_loc, __shfl_xor_sync(0xffffffffu, lo_loc, o)); hi_loc = fmaxf(hi_loc, __shfl_xor_sync(0xffffffffu, hi_loc, o));
        lo_l = fminf(lo_l, __shfl_xor_sync(0xffffffffu, lo_l, o)); hi_l = fmaxf(hi_l, __shfl_xor_sync(0xffffffffu, hi_l, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) red[k][warp] = a[k];
        redm[0][warp] = lo_loc; redm[1][warp] = hi_loc; redm[2][warp] = lo_l; redm[3][warp] = hi_l;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
        atomicAdd(sums + threadIdx.x, t);
    } else if (threadIdx.x >= 32 && threadIdx.x < 36) {
        const int k = threadIdx.x - 32;
        float t = redm[k][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = (k & 1) ? fmaxf(t, redm[k][w]) : fminf(t, redm[k][w]);
        if (t == t && fabsf(t) != __int_as_float(0x7f800000)) { if (k & 1) atomic_max_f32(minmax + k, t); else atomic_min_f32(minmax + k, t); }
    }
}

// ---- exact medians of the two logged tensors (utils/logging_helpers.py:13: torch.median = the LOWER middle element) ----
// Radix select, four passes of 8 bits over an order-preserving 32-bit key of the value; a pass recomputes loc and loss from
// (decoded, x) like the statistics kernel, so that neither tensor is ever materialised.  State per quantity q (0 = loc,
// 1 = loss) in the workspace: [0] key prefix found so far, [1] rank still to go inside that prefix (u64 in [2..3]),
// [4] number of NaNs (u64 in [4..5]); then hist[q][256] (u64).
constexpr int kMedStateWords = 8;
__device__ __forceinline__ unsigned med_key(float v) {          // ascending floats <-> ascending keys
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float med_value(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__global__ void __launch_bounds__(256)
huber_elu_mask_median_pass_kernel(const float *__restrict__ dec, const float *__restrict__ x, const int *__restrict__ num_valid,
                                  const uint8_t *__restrict__ mask_hw, int64_t B, int HW, int Z, int pass, unsigned *state,
                                  unsigned long long *hist) {
    __shared__ unsigned wh[8][2][256];           // per warp and quantity: no contention between warps
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 2 * 256; i += blockDim.x) (&wh[0][0][0])[i] = 0u;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const unsigned pmask = pass == 0 ? 0u : 0xffffffffu << (shift + 8);
    const unsigned prefix[2] = {state[0], state[kMedStateWords]};
    unsigned long long nan_n[2] = {0ull, 0ull};
    const int64_t total = B * (int64_t)HW * Z;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % Z);
        const int64_t r = i / Z;
        const int hw = (int)(r % HW);
        const int b = (int)(r / HW);
        if (mask_hw && !mask_hw[hw]) continue;
        float loc = elu1(dec[i]);
        if (num_valid && z >= num_valid[b]) loc = 0.0f;
        const float df = loc - x[i], d = fabsf(df);
        const float l = d < 1.0f ? 0.5f * d * d : d - 0.5f;
        const float val[2] = {loc, l};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (val[q] != val[q]) { if (pass == 0) ++nan_n[q]; continue; }        // torch.median: any NaN -> NaN
            const unsigned key = med_key(val[q]);
            if ((key & pmask) == prefix[q]) atomicAdd(&wh[warp][q][(key >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 256; i += blockDim.x) {
        unsigned long long t = 0ull;
        for (int w = 0; w < 8; ++w) t += wh[w][i >> 8][i & 255];
        if (t) atomicAdd(hist + i, t);
    }
    if (pass == 0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            unsigned long long t = nan_n[q];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0 && t) atomicAdd(reinterpret_cast<unsigned long long *>(state + q * kMedStateWords + 4), t);
        }
    }
}

// after a pass: the digit whose bucket holds the wanted rank extends the prefix; the last pass writes the values
__global__ void huber_elu_mask_median_step_kernel(int pass, unsigned *state, unsigned long long *hist, float *medians) {
    const int q = threadIdx.x;
    if (q >= 2) return;
    unsigned *st = state + q * kMedStateWords;
    unsigned long long *h = hist + q * 256;
    unsigned long long *rank = reinterpret_cast<unsigned long long *>(st + 2);
    const unsigned long long nans = *reinterpret_cast<unsigned long long *>(st + 4);
    if (pass == 0) {
        unsigned long long n = nans;
        for (int d = 0; d < 256; ++d) n += h[d];
        *rank = n ? (n - 1) / 2 : 0;             // index of torch.median's element in sorted order
        if (n == 0) st[6] = 1u;                  // empty selection
    }
    unsigned long long k = *rank, acc = 0;
    int d = 0;
    for (; d < 255; ++d) {
        if (acc + h[d] > k) break;
        acc += h[d];
    }
    *rank = k - acc;
    st[0] |= (unsigned)d << (24 - 8 * pass);
    for (int e = 0; e < 256; ++e) h[e] = 0ull;
    if (pass == 3) medians[q] = (nans || st[6]) ? __int_as_float(0x7fc00000) : med_value(st[0]);
}

}  // namespace vq3d

using namespace vq3d;

namespace vq3d {
// decode_embeddings.py:43-47 fused: HU = rint(ELU(decoded) * scale - offset) as int64 (np.rint: round half to even)
__global__ void __launch_bounds__(256)
elu_hu_rint_kernel(const float *__restrict__ x, int64_t n, float scale, float offset, int64_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __fsub_rn(__fmul_rn(elu1(x[i]), scale), offset);       // two roundings like the numpy expression
        out[i] = (int64_t)rintf(v);
    }
}

// the same epilogue with a 16-bit result (CT Hounsfield units fit int16: the reference clips its inputs to [-1500, 3000],
// utils/load_nrrd_dataset.py:73-81); values outside int16 saturate.  Four voxels per thread: one 16-byte load, one 8-byte store.
__global__ void __launch_bounds__(256)
elu_hu_rint_i16_kernel(const float *__restrict__ x, int64_t n, float scale, float offset, int16_t *__restrict__ out) {
    auto one = [&](float d) -> int {
        const float v = __fsub_rn(__fmul_rn(elu1(d), scale), offset);
        return (int)fminf(fmaxf(rintf(v), -32768.0f), 32767.0f);
    };
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(reinterpret_cast<const float4 *>(x) + i);
        const int a = one(v.x), b = one(v.y), c = one(v.z), d = one(v.w);
        uint2 pk;
        pk.x = (uint32_t)(a & 0xffff) | ((uint32_t)(b & 0xffff) << 16);
        pk.y = (uint32_t)(c & 0xffff) | ((uint32_t)(d & 0xffff) << 16);
        __stcs(reinterpret_cast<uint2 *>(out) + i, pk);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[(n4 << 2) + threadIdx.x] = (int16_t)one(x[(n4 << 2) + threadIdx.x]);
}

// CT front end, utils/load_nrrd_dataset.py:73-81 (monai ThresholdIntensity x2, ScaleIntensity(factor = -1 + 1/1000),
// ShiftIntensity(1)) on raw int16 Hounsfield units: out = clip(hu, lo, hi) * mul + add in fp32, two roundings like the
// reference's separate transforms.
__global__ void __launch_bounds__(256)
hu_to_network_kernel(const int16_t *__restrict__ hu, int64_t n, float lo, float hi, float mul, float add, float *__restrict__ out) {
    auto one = [&](int h) -> float { return __fadd_rn(__fmul_rn(fminf(fmaxf((float)h, lo), hi), mul), add); };
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 pk = __ldcs(reinterpret_cast<const uint2 *>(hu) + i);
        float4 v;
        v.x = one((int16_t)(pk.x & 0xffff)); v.y = one((int16_t)(pk.x >> 16));
        v.z = one((int16_t)(pk.y & 0xffff)); v.w = one((int16_t)(pk.y >> 16));
        reinterpret_cast<float4 *>(out)[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[(n4 << 2) + threadIdx.x] = one(hu[(n4 << 2) + threadIdx.x]);
}
}  // namespace vq3d

extern "C" int vq3d_elu_hu_rint_i16(const float *decoded, int64_t n, double scale, double offset, int16_t *out, void *stream) {
    if (!decoded || !out || n < 0) return vq3d::fail(VQ3D_ERR_INVALID, "elu_hu_rint_i16: bad arguments");
    if ((reinterpret_cast<uintptr_t>(decoded) & 15) || (reinterpret_cast<uintptr_t>(out) & 7))
        return vq3d::fail(VQ3D_ERR_INVALID, "elu_hu_rint_i16: decoded must be 16-byte and out 8-byte aligned");
    if (n == 0) return VQ3D_OK;
    int64_t blocks = vq3d::ceil_div(n, 256 * 4 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    return vq3d::launch("elu_hu_rint_i16", vq3d::elu_hu_rint_i16_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, decoded, n,
                        (float)scale, (float)offset, out);
}

extern "C" int vq3d_hu_to_network(const int16_t *hu, int64_t n, double min_hu, double max_hu, double mul, double add, float *out, void *stream) {
    if (!hu || !out || n < 0 || !(min_hu <= max_hu)) return vq3d::fail(VQ3D_ERR_INVALID, "hu_to_network: bad arguments");
    if ((reinterpret_cast<uintptr_t>(hu) & 7) || (reinterpret_cast<uintptr_t>(out) & 15))
        return vq3d::fail(VQ3D_ERR_INVALID, "hu_to_network: hu must be 8-byte and out 16-byte aligned");
    if (n == 0) return VQ3D_OK;
    int64_t blocks = vq3d::ceil_div(n, 256 * 4 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    return vq3d::launch("hu_to_network", vq3d::hu_to_network_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, hu, n, (float)min_hu,
                        (float)max_hu, (float)mul, (float)add, out);
}

extern "C" int vq3d_elu_hu_rint(const float *decoded, int64_t n, double scale, double offset, int64_t *out, void *stream) {
    if (!decoded || !out || n < 0) return vq3d::fail(VQ3D_ERR_INVALID, "elu_hu_rint: bad arguments");
    if (n == 0) return VQ3D_OK;
    int64_t blocks = vq3d::ceil_div(n, 256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    return vq3d::launch("elu_hu_rint", vq3d::elu_hu_rint_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, decoded, n, (float)scale,
                        (float)offset, out);
}

extern "C" int vq3d_abi_version(void) { return VQ3D_ABI_VERSION; }
extern "C" const char *vq3d_last_error(void) { return err_buf(); }
extern "C" int vq3d_is_cuda_build(void) {
#ifdef VQ3D_EMU
    return 0;
#else
    return 1;
#endif
}

extern "C" int vq3d_conv3d(const vq3d_conv_desc *d, void *stream) {
    if (!d) return fail(VQ3D_ERR_INVALID, "conv3d: null descriptor");
    if (!d->x1 || !d->w || !d->y) return fail(VQ3D_ERR_INVALID, "conv3d: null x1/w/y");
    if (d->C2 > 0 && !d->x2) return fail(VQ3D_ERR_INVALID, "conv3d: C2 > 0 but x2 is NULL");
    if (d->B < 1 || d->H < 1 || d->W < 1 || d->Z < 1 || d->C1 < 1 || d->C2 < 0 || d->Cout < 1)
        return fail(VQ3D_ERR_INVALID, "conv3d: bad sizes");
    if (d->k < 1 || d->k > 4 || (d->stride != 1 && d->stride != 2) || d->pad < 0 || d->pad >= d->k)
        return fail(VQ3D_ERR_INVALID, "conv3d: unsupported k=%d stride=%d pad=%d", d->k, d->stride, d->pad);
    if (d->pad_circular && (d->pad > d->H || d->pad > d->W || d->pad > d->Z))
        return fail(VQ3D_ERR_INVALID, "conv3d: circular padding larger than the input");
    ConvParams p;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z; p.C1 = d->C1; p.C2 = d->C2; p.Cout = d->Cout;
    p.k = d->k; p.stride = d->stride; p.pad = d->pad; p.circ = d->pad_circular; p.pre_act = d->pre_act; p.post_act = d->post_act;
    p.Ho = (d->H + 2 * d->pad - d->k) / d->stride + 1;
    p.Wo = (d->W + 2 * d->pad - d->k) / d->stride + 1;
    p.Zo = (d->Z + 2 * d->pad - d->k) / d->stride + 1;
    if (p.Ho < 1 || p.Wo < 1 || p.Zo < 1) return fail(VQ3D_ERR_INVALID, "conv3d: empty output");
    p.x1 = d->x1; p.x2 = d->x2; p.w = d->w; p.bias = d->bias; p.pre_a = d->pre_a; p.pre_b = d->pre_b;
    p.post_scale = d->post_scale; p.post_b = d->post_b; p.residual = d->residual; p.y = d->y;
    const int64_t total = (int64_t)p.B * p.Ho * p.Wo * p.Zo;
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (d->k == 1 && d->stride == 1 && d->pad == 0 && ((int64_t)d->H * d->W * d->Z) % 4 == 0 && d->C1 + d->C2 <= kPwMaxCin &&
        total >= 65536 && al16(d->x1) && al16(d->x2) && al16(d->y) && al16(d->residual)) {
        if (p.Cout % 9 == 0) return launch_pointwise<9>(p, stream);
        if (p.Cout <= 4) return launch_pointwise<4>(p, stream);
        if (p.Cout <= 8) return launch_pointwise<8>(p, stream);
        return launch_pointwise<16>(p, stream);
    }
    if (d->k == 3 && d->stride == 1 && d->pad == 1 && d->C1 + d->C2 <= kCtMaxCin && d->H >= 3 && d->W >= 3 && d->Z >= 3 && total >= 16384) {
        if (p.Cout % 9 == 0) return launch_conv_tiled<9>(p, stream);        // the 9- / 18- / 36-channel branches: no padded second pass
        if (p.Cout >= 8) return launch_conv_tiled<8>(p, stream);
        if (p.Cout >= 3) return launch_conv_tiled<4>(p, stream);
        if (p.Cout == 2) return launch_conv_tiled<2>(p, stream);
        return launch_conv_tiled<1>(p, stream);
    }
    const unsigned gx = (unsigned)ceil_div(total, kConvThreads);
    if (p.Cout >= 8) return launch("conv3d<8>", conv3d_generic_kernel<8>, dim3(gx, (unsigned)ceil_div(p.Cout, 8)), dim3(kConvThreads), 0, stream, p);
    if (p.Cout >= 3) return launch("conv3d<4>", conv3d_generic_kernel<4>, dim3(gx, (unsigned)ceil_div(p.Cout, 4)), dim3(kConvThreads), 0, stream, p);
    if (p.Cout == 2) return launch("conv3d<2>", conv3d_generic_kernel<2>, dim3(gx, 1), dim3(kConvThreads), 0, stream, p);
    return launch("conv3d<1>", conv3d_generic_kernel<1>, dim3(gx, 1), dim3(kConvThreads), 0, stream, p);
}

extern "C" int vq3d_upsample2x(const float *x, int64_t B, int C, int H, int W, int Z, int pre_act, const float *pre_a,
                               const float *pre_b, float *y, void *stream) {
    if (!x || !y || B < 1 || C < 1 || H < 1 || W < 1 || Z < 1) return fail(VQ3D_ERR_INVALID, "upsample2x: bad arguments");
    const int64_t total = B * C * (int64_t)H * W * Z * 8;
    int64_t blocks = ceil_div(total, 256);
    if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
    return launch("upsample2x", upsample2x_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, x, B * C, H, W, Z, pre_act, pre_a, pre_b, y);
}

extern "C" int vq3d_huber_elu_mask(const float *decoded, const float *x, const int32_t *num_valid, const uint8_t *mask_hw,
                                   int64_t B, int H, int W, int Z, double *sum, double *count, void *stream) {
    if (!decoded || !x || !sum || !count || B < 1 || H < 1 || W < 1 || Z < 1) return fail(VQ3D_ERR_INVALID, "huber: bad arguments");
    const int64_t total = B * (int64_t)H * W * Z;
    int64_t blocks = ceil_div(total, 256 * 4);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    return launch("huber_elu_mask", huber_elu_mask_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, decoded, x,
                  (const int *)num_valid, mask_hw, B, H * W, Z, sum, count);
}

extern "C" int vq3d_huber_elu_mask_stats(const float *decoded, const float *x, const int32_t *num_valid, const uint8_t *mask_hw,
                                         int64_t B, int H, int W, int Z, double *sums, float *minmax, void *stream) {
    if (!decoded || !x || !sums || !minmax || B < 1 || H < 1 || W < 1 || Z < 1) return fail(VQ3D_ERR_INVALID, "huber stats: bad arguments");
    const int64_t total = B * (int64_t)H * W * Z;
    int64_t blocks = ceil_div(total, 256 * 4);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    return launch("huber_elu_mask_stats", huber_elu_mask_stats_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, decoded, x,
                  (const int *)num_valid, mask_hw, B, H * W, Z, sums, minmax);
}

extern "C" size_t vq3d_huber_elu_mask_medians_workspace(void) { return (size_t)2 * kMedStateWords * 4 + (size_t)2 * 256 * 8; }

extern "C" int vq3d_huber_elu_mask_medians(const float *decoded, const float *x, const int32_t *num_valid, const uint8_t *mask_hw,
                                           int64_t B, int H, int W, int Z, float *medians, void *ws, size_t ws_bytes, void *stream) {
    if (!decoded || !x || !medians || !ws || B < 1 || H < 1 || W < 1 || Z < 1) return fail(VQ3D_ERR_INVALID, "huber medians: bad arguments");
    if (ws_bytes < vq3d_huber_elu_mask_medians_workspace() || (reinterpret_cast<uintptr_t>(ws) & 7) != 0)
        return fail(VQ3D_ERR_INVALID, "huber medians: workspace too small or not 8-byte aligned");
    unsigned *state = static_cast<unsigned *>(ws);
    unsigned long long *hist = reinterpret_cast<unsigned long long *>(state + 2 * kMedStateWords);
    cudaError_t e = cudaMemsetAsync(ws, 0, vq3d_huber_elu_mask_medians_workspace(), static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return check_cuda(e, "huber medians(memset)");
    const int64_t total = B * (int64_t)H * W * Z;
    int64_t blocks = ceil_div(total, 256 * 4);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    for (int pass = 0; pass < 4; ++pass) {
        int rc = launch("huber_elu_mask_median_pass", huber_elu_mask_median_pass_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, decoded, x,
                        (const int *)num_valid, mask_hw, B, H * W, Z, pass, state, hist);
        if (rc != VQ3D_OK) return rc;
        rc = launch("huber_elu_mask_median_step", huber_elu_mask_median_step_kernel, dim3(1), dim3(32), 0, stream, pass, state, hist, medians);
        if (rc != VQ3D_OK) return rc;
    }
    return VQ3D_OK;
}
