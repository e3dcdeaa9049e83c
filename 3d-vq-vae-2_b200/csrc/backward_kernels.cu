// backward_kernels.cu -- gradients of the generic fused convolution (vq3d_conv3d), of the trilinear x2
// upsampling and of the Huber epilogue: what torch.autograd derives for the reference's nn.Conv3d / F.pad /
// F.elu / nn.Upsample / F.smooth_l1_loss chains (vqvae/layers.py:176-195,591-597; model.py:120-163).
// Shape-generic fp32 SIMT kernels (correctness first: the training step composes blocks from them).
//
//   forward:  u = pre_act ? ELU(x + a) + b : x + b ;  raw = conv(u, w) ;  y = raw * s + pb + bias[co] + residual
//   backward: g_raw = gy * s
//             gu[ci, i]   = sum_{co, taps that read i} w[co, ci, tap] * g_raw[co, o]           (dgrad, gather form)
//             gx          = gu * (pre_act ? ELU'(x + a) : 1);   d a = sum gx;   d b = sum gu
//             gw[co,ci,t] = sum_o g_raw[co, o] * u[ci, in(o, t)]                                (wgrad)
//             d bias[co]  = sum_o gy[co, o];   d pb = sum gy;   d s = sum gy * raw;   d residual = gy
#include "vq3d_rt.h"

namespace vq3d {

struct BwdParams {
    int B, H, W, Z, C1, C2, Cout, k, stride, pad, circ, pre_act;
    int Ho, Wo, Zo;
    const float *x1, *x2, *w, *pre_a, *pre_b, *post_scale;
    const float *gy, *raw;
    float *gx1, *gx2, *gw, *gbias, *gscal;      // gscal: [d pre_a, d pre_b, d post_scale, d post_b]
    int gscal_want_pb;                          // conv1x1_bwd_kernel: the forward had a post_b
};

__device__ __forceinline__ float block_sum(float v, float *red /* [32] shared */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) t += red[w];
    return t;    // valid in thread 0
}

// outputs o (per axis) whose window tap kk reads input index i: o*stride - pad + kk == i (mod n when circular)
__device__ __forceinline__ int taps_of(int i, int n_in, int n_out, int k, int stride, int pad, int circ, int *o_list, int *k_list) {
    int cnt = 0;
    for (int kk = 0; kk < k; ++kk) {
        for (int sh = (circ ? -1 : 0); sh <= (circ ? 1 : 0); ++sh) {
            const int num = i + sh * n_in + pad - kk;
            if (num < 0 || num % stride != 0) continue;
            const int o = num / stride;
            if (o >= n_out) continue;
            const int src = o * stride - pad + kk;          // the un-wrapped coordinate the forward pass read
            if (circ ? (src < -n_in || src >= 2 * n_in) : (src < 0 || src >= n_in)) continue;
            if (cnt < 12) { o_list[cnt] = o; k_list[cnt] = kk; ++cnt; }
        }
    }
    return cnt;
}

// one thread = one input voxel x one input channel (blockIdx.y)
__global__ void __launch_bounds__(128)
conv3d_dgrad_kernel(BwdParams p) {
    __shared__ float red[32];
    const int Cin = p.C1 + p.C2, k = p.k, k3 = k * k * k;
    const int ci = blockIdx.y;
    const int64_t S = (int64_t)p.H * p.W * p.Z, So = (int64_t)p.Ho * p.Wo * p.Zo;
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = v < (int64_t)p.B * S;
    const float sc = ld_scalar(p.post_scale, 1.0f);
    float gu = 0.0f, gx = 0.0f;
    if (active) {
        const int b = (int)(v / S);
        int64_t r = v - (int64_t)b * S;
        const int ih = (int)(r / ((int64_t)p.W * p.Z));
        r -= (int64_t)ih * p.W * p.Z;
        const int iw = (int)(r / p.Z), iz = (int)(r - (int64_t)iw * p.Z);
        int oh[12], kh[12], ow[12], kw[12], oz[12], kz[12];
        const int nh = taps_of(ih, p.H, p.Ho, k, p.stride, p.pad, p.circ, oh, kh);
        const int nw = taps_of(iw, p.W, p.Wo, k, p.stride, p.pad, p.circ, ow, kw);
        const int nz = taps_of(iz, p.Z, p.Zo, k, p.stride, p.pad, p.circ, oz, kz);
        for (int co = 0; co < p.Cout; ++co) {
            const float *g = p.gy + ((size_t)b * p.Cout + co) * So;
            const float *wp = p.w + ((size_t)co * Cin + ci) * k3;
            for (int a = 0; a < nh; ++a)
                for (int c = 0; c < nw; ++c)
                    for (int e = 0; e < nz; ++e)
                        gu = __fmaf_rn(wp[(kh[a] * k + kw[c]) * k + kz[e]], g[((size_t)oh[a] * p.Wo + ow[c]) * p.Zo + oz[e]], gu);
        }
        gu *= sc;
        const bool first = ci < p.C1;
        const float *xs = first ? p.x1 + ((size_t)b * p.C1 + ci) * S : p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S;
        float *gxs = first ? (p.gx1 ? p.gx1 + ((size_t)b * p.C1 + ci) * S : nullptr) : (p.gx2 ? p.gx2 + ((size_t)b * p.C2 + (ci - p.C1)) * S : nullptr);
        const int64_t off = v - (int64_t)b * S;
        gx = gu;
        if (p.pre_act) {
            const float t = xs[off] + ld_scalar(p.pre_a, 0.0f);
            gx = t > 0.0f ? gu : gu * __expf(t);
        }
        if (gxs) gxs[off] = gx;
    }
    if (p.gscal) {
        const float sa = block_sum(p.pre_act ? gx : 0.0f, red);
        const float sb = block_sum(gu, red);
        if (threadIdx.x == 0) {
            if (p.pre_act && p.pre_a) atomicAdd(p.gscal + 0, sa);
            if (p.pre_b) atomicAdd(p.gscal + 1, sb);
        }
    }
}

// Input gradient of the two stride-2 convolutions of a 'down' block (layers.py:124-132: k4 s2 p1 branch convolution,
// k2 s2 p0 skip) in gather form without searching: along each axis input index i is read by at most two outputs of the
// k4 convolution (taps k = par and par + 2 with par = (i + 1) & 1, outputs (i + 1 - par) / 2 and one less, wrapped when the
// padding is circular) and by exactly one of the k2 convolution (tap i & 1, output i >> 1).  Thread = input voxel, four input
// channels per pass, weights in shared memory.  (The shape-generic kernel above spent 13 ms on the 4 -> 4 k4 s2 layer of one
// 512 x 512 x 128 volume; this one is bound by writing gx.)
template <int K>
__global__ void __launch_bounds__(256)
conv3d_dgrad_s2_kernel(BwdParams p) {
    VQ3D_DYN_SMEM(float, sw);                          // [Cout][Cin][K^3]
    __shared__ float red[32];
    constexpr int K3 = K * K * K, NC = K == 4 ? 2 : 1;
    const int Cin = p.C1 + p.C2, Cout = p.Cout;
    for (int i = threadIdx.x; i < Cout * Cin * K3; i += blockDim.x) sw[i] = __ldg(p.w + i);
    __syncthreads();
    const int64_t S = (int64_t)p.H * p.W * p.Z, So = (int64_t)p.Ho * p.Wo * p.Zo;
    const float sc = ld_scalar(p.post_scale, 1.0f), pa = ld_scalar(p.pre_a, 0.0f);
    float s_a = 0.0f, s_b = 0.0f;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < (int64_t)p.B * S; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = v / S;
        int64_t r = v - b * S;
        const int ih = (int)(r / ((int64_t)p.W * p.Z));
        r -= (int64_t)ih * p.W * p.Z;
        const int iw = (int)(r / p.Z), iz = (int)(r - (int64_t)iw * p.Z);
        int oh[NC], kh[NC], ow[NC], kw[NC], oz[NC], kz[NC];
        bool vh[NC], vw[NC], vz[NC];
        auto axis = [&](int i, int n_out, int *o, int *k, bool *ok) {
            if (K == 4) {
                const int par = (i + 1) & 1;
                o[0] = (i + 1 - par) >> 1; k[0] = par;
                o[1 % NC] = o[0] - 1; k[1 % NC] = par + 2;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (p.circ) { o[c] = o[c] < 0 ? o[c] + n_out : (o[c] >= n_out ? o[c] - n_out : o[c]); ok[c] = true; }
                    else ok[c] = o[c] >= 0 && o[c] < n_out;
                }
            } else {
                o[0] = i >> 1; k[0] = i & 1; ok[0] = o[0] < n_out;
            }
        };
        axis(ih, p.Ho, oh, kh, vh); axis(iw, p.Wo, ow, kw, vw); axis(iz, p.Zo, oz, kz, vz);
        const float *gyb = p.gy + (size_t)b * Cout * So;
        for (int c0 = 0; c0 < Cin; c0 += 4) {
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int a = 0; a < NC; ++a)
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int e = 0; e < NC; ++e) {
                        if (!(vh[a] && vw[c] && vz[e])) continue;
                        const size_t go = ((size_t)oh[a] * p.Wo + ow[c]) * p.Zo + oz[e];
                        const int tap = (kh[a] * K + kw[c]) * K + kz[e];
                        for (int co = 0; co < Cout; ++co) {
                            const float g = __ldg(gyb + (size_t)co * So + go);
                            const float *wr = sw + ((size_t)co * Cin + c0) * K3 + tap;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (c0 + j < Cin) acc[j] = __fmaf_rn(wr[j * K3], g, acc[j]);
                        }
                    }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = c0 + j;
                if (ci >= Cin) break;
                const float gu = acc[j] * sc;
                const bool first = ci < p.C1;
                const size_t off = first ? ((size_t)b * p.C1 + ci) * S + (v - b * S) : ((size_t)b * p.C2 + (ci - p.C1)) * S + (v - b * S);
                float gx = gu;
                if (p.pre_act) {
                    const float t = (first ? p.x1 : p.x2)[off] + pa;
                    gx = t > 0.0f ? gu : gu * __expf(t);
                }
                float *dst = first ? p.gx1 : p.gx2;
                if (dst) dst[off] = gx;
                s_b += gu;
                s_a += gx;
            }
        }
    }
    if (p.gscal) {
        const float ta = block_sum(p.pre_act ? s_a : 0.0f, red);
        const float tb = block_sum(s_b, red);
        if (threadIdx.x == 0) {
            if (p.pre_act && p.pre_a) atomicAdd(p.gscal + 0, ta);
            if (p.pre_b) atomicAdd(p.gscal + 1, tb);
        }
    }
}

// grid (voxel chunks, Cin, Cout): every thread accumulates the k^3 taps of one (co, ci) pair over its voxels
__global__ void __launch_bounds__(128)
conv3d_wgrad_kernel(BwdParams p) {
    __shared__ float red[32];
    const int Cin = p.C1 + p.C2, k = p.k, k3 = k * k * k;
    const int ci = blockIdx.y, co = blockIdx.z;
    const int64_t S = (int64_t)p.H * p.W * p.Z, So = (int64_t)p.Ho * p.Wo * p.Zo;
    const int64_t total = (int64_t)p.B * So;
    const float sc = ld_scalar(p.post_scale, 1.0f), pa = ld_scalar(p.pre_a, 0.0f), pb = ld_scalar(p.pre_b, 0.0f);
    float acc[64];
#pragma unroll
    for (int t = 0; t < 64; ++t) acc[t] = 0.0f;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(v / So);
        int64_t r = v - (int64_t)b * So;
        const int oh = (int)(r / ((int64_t)p.Wo * p.Zo));
        r -= (int64_t)oh * p.Wo * p.Zo;
        const int ow = (int)(r / p.Zo), oz = (int)(r - (int64_t)ow * p.Zo);
        const float g = p.gy[((size_t)b * p.Cout + co) * So + (v - (int64_t)b * So)] * sc;
        const float *src = ci < p.C1 ? p.x1 + ((size_t)b * p.C1 + ci) * S : p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S;
#pragma unroll 1
        for (int kh = 0; kh < k; ++kh) {
            int ih = oh * p.stride - p.pad + kh;
            if (p.circ) ih = wrap(ih, p.H); else if (ih < 0 || ih >= p.H) continue;
#pragma unroll 1
            for (int kw = 0; kw < k; ++kw) {
                int iw = ow * p.stride - p.pad + kw;
                if (p.circ) iw = wrap(iw, p.W); else if (iw < 0 || iw >= p.W) continue;
                for (int kz = 0; kz < 4; ++kz) {
                    if (kz >= k) break;
                    int iz = oz * p.stride - p.pad + kz;
                    if (p.circ) iz = wrap(iz, p.Z); else if (iz < 0 || iz >= p.Z) continue;
                    float u = src[((size_t)ih * p.W + iw) * p.Z + iz];
                    u = p.pre_act ? elu1(u + pa) + pb : u + pb;
                    const int t = (kh * k + kw) * k + kz;
                    // static indexing keeps acc in registers
#pragma unroll
                    for (int tt = 0; tt < 64; ++tt)
                        if (tt == t) acc[tt] = __fmaf_rn(g, u, acc[tt]);
                }
            }
        }
    }
    float *gw = p.gw + ((size_t)co * Cin + ci) * k3;
#pragma unroll
    for (int t = 0; t < 64; ++t) {
        if (t < k3) {
            const float s = block_sum(acc[t], red);
            if (threadIdx.x == 0) atomicAdd(gw + t, s);
        }
    }
}

// grid (chunks, Cout, B): d bias[co], d post_b, d post_scale
__global__ void __launch_bounds__(256)
conv3d_outgrads_kernel(const float *__restrict__ gy, const float *__restrict__ raw, int Cout, int64_t So, float *gbias, float *gscal,
                       int want_scale, int want_pb) {
    __shared__ float red[32];
    const int co = blockIdx.y, b = blockIdx.z;
    const size_t base = ((size_t)b * Cout + co) * So;
    float s1 = 0.0f, s2 = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < So; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = gy[base + i];
        s1 += g;
        if (raw) s2 = __fmaf_rn(g, raw[base + i], s2);
    }
    const float t1 = block_sum(s1, red);
    const float t2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        if (gbias) atomicAdd(gbias + co, t1);
        if (gscal && want_pb) atomicAdd(gscal + 3, t1);
        if (gscal && want_scale && raw) atomicAdd(gscal + 2, t2);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused backward of a pointwise (k1 s1) convolution: input gradient with the pre-transform chain, weight gradient, bias
// gradient and the four Fixup scalar gradients in ONE launch (the generic path needs a forward-form dgrad convolution, its
// finish kernel, the tiled wgrad kernel, the out-grads kernel and -- for d scale -- a recomputed forward convolution).
// Two of the three convolutions of every PreActFixupResBlock are pointwise (layers.py:134-160).
//
// Persistent CTAs loop over tiles of T voxels (thread = voxel): x, the transformed input u, ELU' and gy are staged in shared
// memory once per tile; each thread then forms gu = scale * W^T gy (-> gx, d pre_a, d pre_b) and raw = W u (-> d scale) for
// its voxel, and the weight gradient is a [Cout x T] x [T x Cin] product out of the same tiles: every (co, ci) pair -- plus
// one pseudo pair per output channel for the bias sums -- is owned by one thread, or by T / R threads that split the tile's
// voxels when there are fewer pairs than threads (the thin 512^3 layers have 8).  Pair sums stay in registers across tiles;
// one atomicAdd per owner at the end.
constexpr int kPwT = 128;          // voxels per tile = threads
constexpr int kPwLD = kPwT + 4;    // shared-memory row length of the staged tile
constexpr int kPwMaxRows = 24;     // (co, ci) + bias rows per thread: Cin * Cout + Cout <= 24 * 128

// NR: (co, ci) + bias rows per thread (2 for the 9 <-> 18 pair, 24 at most): the accumulators and the unrolled row loop are sized by it
template <int NR>
__global__ void __launch_bounds__(kPwT)
conv1x1_bwd_kernel(BwdParams p, int64_t total, int64_t ntiles, int want_scale) {
    VQ3D_DYN_SMEM(float, sm);
    __shared__ float red[32];
    __shared__ float s_pbsum;
    if (threadIdx.x == 0) s_pbsum = 0.0f;
    __syncthreads();
    // rows of 132 floats: 16-byte aligned (128-bit loads along the voxels in the pair phase; 8 consecutive rows fall into 8
    // different 16-byte bank groups), weights padded to whole float4s in both orientations (128-bit broadcasts)
    const int Cin = p.C1 + p.C2, Cout = p.Cout, T = kPwT, LD = kPwLD, CinP = (Cin + 3) & ~3, CoutP = (Cout + 3) & ~3;
    float *sW = sm;                               // [Cout][CinP]
    float *sWt = sW + Cout * CinP;                // [Cin][CoutP]
    float *sU = sWt + Cin * CoutP;                // [Cin][LD]  transformed input
    float *sD = sU + Cin * LD;                    // [Cin][LD]  ELU' (1 without pre-activation)
    float *sG = sD + Cin * LD;                    // [Cout][LD] gy
    const int tid = threadIdx.x;
    for (int i = tid; i < Cout * CinP; i += T) { const int co = i / CinP, ci = i - co * CinP; sW[i] = ci < Cin ? __ldg(p.w + co * Cin + ci) : 0.0f; }
    for (int i = tid; i < Cin * CoutP; i += T) { const int ci = i / CoutP, co = i - ci * CoutP; sWt[i] = co < Cout ? __ldg(p.w + co * Cin + ci) : 0.0f; }
    const float sc = ld_scalar(p.post_scale, 1.0f), pa = ld_scalar(p.pre_a, 0.0f), pb = ld_scalar(p.pre_b, 0.0f);
    const int64_t S = (int64_t)p.H * p.W * p.Z;
    const int R = Cin * Cout + Cout;              // rows of the pair phase: (co, ci) pairs, then one bias row per co
    const int nsplit = R < T ? T / R : 1;
    const int my_row0 = R < T ? (tid < R * nsplit ? tid % R : -1) : tid;
    const int my_part = R < T ? tid / R : 0;
    float acc[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = 0.0f;
    float s_a = 0.0f, s_b = 0.0f, s_s = 0.0f;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t v = tile * T + tid;
        const bool active = v < total;
        const int64_t b = active ? v / S : 0, r = active ? v - b * S : 0;
        __syncthreads();                          // the previous tile's pair phase has read sU / sG
        // stage the tile: all global loads of a thread are independent -> issue them four channels at a time
        for (int c0 = 0; c0 < Cin; c0 += 4) {
            float xv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = c0 + j;
                xv[j] = 0.0f;
                if (active && ci < Cin)
                    xv[j] = ci < p.C1 ? __ldg(p.x1 + ((size_t)b * p.C1 + ci) * S + r) : __ldg(p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S + r);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = c0 + j;
                if (ci >= Cin) break;
                float u = 0.0f, dd = 0.0f;
                if (active) {
                    if (p.pre_act) {
                        const float t = xv[j] + pa;
                        const float e = __expf(t);
                        u = (t > 0.0f ? t : e - 1.0f) + pb;
                        dd = t > 0.0f ? 1.0f : e;
                    } else {
                        u = xv[j] + pb;
                        dd = 1.0f;
                    }
                }
                sU[ci * LD + tid] = u;
                sD[ci * LD + tid] = dd;
            }
        }
        for (int c0 = 0; c0 < Cout; c0 += 4) {
            float gv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) gv[j] = (active && c0 + j < Cout) ? __ldg(p.gy + ((size_t)b * Cout + c0 + j) * S + r) : 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < Cout) sG[(c0 + j) * LD + tid] = gv[j];
        }
        __syncthreads();
        if (active) {
            const bool want_in = p.gx1 || p.gx2 || (p.gscal && (p.pre_a || p.pre_b));
            if (want_in) {
                for (int c0 = 0; c0 < Cin; c0 += 4) {                 // four input channels per pass: one gy read feeds four FMAs
                    float gu[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    for (int co = 0; co < Cout; ++co) {
                        const float g = sG[co * LD + tid];
                        const float4 w4 = *reinterpret_cast<const float4 *>(sW + co * CinP + c0);      // padded with zeros
                        gu[0] = __fmaf_rn(w4.x, g, gu[0]); gu[1] = __fmaf_rn(w4.y, g, gu[1]);
                        gu[2] = __fmaf_rn(w4.z, g, gu[2]); gu[3] = __fmaf_rn(w4.w, g, gu[3]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ci = c0 + j;
                        if (ci >= Cin) break;
                        const float guj = gu[j] * sc;
                        const float gx = guj * sD[ci * LD + tid];
                        s_b += guj;
                        s_a += gx;
                        float *dst = ci < p.C1 ? (p.gx1 ? p.gx1 + ((size_t)b * p.C1 + ci) * S + r : nullptr)
                                               : (p.gx2 ? p.gx2 + ((size_t)b * p.C2 + (ci - p.C1)) * S + r : nullptr);
                        if (dst) *dst = gx;
                    }
                }
            }
            if (want_scale) {
                for (int c0 = 0; c0 < Cout; c0 += 4) {
                    float raw[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    for (int ci = 0; ci < Cin; ++ci) {
                        const float u = sU[ci * LD + tid];
                        const float4 w4 = *reinterpret_cast<const float4 *>(sWt + ci * CoutP + c0);    // padded with zeros
                        raw[0] = __fmaf_rn(w4.x, u, raw[0]); raw[1] = __fmaf_rn(w4.y, u, raw[1]);
                        raw[2] = __fmaf_rn(w4.z, u, raw[2]); raw[3] = __fmaf_rn(w4.w, u, raw[3]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c0 + j < Cout) s_s = __fmaf_rn(sG[(c0 + j) * LD + tid], raw[j], s_s);
                }
            }
        }
        if (p.gw || p.gbias || (p.gscal && p.gscal_want_pb)) {
            // pair phase: acc[row] += sum over the tile's voxels of gy[co] * u[ci]   (bias rows: u == 1)
            if (R < T) {
                if (my_row0 >= 0) {
                    const int row = my_row0;
                    float a = 0.0f;
                    if (row < Cin * Cout) {
                        const float *g = sG + (row / Cin) * LD, *u = sU + (row % Cin) * LD;
                        for (int vv = my_part; vv < T; vv += nsplit) a = __fmaf_rn(g[vv], u[vv], a);
                    } else {
                        const float *g = sG + (row - Cin * Cout) * LD;
                        for (int vv = my_part; vv < T; vv += nsplit) a += g[vv];
                    }
                    acc[0] += a;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const int row = tid + j * T;
                    if (row < R) {
                        float a = 0.0f;
                        if (row < Cin * Cout) {
                            const float4 *g = reinterpret_cast<const float4 *>(sG + (row / Cin) * LD), *u = reinterpret_cast<const float4 *>(sU + (row % Cin) * LD);
#pragma unroll 4
                            for (int vv = 0; vv < T / 4; ++vv) {          // same summation order as the scalar loop
                                const float4 g4 = g[vv], u4 = u[vv];
                                a = __fmaf_rn(g4.x, u4.x, a); a = __fmaf_rn(g4.y, u4.y, a);
                                a = __fmaf_rn(g4.z, u4.z, a); a = __fmaf_rn(g4.w, u4.w, a);
                            }
                        } else {
                            const float *g = sG + (row - Cin * Cout) * LD;
#pragma unroll 4
                            for (int vv = 0; vv < T; ++vv) a += g[vv];
                        }
                        acc[j] += a;
                    }
                }
            }
        }
    }
    // ---- flush: one atomicAdd per row and CTA (split mode: the T / R partial sums of a row meet in shared memory first;
    // per-thread atomics on a handful of addresses were 80 % of this kernel's time on the thin layers) ----
    if (R < T) {
        __syncthreads();
        if (my_row0 >= 0) sU[my_part * R + my_row0] = acc[0];          // nsplit * R <= T floats
        __syncthreads();
        if (tid < R) {
            float a = 0.0f;
            for (int part = 0; part < nsplit; ++part) a += sU[part * R + tid];
            acc[0] = a;
        }
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        const int row = R < T ? ((j == 0 && tid < R) ? tid : -1) : tid + j * T;
        if (row < 0 || row >= R) continue;
        if (row < Cin * Cout) {
            if (p.gw) atomicAdd(p.gw + row, acc[j] * sc);                 // gw [Cout][Cin]: row = co * Cin + ci
        } else {
            if (p.gbias) atomicAdd(p.gbias + (row - Cin * Cout), acc[j]);
            if (p.gscal && p.gscal_want_pb) atomicAdd(&s_pbsum, acc[j]);     // d post_b = the sum of the bias rows: ONE global atomic per CTA below
        }                                                                      // (C_out atomics per CTA on one address were half of this kernel's time)
    }
    if (p.gscal && p.gscal_want_pb) {
        __syncthreads();
        if (tid == 0) atomicAdd(p.gscal + 3, s_pbsum);
    }
    if (p.gscal) {
        const float ta = block_sum(p.pre_act ? s_a : 0.0f, red);
        const float tb = block_sum(s_b, red);
        const float ts = block_sum(s_s, red);
        if (tid == 0) {
            if (p.pre_act && p.pre_a) atomicAdd(p.gscal + 0, ta);
            if (p.pre_b) atomicAdd(p.gscal + 1, tb);
            if (want_scale) atomicAdd(p.gscal + 2, ts);
        }
    }
}

// Thin pointwise convolutions (C_in, C_out <= 8: every 1x1 of the 512^3 / 256^3 levels and of the 2- / 4- / 8-channel
// stacks): a streaming variant without shared-memory staging.  Thread = voxel (grid-stride); the (co, ci) pair sums and the
// bias sums live in registers for the whole kernel and meet in one warp-shuffle + shared-memory reduction at the end, so the
// kernel reads x and gy once, writes gx once, and issues Cin * Cout + Cout + 3 atomics per CTA.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
conv1x1_bwd_thin_kernel(BwdParams p, int64_t total, int want_scale) {
    __shared__ float sW[COUT * CIN];
    __shared__ float red[8][CIN * COUT + COUT + 3];
    const int tid = threadIdx.x;
    for (int i = tid; i < COUT * CIN; i += blockDim.x) sW[i] = __ldg(p.w + i);
    __syncthreads();
    float w[COUT][CIN];
#pragma unroll
    for (int co = 0; co < COUT; ++co)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) w[co][ci] = sW[co * CIN + ci];
    const float sc = ld_scalar(p.post_scale, 1.0f), pa = ld_scalar(p.pre_a, 0.0f), pb = ld_scalar(p.pre_b, 0.0f);
    const int64_t S = (int64_t)p.H * p.W * p.Z;
    float acc[COUT][CIN], bacc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        bacc[co] = 0.0f;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) acc[co][ci] = 0.0f;
    }
    float s_a = 0.0f, s_b = 0.0f, s_s = 0.0f;
    const bool want_in = p.gx1 || p.gx2 || (p.gscal && (p.pre_a || p.pre_b));
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + tid; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = v / S, r = v - b * S;
        float xv[CIN], g[COUT], u[CIN], dd[CIN];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
            xv[ci] = ci < p.C1 ? __ldcs(p.x1 + ((size_t)b * p.C1 + ci) * S + r) : __ldcs(p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S + r);
#pragma unroll
        for (int co = 0; co < COUT; ++co) g[co] = __ldcs(p.gy + ((size_t)b * COUT + co) * S + r);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            if (p.pre_act) {
                const float t = xv[ci] + pa, e = __expf(t);
                u[ci] = (t > 0.0f ? t : e - 1.0f) + pb;
                dd[ci] = t > 0.0f ? 1.0f : e;
            } else {
                u[ci] = xv[ci] + pb;
                dd[ci] = 1.0f;
            }
        }
        if (want_in) {
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                float gu = 0.0f;
#pragma unroll
                for (int co = 0; co < COUT; ++co) gu = __fmaf_rn(w[co][ci], g[co], gu);
                gu *= sc;
                const float gx = gu * dd[ci];
                s_b += gu;
                s_a += gx;
                float *dst = ci < p.C1 ? (p.gx1 ? p.gx1 + ((size_t)b * p.C1 + ci) * S + r : nullptr)
                                       : (p.gx2 ? p.gx2 + ((size_t)b * p.C2 + (ci - p.C1)) * S + r : nullptr);
                if (dst) __stcs(dst, gx);
            }
        }
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            if (want_scale) {
                float raw = 0.0f;
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) raw = __fmaf_rn(w[co][ci], u[ci], raw);
                s_s = __fmaf_rn(g[co], raw, s_s);
            }
            bacc[co] += g[co];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) acc[co][ci] = __fmaf_rn(g[co], u[ci], acc[co][ci]);
        }
    }
    // ---- one reduction per CTA ----
    constexpr int NV = CIN * COUT + COUT + 3;
    float vals[NV];
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) vals[co * CIN + ci] = acc[co][ci];
        vals[CIN * COUT + co] = bacc[co];
    }
    vals[NV - 3] = p.pre_act ? s_a : 0.0f; vals[NV - 2] = s_b; vals[NV - 1] = s_s;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], o);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[warp][i] = vals[i];
    }
    __syncthreads();
    if (tid < NV) {
        float t = 0.0f;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) t += red[wv][tid];
        if (tid < CIN * COUT) {
            if (p.gw) atomicAdd(p.gw + tid, t * sc);
        } else if (tid < CIN * COUT + COUT) {
            if (p.gbias) atomicAdd(p.gbias + (tid - CIN * COUT), t);
            if (p.gscal && p.gscal_want_pb) atomicAdd(p.gscal + 3, t);
        } else if (p.gscal) {
            const int k = tid - (CIN * COUT + COUT);
            if (k == 0 && p.pre_act && p.pre_a) atomicAdd(p.gscal + 0, t);
            if (k == 1 && p.pre_b) atomicAdd(p.gscal + 1, t);
            if (k == 2 && want_scale) atomicAdd(p.gscal + 2, t);
        }
    }
}

template <int CIN>
static int launch_pw_thin(const BwdParams &p, int Cout, int64_t total, int want_scale, void *stream) {
    int64_t blocks = ceil_div(total, 256);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    const dim3 grid((unsigned)blocks), block(256);
    switch (Cout) {
        case 1: return launch("conv1x1_bwd_thin", conv1x1_bwd_thin_kernel<CIN, 1>, grid, block, 0, stream, p, total, want_scale);
        case 2: return launch("conv1x1_bwd_thin", conv1x1_bwd_thin_kernel<CIN, 2>, grid, block, 0, stream, p, total, want_scale);
        case 4: return launch("conv1x1_bwd_thin", conv1x1_bwd_thin_kernel<CIN, 4>, grid, block, 0, stream, p, total, want_scale);
        case 8: return launch("conv1x1_bwd_thin", conv1x1_bwd_thin_kernel<CIN, 8>, grid, block, 0, stream, p, total, want_scale);
        default: return -1;
    }
}

__device__ __forceinline__ void up_taps_b(int o, int n, int &i0, int &i1, float &l1) {
    float src = 0.5f * (float)o - 0.25f;
    if (src < 0.0f) src = 0.0f;
    i0 = (int)src;
    l1 = src - (float)i0;
    i1 = i0 + (i0 < n - 1 ? 1 : 0);
}

// transposed trilinear x2 (align_corners=False): thread = one low-res element, gathers the <= 4^3 hi-res outputs that read it.
// The optional input transform u = pre_act ? ELU(x+a)+b : x+b of vq3d_upsample2x is differentiated as in dgrad.
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const float *__restrict__ gy, const float *__restrict__ x, int64_t BC, int H, int W, int Z, int pre_act,
                      const float *pre_a, const float *pre_b, float *__restrict__ gx, float *gscal) {
    __shared__ float red[32];
    const int Ho = 2 * H, Wo = 2 * W, Zo = 2 * Z;
    const int64_t S = (int64_t)H * W * Z, So = 8 * S;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float gu = 0.0f, g = 0.0f;
    if (i < BC * S) {
        const int64_t bc = i / S;
        int64_t r = i - bc * S;
        const int ih = (int)(r / ((int64_t)W * Z));
        r -= (int64_t)ih * W * Z;
        const int iw = (int)(r / Z), iz = (int)(r - (int64_t)iw * Z);
        const float *gp = gy + bc * So;
        for (int oh = max(2 * ih - 2, 0); oh <= min(2 * ih + 2, Ho - 1); ++oh) {
            int h0, h1; float lh;
            up_taps_b(oh, H, h0, h1, lh);
            const float ch = (h0 == ih ? 1.0f - lh : 0.0f) + (h1 == ih ? lh : 0.0f);
            if (ch == 0.0f) continue;
            for (int ow = max(2 * iw - 2, 0); ow <= min(2 * iw + 2, Wo - 1); ++ow) {
                int w0, w1; float lw;
                up_taps_b(ow, W, w0, w1, lw);
                const float cw = (w0 == iw ? 1.0f - lw : 0.0f) + (w1 == iw ? lw : 0.0f);
                if (cw == 0.0f) continue;
                for (int oz = max(2 * iz - 2, 0); oz <= min(2 * iz + 2, Zo - 1); ++oz) {
                    int z0, z1; float lz;
                    up_taps_b(oz, Z, z0, z1, lz);
                    const float cz = (z0 == iz ? 1.0f - lz : 0.0f) + (z1 == iz ? lz : 0.0f);
                    if (cz == 0.0f) continue;
                    gu = __fmaf_rn(ch * cw * cz, gp[((size_t)oh * Wo + ow) * Zo + oz], gu);
                }
            }
        }
        g = gu;
        if (pre_act) {
            const float t = x[i] + ld_scalar(pre_a, 0.0f);
            g = t > 0.0f ? gu : gu * __expf(t);
        }
        gx[i] = g;
    }
    if (gscal) {
        const float sa = block_sum(pre_act ? g : 0.0f, red);
        const float sb = block_sum(gu, red);
        if (threadIdx.x == 0) {
            if (pre_act && pre_a) atomicAdd(gscal + 0, sa);
            if (pre_b) atomicAdd(gscal + 1, sb);
        }
    }
}

// d/d decoded of mean smooth_l1(mask(ELU(decoded)), x): model.py:120-152
__global__ void __launch_bounds__(256)
huber_elu_mask_bwd_kernel(const float *__restrict__ dec, const float *__restrict__ x, const int *__restrict__ num_valid,
                          const uint8_t *__restrict__ mask_hw, int64_t B, int HW, int Z, const double *__restrict__ count,
                          const float *__restrict__ gloss, float *__restrict__ gdec) {
    const int64_t total = B * (int64_t)HW * Z;
    const float scale = __ldg(gloss) / (float)(*count);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % Z);
        const int64_t r = i / Z;
        const int hw = (int)(r % HW), b = (int)(r / HW);
        float g = 0.0f;
        if ((!mask_hw || mask_hw[hw]) && !(num_valid && z >= num_valid[b])) {
            const float d = dec[i];
            const float loc = elu1(d);
            const float diff = loc - x[i];
            const float dl = fabsf(diff) < 1.0f ? diff : (diff > 0.0f ? 1.0f : -1.0f);
            g = dl * (d > 0.0f ? 1.0f : __expf(d)) * scale;
        }
        gdec[i] = g;
    }
}

__global__ void __launch_bounds__(256)
adam_amsgrad_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, float *__restrict__ vmax,
                    int64_t n, float lr, float b1, float b2, float eps, float bc1, float bc2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        const float vm = fmaxf(vmax[i], vi);
        m[i] = mi; v[i] = vi; vmax[i] = vm;
        p[i] -= (lr / bc1) * mi / (sqrtf(vm) / sqrtf(bc2) + eps);       // torch.optim.Adam(amsgrad=True), model.py:91-93
    }
}

// CUDA-graph friendly variant: the step counter and the two bias corrections live on the device (st[0] = step, st[1] = bc1,
// st[2] = bc2), advanced by a one-thread kernel in front of the update, so a captured step replays correctly
__global__ void adam_advance_kernel(double *st, double b1, double b2) {
    st[0] += 1.0;
    st[1] = 1.0 - pow(b1, st[0]);
    st[2] = 1.0 - pow(b2, st[0]);
}

__global__ void __launch_bounds__(256)
adam_amsgrad_dev_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, float *__restrict__ vmax,
                        int64_t n, float lr, float b1, float b2, float eps, const double *__restrict__ st) {
    const float bc1 = (float)st[1], bc2 = (float)st[2];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        const float vm = fmaxf(vmax[i], vi);
        m[i] = mi; v[i] = vi; vmax[i] = vm;
        p[i] -= (lr / bc1) * mi / (sqrtf(vm) / sqrtf(bc2) + eps);
    }
}

// Weight gradient of a stride-1 "same" convolution from shared-memory tiles: a CTA stages the pre-transformed input halo
// tile (all C_in) and the scaled output-gradient tile (all C_out) of a 4 x 4 x 32 output box once, then every thread
// accumulates its APT (co, ci, tap) products over the box (2 LDS + 1 FMA per product; the generic kernel re-reads both
// operands from global memory for every (co, ci) pair) and adds them to gw with one atomic each.
// The box is 512 voxels whatever the depth: 4 x 4 x 32 for deep tensors, 8 x 8 x 8 and 16 x 16 x 2 for the shallow top levels
// (a 32-deep box on an 8 x 8 x 2 volume would be 15/16 padding).
constexpr int kWgThreads = 256;
template <int TZ> struct WgTile { static constexpr int TH = TZ == 32 ? 4 : (TZ == 8 ? 8 : 16), TW = TH; };

constexpr int kWgCOB = 4, kWgTB = 4;      // accumulator block of a thread: 4 output channels x 4 (input channel, tap) pairs

template <int kWgTZ, int ST>              // ST: convolution stride (1 or 2); the box is a box of OUTPUT voxels
__global__ void __launch_bounds__(kWgThreads)
conv3d_wgrad_tiled_kernel(BwdParams p, int tilesH, int tilesW, int tilesZ, int cic) {
    constexpr int kWgTH = WgTile<kWgTZ>::TH, kWgTW = WgTile<kWgTZ>::TW;
    VQ3D_DYN_SMEM(float, smem);
    // blockIdx.y selects a chunk of `cic` input channels (the accumulators of one chunk fit the CTA's registers and its
    // halo tile fits shared memory); Cin below is the chunk's channel count, ci0 its first channel
    const int CinAll = p.C1 + p.C2, ci0 = blockIdx.y * cic;
    const int Cin = min(cic, CinAll - ci0), k = p.k, k3 = k * k * k, pad = p.pad;
    const int HH = (kWgTH - 1) * ST + k, HW = (kWgTW - 1) * ST + k, HZ = (kWgTZ - 1) * ST + k, HV = HH * HW * HZ, TV = kWgTH * kWgTW * kWgTZ;
    float *su = smem, *sg = smem + (size_t)Cin * HV;
    int tile = blockIdx.x;
    const int tz = tile % tilesZ; tile /= tilesZ;
    const int tw = tile % tilesW; tile /= tilesW;
    const int th = tile % tilesH;
    const int b = tile / tilesH;
    const int h0 = th * kWgTH, w0 = tw * kWgTW, z0 = tz * kWgTZ;
    const int64_t S = (int64_t)p.H * p.W * p.Z, So = (int64_t)p.Ho * p.Wo * p.Zo;
    const float sc = ld_scalar(p.post_scale, 1.0f), pa = ld_scalar(p.pre_a, 0.0f), pb = ld_scalar(p.pre_b, 0.0f);
    for (int i = threadIdx.x; i < Cin * HV; i += kWgThreads) {
        const int ci = i / HV;
        int r = i - ci * HV;
        const int hh = r / (HW * HZ); r -= hh * HW * HZ;
        const int ww = r / HZ, zz = r - ww * HZ;
        int ih = h0 * ST - pad + hh, iw = w0 * ST - pad + ww, iz = z0 * ST - pad + zz;
        if (p.circ) {                      // boxes may reach far past the edge (partial tiles): full modulo
            ih %= p.H; if (ih < 0) ih += p.H;
            iw %= p.W; if (iw < 0) iw += p.W;
            iz %= p.Z; if (iz < 0) iz += p.Z;
        }
        const bool ok = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W && iz >= 0 && iz < p.Z;
        float u = 0.0f;
        if (ok) {
            const int cg = ci0 + ci;
            const float *src = cg < p.C1 ? p.x1 + ((size_t)b * p.C1 + cg) * S : p.x2 + ((size_t)b * p.C2 + (cg - p.C1)) * S;
            u = src[((size_t)ih * p.W + iw) * p.Z + iz];
            u = p.pre_act ? elu1(u + pa) + pb : u + pb;
        }
        su[i] = u;
    }
    for (int i = threadIdx.x; i < p.Cout * TV; i += kWgThreads) {
        const int co = i / TV;
        int r = i - co * TV;
        const int dh = r / (kWgTW * kWgTZ); r -= dh * kWgTW * kWgTZ;
        const int dw = r / kWgTZ, dz = r - dw * kWgTZ;
        const int oh = h0 + dh, ow = w0 + dw, oz = z0 + dz;
        sg[i] = (oh < p.Ho && ow < p.Wo && oz < p.Zo) ? p.gy[((size_t)b * p.Cout + co) * So + ((size_t)oh * p.Wo + ow) * p.Zo + oz] * sc : 0.0f;
    }
    __syncthreads();
    // thread -> (block of kWgCOB output channels, block of kWgTB (input channel, tap) pairs): per voxel kWgCOB + kWgTB shared
    // loads feed kWgCOB * kWgTB FMAs
    const int nrt = Cin * k3;                                   // (ci, tap) pairs of this chunk
    const int nrb = (nrt + kWgTB - 1) / kWgTB;
    const int cb = threadIdx.x / nrb, rb = threadIdx.x - cb * nrb;
    const bool worker = cb * kWgCOB < p.Cout;                    // the host sizes the chunk so that every block has a thread
    int gofs[kWgCOB], uofs[kWgTB];
    float acc[kWgCOB][kWgTB];
#pragma unroll
    for (int jc = 0; jc < kWgCOB; ++jc) {
        const int co = min(cb * kWgCOB + jc, p.Cout - 1);
        gofs[jc] = co * TV;
#pragma unroll
        for (int jt = 0; jt < kWgTB; ++jt) acc[jc][jt] = 0.0f;
    }
#pragma unroll
    for (int jt = 0; jt < kWgTB; ++jt) {
        const int rt = min(rb * kWgTB + jt, nrt - 1);
        const int ci = rt / k3;
        int rem = rt - ci * k3;
        const int kh = rem / (k * k); rem -= kh * k * k;
        const int kw = rem / k, kz = rem - kw * k;
        uofs[jt] = ci * HV + (kh * HW + kw) * HZ + kz;
    }
    if (worker) {
#pragma unroll 1
        for (int dh = 0; dh < kWgTH; ++dh)
#pragma unroll 1
            for (int dw = 0; dw < kWgTW; ++dw) {
                const float *gb = sg + (dh * kWgTW + dw) * kWgTZ, *ub = su + (dh * ST * HW + dw * ST) * HZ;
#pragma unroll(kWgTZ < 4 ? kWgTZ : 4)
                for (int dz = 0; dz < kWgTZ; ++dz) {
                    float gv[kWgCOB], uv[kWgTB];
#pragma unroll
                    for (int jc = 0; jc < kWgCOB; ++jc) gv[jc] = gb[gofs[jc] + dz];
#pragma unroll
                    for (int jt = 0; jt < kWgTB; ++jt) uv[jt] = ub[uofs[jt] + dz * ST];
#pragma unroll
                    for (int jc = 0; jc < kWgCOB; ++jc)
#pragma unroll
                        for (int jt = 0; jt < kWgTB; ++jt) acc[jc][jt] = __fmaf_rn(gv[jc], uv[jt], acc[jc][jt]);
                }
            }
#pragma unroll
        for (int jc = 0; jc < kWgCOB; ++jc) {
            const int co = cb * kWgCOB + jc;
#pragma unroll
            for (int jt = 0; jt < kWgTB; ++jt) {
                const int rt = rb * kWgTB + jt;
                if (co < p.Cout && rt < nrt) atomicAdd(p.gw + ((size_t)co * CinAll + ci0) * k3 + rt, acc[jc][jt]);     // (co, local ci, tap) -> the weight's layout
            }
        }
    }
}

// Weight gradient of k3 s1 "same" convolutions with few output channels (the 9-, 4-, 2- and 1-channel conv2 of the residual
// branches: the bulk of a training step's conv3d_backward time).  Same staging as conv3d_wgrad_tiled_kernel (input halo tile +
// scaled output-gradient tile of a 4 x 4 x 32 box), different work split: a thread owns one (ci, kh, kw) and a subset of the 16
// (dh, dw) rows, slides a three-value window along z (ONE new shared-memory load per voxel) and accumulates all COUT x 3 kz
// products -- 1 + COUT loads feed 3 COUT FMAs, the output-gradient loads are broadcasts, and the window loads of a warp fall
// into different banks (the tiled kernel's (co, tap)-block split needs 8 loads per 16 FMAs and its window loads conflict).
// The groups' partial sums meet in shared memory; one global atomic per (ci, tap, co) and CTA.
template <int COUT>
__global__ void __launch_bounds__(kWgThreads)
conv3d_wgrad_rows_kernel(BwdParams p, int tilesH, int tilesW, int tilesZ, int cic) {
    constexpr int TH = 4, TW = 4, TZ = 32, HH = TH + 2, HW = TW + 2, HZ = TZ + 2, HV = HH * HW * HZ, TV = TH * TW * TZ;
    VQ3D_DYN_SMEM(float, smem);
    const int CinAll = p.C1 + p.C2, ci0 = blockIdx.y * cic;
    const int Cin = min(cic, CinAll - ci0);
    float *su = smem, *sg = smem + (size_t)Cin * HV;
    int tile = blockIdx.x;
    const int tz = tile % tilesZ; tile /= tilesZ;
    const int tw = tile % tilesW; tile /= tilesW;
    const int th = tile % tilesH;
    const int b = tile / tilesH;
    const int h0 = th * TH, w0 = tw * TW, z0 = tz * TZ;
    const int64_t S = (int64_t)p.H * p.W * p.Z;
    const float sc = ld_scalar(p.post_scale, 1.0f), pa = ld_scalar(p.pre_a, 0.0f), pb = ld_scalar(p.pre_b, 0.0f);
    for (int i = threadIdx.x; i < Cin * HV; i += kWgThreads) {
        const int ci = i / HV;
        int r = i - ci * HV;
        const int hh = r / (HW * HZ); r -= hh * HW * HZ;
        const int ww = r / HZ, zz = r - ww * HZ;
        int ih = h0 - 1 + hh, iw = w0 - 1 + ww, iz = z0 - 1 + zz;
        if (p.circ) {                      // boxes may reach far past the edge (partial tiles): full modulo
            ih %= p.H; if (ih < 0) ih += p.H;
            iw %= p.W; if (iw < 0) iw += p.W;
            iz %= p.Z; if (iz < 0) iz += p.Z;
        }
        const bool ok = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W && iz >= 0 && iz < p.Z;
        float u = 0.0f;
        if (ok) {
            const int cg = ci0 + ci;
            const float *src = cg < p.C1 ? p.x1 + ((size_t)b * p.C1 + cg) * S : p.x2 + ((size_t)b * p.C2 + (cg - p.C1)) * S;
            u = src[((size_t)ih * p.W + iw) * p.Z + iz];
            u = p.pre_act ? elu1(u + pa) + pb : u + pb;
        }
        su[i] = u;
    }
    for (int i = threadIdx.x; i < COUT * TV; i += kWgThreads) {
        const int co = i / TV;
        int r = i - co * TV;
        const int dh = r / (TW * TZ); r -= dh * TW * TZ;
        const int dw = r / TZ, dz = r - dw * TZ;
        const int oh = h0 + dh, ow = w0 + dw, oz = z0 + dz;
        sg[i] = (oh < p.H && ow < p.W && oz < p.Z) ? p.gy[((size_t)b * COUT + co) * S + ((size_t)oh * p.W + ow) * p.Z + oz] * sc : 0.0f;
    }
    __syncthreads();
    const int NI = Cin * 9;                              // (ci, kh, kw) items; the host keeps NI <= kWgThreads
    const int G = min(kWgThreads / NI, TH * TW);          // row groups
    const int item = threadIdx.x % NI, grp = threadIdx.x / NI;
    float acc[COUT][3];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co][0] = acc[co][1] = acc[co][2] = 0.0f;
    const int ci = item / 9, kh = (item % 9) / 3, kw = item % 3;
    if (grp < G) {
#pragma unroll 1
        for (int row = grp; row < TH * TW; row += G) {
            const int dh = row / TW, dw = row - dh * TW;
            const float *ub = su + (size_t)ci * HV + ((dh + kh) * HW + dw + kw) * HZ;
            const float *gb = sg + row * TZ;
            float u0 = ub[0], u1 = ub[1];
#pragma unroll 4
            for (int dz = 0; dz < TZ; ++dz) {
                const float u2 = ub[dz + 2];
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    const float g = gb[co * TV + dz];
                    acc[co][0] = __fmaf_rn(g, u0, acc[co][0]);
                    acc[co][1] = __fmaf_rn(g, u1, acc[co][1]);
                    acc[co][2] = __fmaf_rn(g, u2, acc[co][2]);
                }
                u0 = u1; u1 = u2;
            }
        }
    }
    __syncthreads();                                      // the tiles have been read: their memory holds the groups' sums now
    float *sacc = smem;                                   // [NI][COUT * 3]
    for (int i = threadIdx.x; i < NI * COUT * 3; i += kWgThreads) sacc[i] = 0.0f;
    __syncthreads();
    if (grp < G) {
#pragma unroll
        for (int co = 0; co < COUT; ++co)
#pragma unroll
            for (int kz = 0; kz < 3; ++kz) atomicAdd(&sacc[(item * COUT + co) * 3 + kz], acc[co][kz]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NI * COUT * 3; i += kWgThreads) {
        const int kz = i % 3, co = (i / 3) % COUT, it = i / (3 * COUT);
        const int cl = it / 9, hw = it % 9;
        atomicAdd(p.gw + ((size_t)co * CinAll + ci0 + cl) * 27 + hw * 3 + kz, sacc[i]);
    }
}

// grid (chunks, Cin, B): finish of the forward-convolution form of dgrad (see vq3d_conv3d_dgrad_finish)
__global__ void __launch_bounds__(256)
conv3d_dgrad_finish_kernel(BwdParams p, const float *__restrict__ gu_all) {
    __shared__ float red[32];
    const int ci = blockIdx.y, b = blockIdx.z, Cin = p.C1 + p.C2;
    const int64_t S = (int64_t)p.H * p.W * p.Z;
    const float sc = ld_scalar(p.post_scale, 1.0f), pa = ld_scalar(p.pre_a, 0.0f);
    const bool first = ci < p.C1;
    const float *xs = first ? p.x1 + ((size_t)b * p.C1 + ci) * S : p.x2 + ((size_t)b * p.C2 + (ci - p.C1)) * S;
    float *gxs = first ? (p.gx1 ? p.gx1 + ((size_t)b * p.C1 + ci) * S : nullptr) : (p.gx2 ? p.gx2 + ((size_t)b * p.C2 + (ci - p.C1)) * S : nullptr);
    const float *gsrc = gu_all + ((size_t)b * Cin + ci) * S;
    float sa = 0.0f, sb = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
        const float gu = gsrc[i] * sc;
        float gx = gu;
        if (p.pre_act) {
            const float t = xs[i] + pa;
            gx = t > 0.0f ? gu : gu * __expf(t);
        }
        if (gxs) gxs[i] = gx;
        sa += p.pre_act ? gx : 0.0f;
        sb += gu;
    }
    if (p.gscal) {
        const float ta = block_sum(sa, red);
        const float tb = block_sum(sb, red);
        if (threadIdx.x == 0) {
            if (p.pre_act && p.pre_a) atomicAdd(p.gscal + 0, ta);
            if (p.pre_b) atomicAdd(p.gscal + 1, tb);
        }
    }
}

// trailing ELU from its output: d ELU(u) / d u = 1 (u > 0), exp(u) = y + 1 (u <= 0)
__global__ void __launch_bounds__(256)
elu_backward_kernel(const float *__restrict__ gy, const float *__restrict__ y, float *__restrict__ gx, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float yv = y[i];
        gx[i] = yv > 0.0f ? gy[i] : gy[i] * (yv + 1.0f);
    }
}

}  // namespace vq3d

using namespace vq3d;

extern "C" int vq3d_conv3d_backward(const vq3d_conv_desc *d, const vq3d_conv_bwd *g, void *stream) {
    if (!d || !g || !g->gy) return fail(VQ3D_ERR_INVALID, "conv3d_backward: null descriptor / gy");
    if (!d->x1 || !d->w) return fail(VQ3D_ERR_INVALID, "conv3d_backward: null x1/w");
    if (d->post_act) return fail(VQ3D_ERR_UNSUPPORTED, "conv3d_backward: post_act (FixupResBlock) has no backward in this build");
    if (d->k < 1 || d->k > 4 || d->stride < 1 || d->stride > 2 || d->pad < 0 || d->pad >= d->k) return fail(VQ3D_ERR_INVALID, "conv3d_backward: unsupported geometry");
    BwdParams p;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z; p.C1 = d->C1; p.C2 = d->C2; p.Cout = d->Cout;
    p.k = d->k; p.stride = d->stride; p.pad = d->pad; p.circ = d->pad_circular; p.pre_act = d->pre_act;
    p.Ho = (d->H + 2 * d->pad - d->k) / d->stride + 1;
    p.Wo = (d->W + 2 * d->pad - d->k) / d->stride + 1;
    p.Zo = (d->Z + 2 * d->pad - d->k) / d->stride + 1;
    p.x1 = d->x1; p.x2 = d->x2; p.w = d->w; p.pre_a = d->pre_a; p.pre_b = d->pre_b; p.post_scale = d->post_scale;
    p.gy = g->gy; p.raw = g->raw; p.gx1 = g->gx1; p.gx2 = g->gx2; p.gw = g->gw; p.gbias = g->gbias; p.gscal = g->gscalars;
    const int Cin = d->C1 + d->C2;
    const int64_t S = (int64_t)d->H * d->W * d->Z, So = (int64_t)p.Ho * p.Wo * p.Zo;
    int rc;
    const bool s2_k4 = d->stride == 2 && d->k == 4 && d->pad == 1, s2_k2 = d->stride == 2 && d->k == 2 && d->pad == 0;
    const size_t s2_smem = (size_t)d->Cout * Cin * d->k * d->k * d->k * sizeof(float);
    if (!g->skip_input_grads && (p.gx1 || p.gx2 || (p.gscal && (d->pre_a || d->pre_b))) && (s2_k4 || s2_k2) && s2_smem <= 96 * 1024 &&
        d->H % 2 == 0 && d->W % 2 == 0 && d->Z % 2 == 0) {
        int64_t blocks = ceil_div((int64_t)d->B * S, 256);
        if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
        rc = s2_k4 ? launch("conv3d_dgrad_s2", conv3d_dgrad_s2_kernel<4>, dim3((unsigned)blocks), dim3(256), s2_smem, stream, p)
                   : launch("conv3d_dgrad_s2", conv3d_dgrad_s2_kernel<2>, dim3((unsigned)blocks), dim3(256), s2_smem, stream, p);
        if (rc) return rc;
    } else if (!g->skip_input_grads && (p.gx1 || p.gx2 || (p.gscal && (d->pre_a || d->pre_b)))) {
        rc = launch("conv3d_dgrad", conv3d_dgrad_kernel, dim3((unsigned)ceil_div((int64_t)d->B * S, 128), (unsigned)Cin), dim3(128), 0, stream, p);
        if (rc) return rc;
    }
    // tiled weight gradient: input channels in chunks whose accumulators (<= 16 per thread) and halo tile (<= 150 KB with the
    // output-gradient tile) fit one CTA
    const int k3 = d->k * d->k * d->k;
    const int wtz = p.Zo >= 32 ? 32 : (p.Zo >= 8 ? 8 : 2);
    const int wth = wtz == 32 ? 4 : (wtz == 8 ? 8 : 16);
    const int st = d->stride;
    const size_t wg_halo = (size_t)((wth - 1) * st + d->k) * ((wth - 1) * st + d->k) * ((wtz - 1) * st + d->k) * sizeof(float);
    const size_t wg_gtile = (size_t)d->Cout * 512 * sizeof(float);
    const int co_blocks = (d->Cout + kWgCOB - 1) / kWgCOB;
    int cic = co_blocks <= kWgThreads ? ((kWgThreads / co_blocks) * kWgTB) / k3 : 0;     // (ci, tap) blocks per output-channel block
    if (cic > Cin) cic = Cin;
    while (cic > 0 && wg_gtile + (size_t)cic * wg_halo > 150 * 1024) --cic;
    // k3 s1 with few output channels on deep tensors: the row-sliding kernel (input channels in chunks of <= 28 = 256 / 9 items)
    int ric = Cin < kWgThreads / 9 ? Cin : kWgThreads / 9;
    while (ric > 0 && wg_gtile + (size_t)ric * wg_halo > 150 * 1024) --ric;
    const bool rows_ok = p.gw && d->k == 3 && st == 1 && d->pad == 1 && wtz == 32 && ric > 0 && getenv("VQ3D_WGRAD_TILED") == nullptr &&
                         (d->Cout == 9 || d->Cout == 4 || d->Cout == 2 || d->Cout == 1 || d->Cout == 8) &&
                         (!d->pad_circular || (d->H >= 3 && d->W >= 3 && d->Z >= 3));
    if (rows_ok) {
        const int tH = (int)ceil_div(p.Ho, 4), tW = (int)ceil_div(p.Wo, 4), tZ = (int)ceil_div(p.Zo, 32);
        const dim3 grid((unsigned)((int64_t)d->B * tH * tW * tZ), (unsigned)ceil_div(Cin, ric));
        size_t rsmem = wg_gtile + (size_t)ric * wg_halo;
        const size_t racc = (size_t)ric * 9 * d->Cout * 3 * sizeof(float);
        if (rsmem < racc) rsmem = racc;
#define VQ3D_WR_LAUNCH(CO) launch("conv3d_wgrad_rows", conv3d_wgrad_rows_kernel<CO>, grid, dim3(kWgThreads), rsmem, stream, p, tH, tW, tZ, ric)
        rc = d->Cout == 9 ? VQ3D_WR_LAUNCH(9) : (d->Cout == 8 ? VQ3D_WR_LAUNCH(8) : (d->Cout == 4 ? VQ3D_WR_LAUNCH(4) : (d->Cout == 2 ? VQ3D_WR_LAUNCH(2) : VQ3D_WR_LAUNCH(1))));
#undef VQ3D_WR_LAUNCH
        if (rc) return rc;
    } else
    if (p.gw && cic > 0 && (!d->pad_circular || (d->H >= d->k && d->W >= d->k && d->Z >= d->k))) {
        const int tH = (int)ceil_div(p.Ho, wth), tW = (int)ceil_div(p.Wo, wth), tZ = (int)ceil_div(p.Zo, wtz);
        const dim3 grid((unsigned)((int64_t)d->B * tH * tW * tZ), (unsigned)ceil_div(Cin, cic));
        const size_t wg_smem = wg_gtile + (size_t)cic * wg_halo;
#define VQ3D_WG_LAUNCH(TZ, ST) launch("conv3d_wgrad_tiled", conv3d_wgrad_tiled_kernel<TZ, ST>, grid, dim3(kWgThreads), wg_smem, stream, p, tH, tW, tZ, cic)
#define VQ3D_WG_BY_TZ(ST) (wtz == 32 ? VQ3D_WG_LAUNCH(32, ST) : (wtz == 8 ? VQ3D_WG_LAUNCH(8, ST) : VQ3D_WG_LAUNCH(2, ST)))
        rc = st == 1 ? VQ3D_WG_BY_TZ(1) : VQ3D_WG_BY_TZ(2);
#undef VQ3D_WG_BY_TZ
#undef VQ3D_WG_LAUNCH
        if (rc) return rc;
    } else if (p.gw) {
        int64_t chunks = ceil_div((int64_t)d->B * So, 128 * 8);
        const int64_t cap = (int64_t)kNumSMs * 8 / ((int64_t)Cin * d->Cout) + 1;
        if (chunks > cap) chunks = cap;
        rc = launch("conv3d_wgrad", conv3d_wgrad_kernel, dim3((unsigned)chunks, (unsigned)Cin, (unsigned)d->Cout), dim3(128), 0, stream, p);
        if (rc) return rc;
    }
    const int want_scale = d->post_scale != nullptr, want_pb = d->post_b != nullptr;
    if (p.gbias || (p.gscal && (want_scale || want_pb))) {
        if (want_scale && p.gscal && !g->raw) return fail(VQ3D_ERR_INVALID, "conv3d_backward: d post_scale needs the raw convolution output");
        int64_t chunks = ceil_div(So, 256 * 8);
        if (chunks > 256) chunks = 256;
        rc = launch("conv3d_outgrads", conv3d_outgrads_kernel, dim3((unsigned)chunks, (unsigned)d->Cout, (unsigned)d->B), dim3(256), 0, stream,
                    g->gy, g->raw, d->Cout, So, g->gbias, g->gscalars, want_scale, want_pb);
        if (rc) return rc;
    }
    return VQ3D_OK;
}

extern "C" int vq3d_conv1x1_backward(const vq3d_conv_desc *d, const vq3d_conv_bwd *g, void *stream) {
    if (!d || !g || !g->gy) return fail(VQ3D_ERR_INVALID, "conv1x1_backward: null descriptor / gy");
    if (!d->x1 || !d->w) return fail(VQ3D_ERR_INVALID, "conv1x1_backward: null x1/w");
    if (d->k != 1 || d->stride != 1 || d->pad != 0) return fail(VQ3D_ERR_INVALID, "conv1x1_backward: pointwise (k1 s1 p0) convolutions only");
    if (d->post_act) return fail(VQ3D_ERR_UNSUPPORTED, "conv1x1_backward: post_act has no backward in this build");
    if (d->C2 > 0 && !d->x2) return fail(VQ3D_ERR_INVALID, "conv1x1_backward: C2 > 0 but x2 is NULL");
    const int Cin = d->C1 + d->C2;
    if (Cin * d->Cout + d->Cout > kPwMaxRows * kPwT) return fail(VQ3D_ERR_UNSUPPORTED, "conv1x1_backward: %d x %d channels exceed the fused kernel", Cin, d->Cout);
    const size_t smem = ((size_t)d->Cout * ((Cin + 3) & ~3) + (size_t)Cin * ((d->Cout + 3) & ~3) + (size_t)(2 * Cin + d->Cout) * kPwLD) * sizeof(float);
    if (smem > 200 * 1024) return fail(VQ3D_ERR_UNSUPPORTED, "conv1x1_backward: channel counts exceed shared memory");
    BwdParams p = {};
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z; p.C1 = d->C1; p.C2 = d->C2; p.Cout = d->Cout;
    p.k = 1; p.stride = 1; p.pad = 0; p.circ = 0; p.pre_act = d->pre_act;
    p.Ho = d->H; p.Wo = d->W; p.Zo = d->Z;
    p.x1 = d->x1; p.x2 = d->x2; p.w = d->w; p.pre_a = d->pre_a; p.pre_b = d->pre_b; p.post_scale = d->post_scale;
    p.gy = g->gy; p.raw = nullptr; p.gx1 = g->gx1; p.gx2 = g->gx2; p.gw = g->gw; p.gbias = g->gbias; p.gscal = g->gscalars;
    p.gscal_want_pb = d->post_b != nullptr;
    const int64_t total = (int64_t)d->B * d->H * d->W * d->Z;
    {   // thin layers: streaming kernel, pair sums in registers
        const int ws = (int)(d->post_scale != nullptr && g->gscalars != nullptr);
        int rc = -1;
        if (Cin * d->Cout <= 32)                 // 8 x 8 needs 255 registers: left to the tiled kernel
        switch (Cin) {
            case 1: rc = launch_pw_thin<1>(p, d->Cout, total, ws, stream); break;
            case 2: rc = launch_pw_thin<2>(p, d->Cout, total, ws, stream); break;
            case 4: rc = launch_pw_thin<4>(p, d->Cout, total, ws, stream); break;
            case 8: rc = launch_pw_thin<8>(p, d->Cout, total, ws, stream); break;
            default: break;
        }
        if (rc >= 0) return rc;
    }
    const int64_t ntiles = ceil_div(total, kPwT);
    int per_sm = (int)(200 * 1024 / (smem > 1024 ? smem : 1024));
    if (per_sm > 12) per_sm = 12;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)kNumSMs * per_sm;
    if (grid > ntiles) grid = ntiles;
    const int rows = (Cin * d->Cout + d->Cout + kPwT - 1) / kPwT, want_scale = (int)(d->post_scale != nullptr && g->gscalars != nullptr);
#define VQ3D_PW_LAUNCH(NR) launch("conv1x1_bwd", conv1x1_bwd_kernel<NR>, dim3((unsigned)grid), dim3(kPwT), smem, stream, p, total, ntiles, want_scale)
    const int rc = rows <= 2 ? VQ3D_PW_LAUNCH(2) : (rows <= 6 ? VQ3D_PW_LAUNCH(6) : (rows <= 12 ? VQ3D_PW_LAUNCH(12) : VQ3D_PW_LAUNCH(kPwMaxRows)));
#undef VQ3D_PW_LAUNCH
    return rc;
}

extern "C" int vq3d_conv3d_dgrad_finish(const vq3d_conv_desc *d, const float *gu_all, float *gx1, float *gx2, float *gscalars, void *stream) {
    if (!d || !gu_all || !d->x1) return fail(VQ3D_ERR_INVALID, "conv3d_dgrad_finish: null descriptor / gu_all / x1");
    if (d->stride != 1 || (d->k & 1) == 0 || d->pad != (d->k - 1) / 2) return fail(VQ3D_ERR_INVALID, "conv3d_dgrad_finish: stride-1 same convolutions only");
    if (d->C2 > 0 && !d->x2) return fail(VQ3D_ERR_INVALID, "conv3d_dgrad_finish: C2 > 0 but x2 is NULL");
    BwdParams p = {};
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z; p.C1 = d->C1; p.C2 = d->C2; p.Cout = d->Cout;
    p.k = d->k; p.stride = 1; p.pad = d->pad; p.circ = d->pad_circular; p.pre_act = d->pre_act;
    p.Ho = d->H; p.Wo = d->W; p.Zo = d->Z;
    p.x1 = d->x1; p.x2 = d->x2; p.w = d->w; p.pre_a = d->pre_a; p.pre_b = d->pre_b; p.post_scale = d->post_scale;
    p.gx1 = gx1; p.gx2 = gx2; p.gscal = gscalars;
    const int64_t S = (int64_t)d->H * d->W * d->Z;
    int64_t chunks = ceil_div(S, 256 * 4);
    if (chunks > 1024) chunks = 1024;
    return launch("conv3d_dgrad_finish", conv3d_dgrad_finish_kernel, dim3((unsigned)chunks, (unsigned)(d->C1 + d->C2), (unsigned)d->B), dim3(256), 0,
                  stream, p, gu_all);
}

extern "C" int vq3d_upsample2x_backward(const float *gy, const float *x, int64_t B, int C, int H, int W, int Z, int pre_act,
                                        const float *pre_a, const float *pre_b, float *gx, float *gscalars, void *stream) {
    if (!gy || !gx || B < 1 || C < 1 || H < 1 || W < 1 || Z < 1) return fail(VQ3D_ERR_INVALID, "upsample2x_backward: bad arguments");
    if (pre_act && !x) return fail(VQ3D_ERR_INVALID, "upsample2x_backward: pre_act needs x");
    const int64_t total = B * C * (int64_t)H * W * Z;
    return launch("upsample2x_bwd", upsample2x_bwd_kernel, dim3((unsigned)ceil_div(total, 256)), dim3(256), 0, stream, gy, x, B * C, H, W, Z,
                  pre_act, pre_a, pre_b, gx, gscalars);
}

extern "C" int vq3d_huber_elu_mask_backward(const float *decoded, const float *x, const int32_t *num_valid, const uint8_t *mask_hw,
                                            int64_t B, int H, int W, int Z, const double *count, const float *grad_loss, float *grad_decoded,
                                            void *stream) {
    if (!decoded || !x || !count || !grad_loss || !grad_decoded || B < 1 || H < 1 || W < 1 || Z < 1) return fail(VQ3D_ERR_INVALID, "huber backward: bad arguments");
    const int64_t total = B * (int64_t)H * W * Z;
    int64_t blocks = ceil_div(total, 256 * 4);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    return launch("huber_elu_mask_bwd", huber_elu_mask_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, decoded, x, (const int *)num_valid,
                  mask_hw, B, H * W, Z, count, grad_loss, grad_decoded);
}

extern "C" int vq3d_adam_amsgrad_step_dev(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *max_exp_avg_sq, int64_t n,
                                         double lr, double beta1, double beta2, double eps, double *step_state, void *stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || !max_exp_avg_sq || !step_state || n < 0) return fail(VQ3D_ERR_INVALID, "adam_amsgrad_step_dev: bad arguments");
    int rc = launch("adam_advance", adam_advance_kernel, dim3(1), dim3(1), 0, stream, step_state, beta1, beta2);
    if (rc || n == 0) return rc;
    int64_t blocks = ceil_div(n, 256 * 4);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    return launch("adam_amsgrad_dev", adam_amsgrad_dev_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, param, grad, exp_avg, exp_avg_sq,
                  max_exp_avg_sq, n, (float)lr, (float)beta1, (float)beta2, (float)eps, (const double *)step_state);
}

extern "C" int vq3d_elu_backward(const float *gy, const float *y, float *gx, int64_t n, void *stream) {
    if (!gy || !y || !gx || n < 0) return fail(VQ3D_ERR_INVALID, "elu_backward: bad arguments");
    if (n == 0) return VQ3D_OK;
    int64_t blocks = ceil_div(n, 256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    return launch("elu_backward", elu_backward_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, gy, y, gx, n);
}

extern "C" int vq3d_adam_amsgrad_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *max_exp_avg_sq, int64_t n,
                                      double lr, double beta1, double beta2, double eps, int64_t step, void *stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || !max_exp_avg_sq || n < 0 || step < 1) return fail(VQ3D_ERR_INVALID, "adam: bad arguments");
    if (n == 0) return VQ3D_OK;
    int64_t blocks = ceil_div(n, 256 * 4);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    return launch("adam_amsgrad", adam_amsgrad_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq,
                  n, (float)lr, (float)beta1, (float)beta2, (float)eps, (float)bc1, (float)bc2);
}
