// evonorm_kernels.cu -- EvoNorm3D-S0 (vqvae/evonorm.py:12-26,59-76) for `--block-type evonorm`:
//   y = x * sigmoid(v_c * x) * gamma_c / sqrt(var_g(x) + eps) + beta_c
// var_g = UNBIASED variance over the (C/groups, H, W, Z) elements of channel group g, groups =
// max(C // 8, 1); the reference only supports batch 1 (evonorm.py:24 reshapes std to (1, C, 1, 1, 1)).
// Two launches: group statistics (double accumulation: one pass over x) and the elementwise apply.
#include "vq3d_rt.h"

namespace vq3d {

// grid (chunks, groups): sum and sum of squares of one slice of a group's contiguous [cpg * S] elements
__global__ void __launch_bounds__(256)
evonorm_sums_kernel(const float *__restrict__ x, int64_t group_elems, double *__restrict__ scratch) {
    __shared__ double red1[32], red2[32];
    const float *p = x + (size_t)blockIdx.y * group_elems;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < group_elems; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)p[i];
        s1 += v;
        s2 += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red1[warp] = s1; red2[warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red1[w]; b += red2[w]; }
        atomicAdd(&scratch[2 * blockIdx.y], a);
        atomicAdd(&scratch[2 * blockIdx.y + 1], b);
    }
}

__global__ void evonorm_finish_kernel(const double *__restrict__ scratch, int C, int cpg, double n, double eps, float *__restrict__ std_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        const int g = c / cpg;
        const double mean = scratch[2 * g] / n;
        double var = (scratch[2 * g + 1] - scratch[2 * g] * mean) / (n - 1.0);     // torch.var: unbiased
        if (var < 0.0) var = 0.0;
        std_out[c] = (float)sqrt((double)(float)var + eps);
    }
}

__global__ void __launch_bounds__(256)
evonorm_apply_kernel(const float *__restrict__ x, const float *__restrict__ v, const float *__restrict__ gamma,
                     const float *__restrict__ beta, const float *__restrict__ stdv, int64_t S, float *__restrict__ y) {
    const int c = blockIdx.y;
    const float vc = __ldg(v + c), gc = __ldg(gamma + c), bc = __ldg(beta + c), sd = __ldg(stdv + c);
    const float *px = x + (size_t)c * S;
    float *py = y + (size_t)c * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
        const float xv = px[i];
        const float sig = 1.0f / (1.0f + __expf(-xv * vc));
        py[i] = __fdiv_rn(xv * sig * gc, sd) + bc;      // num * gamma / std + beta, evonorm.py:76
    }
}

// grid (chunks, C): per-channel sums for the backward (double accumulation)
__global__ void __launch_bounds__(256)
evonorm_bwd_sums_kernel(const float *__restrict__ x, const float *__restrict__ gy, const float *__restrict__ v, int64_t S,
                        double *__restrict__ sums) {
    __shared__ double red[3][32];
    const int c = blockIdx.y;
    const float vc = __ldg(v + c);
    const float *px = x + (size_t)c * S, *pg = gy + (size_t)c * S;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
        const float xv = px[i], g = pg[i];
        const float sig = 1.0f / (1.0f + __expf(-xv * vc));
        s0 += (double)g;
        s1 += (double)(g * xv * sig);
        s2 += (double)(g * xv * xv * sig * (1.0f - sig));
    }
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, d = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; d += red[2][w]; }
        atomicAdd(&sums[3 * c], a);
        atomicAdd(&sums[3 * c + 1], b);
        atomicAdd(&sums[3 * c + 2], d);
    }
}

__global__ void __launch_bounds__(256)
evonorm_bwd_apply_kernel(const float *__restrict__ x, const float *__restrict__ gy, const float *__restrict__ v,
                         const float *__restrict__ coef_a, const float *__restrict__ coef_b, const float *__restrict__ mean,
                         int64_t S, float *__restrict__ gx) {
    const int c = blockIdx.y;
    const float vc = __ldg(v + c), ca = __ldg(coef_a + c), cb = __ldg(coef_b + c), mu = __ldg(mean + c);
    const float *px = x + (size_t)c * S, *pg = gy + (size_t)c * S;
    float *po = gx + (size_t)c * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
        const float xv = px[i];
        const float sig = 1.0f / (1.0f + __expf(-xv * vc));
        po[i] = pg[i] * ca * (sig + xv * vc * sig * (1.0f - sig)) + cb * (xv - mu);      // evonorm.py:41-44 + the group-std term
    }
}

}  // namespace vq3d

using namespace vq3d;

extern "C" int vq3d_evonorm_s0_backward_sums(const float *x, const float *gy, const float *v, int C, int64_t S, double *sums, void *stream) {
    if (!x || !gy || !v || !sums || C < 1 || C > 65535 || S < 1) return fail(VQ3D_ERR_INVALID, "evonorm_s0_backward_sums: bad arguments");
    int rc = check_cuda(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * C, (cudaStream_t)stream), "evonorm_s0_backward_sums memset");
    if (rc) return rc;
    int64_t chunks = ceil_div(S, 256 * 16);
    if (chunks > 1024) chunks = 1024;
    return launch("evonorm_bwd_sums", evonorm_bwd_sums_kernel, dim3((unsigned)chunks, (unsigned)C), dim3(256), 0, stream, x, gy, v, S, sums);
}

extern "C" int vq3d_evonorm_s0_backward_apply(const float *x, const float *gy, const float *v, const float *coef_a, const float *coef_b,
                                              const float *mean, int C, int64_t S, float *gx, void *stream) {
    if (!x || !gy || !v || !coef_a || !coef_b || !mean || !gx || C < 1 || C > 65535 || S < 1)
        return fail(VQ3D_ERR_INVALID, "evonorm_s0_backward_apply: bad arguments");
    int64_t chunks = ceil_div(S, 256 * 4);
    if (chunks > 4096) chunks = 4096;
    return launch("evonorm_bwd_apply", evonorm_bwd_apply_kernel, dim3((unsigned)chunks, (unsigned)C), dim3(256), 0, stream,
                  x, gy, v, coef_a, coef_b, mean, S, gx);
}

extern "C" int vq3d_evonorm_s0_stats(const float *x, int C, int64_t S, int groups, double eps, double *scratch, float *std_out, void *stream) {
    if (!x || !scratch || !std_out || C < 1 || S < 1 || groups < 1 || C % groups != 0) return fail(VQ3D_ERR_INVALID, "evonorm_s0_stats: bad arguments");
    const int cpg = C / groups;
    const int64_t ge = (int64_t)cpg * S;
    if (ge < 2) return fail(VQ3D_ERR_INVALID, "evonorm_s0_stats: unbiased variance needs at least 2 elements per group");
    int rc = check_cuda(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * groups, (cudaStream_t)stream), "evonorm_s0_stats memset");
    if (rc) return rc;
    int64_t chunks = ceil_div(ge, 256 * 16);
    if (chunks > 2048) chunks = 2048;
    rc = launch("evonorm_sums", evonorm_sums_kernel, dim3((unsigned)chunks, (unsigned)groups), dim3(256), 0, stream, x, ge, scratch);
    if (rc) return rc;
    return launch("evonorm_finish", evonorm_finish_kernel, dim3((unsigned)ceil_div(C, 128)), dim3(128), 0, stream,
                  (const double *)scratch, C, cpg, (double)ge, eps, std_out);
}

extern "C" int vq3d_evonorm_s0_apply(const float *x, const float *v, const float *gamma, const float *beta, const float *std_in,
                                     int C, int64_t S, float *y, void *stream) {
    if (!x || !v || !gamma || !beta || !std_in || !y || C < 1 || C > 65535 || S < 1) return fail(VQ3D_ERR_INVALID, "evonorm_s0_apply: bad arguments");
    int64_t chunks = ceil_div(S, 256 * 4);
    if (chunks > 4096) chunks = 4096;
    return launch("evonorm_apply", evonorm_apply_kernel, dim3((unsigned)chunks, (unsigned)C), dim3(256), 0, stream, x, v, gamma, beta, std_in, S, y);
}
