/*
 * vqvae3d_b200.h -- C ABI of libvqvae3d_b200.so: sm_100a kernels for the 3D VQ-VAE-2
 * encode -> quantize -> decode hot path of sara-nl/3D-VQ-VAE-2.
 *
 * The reference has no FFI layer: its boundary is the Python nn.Module API of
 * vqvae/layers.py / vqvae/evonorm.py.  Each entry point below states which
 * reference function (file:line under the reference repo) it replaces.  The Python mirror
 * of those modules (3d-vq-vae-2_b200/vqvae/) binds this library with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated otherwise; the caller owns all memory;
 *  - activations are fp32, contiguous, in the reference's layout (B, C, H, W, Z) with the
 *    CT depth axis Z innermost; S = H*W*Z;
 *  - convolution weights are the reference's (C_out, C_in, k, k, k) fp32 tensors, unpacked;
 *  - the Fixup scalars (bias1a ... bias4, scale: nn.Parameter of shape (1,)) are passed as
 *    device pointers to one float; NULL means "absent" (0 for biases, 1 for scale);
 *  - `stream` is a cudaStream_t passed as void* (0 = default stream); calls only enqueue work;
 *  - return value 0 = success; non-zero = error, message in vq3d_last_error() (thread-local).
 *    No global state, re-entrant per stream.
 */
#ifndef VQVAE3D_B200_H
#define VQVAE3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQ3D_ABI_VERSION 13

int vq3d_abi_version(void);
const char *vq3d_last_error(void);
/* 1 if the library was built from the CUDA sources for sm_100a (always, for the product). */
int vq3d_is_cuda_build(void);

/* ------------------------------------------------------------------------------------
 * Quantizer (vqvae/layers.py:602-728)
 * ------------------------------------------------------------------------------------ */

/*
 * Fused Quantizer.forward body, layers.py:687-720 (+ the one-hot sums of _update_ema,
 * :638-643, when counts/dw are given):
 *   idx[b,s]   = first k minimising sqrt(sum_d (x[b,d,s]-embed[k,d])^2), fp32, in the
 *                summation order of torch.cdist(compute_mode='donot_use_mm_for_euclid_dist')
 *   quant[b,d,s] = x + (embed[idx[b,s], d] - x) (straight-through value, layers.py:720, same two
 *                roundings; the codebook BEFORE this step's EMA update)
 *   *sqerr     = sum (embed[idx] - x)^2             (double accumulator; zeroed by the call itself)
 *   counts[k] += #{idx == k};  dw[k,d] += sum_{idx==k} x[.,d,.]   (fp32; caller zeroes)
 * x, quant: [B, D, S]; idx: [B, S] int64; embed: [K, D]; counts/dw may both be NULL (eval).
 */
int vq3d_vq_assign(const float *x, const float *embed, int64_t B, int D, int64_t S, int K,
                   float *quant, int64_t *idx, double *sqerr, float *counts, float *dw,
                   void *stream);

/*
 * Same contract and bit-identical results as vq3d_vq_assign, for the large problems of the quantizer
 * sweep (embedding_dim 32, 64 or 128; B*S latent vectors in the tens of thousands and up): the distance
 * matrix is a tcgen05 candidate pass (bf16 hi/lo split operands, fp32 accumulation in TMEM) and only the
 * codes within a proven error margin of each row minimum are re-ranked with the reference's exact fp32
 * arithmetic (rows with too many candidates fall back to the exact full scan).  ws: 256-byte aligned
 * device buffer of vq3d_vq_assign_tc_workspace(D, K) bytes (codebook operand images); 0 = unsupported D/K.
 */
size_t vq3d_vq_assign_tc_workspace(int D, int K);
int vq3d_vq_assign_tc(const float *x, const float *embed, int64_t B, int D, int64_t S, int K, float *quant, int64_t *idx,
                      double *sqerr, float *counts, float *dw, void *ws, size_t ws_bytes, void *stream);

/* loss = commitment_cost * sqerr / numel  (F.mse_loss * commitment_cost, layers.py:716-717) */
int vq3d_vq_loss(const double *sqerr, double commitment_cost, int64_t numel, float *loss, void *stream);

/*
 * Tail of Quantizer._update_ema, layers.py:649-663, in place on the three buffers:
 *   cluster_size = decay*cluster_size + (1-decay)*counts
 *   embed_avg    = decay*embed_avg    + (1-decay)*dw
 *   n = sum(cluster_size); smoothed = n*(cluster_size+alpha)/(n+K*alpha); embed = embed_avg/smoothed
 * counts/dw are the (already all-reduced) outputs of vq3d_vq_assign.
 */
int vq3d_vq_ema_update(const float *counts, const float *dw, int K, int D, double decay, double laplace_alpha,
                       float *cluster_size, float *embed_avg, float *embed, void *stream);

/*
 * First half of Quantizer._init_ema, layers.py:666-667: mean[d], unbiased std[d] of the
 * B*S latent vectors.  meanstd: [2, D] fp32 out (row 0 mean, row 1 std); scratch: [2*D]
 * doubles, zeroed by the call.  The caller all-reduces/averages meanstd across ranks
 * (layers.py:670-674) before vq3d_vq_init_apply.
 */
int vq3d_vq_init_stats(const float *x, int64_t B, int D, int64_t S, double *scratch, float *meanstd, void *stream);

/*
 * Second half of _init_ema, layers.py:678-683: embed = embed*std + mean; embed_avg = embed;
 * cluster_size += total_vectors / K; *first_pass = 0 (int64 scalar buffer).
 */
int vq3d_vq_init_apply(const float *meanstd, int K, int D, double total_vectors, float *embed, float *embed_avg,
                       float *cluster_size, int64_t *first_pass, void *stream);

/* Quantizer.embed_code, layers.py:633-634 (F.embedding): out[n, :] = embed[idx[n], :]. */
int vq3d_vq_embed_code(const int64_t *idx, const float *embed, int64_t n, int D, int K, float *out, void *stream);

/*
 * Backward of the straight-through estimator + commitment loss, layers.py:716-720:
 *   grad_x = grad_quant + grad_loss * 2*commitment_cost/numel * (x - quant)
 * grad_quant may be NULL (treated as 0); grad_loss is a device scalar.
 */
int vq3d_vq_backward(const float *grad_quant, const float *grad_loss, const float *x, const float *quant,
                     int64_t numel, double commitment_cost, float *grad_x, void *stream);

/* ------------------------------------------------------------------------------------
 * Convolution building blocks (every nn.Conv3d / ResizeConv3D call site of layers.py)
 * ------------------------------------------------------------------------------------ */

typedef struct vq3d_conv_desc {
    /* geometry */
    int32_t B, H, W, Z;          /* input spatial size */
    int32_t C1, C2;              /* channels of x1 and x2; C_in = C1 + C2 (implicit torch.cat, layers.py:385,512) */
    int32_t Cout;
    int32_t k, stride, pad;      /* cubic kernel; out = (in + 2*pad - k)/stride + 1 */
    int32_t pad_circular;        /* 1: padding_mode='circular' (layers.py:109), 0: zeros */
    /* input transform, applied to every input element before padding:
     *   pre_act ? ELU(x + *pre_a) + *pre_b : x + *pre_b                          (layers.py:178-185) */
    int32_t pre_act;
    /* output transform: y = conv * (*post_scale) + *post_b + bias[co] + residual; then ELU if post_act */
    int32_t post_act;
    const float *x1, *x2;        /* [B, C1, S], [B, C2, S]; x2 may be NULL when C2 == 0 */
    const float *w;              /* [Cout, C1+C2, k, k, k] */
    const float *bias;           /* [Cout] or NULL */
    const float *pre_a, *pre_b;  /* device scalars or NULL */
    const float *post_scale, *post_b;
    const float *residual;       /* [B, Cout, S_out] or NULL */
    float *y;                    /* [B, Cout, S_out] */
} vq3d_conv_desc;

/* Direct convolution for any shape on the path (fallback for layers without a fused kernel). */
int vq3d_conv3d(const vq3d_conv_desc *desc, void *stream);

/*
 * Same contract as vq3d_conv3d, computed as an implicit GEMM on the tcgen05 tensor cores
 * (bf16 operands, fp32 accumulation in TMEM).  For GEMM-shaped layers only: returns
 * VQ3D_ERR_UNSUPPORTED when C_in*k^3 < 32, C_out < 8 or C_out > 256 (callers fall back to
 * vq3d_conv3d).  Results agree with the fp32 kernel to bf16 operand rounding (~1e-2 relative).
 * Layers with few output voxels and a long reduction (128 -> 128 k3 at 8x8x2 ...) are split along K
 * across CTAs: every K slice stores its partial tile into its own slab of `ws` and a second kernel adds
 * the slabs in slice order and applies the epilogue (no atomics: bit-reproducible run to run).
 * Pass a 16-byte aligned device buffer of vq3d_conv3d_tc_workspace(desc) bytes (0 = no split wanted
 * for this shape); with ws == NULL the layer runs unsplit.
 */
size_t vq3d_conv3d_tc_workspace(const vq3d_conv_desc *desc);
int vq3d_conv3d_tc(const vq3d_conv_desc *desc, void *ws, size_t ws_bytes, void *stream);

/*
 * Backward of vq3d_conv3d (what autograd derives for the reference's Conv3d / F.pad / ELU / Fixup-scalar chain,
 * layers.py:176-195), any shape, fp32.  desc is the FORWARD descriptor (residual / y / post_b values are not read):
 *   gy [B, Cout, S_out]                 gradient wrt y (also the gradient wrt the residual)
 *   gx1 / gx2                           out: gradient wrt x1 / x2 (NULL = not wanted)
 *   gw [Cout, C1+C2, k, k, k], gbias [Cout]   ACCUMULATED into (caller zeroes) or NULL
 *   gscalars float[4]                   ACCUMULATED: d pre_a, d pre_b, d post_scale, d post_b (entries whose forward
 *                                       pointer is NULL are left alone); NULL = not wanted
 *   raw [B, Cout, S_out]                the convolution output before the post transform (vq3d_conv3d with
 *                                       post_scale = post_b = bias = residual = NULL); required iff d post_scale is wanted
 * post_act (FixupResBlock) is not differentiated: VQ3D_ERR_UNSUPPORTED.
 */
typedef struct vq3d_conv_bwd {
    const float *gy;
    const float *raw;
    float *gx1, *gx2;
    float *gw, *gbias;
    float *gscalars;
    int32_t skip_input_grads;     /* 1: gx1/gx2 and d pre_a / d pre_b come from vq3d_conv3d_dgrad_finish (below) */
} vq3d_conv_bwd;
int vq3d_conv3d_backward(const vq3d_conv_desc *desc, const vq3d_conv_bwd *grads, void *stream);

/*
 * The same gradients for a POINTWISE convolution (k = 1, stride 1, no padding: branch_conv1 / branch_conv3 / the 1x1 skip and
 * proj / parse_input / out convolutions, layers.py:134-160,377,490,508,535) in ONE launch: gx1 / gx2, gw, gbias and all four
 * scalar gradients, with the convolution output that d post_scale needs recomputed inside the kernel (grads->raw and
 * grads->skip_input_grads are ignored).  Same accumulate / overwrite conventions as vq3d_conv3d_backward.
 * VQ3D_ERR_UNSUPPORTED when (C1 + C2) * Cout + Cout > 3072 (callers use vq3d_conv3d_backward).
 */
int vq3d_conv1x1_backward(const vq3d_conv_desc *desc, const vq3d_conv_bwd *grads, void *stream);

/*
 * Input gradient of a stride-1 "same" convolution (k odd, pad = (k-1)/2) as a FORWARD convolution: the caller runs
 * vq3d_conv3d[_tc] on gy with the flipped / transposed weight (C_in <-> C_out, every tap mirrored, same padding mode),
 * which yields gu_all [B, C1+C2, H, W, Z] = d loss / d (conv input) up to the post scale; this entry point finishes it:
 * gu = gu_all * post_scale, gx = gu * (pre_act ? ELU'(x + pre_a) : 1) split into gx1 / gx2 (either may be NULL),
 * gscalars[0] += sum(gx) (pre_act only), gscalars[1] += sum(gu).  desc: the forward call's descriptor.
 */
int vq3d_conv3d_dgrad_finish(const vq3d_conv_desc *desc, const float *gu_all, float *gx1, float *gx2, float *gscalars, void *stream);

/* Backward of vq3d_upsample2x: gx [B, C, H, W, Z] from gy [B, C, 2H, 2W, 2Z]; gscalars as above (entries 0, 1). */
int vq3d_upsample2x_backward(const float *gy, const float *x, int64_t B, int C, int H, int W, int Z, int pre_act,
                             const float *pre_a, const float *pre_b, float *gx, float *gscalars, void *stream);

/* Backward of vq3d_huber_elu_mask's mean: grad_decoded = grad_loss / count * d smooth_l1(mask(ELU(decoded)), x) / d decoded. */
int vq3d_huber_elu_mask_backward(const float *decoded, const float *x, const int32_t *num_valid, const uint8_t *mask_hw,
                                 int64_t B, int H, int W, int Z, const double *count, const float *grad_loss,
                                 float *grad_decoded, void *stream);

/* The same step with the step counter on the device (step_state: 3 doubles = step, 1 - beta1^step, 1 - beta2^step; zero it once):
 * the call first advances the counter, so a step captured in a CUDA graph replays with the right bias corrections. */
int vq3d_adam_amsgrad_step_dev(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *max_exp_avg_sq, int64_t n,
                               double lr, double beta1, double beta2, double eps, double *step_state, void *stream);

/* Backward of a trailing ELU given its OUTPUT y (FixupResBlock's last activation, vqvae/layers.py:288-289):
 * gx = gy * (y > 0 ? 1 : y + 1), n elements. */
int vq3d_elu_backward(const float *gy, const float *y, float *gx, int64_t n, void *stream);

/* One Adam(amsgrad=True) step on a flat fp32 tensor (model.py:91-93; torch.optim.Adam defaults otherwise). step >= 1. */
int vq3d_adam_amsgrad_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *max_exp_avg_sq, int64_t n,
                           double lr, double beta1, double beta2, double eps, int64_t step, void *stream);

/*
 * nn.Upsample(scale_factor=2, mode='trilinear', align_corners=False) of ResizeConv3D
 * (layers.py:591-597) with the same optional input transform as vq3d_conv_desc:
 * y[B, C, 2H, 2W, 2Z] = upsample(pre_act ? ELU(x + *pre_a) + *pre_b : x + *pre_b).
 */
int vq3d_upsample2x(const float *x, int64_t B, int C, int H, int W, int Z, int pre_act,
                    const float *pre_a, const float *pre_b, float *y, void *stream);

/*
 * One whole PreActFixupResBlock.forward (layers.py:176-195) in ONE launch:
 *   o = conv1x1(ELU(x+b1a)+b1b); o = conv_k(ELU(o+b2a)+b2b) [circular; after a trilinear x2
 *   upsample when mode==up]; o = conv1x1(ELU(o+b3a)+b3b); o = o*scale + b4;
 *   y = o + (skip(x+b1c)+b1d | x)
 * mode: 0 same/out (k3 s1), 1 down (k4 s2, skip k2 s2), 2 up (ResizeConv3D k3, skip ResizeConv3D k1).
 * Returns VQ3D_ERR_UNSUPPORTED (2) when no fused instantiation covers the channel counts;
 * the caller then composes the block from vq3d_conv3d / vq3d_upsample2x.
 */
typedef struct vq3d_preact_desc {
    int32_t B, H, W, Z;          /* input spatial size */
    int32_t Cin, Cb, Cout;
    int32_t mode;
    const float *x;              /* [B, Cin, S] */
    const float *w1, *w2, *w3;   /* branch_conv1/2/3.weight */
    const float *wskip;          /* skip_conv.weight or NULL */
    const float *b1a, *b1b, *b2a, *b2b, *b3a, *b3b, *b4, *scale, *b1c, *b1d;
    float *y;                    /* [B, Cout, S_out] */
    /* optional trailing 1x1 convolution fused into the block's epilogue (the decoder's `out` conv,
     * layers.py:508,516: Cout -> 1 channel with bias): when out_w != NULL the block writes
     * out_y[B, 1, S_out] = sum_c out_w[c] * y[c] + *out_b instead of y.  Honoured by vq3d_preact_block /
     * the last block of vq3d_preact_stack for 'same' blocks the row kernel covers; otherwise
     * VQ3D_ERR_UNSUPPORTED (run the block and the convolution separately). */
    const float *out_w, *out_b;
    float *out_y;
    /* optional leading 1x1 convolution fused into the block's loads (the encoder's parse_input, layers.py:535,578:
     * 1 -> Cin channels with bias): when pre_w != NULL, x is a ONE-channel tensor [B, 1, S] and the block's input is
     * pre_w[c] * x + pre_b[c].  Honoured by vq3d_preact_block for the 'down' shapes the row kernel covers;
     * otherwise VQ3D_ERR_UNSUPPORTED. */
    const float *pre_w, *pre_b;
} vq3d_preact_desc;

#define VQ3D_OK 0
#define VQ3D_ERR_INVALID 1
#define VQ3D_ERR_UNSUPPORTED 2
#define VQ3D_ERR_CUDA 3

int vq3d_preact_block(const vq3d_preact_desc *desc, void *stream);

/*
 * A run of n consecutive 'same' PreActFixupResBlocks with equal channel counts (the
 * nn.Sequential stacks of layers.py:492-494,566-569) in as few launches as possible.
 * blocks[i] describes block i; blocks[i].x / .y are ignored except blocks[0].x (input) and
 * blocks[n-1].y (output); tmp is a scratch activation buffer of the same size (ping-pong).
 */
int vq3d_preact_stack(const vq3d_preact_desc *blocks, int n, float *tmp, void *stream);

/*
 * Same contract as vq3d_preact_stack (n >= 1 equal-shape 'same' blocks, layers.py:492-494,566-569) with
 * the k3 convolution on the tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM) and up to
 * 24 blocks per launch: persistent warp-specialised CTAs pass a grid barrier between consecutive
 * blocks (cooperative launch).  blocks[0].x is the input, blocks[n-1].y the output (updated in place by
 * blocks 1..n-1; it must not alias the input).  ws: device workspace of at least
 * vq3d_preact_stack_tc_workspace(&blocks[0]) bytes, 256-byte aligned (bf16 intermediate of the next
 * block + the barrier counter); contents need not be preserved between calls.  Results agree with the
 * fp32 kernels to bf16 operand rounding.  Returns VQ3D_ERR_UNSUPPORTED when no instantiation covers
 * (Cin, Cb) -- callers fall back to vq3d_preact_stack; the workspace query then returns 0.
 */
size_t vq3d_preact_stack_tc_workspace(const vq3d_preact_desc *first_block);
int vq3d_preact_stack_tc(const vq3d_preact_desc *blocks, int n, void *ws, size_t ws_bytes, void *stream);

/*
 * Fused backward of a 'same' PreActFixupResBlock without skip convolution (autograd of layers.py:176-195 for the blocks of
 * the 50 / 150-deep stacks): the training forward keeps only the block input x; two tiled kernels recompute the intermediates
 * (a2 on a 1-voxel halo, c2, the activations) and produce every gradient, exchanging gt3 and c1 through `ws`; d W2 comes from
 * the tiled weight-gradient kernel of vq3d_conv3d_backward.  desc: the forward descriptor (mode 0, wskip NULL, Cin == Cout
 * <= 32, Cb <= 16; y / out_* / pre_* unused).  gy [B, C, S]; gx [B, C, S] written (may be NULL); gw1 [Cb, C], gw2 [Cb, Cb, 27],
 * gw3 [C, Cb] and gscalars[8] = d bias1a, bias1b, bias2a, bias2b, bias3a, bias3b, bias4, scale are ACCUMULATED (caller zeroes;
 * any may be NULL).  ws: vq3d_preact_same_backward_workspace(desc) bytes (0 = shape not covered: compose the block from the
 * generic entry points).  fp32 throughout.
 */
size_t vq3d_preact_same_backward_workspace(const vq3d_preact_desc *desc);
int vq3d_preact_same_backward(const vq3d_preact_desc *desc, const float *gy, void *ws, size_t ws_bytes, float *gx, float *gw1, float *gw2,
                              float *gw3, float *gscalars, void *stream);

/*
 * One 'up' PreActFixupResBlock (mode 2: trilinear x2 + k3 circular convolution, skip = ResizeConv3D k1; layers.py:124-132,
 * 176-195,591-597) with the k3 convolution on the tensor cores.  Same descriptor as vq3d_preact_block (mode must be 2).
 * The pointwise low-resolution stage (conv1 and the 1x1 skip convolution, which commutes with the interpolation) and the
 * trilinear expansion run as two small fp32 kernels; the expansion writes the bf16 branch activation straight into the
 * tensor-core operand layout and the interpolated skip path into y, and the k3 convolution + conv3 + residual run on the
 * persistent tcgen05 kernel of vq3d_preact_stack_tc.  ws: 256-byte aligned device buffer of vq3d_preact_up_tc_workspace(desc)
 * bytes (0 = no instantiation for (Cout, Cb): callers fall back to vq3d_preact_block).  Results agree with the fp32 kernels to
 * bf16 operand rounding of the branch; the skip path stays fp32.
 */
size_t vq3d_preact_up_tc_workspace(const vq3d_preact_desc *desc);
int vq3d_preact_up_tc(const vq3d_preact_desc *desc, void *ws, size_t ws_bytes, void *stream);

/*
 * The thin 'same' blocks (4 -> 2 -> 4, 8 -> 4 -> 8; full-depth tiles: Z in {32, 64, 128}, H and W multiples of 8) with
 * conv2 as ONE merged-tap GEMM per 128 voxels on the tensor cores (K = 27 * C_b; bf16 operands, fp32 accumulation).
 * Same arguments and ping-pong convention as vq3d_preact_stack (tmp may be NULL for n == 1); a last block that carries
 * the fused out convolution runs on the fp32 row kernel.  VQ3D_ERR_UNSUPPORTED for any other shape.
 */
int vq3d_preact_stack_thin_tc(const vq3d_preact_desc *blocks, int n, float *tmp, void *stream);

/*
 * EvoNorm3D-S0 (vqvae/evonorm.py:12-26,59-76; batch 1 like the reference): per-channel std[c] =
 * sqrt(unbiased_var(x over the channel group of c) + eps), groups = max(C // 8, 1) (scratch: 2*groups doubles),
 * then y = x * sigmoid(v[c] * x) * gamma[c] / std[c] + beta[c].  x, y: [C, S].
 */
int vq3d_evonorm_s0_stats(const float *x, int C, int64_t S, int groups, double eps, double *scratch, float *std_out, void *stream);
int vq3d_evonorm_s0_apply(const float *x, const float *v, const float *gamma, const float *beta, const float *std_in,
                          int C, int64_t S, float *y, void *stream);

/*
 * Backward of EvoNorm3D-S0 (autograd of evonorm.py:12-26,59-76 incl. SiLUVelocityFunc.backward :36-47), batch 1, x/gy/gx [C, S].
 * _sums: per channel c the doubles sums[3c..3c+2] = sum(gy), sum(gy * x sigmoid(v x)), sum(gy * x^2 s (1 - s)) (caller need
 * not zero them).  The C-sized algebra in between (d gamma, d beta, d v, the group coefficients) is host code.
 * _apply: gx = gy * coef_a[c] * (s + x v s (1 - s)) + coef_b[c] * (x - mean[c]), with coef_a = gamma / std,
 * coef_b = -(sum over the group of gamma * sums[3c+1]) / (std^3 (n - 1)), mean = the group mean.
 */
int vq3d_evonorm_s0_backward_sums(const float *x, const float *gy, const float *v, int C, int64_t S, double *sums, void *stream);
int vq3d_evonorm_s0_backward_apply(const float *x, const float *gy, const float *v, const float *coef_a, const float *coef_b,
                                   const float *mean, int C, int64_t S, float *gx, void *stream);

/*
 * Loss epilogue of VQVAE.loc_metric, model.py:120-152, fused: loc = ELU(decoded), zero where
 * z >= num_valid[b], optional centre-cylinder mask over (H, W) (utils/load_nrrd_dataset.py:288-300,
 * mask: [H*W] uint8 or NULL), smooth-L1 (beta 1) against x, summed into *sum (double, caller
 * zeroes) with the element count into *count.
 */
int vq3d_huber_elu_mask(const float *decoded, const float *x, const int32_t *num_valid, const uint8_t *mask_hw,
                        int64_t B, int H, int W, int Z, double *sum, double *count, void *stream);

/*
 * The same fused pass with every statistic the reference logs beside the loss (model.py:143-149: sub_metric_log_dict of
 * the unreduced loss and of the masked reconstruction, nmse / psnr of metrics/evaluate.py:18-24) as partial sums, so that the
 * validation epilogue reads the two volumes once and nothing is materialised:
 *   sums[0..6]  += sum loss, count, sum (loc - x)^2, sum x^2, sum loc, sum loc^2, sum loss^2   (doubles; caller zeroes 8 entries)
 *   minmax[0..3] = min loc, max loc, min loss, max loss      (caller initialises to +inf, -inf, +inf, -inf)
 * loc = mask(ELU(decoded)); voxels outside mask_hw are skipped as in vq3d_huber_elu_mask.  The median entries of the
 * reference's log come from vq3d_huber_elu_mask_medians below.
 */
int vq3d_huber_elu_mask_stats(const float *decoded, const float *x, const int32_t *num_valid, const uint8_t *mask_hw,
                              int64_t B, int H, int W, int Z, double *sums, float *minmax, void *stream);

/*
 * The median entries of the same log (utils/logging_helpers.py:13, torch.median = the LOWER middle element of the flattened
 * tensor; NaN when it holds a NaN): medians[0] = median of loc, medians[1] = median of the unreduced smooth-L1 loss, over the
 * same voxels as vq3d_huber_elu_mask_stats.  Exact (the result is an element of the tensor): a radix select of four 8-bit
 * passes over an order-preserving key, each pass recomputing loc and loss from (decoded, x), so nothing is materialised or
 * sorted.  ws: vq3d_huber_elu_mask_medians_workspace() bytes of device scratch, 8-byte aligned (zeroed by the call).
 */
size_t vq3d_huber_elu_mask_medians_workspace(void);
int vq3d_huber_elu_mask_medians(const float *decoded, const float *x, const int32_t *num_valid, const uint8_t *mask_hw,
                                int64_t B, int H, int W, int Z, float *medians, void *ws, size_t ws_bytes, void *stream);

/*
 * Output epilogue of vqvae/decode_embeddings.py:43-47, fused: out[i] = rint(ELU(decoded[i]) * scale - offset) as int64
 * (the reference: F.elu, then numpy `res * 1000 - 1000`, np.rint, astype(int)); n elements.
 */
int vq3d_elu_hu_rint(const float *decoded, int64_t n, double scale, double offset, int64_t *out, void *stream);

/*
 * The same epilogue with an int16 result (saturating): Hounsfield units of CT volumes fit 16 bits (the reference clips its
 * inputs to [-1500, 3000], utils/load_nrrd_dataset.py:73-81), so a volume leaves the device in 2 bytes per voxel instead of
 * the 8 of `astype(int)` (decode_embeddings.py:47).  decoded 16-byte aligned, out 8-byte aligned.
 */
int vq3d_elu_hu_rint_i16(const float *decoded, int64_t n, double scale, double offset, int16_t *out, void *stream);

/*
 * CT front end of the reference's data module, utils/load_nrrd_dataset.py:73-81, on raw int16 Hounsfield units:
 *   out[i] = clip(hu[i], min_hu, max_hu) * mul + add        (fp32, product and sum rounded separately)
 * = ThresholdIntensity(3000) / ThresholdIntensity(-1500) / ScaleIntensity(factor = -1 + 1/1000) / ShiftIntensity(1) with
 * (min_hu, max_hu, mul, add) = (-1500, 3000, 0.001f, 1).  A volume enters the device in 2 bytes per voxel and becomes the
 * network's fp32 (B, 1, H, W, D) input there.  hu 8-byte aligned, out 16-byte aligned; n elements.
 */
int vq3d_hu_to_network(const int16_t *hu, int64_t n, double min_hu, double max_hu, double mul, double add, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VQVAE3D_B200_H */
