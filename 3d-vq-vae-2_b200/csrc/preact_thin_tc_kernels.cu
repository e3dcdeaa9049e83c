// preact_thin_tc_kernels.cu -- the thin pre-activation 'same' blocks (4 -> 2 -> 4 at 512 x 512 x 128, 8 -> 4 -> 8 at
// 256 x 256 x 64: PreActFixupResBlock, vqvae/layers.py:176-195, circular 3^3 conv2) with conv2 on the tensor cores.
//
// These blocks are compute bound on the SIMT pipes (16-25 TFLOP/s of fp32 FMAs, ~380 instructions per voxel in
// preact_row_kernel) although they move only 32-64 bytes per voxel.  One tcgen05.mma per tap (what preact_tc_kernel does
// for the wide blocks) would not help: with 2-4 branch channels an M128 x N16 x K16 instruction is 7/8 padding and still
// costs its fixed ~45 cycles.  Here the 27 taps are MERGED into the reduction dimension instead: K = 27 * C_b (54 -> 64,
// 108 -> 112), i.e. 4 (7) MMAs per 128 voxels, fed by an im2col image that the CTA's threads build in shared memory
// from the haloed bf16 conv1 tile (27 shared loads + K/8 128-bit stores per voxel).
//
// Per CTA (persistent, one per SM; a tile = 8 x 8 columns x the full depth Z):
//   1. all threads: conv1 for the (8+2) x (8+2) x Z halo box -- x (fp32, global, coalesced along z) -> ELU(x + b1a) + b1b
//      -> 1x1 conv -> ELU(. + b2a) + b2b -> bf16 -> shared t1[hh][ww][z][C_b]  (z wraps inside the tile: circular padding)
//   2. four warp groups walk the tile's 128-voxel M-blocks: build the K-major im2col A image (row = voxel), one thread
//      issues the K/16 MMAs into the group's TMEM accumulator (fp32), the group reads it back (tcgen05.ld, lane = voxel)
//      and finishes: ELU(t2 + b3a) + b3b -> 1x1 conv3 -> * scale + b4 + x -> y (fp32, coalesced along z).
// bf16 operands, fp32 accumulation: same numerics class as the other tensor-core kernels (1e-2 relative, tests).
#include "vq3d_rt.h"

#ifndef VQ3D_EMU
#include <cuda_bf16.h>

namespace vq3d {

#include "tc_common.cuh"

int preact_row_dispatch(const vq3d_preact_desc *d, void *stream, bool *handled);    // preact_row_kernels.cu

struct ThinParams {
    const float *x;
    float *y;
    const float *w1, *w2, *w3;
    const float *b1a, *b1b, *b2a, *b2b, *b3a, *b3b, *b4, *scale;
    int B, H, W, Z;
    int tilesH, tilesW;
};

template <int C, int CB>
struct ThinCfg {
    static constexpr int TH = 8, TW = 8, HH = TH + 2, HW = TW + 2;
    static constexpr int KP = ((27 * CB + 15) / 16) * 16;          // merged reduction length: 64 (C_b 2), 112 (C_b 4)
    static constexpr int KC = KP / 8, KS = KP / 16;
    static constexpr int NWG = CB == 2 ? 8 : 4, THREADS = NWG * 128;     // warp groups: independent build -> MMA -> epilogue chains
    static constexpr uint32_t ABYTES = 128u * KP * 2, BBYTES = 16u * KP * 2;
    static constexpr uint32_t LBO_A = 128 * 16, LBO_B = 16 * 16;
    static size_t smem(int Z) { return 128 + (size_t)HH * HW * Z * CB * 2 + BBYTES + (size_t)NWG * ABYTES + (size_t)(CB * C * 2) * 4; }
};

template <int C, int CB, int Z>          // Z (the full depth) is a compile-time constant: every index split and shared-memory offset folds
__global__ void __launch_bounds__(ThinCfg<C, CB>::THREADS, 1)
preact_thin_tc_kernel(const __grid_constant__ ThinParams p) {
    using Cfg = ThinCfg<C, CB>;
    constexpr int TH = Cfg::TH, TW = Cfg::TW, HW = Cfg::HW, KC = Cfg::KC, KS = Cfg::KS, NWG = Cfg::NWG;
    VQ3D_DYN_SMEM(unsigned char, smem_raw);
    __shared__ __align__(8) uint64_t done[NWG];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, r = tid & 127;
    const uint32_t base = (s_u32(smem_raw) + 127u) & ~127u;
    unsigned char *smem = smem_raw + (base - s_u32(smem_raw));
    const size_t t1_bytes = (size_t)Cfg::HH * HW * Z * CB * 2;
    unsigned char *s_t1 = smem;
    unsigned char *s_b = smem + t1_bytes;
    unsigned char *s_a = s_b + Cfg::BBYTES;
    float *s_w13 = reinterpret_cast<float *>(s_a + (size_t)NWG * Cfg::ABYTES);         // w1 [CB][C] then w3 [C][CB]
    const uint32_t b_addr = base + (uint32_t)t1_bytes, a_addr = b_addr + Cfg::BBYTES + (uint32_t)wg * Cfg::ABYTES;
    unsigned char *my_a = s_a + (size_t)wg * Cfg::ABYTES;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"((uint32_t)(NWG * 16)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int g = 0; g < NWG; ++g) mbarrier_init(&done[g], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // conv2 weights as the K-major B operand: B[n][k = tap * CB + c] = w2[n][c][tap], rows n >= CB and the K padding are zero
    for (int i = tid; i < 16 * Cfg::KP; i += Cfg::THREADS) {
        const int n = i / Cfg::KP, k = i - n * Cfg::KP;
        float v = 0.0f;
        if (n < CB && k < 27 * CB) v = __ldg(p.w2 + ((size_t)n * CB + (k % CB)) * 27 + k / CB);
        *reinterpret_cast<__nv_bfloat16 *>(s_b + (size_t)(k >> 3) * Cfg::LBO_B + (size_t)n * 16 + (size_t)(k & 7) * 2) = __float2bfloat16_rn(v);
    }
    for (int i = tid; i < 2 * CB * C; i += Cfg::THREADS) s_w13[i] = i < CB * C ? __ldg(p.w1 + i) : __ldg(p.w3 + i - CB * C);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot + (uint32_t)(wg * 16);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const float b1a = ld_scalar(p.b1a, 0.f), b1b = ld_scalar(p.b1b, 0.f), b2a = ld_scalar(p.b2a, 0.f), b2b = ld_scalar(p.b2b, 0.f);
    const float b3a = ld_scalar(p.b3a, 0.f), b3b = ld_scalar(p.b3b, 0.f), b4 = ld_scalar(p.b4, 0.f), scale = ld_scalar(p.scale, 1.f);
    const int64_t S = (int64_t)p.H * p.W * Z;
    const int tiles = p.B * p.tilesH * p.tilesW;
    const int mblocks = TH * TW * Z / 128;
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int tw = tile % p.tilesW, th = (tile / p.tilesW) % p.tilesH, b = tile / (p.tilesW * p.tilesH);
        const int h0 = th * TH, w0 = tw * TW;
        const float *xb = p.x + (size_t)b * C * S;
        float *yb = p.y + (size_t)b * C * S;
        // ---- 1. conv1 on the halo box (two voxels per iteration: their 2 * C loads are in flight together) ----
        constexpr int HALO = Cfg::HH * HW * Z;
        for (int i0 = tid; i0 < HALO; i0 += 2 * Cfg::THREADS) {
            float xin2[2][C];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = i0 + q * Cfg::THREADS;
                const int z = i % Z, ww = (i / Z) % HW, hh = i / (Z * HW);
                const int ih = wrap(h0 - 1 + hh, p.H), iw = wrap(w0 - 1 + ww, p.W);
                const float *src = xb + ((size_t)ih * p.W + iw) * Z + z;
#pragma unroll
                for (int c = 0; c < C; ++c) xin2[q][c] = i < HALO ? __ldg(src + (size_t)c * S) : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = i0 + q * Cfg::THREADS;
                if (i >= HALO) break;
                float u[C];
#pragma unroll
                for (int c = 0; c < C; ++c) u[c] = elu_bl(xin2[q][c] + b1a) + b1b;
                float t[CB];
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    float a = 0.0f;
#pragma unroll
                    for (int c = 0; c < C; ++c) a = __fmaf_rn(s_w13[j * C + c], u[c], a);
                    t[j] = elu_bl(a + b2a) + b2b;
                }
                if (CB == 2) {
                    *reinterpret_cast<uint32_t *>(s_t1 + (size_t)i * 4) = bf16x2(t[0], t[1]);
                } else {
                    uint2 v;
                    v.x = bf16x2(t[0], t[1]); v.y = bf16x2(t[CB > 2 ? 2 : 0], t[CB > 3 ? 3 : 0]);
                    *reinterpret_cast<uint2 *>(s_t1 + (size_t)i * 8) = v;
                }
            }
        }
        __syncthreads();
        // ---- 2. M-blocks of this warp group ----
        for (int mb = wg; mb < mblocks; mb += NWG) {
            const int v = mb * 128 + r;
            const int z = v % Z, dw = (v / Z) % TW, dh = v / (Z * TW);
            const int zs[3] = {z == 0 ? Z - 1 : z - 1, z, z == Z - 1 ? 0 : z + 1};
            // im2col: row r, k = tap * CB + c (tap = (kh * 3 + kw) * 3 + kz); 16-byte chunks of 8 k's
            const unsigned char *t1row = s_t1 + ((size_t)(dh * HW + dw) * Z) * (CB * 2);
            unsigned char *arow = my_a + (size_t)r * 16;
            if (CB == 2) {
                uint32_t tv[28];
#pragma unroll
                for (int t = 0; t < 27; ++t) {
                    const int kh = t / 9, kw = (t / 3) % 3, kz = t % 3;
                    tv[t] = *reinterpret_cast<const uint32_t *>(t1row + ((size_t)(kh * HW + kw) * Z + zs[kz]) * 4);
                }
                tv[27] = 0u;
#pragma unroll
                for (int kc = 0; kc < 7; ++kc) {
                    uint4 q;
                    q.x = tv[4 * kc]; q.y = tv[4 * kc + 1]; q.z = tv[4 * kc + 2]; q.w = tv[4 * kc + 3];
                    *reinterpret_cast<uint4 *>(arow + (size_t)kc * Cfg::LBO_A) = q;
                }
                *reinterpret_cast<uint4 *>(arow + (size_t)7 * Cfg::LBO_A) = make_uint4(0u, 0u, 0u, 0u);
            } else {
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
                    const int t0 = 2 * kc, t1i = 2 * kc + 1;
                    if (t0 < 27) lo = *reinterpret_cast<const uint2 *>(t1row + ((size_t)((t0 / 9) * HW + (t0 / 3) % 3) * Z + zs[t0 % 3]) * 8);
                    if (t1i < 27) hi = *reinterpret_cast<const uint2 *>(t1row + ((size_t)((t1i / 9) * HW + (t1i / 3) % 3) * Z + zs[t1i % 3]) * 8);
                    *reinterpret_cast<uint4 *>(arow + (size_t)kc * Cfg::LBO_A) = make_uint4(lo.x, lo.y, hi.x, hi.y);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
            if (r == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                    umma_f16(tmem_d, umma_desc(a_addr + (uint32_t)(2 * ks) * Cfg::LBO_A, Cfg::LBO_A, 128),
                             umma_desc(b_addr + (uint32_t)(2 * ks) * Cfg::LBO_B, Cfg::LBO_B, 128), idesc, ks > 0 ? 1u : 0u);
                umma_commit_to(&done[wg]);
            }
            // the residual input is fetched while the MMAs run
            const float *xv = xb + ((size_t)(h0 + dh) * p.W + (w0 + dw)) * Z + z;
            float xin[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xin[c] = __ldg(xv + (size_t)c * S);
            mbarrier_wait(&done[wg], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float acc[16];
            tmem_ld16(tmem_d + ((uint32_t)((warp & 3) * 32) << 16), acc);
            float u3[CB];
#pragma unroll
            for (int j = 0; j < CB; ++j) u3[j] = elu_bl(acc[j] + b3a) + b3b;
            float *yv = yb + ((size_t)(h0 + dh) * p.W + (w0 + dw)) * Z + z;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float o = 0.0f;
#pragma unroll
                for (int j = 0; j < CB; ++j) o = __fmaf_rn(s_w13[CB * C + c * CB + j], u3[j], o);
                yv[(size_t)c * S] = __fmaf_rn(o, scale, b4) + xin[c];
            }
            phase ^= 1u;
            // the accumulator and the A image are rewritten by the next M-block: every read of this one is complete
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        }
        __syncthreads();        // t1 is rewritten by the next tile
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"((uint32_t)(NWG * 16)) : "memory");
    }
}

template <int C, int CB, int Z>
static int launch_thin_z(const vq3d_preact_desc *d, void *stream) {
    using Cfg = ThinCfg<C, CB>;
    ThinParams p;
    p.x = d->x; p.y = d->y; p.w1 = d->w1; p.w2 = d->w2; p.w3 = d->w3;
    p.b1a = d->b1a; p.b1b = d->b1b; p.b2a = d->b2a; p.b2b = d->b2b; p.b3a = d->b3a; p.b3b = d->b3b; p.b4 = d->b4; p.scale = d->scale;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z;
    p.tilesH = d->H / Cfg::TH; p.tilesW = d->W / Cfg::TW;
    const int64_t tiles = (int64_t)d->B * p.tilesH * p.tilesW;
    int grid = kNumSMs;
    if (grid > tiles) grid = (int)tiles;
    return launch("preact_thin_tc", preact_thin_tc_kernel<C, CB, Z>, dim3((unsigned)grid), dim3(Cfg::THREADS), Cfg::smem(Z), stream, p);
}

template <int C, int CB>
static int launch_thin(const vq3d_preact_desc *d, void *stream) {
    switch (d->Z) {
        case 128: return launch_thin_z<C, CB, 128>(d, stream);
        case 64: return launch_thin_z<C, CB, 64>(d, stream);
        default: return launch_thin_z<C, CB, 32>(d, stream);
    }
}

static bool thin_supported(const vq3d_preact_desc *d) {
    if (d->mode != 0 || d->wskip || d->out_w || d->pre_w || d->Cin != d->Cout) return false;
    if (!((d->Cin == 4 && d->Cb == 2) || (d->Cin == 8 && d->Cb == 4))) return false;
    if (d->Z != 32 && d->Z != 64 && d->Z != 128) return false;
    if (d->H % 8 != 0 || d->W % 8 != 0 || d->H < 8 || d->W < 8) return false;
    const size_t smem = d->Cb == 2 ? ThinCfg<4, 2>::smem(d->Z) : ThinCfg<8, 4>::smem(d->Z);
    return smem <= 225 * 1024;
}

}  // namespace vq3d
#endif  // !VQ3D_EMU

using namespace vq3d;

extern "C" int vq3d_preact_stack_thin_tc(const vq3d_preact_desc *blocks, int n, float *tmp, void *stream) {
#ifdef VQ3D_EMU
    (void)blocks; (void)n; (void)tmp; (void)stream;
    return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_thin_tc: tensor-core kernels cannot run in the host emulator");
#else
    if (!blocks || n < 1 || (n > 1 && !tmp)) return fail(VQ3D_ERR_INVALID, "preact_stack_thin_tc: bad arguments");
    for (int i = 0; i < n; ++i) {
        const vq3d_preact_desc *d = &blocks[i];
        if (!d->x || !d->y || !d->w1 || !d->w2 || !d->w3) return fail(VQ3D_ERR_INVALID, "preact_stack_thin_tc: null pointer");
        if (d->Cin != blocks[0].Cin || d->Cb != blocks[0].Cb || d->mode != 0 || d->wskip) return fail(VQ3D_ERR_INVALID, "preact_stack_thin_tc: blocks must be equal-shape 'same' blocks");
        if (i + 1 < n && d->out_w) return fail(VQ3D_ERR_INVALID, "preact_stack_thin_tc: only the last block may carry a trailing 1x1 convolution");
    }
    vq3d_preact_desc probe = blocks[0];
    probe.out_w = nullptr; probe.out_b = nullptr; probe.out_y = nullptr;
    if (!thin_supported(&probe)) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_thin_tc: shape not covered (C=%d, Cb=%d, %dx%dx%d)", probe.Cin, probe.Cb, probe.H, probe.W, probe.Z);
    // same ping-pong convention as vq3d_preact_stack
    const float *src = blocks[0].x;
    float *out = blocks[n - 1].y;
    for (int i = 0; i < n; ++i) {
        vq3d_preact_desc d = blocks[i];
        d.x = src;
        d.y = ((n - 1 - i) % 2 == 0) ? out : tmp;
        int rc;
        if (d.out_w) {               // the block that carries the decoder's out conv stays on the row kernel (fused 1x1)
            bool handled = false;
            rc = preact_row_dispatch(&d, stream, &handled);
            if (!rc && !handled) rc = fail(VQ3D_ERR_UNSUPPORTED, "preact_stack_thin_tc: no kernel fuses the trailing 1x1 convolution for this shape");
        } else {
            rc = d.Cb == 2 ? launch_thin<4, 2>(&d, stream) : launch_thin<8, 4>(&d, stream);
        }
        if (rc) return rc;
        src = d.y;
    }
    return VQ3D_OK;
#endif
}
