// preact_row_kernels.cu -- 'same' PreActFixupResBlocks (vqvae/layers.py:102-216) with very few channels
// (4 -> 2 -> 4 at 512x512x128, 8 -> 4 -> 8 at 256x256x64, 2 -> 1 -> 2): the layers whose tensors are
// hundreds of MB and whose arithmetic is far too thin for the tensor cores (K = 2..4 of a K16 MMA).
//
// They are bound by HBM traffic and, in practice, by how many instructions the SIMT pipes spend per
// voxel, so this kernel is built around the contiguous depth axis: a thread owns 4 consecutive z
// (float4 loads/stores, one LDS.128 per conv2 row window), a tile is (th, tw) FULL depth rows, so the
// circular wrap along z needs no halo traffic at all (two extra shared-memory entries per row), and no
// per-voxel integer division is left anywhere.
//
//   stage A  t1 = ELU(conv1(ELU(x+b1a)+b1b)+b2a)+b2b for the (th+2)(tw+2) haloed rows -> shared memory
//   stage B  3x3x3 circular conv from shared memory, 4 z per thread in registers
//   stage C  ELU, conv3, *scale + b4 + x  [+ the decoder's final 1x1 `out` convolution (layers.py:508,516)
//            when asked to: the last block then writes 1 channel instead of 4]
#include "vq3d_rt.h"

namespace vq3d {

#ifndef VQ3D_ROW_PACK_C8
#define VQ3D_ROW_PACK_C8 1
#endif
template <int C> struct RowThreads { static constexpr int value = C >= 8 ? 256 : 512; };   // wide variants need the registers

struct RowParams {
    int B, H, W, Z;
    int th, tw, nth, ntw;
    const float *x, *w1, *w2, *w3;
    const float *b1a, *b1b, *b2a, *b2b, *b3a, *b3b, *b4, *scale;
    const float *out_w, *out_b;       // fused trailing 1x1 conv (C -> 1) or NULL
    float *y;                         // [B, C, S], or [B, 1, S] when out_w is given
};

__host__ __device__ __forceinline__ int rmod(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}

template <int C, int CB>
struct RowSmem {
    static constexpr int w1 = 0;                        // [C][CB]
    static constexpr int w2 = w1 + C * CB;              // [CB ci][9 (kh,kw)][3 kz][CB co]
    static constexpr int w3 = w2 + CB * 27 * CB;        // [CB][C]
    static constexpr int wo = w3 + CB * C;              // [C] + bias
    static constexpr int tile = (wo + C + 1 + 3) & ~3;  // [CB][(th+2)(tw+2)][Z+8], 16-byte aligned rows
    static size_t floats(int th, int tw, int Z) { return tile + (size_t)CB * (th + 2) * (tw + 2) * (Z + 8); }
};

template <int C, int CB, bool OUTC, int NZ, int NW>
__global__ void __launch_bounds__(RowThreads<C>::value, (C >= 8 || NZ == 8 || NW == 2) ? 2 : 0)
preact_row_kernel(RowParams p) {
    using SM = RowSmem<C, CB>;
    constexpr int kRowThreads = RowThreads<C>::value;
    constexpr bool kPackAC = CB % 2 == 0 && (C < 8 || VQ3D_ROW_PACK_C8);   // FFMA2 in conv1 / conv3 too (conv2 always when CB is even)
    VQ3D_DYN_SMEM(float, smem);
    float *s_w1 = smem + SM::w1, *s_w2 = smem + SM::w2, *s_w3 = smem + SM::w3, *s_wo = smem + SM::wo, *s_t1 = smem + SM::tile;
    const int tid = threadIdx.x;
    const int Z = p.Z, ZQ = Z >> 2, ZP = Z + 8;
    const int IW = p.tw + 2, nrows_in = (p.th + 2) * IW, nrows_out = p.th * p.tw;
    const int zq = tid % ZQ, slot = tid / ZQ, nslots = kRowThreads / ZQ;
    const int64_t S = (int64_t)p.H * p.W * Z;

    int bid = blockIdx.x;
    const int twi = bid % p.ntw; bid /= p.ntw;
    const int thi = bid % p.nth; bid /= p.nth;
    const int b = bid, oh0 = thi * p.th, ow0 = twi * p.tw;
    const float *xb = p.x + (size_t)b * C * S;

    for (int i = tid; i < C * CB; i += kRowThreads) s_w1[i] = p.w1[(i % CB) * C + i / CB];
    for (int i = tid; i < CB * 27 * CB; i += kRowThreads) {
        const int co = i % CB, t = (i / CB) % 27, ci = i / (CB * 27);
        s_w2[i] = p.w2[((size_t)co * CB + ci) * 27 + t];
    }
    for (int i = tid; i < CB * C; i += kRowThreads) s_w3[i] = p.w3[(i % C) * CB + i / C];
    if (OUTC) {
        if (tid < C) s_wo[tid] = p.out_w[tid];
        if (tid == C) s_wo[C] = p.out_b ? p.out_b[0] : 0.0f;
    }
    const float b1a = ld_scalar(p.b1a, 0.f), b1b = ld_scalar(p.b1b, 0.f), b2a = ld_scalar(p.b2a, 0.f), b2b = ld_scalar(p.b2b, 0.f);
    __syncthreads();

    // ---- stage A (the loads of the next row slot are issued before the arithmetic of the current one) ------
    {
        float4 nxt[C];
        auto issue = [&](int rs) {
            const int lh = rs / IW, lw = rs - lh * IW;
            const int gh = rmod(oh0 - 1 + lh, p.H), gw = rmod(ow0 - 1 + lw, p.W);
            const float *px = xb + ((size_t)gh * p.W + gw) * Z + 4 * zq;
#pragma unroll
            for (int c = 0; c < C; ++c) nxt[c] = __ldg(reinterpret_cast<const float4 *>(px + (size_t)c * S));
        };
        if (slot < nrows_in) issue(slot);
        for (int rs = slot; rs < nrows_in; rs += nslots) {
            float4 cur[C];
#pragma unroll
            for (int c = 0; c < C; ++c) cur[c] = nxt[c];
            if (rs + nslots < nrows_in) issue(rs + nslots);
            float t[CB][4];
            if constexpr (kPackAC) {
                float2 t2[CB / 2][4];
#pragma unroll
                for (int cp = 0; cp < CB / 2; ++cp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) t2[cp][k] = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 v = cur[c];
                    const float a[4] = {elu1(v.x + b1a) + b1b, elu1(v.y + b1a) + b1b, elu1(v.z + b1a) + b1b, elu1(v.w + b1a) + b1b};
#pragma unroll
                    for (int cp = 0; cp < CB / 2; ++cp) {
                        const float2 w = *reinterpret_cast<const float2 *>(s_w1 + c * CB + 2 * cp);
#pragma unroll
                        for (int k = 0; k < 4; ++k) t2[cp][k] = ffma2_bcast(w, a[k], t2[cp][k]);
                    }
                }
#pragma unroll
                for (int cp = 0; cp < CB / 2; ++cp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) { t[2 * cp][k] = t2[cp][k].x; t[2 * cp + 1][k] = t2[cp][k].y; }
            } else {
#pragma unroll
            for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                for (int k = 0; k < 4; ++k) t[cb][k] = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float4 v = cur[c];
                const float a0 = elu1(v.x + b1a) + b1b, a1 = elu1(v.y + b1a) + b1b, a2 = elu1(v.z + b1a) + b1b, a3 = elu1(v.w + b1a) + b1b;
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) {
                    const float w = s_w1[c * CB + cb];
                    t[cb][0] = __fmaf_rn(w, a0, t[cb][0]); t[cb][1] = __fmaf_rn(w, a1, t[cb][1]);
                    t[cb][2] = __fmaf_rn(w, a2, t[cb][2]); t[cb][3] = __fmaf_rn(w, a3, t[cb][3]);
                }
            }
            }
#pragma unroll
            for (int cb = 0; cb < CB; ++cb) {
                float4 o;
                o.x = elu1(t[cb][0] + b2a) + b2b; o.y = elu1(t[cb][1] + b2a) + b2b;
                o.z = elu1(t[cb][2] + b2a) + b2b; o.w = elu1(t[cb][3] + b2a) + b2b;
                float *row = s_t1 + ((size_t)cb * nrows_in + rs) * ZP;
                *reinterpret_cast<float4 *>(row + 4 + 4 * zq) = o;
                if (zq == 0) row[Z + 4] = o.x;              // circular halo: z = Z  -> z = 0
                if (zq == ZQ - 1) row[3] = o.w;             //                z = -1 -> z = Z-1
            }
        }
    }
    __syncthreads();

    // ---- stage B + C: NZ consecutive z per thread (4; 8 is a measured negative result, see launch_row) ---------
    constexpr int NV = NZ / 4;
    const int ZQB = Z / NZ, zqb = tid % ZQB, slotb = tid / ZQB, nslotsb = kRowThreads / ZQB;
    const float b3a = ld_scalar(p.b3a, 0.f), b3b = ld_scalar(p.b3b, 0.f), b4 = ld_scalar(p.b4, 0.f), sc = ld_scalar(p.scale, 1.f);
    auto stage_c = [&](float (&acc)[CB][NZ], const size_t off) {   // ELU, conv3, *scale + b4 + x [, out conv], store
#pragma unroll
        for (int co = 0; co < CB; ++co)
#pragma unroll
            for (int k = 0; k < NZ; ++k) acc[co][k] = elu1(acc[co][k] + b3a) + b3b;
        float o4[NZ];
#pragma unroll
        for (int k = 0; k < NZ; ++k) o4[k] = 0.f;
        auto finish = [&](int c, const float *out) {      // *scale + b4 + x, then store (or the fused `out` convolution)
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 xv = __ldg(reinterpret_cast<const float4 *>(xb + (size_t)c * S + off + 4 * v));
                float4 yv;
                yv.x = __fmaf_rn(out[4 * v + 0], sc, b4) + xv.x; yv.y = __fmaf_rn(out[4 * v + 1], sc, b4) + xv.y;
                yv.z = __fmaf_rn(out[4 * v + 2], sc, b4) + xv.z; yv.w = __fmaf_rn(out[4 * v + 3], sc, b4) + xv.w;
                if (OUTC) {
                    const float w = s_wo[c];
                    o4[4 * v + 0] = __fmaf_rn(w, yv.x, o4[4 * v + 0]); o4[4 * v + 1] = __fmaf_rn(w, yv.y, o4[4 * v + 1]);
                    o4[4 * v + 2] = __fmaf_rn(w, yv.z, o4[4 * v + 2]); o4[4 * v + 3] = __fmaf_rn(w, yv.w, o4[4 * v + 3]);
                } else {
                    *reinterpret_cast<float4 *>(p.y + (size_t)b * C * S + (size_t)c * S + off + 4 * v) = yv;
                }
            }
        };
        if constexpr (kPackAC) {
#pragma unroll
            for (int cq = 0; cq < C / 2; ++cq) {
                float2 out2[NZ];                // conv3 for the output channel pair (2cq, 2cq+1)
#pragma unroll
                for (int k = 0; k < NZ; ++k) out2[k] = make_float2(0.f, 0.f);
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) {
                    const float2 w = make_float2(s_w3[cb * C + 2 * cq], s_w3[cb * C + 2 * cq + 1]);
#pragma unroll
                    for (int k = 0; k < NZ; ++k) out2[k] = ffma2_bcast(w, acc[cb][k], out2[k]);
                }
                float oa[NZ], ob[NZ];
#pragma unroll
                for (int k = 0; k < NZ; ++k) { oa[k] = out2[k].x; ob[k] = out2[k].y; }
                finish(2 * cq, oa);
                finish(2 * cq + 1, ob);
            }
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float out[NZ];
#pragma unroll
                for (int k = 0; k < NZ; ++k) out[k] = 0.f;
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) {
                    const float w = s_w3[cb * C + c];
#pragma unroll
                    for (int k = 0; k < NZ; ++k) out[k] = __fmaf_rn(w, acc[cb][k], out[k]);
                }
                finish(c, out);
            }
        }
        if (OUTC) {
            const float bo = s_wo[C];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                float4 ov;
                ov.x = o4[4 * v + 0] + bo; ov.y = o4[4 * v + 1] + bo; ov.z = o4[4 * v + 2] + bo; ov.w = o4[4 * v + 3] + bo;
                *reinterpret_cast<float4 *>(p.y + (size_t)b * S + off + 4 * v) = ov;
            }
        }
    };
    if constexpr (NW == 2) {
        // two neighbouring output rows (w, w + 1) per thread: the four windows of a (ci, kh) row pair and every weight
        // are loaded once for both, 21 instead of 36 shared-memory instructions per 72 FFMA2
        constexpr int CP = CB / 2;
        const int tw2 = p.tw >> 1;
        for (int ro = slotb; ro < p.th * tw2; ro += nslotsb) {
            const int lh = ro / tw2, lw = 2 * (ro - lh * tw2);
            const int oh = oh0 + lh, ow = ow0 + lw;
            if (oh >= p.H || ow >= p.W) continue;
            float2 acc2[2][CP][4];
#pragma unroll
            for (int rw = 0; rw < 2; ++rw)
#pragma unroll
                for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc2[rw][cp][k] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int ci = 0; ci < CB; ++ci) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    float2 wprev[3][CP];
#pragma unroll
                    for (int kwi = 0; kwi < 4; ++kwi) {
                        const float *row = s_t1 + ((size_t)ci * nrows_in + (lh + kh) * IW + lw + kwi) * ZP + 4 * zqb;
                        const float4 m = *reinterpret_cast<const float4 *>(row + 4);
                        const float r[6] = {row[3], m.x, m.y, m.z, m.w, row[8]};
                        float2 wcur[3][CP];
                        if (kwi < 3) {
                            const float2 *wt = reinterpret_cast<const float2 *>(s_w2 + ((ci * 9 + kh * 3 + kwi) * 3) * CB);
#pragma unroll
                            for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                                for (int cp = 0; cp < CP; ++cp) wcur[kz][cp] = wt[kz * CP + cp];
                        }
#pragma unroll
                        for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                            for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (kwi < 3) acc2[0][cp][k] = ffma2_bcast(wcur[kz][cp], r[k + kz], acc2[0][cp][k]);
                                    if (kwi > 0) acc2[1][cp][k] = ffma2_bcast(wprev[kz][cp], r[k + kz], acc2[1][cp][k]);
                                }
                        if (kwi < 3) {
#pragma unroll
                            for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                                for (int cp = 0; cp < CP; ++cp) wprev[kz][cp] = wcur[kz][cp];
                        }
                    }
                }
            }
#pragma unroll
            for (int rw = 0; rw < 2; ++rw) {
                float acc[CB][NZ];
#pragma unroll
                for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) { acc[2 * cp][k] = acc2[rw][cp][k].x; acc[2 * cp + 1][k] = acc2[rw][cp][k].y; }
                stage_c(acc, ((size_t)oh * p.W + ow + rw) * Z + 4 * zqb);
            }
        }
    } else
    for (int ro = slotb; ro < nrows_out; ro += nslotsb) {
        const int lh = ro / p.tw, lw = ro - lh * p.tw;
        const int oh = oh0 + lh, ow = ow0 + lw;
        if (oh >= p.H || ow >= p.W) continue;
        float acc[CB][NZ];
        auto window = [&](int ci, int kh, int kw, float *r) {       // r[NZ + 2] = t1[z0 - 1 .. z0 + NZ]
            const float *row = s_t1 + ((size_t)ci * nrows_in + (lh + kh) * IW + lw + kw) * ZP + NZ * zqb;
            r[0] = row[3];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 m = *reinterpret_cast<const float4 *>(row + 4 + 4 * v);
                r[1 + 4 * v] = m.x; r[2 + 4 * v] = m.y; r[3 + 4 * v] = m.z; r[4 + 4 * v] = m.w;
            }
            r[NZ + 1] = row[4 + NZ];
        };
        if constexpr (CB % 2 == 0) {
            // two branch channels per FFMA2 (w pair from shared memory, the window value broadcast)
            constexpr int CP = CB / 2;
            float2 acc2[CP][NZ];
#pragma unroll
            for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                for (int k = 0; k < NZ; ++k) acc2[cp][k] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int ci = 0; ci < CB; ++ci) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        float r[NZ + 2];
                        window(ci, kh, kw, r);
                        const float2 *wt = reinterpret_cast<const float2 *>(s_w2 + ((ci * 9 + kh * 3 + kw) * 3) * CB);
#pragma unroll
                        for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                            for (int cp = 0; cp < CP; ++cp) {
                                const float2 w = wt[kz * CP + cp];
#pragma unroll
                                for (int k = 0; k < NZ; ++k) acc2[cp][k] = ffma2_bcast(w, r[k + kz], acc2[cp][k]);
                            }
                    }
                }
            }
#pragma unroll
            for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                for (int k = 0; k < NZ; ++k) { acc[2 * cp][k] = acc2[cp][k].x; acc[2 * cp + 1][k] = acc2[cp][k].y; }
        } else {
#pragma unroll
            for (int co = 0; co < CB; ++co)
#pragma unroll
                for (int k = 0; k < NZ; ++k) acc[co][k] = 0.0f;
#pragma unroll
            for (int ci = 0; ci < CB; ++ci) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        float r[NZ + 2];
                        window(ci, kh, kw, r);
                        const float *wt = s_w2 + ((ci * 9 + kh * 3 + kw) * 3) * CB;
#pragma unroll
                        for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                            for (int co = 0; co < CB; ++co) {
                                const float w = wt[kz * CB + co];
#pragma unroll
                                for (int k = 0; k < NZ; ++k) acc[co][k] = __fmaf_rn(w, r[k + kz], acc[co][k]);
                            }
                    }
                }
            }
        }
        stage_c(acc, ((size_t)oh * p.W + ow) * Z + NZ * zqb);
    }
}

template <int C, int CB, bool OUTC, int NZ, int NW = 1>
static int launch_row_nz(const vq3d_preact_desc *d, void *stream) {
    using SM = RowSmem<C, CB>;
    RowParams p;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z;
    // tile: as large as shared memory allows (halo recompute), but keep >= 2 CTAs per SM in flight
    const size_t cap = 110 * 1024;
    int th = d->H < 8 ? d->H : 8, tw = d->W < 8 ? d->W : 8;
    auto ntiles = [&]() { return (int64_t)d->B * ceil_div(d->H, th) * ceil_div(d->W, tw); };
    while (SM::floats(th, tw, d->Z) * 4 > cap || (ntiles() < 2 * kNumSMs && th * tw > 4)) {
        if (th >= tw && th > 1) th = (th + 1) / 2;
        else if (tw > 1) tw = (tw + 1) / 2;
        else return fail(VQ3D_ERR_UNSUPPORTED, "preact_block(row): tile does not fit shared memory");
    }
    p.th = th; p.tw = tw; p.nth = (int)ceil_div(d->H, th); p.ntw = (int)ceil_div(d->W, tw);
    p.x = d->x; p.w1 = d->w1; p.w2 = d->w2; p.w3 = d->w3;
    p.b1a = d->b1a; p.b1b = d->b1b; p.b2a = d->b2a; p.b2b = d->b2b; p.b3a = d->b3a; p.b3b = d->b3b; p.b4 = d->b4; p.scale = d->scale;
    p.out_w = d->out_w; p.out_b = d->out_b;
    p.y = OUTC ? d->out_y : d->y;
    const int64_t grid = ntiles();
    if (grid > 0x7fffffff) return fail(VQ3D_ERR_INVALID, "preact_block(row): grid too large");
    return launch("preact_row", preact_row_kernel<C, CB, OUTC, NZ, NW>, dim3((unsigned)grid), dim3(RowThreads<C>::value), SM::floats(th, tw, d->Z) * 4, stream, p);
}

// z per thread in stage B/C (VQ3D_ROW_NZ8: 0 = always 4, 1 = 8 for the 4 -> 2 -> 4 variant, 2 = also for 8 -> 4 -> 8).
// Measured (profiles/r01x_ffma2_ab.txt, variants P / B / N2): 8 z per thread halve the shared-memory instructions per FMA
// but the 32-byte lane stride makes every LDS.128 of the window a 2-way bank conflict: 4 -> 2 -> 4 at 512^3 gets 13 %
// SLOWER, 8 -> 4 -> 8 at 256^3 1.6 % slower.  Off by default.
#ifndef VQ3D_ROW_NZ8
#define VQ3D_ROW_NZ8 0
#endif
// output rows per thread in stage B (VQ3D_ROW_NW2: 0 = one, 1 = two for 4 -> 2 -> 4, 2 = also for 8 -> 4 -> 8)
#ifndef VQ3D_ROW_NW2
#define VQ3D_ROW_NW2 2
#endif
template <int C, int CB, bool OUTC>
static int launch_row(const vq3d_preact_desc *d, void *stream) {
    constexpr bool kWide = CB % 2 == 0 && ((C < 8 && VQ3D_ROW_NZ8 >= 1) || VQ3D_ROW_NZ8 >= 2);
    if constexpr (kWide) {
        if (d->Z % 8 == 0 && RowThreads<C>::value % (d->Z / 8) == 0) return launch_row_nz<C, CB, OUTC, 8>(d, stream);
    }
    if constexpr (CB % 2 == 0 && (VQ3D_ROW_NW2 >= 2 || (VQ3D_ROW_NW2 == 1 && C < 8))) {
        if (d->W % 2 == 0 && d->W >= 8 && d->H >= 8) return launch_row_nz<C, CB, OUTC, 4, 2>(d, stream);
    }
    return launch_row_nz<C, CB, OUTC, 4>(d, stream);
}

// ---------------------------------------------------------------------------------------------------
// 'down' blocks at the big levels (4 -> 4 -> 8 at 512^3, 8 -> 8 -> 16 at 256^3; layers.py:124-126,164-171):
// conv1 1x1, conv2 k4 s2 circular, conv3 1x1, skip k2 s2.  Same depth-row organisation: stage A computes t1
// for the (2t+2)^2 input rows of a (t x t) output-row tile over the FULL depth (the skip convolution re-reads its
// raw centre rows from L1/L2); stage B gives each thread 4 consecutive output z of ONE branch channel (the CB
// threads of a voxel sit in adjacent lanes and exchange t2 by shuffles for conv3).  Optionally the encoder's
// parse_input 1x1 convolution (layers.py:535,578) is applied on the fly to a 1-channel input.
constexpr int kDownThreads = 256;

struct DownParams {
    int B, H, W, Z;                   // input extent
    int tho, two, ntho, ntwo;         // output-row tile
    const float *x, *w1, *w2, *w3, *ws;
    const float *b1a, *b1b, *b2a, *b2b, *b3a, *b3b, *b4, *scale, *b1c, *b1d;
    const float *pre_w, *pre_b;       // fused parse_input (1 -> CIN) or NULL
    float *y;
};

template <int CIN, int CB, int COUT, bool PARSE>
struct DownSmem {
    static constexpr int CX = PARSE ? 1 : CIN;
    static constexpr int w1 = 0;                         // [CIN][CB]
    static constexpr int w2 = w1 + CIN * CB;             // [CB ci][16 (kh,kw)][CB co][4 kz]
    static constexpr int w3 = w2 + CB * 16 * CB * 4;     // [CB][COUT]
    static constexpr int ws = w3 + CB * COUT;            // [CIN][8 taps][COUT]
    static constexpr int pw = ws + CIN * 8 * COUT;       // [CIN] w, [CIN] b
    static constexpr int tile = (pw + 2 * CIN + 3) & ~3;
    static size_t floats(int tho, int two, int Z) {
        return tile + (size_t)CB * (2 * tho + 2) * (2 * two + 2) * (Z + 8);
    }
};

template <int CIN, int CB, int COUT, bool PARSE>
__global__ void __launch_bounds__(kDownThreads)
preact_down_row_kernel(DownParams p) {
    using SM = DownSmem<CIN, CB, COUT, PARSE>;
    constexpr int CX = SM::CX, NPC = COUT / CB;
    static_assert(COUT % CB == 0 && (CB & (CB - 1)) == 0 && CB <= 32, "one branch channel per lane, CB lanes per voxel");
    VQ3D_DYN_SMEM(float, smem);
    float *s_w1 = smem + SM::w1, *s_w2 = smem + SM::w2, *s_w3 = smem + SM::w3, *s_ws = smem + SM::ws, *s_pw = smem + SM::pw;
    const int tid = threadIdx.x;
    const int Z = p.Z, ZQ = Z >> 2, ZQo = Z >> 3, ZP = Z + 8, Zo = Z >> 1;
    const int IW = 2 * p.two + 2, nrows_in = (2 * p.tho + 2) * IW, nrows_out = p.tho * p.two;
    float *s_t1 = smem + SM::tile;
    const int64_t S = (int64_t)p.H * p.W * Z, So = S >> 3;
    const int Ho = p.H >> 1, Wo = p.W >> 1;

    int bid = blockIdx.x;
    const int twi = bid % p.ntwo; bid /= p.ntwo;
    const int thi = bid % p.ntho; bid /= p.ntho;
    const int b = bid, oh0 = thi * p.tho, ow0 = twi * p.two;
    const float *xb = p.x + (size_t)b * CX * S;

    for (int i = tid; i < CIN * CB; i += kDownThreads) s_w1[i] = p.w1[(i % CB) * CIN + i / CB];
    for (int i = tid; i < CB * 16 * CB * 4; i += kDownThreads) {
        const int kz = i & 3, co = (i >> 2) % CB, hw = (i / (4 * CB)) % 16, ci = i / (64 * CB);
        s_w2[i] = p.w2[((size_t)co * CB + ci) * 64 + hw * 4 + kz];
    }
    for (int i = tid; i < CB * COUT; i += kDownThreads) s_w3[i] = p.w3[(i % COUT) * CB + i / COUT];
    for (int i = tid; i < CIN * 8 * COUT; i += kDownThreads) {
        const int c = i % COUT, t = (i / COUT) % 8, ci = i / (8 * COUT);
        s_ws[i] = p.ws[((size_t)c * CIN + ci) * 8 + t];
    }
    if (PARSE && tid < CIN) { s_pw[tid] = p.pre_w[tid]; s_pw[CIN + tid] = p.pre_b ? p.pre_b[tid] : 0.0f; }
    const float b1a = ld_scalar(p.b1a, 0.f), b1b = ld_scalar(p.b1b, 0.f), b2a = ld_scalar(p.b2a, 0.f), b2b = ld_scalar(p.b2b, 0.f);
    __syncthreads();

    // ---- stage A: t1 on the haloed input rows, raw centre rows for the skip ---------------------------------
    {
        const int zq = tid % ZQ, slot = tid / ZQ, nslots = kDownThreads / ZQ;
        // (as in preact_row_kernel: the loads of the next row slot are issued before the arithmetic of the current
        //  one, when the registers allow it)
        constexpr bool kPrefetch = CX <= 4;
        float4 nxt[CX];
        auto issue = [&](int rs) {
            const int lh = rs / IW, lw = rs - lh * IW;
            const int gh = rmod(2 * oh0 - 1 + lh, p.H), gw = rmod(2 * ow0 - 1 + lw, p.W);
            const float *px = xb + ((size_t)gh * p.W + gw) * Z + 4 * zq;
#pragma unroll
            for (int c = 0; c < CX; ++c) nxt[c] = __ldg(reinterpret_cast<const float4 *>(px + (size_t)c * S));
        };
        if (kPrefetch && slot < nrows_in) issue(slot);
        for (int rs = slot; rs < nrows_in; rs += nslots) {
            if (!kPrefetch) issue(rs);
            float4 raw[CX];
#pragma unroll
            for (int c = 0; c < CX; ++c) raw[c] = nxt[c];
            if (kPrefetch && rs + nslots < nrows_in) issue(rs + nslots);
            float t[CB][4];
#pragma unroll
            for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                for (int k = 0; k < 4; ++k) t[cb][k] = 0.0f;
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                float4 v;
                if (PARSE) {
                    const float w = s_pw[c], bb = s_pw[CIN + c];
                    v.x = __fmaf_rn(w, raw[0].x, bb); v.y = __fmaf_rn(w, raw[0].y, bb); v.z = __fmaf_rn(w, raw[0].z, bb); v.w = __fmaf_rn(w, raw[0].w, bb);
                } else {
                    v = raw[PARSE ? 0 : c];
                }
                const float a0 = elu1(v.x + b1a) + b1b, a1 = elu1(v.y + b1a) + b1b, a2 = elu1(v.z + b1a) + b1b, a3 = elu1(v.w + b1a) + b1b;
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) {
                    const float w = s_w1[c * CB + cb];
                    t[cb][0] = __fmaf_rn(w, a0, t[cb][0]); t[cb][1] = __fmaf_rn(w, a1, t[cb][1]);
                    t[cb][2] = __fmaf_rn(w, a2, t[cb][2]); t[cb][3] = __fmaf_rn(w, a3, t[cb][3]);
                }
            }
#pragma unroll
            for (int cb = 0; cb < CB; ++cb) {
                float4 o;
                o.x = elu1(t[cb][0] + b2a) + b2b; o.y = elu1(t[cb][1] + b2a) + b2b;
                o.z = elu1(t[cb][2] + b2a) + b2b; o.w = elu1(t[cb][3] + b2a) + b2b;
                float *row = s_t1 + ((size_t)cb * nrows_in + rs) * ZP;
                *reinterpret_cast<float4 *>(row + 4 + 4 * zq) = o;
                if (zq == 0) row[Z + 4] = o.x;
                if (zq == ZQ - 1) row[3] = o.w;
            }
        }
    }
    __syncthreads();

    // ---- stage B + C: thread = (output row slot, 4 output z, ONE branch channel co) ---------------------------
    const float b3a = ld_scalar(p.b3a, 0.f), b3b = ld_scalar(p.b3b, 0.f), b4 = ld_scalar(p.b4, 0.f), sc = ld_scalar(p.scale, 1.f);
    const float b1c = ld_scalar(p.b1c, 0.f), b1d = ld_scalar(p.b1d, 0.f);
    const int co = tid % CB, lane = (tid / CB) % ZQo, slot = tid / (CB * ZQo), nslots = kDownThreads / (CB * ZQo);
    const int nrounds = (nrows_out + nslots - 1) / nslots;
    for (int rd = 0; rd < nrounds; ++rd) {             // every thread runs every round: the shuffles below are warp-wide
        const int ro = rd * nslots + slot;
        const bool live = ro < nrows_out;
        const int lho = live ? ro / p.two : 0, lwo = live ? ro - lho * p.two : 0;
        const int oh = oh0 + lho, ow = ow0 + lwo;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ci = 0; ci < CB; ++ci) {
#pragma unroll
            for (int hw = 0; hw < 16; ++hw) {
                const int kh = hw >> 2, kw = hw & 3;
                const float *row = s_t1 + ((size_t)ci * nrows_in + (2 * lho + kh) * IW + 2 * lwo + kw) * ZP + 8 * lane;
                const float4 m0 = *reinterpret_cast<const float4 *>(row + 4), m1 = *reinterpret_cast<const float4 *>(row + 8);
                const float v[10] = {row[3], m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, row[12]};
                const float4 w = *reinterpret_cast<const float4 *>(s_w2 + (((size_t)ci * 16 + hw) * CB + co) * 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[k] = __fmaf_rn(w.x, v[2 * k], acc[k]); acc[k] = __fmaf_rn(w.y, v[2 * k + 1], acc[k]);
                    acc[k] = __fmaf_rn(w.z, v[2 * k + 2], acc[k]); acc[k] = __fmaf_rn(w.w, v[2 * k + 3], acc[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = elu1(acc[k] + b3a) + b3b;
        // conv3: this lane produces output channels co*NPC .. co*NPC+NPC-1 from the CB lanes' t2
        float out[NPC][4];
#pragma unroll
        for (int j = 0; j < NPC; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) out[j][k] = 0.0f;
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
            float tv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) tv[k] = __shfl_sync(0xffffffffu, acc[k], cb, CB);
            if constexpr (NPC == 2) {               // the lane's two output channels in one packed FMA
                const float2 w = *reinterpret_cast<const float2 *>(s_w3 + cb * COUT + co * NPC);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 o = ffma2_bcast(w, tv[k], make_float2(out[0][k], out[1][k]));
                    out[0][k] = o.x; out[1][k] = o.y;
                }
            } else {
#pragma unroll
            for (int j = 0; j < NPC; ++j) {
                const float w = s_w3[cb * COUT + co * NPC + j];
#pragma unroll
                for (int k = 0; k < 4; ++k) out[j][k] = __fmaf_rn(w, tv[k], out[j][k]);
            }
            }
        }
        if (live && oh < Ho && ow < Wo) {
            // skip: k2 s2 on (x + b1c), zero padding never reached (even extents)
            float sk[NPC][4];
#pragma unroll
            for (int j = 0; j < NPC; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) sk[j][k] = 0.0f;
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) {
                // the raw centre rows were read by stage A a moment ago: L1/L2 hits (keeping them in shared memory
                // instead cost a third of the tile and forced 2 x 1 tiles with a 3x halo recompute)
                const float *rx0 = xb + ((size_t)(2 * oh + (t2 >> 1)) * p.W + 2 * ow + (t2 & 1)) * Z + 8 * lane;
                float4 r0[CX], r1[CX];
#pragma unroll
                for (int c = 0; c < CX; ++c) {
                    const float *rx = rx0 + (size_t)c * S;
                    r0[c] = __ldg(reinterpret_cast<const float4 *>(rx)); r1[c] = __ldg(reinterpret_cast<const float4 *>(rx + 4));
                }
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    float xv[8];
                    const float4 a = r0[PARSE ? 0 : ci], bq = r1[PARSE ? 0 : ci];
                    xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = bq.x; xv[5] = bq.y; xv[6] = bq.z; xv[7] = bq.w;
                    if (PARSE) {
                        const float w = s_pw[ci], bb = s_pw[CIN + ci];
#pragma unroll
                        for (int e = 0; e < 8; ++e) xv[e] = __fmaf_rn(w, xv[e], bb);
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) xv[e] += b1c;
                    if constexpr (NPC == 2) {
                        const float2 w0 = *reinterpret_cast<const float2 *>(s_ws + (ci * 8 + t2 * 2 + 0) * COUT + co * NPC);
                        const float2 w1 = *reinterpret_cast<const float2 *>(s_ws + (ci * 8 + t2 * 2 + 1) * COUT + co * NPC);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 o = ffma2_bcast(w1, xv[2 * k + 1], ffma2_bcast(w0, xv[2 * k], make_float2(sk[0][k], sk[1][k])));
                            sk[0][k] = o.x; sk[1][k] = o.y;
                        }
                    } else {
#pragma unroll
                    for (int j = 0; j < NPC; ++j) {
                        const float w0 = s_ws[(ci * 8 + t2 * 2 + 0) * COUT + co * NPC + j], w1 = s_ws[(ci * 8 + t2 * 2 + 1) * COUT + co * NPC + j];
#pragma unroll
                        for (int k = 0; k < 4; ++k) sk[j][k] = __fmaf_rn(w1, xv[2 * k + 1], __fmaf_rn(w0, xv[2 * k], sk[j][k]));
                    }
                    }
                }
            }
            const size_t off = ((size_t)oh * Wo + ow) * Zo + 4 * lane;
#pragma unroll
            for (int j = 0; j < NPC; ++j) {
                float4 yv;
                yv.x = __fmaf_rn(out[j][0], sc, b4) + (sk[j][0] + b1d); yv.y = __fmaf_rn(out[j][1], sc, b4) + (sk[j][1] + b1d);
                yv.z = __fmaf_rn(out[j][2], sc, b4) + (sk[j][2] + b1d); yv.w = __fmaf_rn(out[j][3], sc, b4) + (sk[j][3] + b1d);
                *reinterpret_cast<float4 *>(p.y + ((size_t)b * COUT + co * NPC + j) * So + off) = yv;
            }
        }
    }
}

template <int CIN, int CB, int COUT, bool PARSE>
static int launch_down_row(const vq3d_preact_desc *d, void *stream) {
    using SM = DownSmem<CIN, CB, COUT, PARSE>;
    DownParams p;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z;
    const int Ho = d->H / 2, Wo = d->W / 2;
    int tho = Ho < 2 ? Ho : 2, two = Wo < 2 ? Wo : 2;
    const size_t cap = 112 * 1024;
    while (SM::floats(tho, two, d->Z) * 4 > cap) {
        if (tho >= two && tho > 1) tho = 1;
        else if (two > 1) two = 1;
        else return fail(VQ3D_ERR_UNSUPPORTED, "preact_block(down row): tile does not fit shared memory");
    }
    p.tho = tho; p.two = two; p.ntho = (int)ceil_div(Ho, tho); p.ntwo = (int)ceil_div(Wo, two);
    p.x = d->x; p.w1 = d->w1; p.w2 = d->w2; p.w3 = d->w3; p.ws = d->wskip;
    p.b1a = d->b1a; p.b1b = d->b1b; p.b2a = d->b2a; p.b2b = d->b2b; p.b3a = d->b3a; p.b3b = d->b3b; p.b4 = d->b4; p.scale = d->scale;
    p.b1c = d->b1c; p.b1d = d->b1d; p.pre_w = d->pre_w; p.pre_b = d->pre_b; p.y = d->y;
    const int64_t grid = (int64_t)d->B * p.ntho * p.ntwo;
    if (grid > 0x7fffffff) return fail(VQ3D_ERR_INVALID, "preact_block(down row): grid too large");
    return launch("preact_down_row", preact_down_row_kernel<CIN, CB, COUT, PARSE>, dim3((unsigned)grid), dim3(kDownThreads),
                  SM::floats(tho, two, d->Z) * 4, stream, p);
}

// ---------------------------------------------------------------------------------------------------
// 'up' blocks at the big levels (8 -> 4 -> 4 at 256^3 -> 512^3, 18 -> 9 -> 8 at 128^3 -> 256^3; layers.py:128-132,
// 591-597): conv1 1x1 and the 1x1 skip are evaluated at LOW resolution (they commute with the trilinear x2
// upsampling), the branch is upsampled into a haloed hi-res shared-memory tile (circular wrap of the k3 padding is
// applied on the hi-res grid, the upsampling clamps at the volume edges -- as the reference), conv2 k3 runs from
// that tile with 4 consecutive hi-res z per thread, the skip is upsampled on the fly in the epilogue.
// Low-res rows are stored with one clamped element of padding on each side, which turns the depth interpolation of
// hi z = 4q..4q+3 into fixed-weight blends of padded entries 2q..2q+3 with no edge cases.
#ifndef VQ3D_UP_NW2
#define VQ3D_UP_NW2 1
#endif
constexpr int kUpThreads = 256;

struct UpParams {
    int B, H, W, Z;                   // low-res input extent
    int tho, two, ntho, ntwo;         // hi-res output-row tile (even)
    const float *x, *w1, *w2, *w3, *ws;
    const float *b1a, *b1b, *b2a, *b2b, *b3a, *b3b, *b4, *scale, *b1c, *b1d;
    float *y;
};

template <int CIN, int CB, int COUT>
struct UpSmem {
    static constexpr int w1 = 0;                         // [CIN][CB]
    static constexpr int w2 = w1 + CIN * CB;             // [CB ci][9][3 kz][CB co]
    static constexpr int w3 = w2 + CB * 27 * CB;         // [CB][COUT]
    static constexpr int ws = w3 + CB * COUT;            // [CIN][COUT]
    static constexpr int tab = (ws + CIN * COUT + 3) & ~3;   // tap tables: 3 floats per hi tile row / column
    __host__ __device__ static int zpl(int Z) { return (Z + 2 + 3) & ~3; }
    static size_t floats(int tho, int two, int Z) {
        const int LR = (tho / 2 + 2) * (two / 2 + 2), HR = (tho + 2) * (two + 2);
        return tab + 3 * (size_t)(tho + 2 + two + 2) + (size_t)(CB + COUT) * LR * zpl(Z) + (size_t)CB * HR * (2 * Z + 8) + 4;
    }
};

__device__ __forceinline__ void up_taps_row(int o, int n, int &i0, int &i1, float &l1) {
    float src = 0.5f * (float)o - 0.25f;
    if (src < 0.0f) src = 0.0f;
    i0 = (int)src;
    l1 = src - (float)i0;
    i1 = i0 + (i0 < n - 1 ? 1 : 0);
}

// depth interpolation of 4 consecutive hi-res z from the padded low-res entries p[0..3] (= padded index 2q..2q+3)
__device__ __forceinline__ void up_z4(const float *p, float *o) {
    o[0] = 0.25f * p[0] + 0.75f * p[1];
    o[1] = 0.75f * p[1] + 0.25f * p[2];
    o[2] = 0.25f * p[1] + 0.75f * p[2];
    o[3] = 0.75f * p[2] + 0.25f * p[3];
}

template <int CIN, int CB, int COUT>
__global__ void __launch_bounds__(kUpThreads)
preact_up_row_kernel(UpParams p) {
    using SM = UpSmem<CIN, CB, COUT>;
    VQ3D_DYN_SMEM(float, smem);
    float *s_w1 = smem + SM::w1, *s_w2 = smem + SM::w2, *s_w3 = smem + SM::w3, *s_ws = smem + SM::ws, *s_tab = smem + SM::tab;
    const int tid = threadIdx.x;
    const int Z = p.Z, Zo = 2 * Z, ZQ = Z >> 2, ZQo = Zo >> 2, ZPL = SM::zpl(Z), ZPH = Zo + 8;
    const int Ho = 2 * p.H, Wo = 2 * p.W;
    const int IH = p.tho + 2, IW = p.two + 2, HR = IH * IW;
    const int LH = p.tho / 2 + 2, LW = p.two / 2 + 2, LR = LH * LW;
    float *s_lo = s_tab + 3 * (IH + IW);                 // [CB][LR][ZPL]
    s_lo = smem + (((s_lo - smem) + 3) & ~3);
    float *s_sk = s_lo + (size_t)CB * LR * ZPL;          // [COUT][LR][ZPL]
    float *s_hi = s_sk + (size_t)COUT * LR * ZPL;        // [CB][HR][ZPH]
    const int64_t S = (int64_t)p.H * p.W * Z, So = 8 * S;

    int bid = blockIdx.x;
    const int twi = bid % p.ntwo; bid /= p.ntwo;
    const int thi = bid % p.ntho; bid /= p.ntho;
    const int b = bid, oh0 = thi * p.tho, ow0 = twi * p.two;
    const int lh0 = oh0 / 2 - 1, lw0 = ow0 / 2 - 1;      // global low-res row / column of low tile index 0 (before wrapping)
    const float *xb = p.x + (size_t)b * CIN * S;

    for (int i = tid; i < CIN * CB; i += kUpThreads) s_w1[i] = p.w1[(i % CB) * CIN + i / CB];
    for (int i = tid; i < CB * 27 * CB; i += kUpThreads) {
        const int co = i % CB, t = (i / CB) % 27, ci = i / (CB * 27);
        s_w2[i] = p.w2[((size_t)co * CB + ci) * 27 + t];
    }
    for (int i = tid; i < CB * COUT; i += kUpThreads) s_w3[i] = p.w3[(i % COUT) * CB + i / COUT];
    for (int i = tid; i < CIN * COUT; i += kUpThreads) s_ws[i] = p.ws[(i % COUT) * CIN + i / COUT];
    // tap tables of the hi tile rows / columns (hi index wraps: conv padding; low taps clamp: upsampling)
    for (int i = tid; i < IH + IW; i += kUpThreads) {
        const bool isw = i >= IH;
        const int q = isw ? i - IH : i;
        const int pos = rmod((isw ? ow0 : oh0) - 1 + q, isw ? Wo : Ho);
        int i0, i1; float l1;
        up_taps_row(pos, isw ? p.W : p.H, i0, i1, l1);
        const int n = isw ? p.W : p.H, l0 = isw ? lw0 : lh0, ext = isw ? LW : LH;
        s_tab[3 * i + 0] = (float)min(rmod(i0 - l0, n), ext - 1);
        s_tab[3 * i + 1] = (float)min(rmod(i1 - l0, n), ext - 1);
        s_tab[3 * i + 2] = l1;
    }
    const float b1a = ld_scalar(p.b1a, 0.f), b1b = ld_scalar(p.b1b, 0.f), b2a = ld_scalar(p.b2a, 0.f), b2b = ld_scalar(p.b2b, 0.f);
    const float b1c = ld_scalar(p.b1c, 0.f);
    __syncthreads();

    // ---- stage A1: low-res t1 and skip rows (padded by one clamped element on each side) ---------------------
    {
        const int zq = tid % ZQ, slot = tid / ZQ, nslots = kUpThreads / ZQ;
        for (int rs = slot; rs < LR; rs += nslots) {
            const int lh = rs / LW, lw = rs - lh * LW;
            const int gh = rmod(lh0 + lh, p.H), gw = rmod(lw0 + lw, p.W);
            const float *px = xb + ((size_t)gh * p.W + gw) * Z + 4 * zq;
            float t[CB][4], sk[COUT][4];
#pragma unroll
            for (int c = 0; c < CB; ++c) t[c][0] = t[c][1] = t[c][2] = t[c][3] = 0.0f;
#pragma unroll
            for (int c = 0; c < COUT; ++c) sk[c][0] = sk[c][1] = sk[c][2] = sk[c][3] = 0.0f;
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(px + (size_t)c * S));
                const float xv[4] = {v.x, v.y, v.z, v.w};
                float a[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) a[k] = elu1(xv[k] + b1a) + b1b;
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) {
                    const float w = s_w1[c * CB + cb];
#pragma unroll
                    for (int k = 0; k < 4; ++k) t[cb][k] = __fmaf_rn(w, a[k], t[cb][k]);
                }
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    const float w = s_ws[c * COUT + co];
#pragma unroll
                    for (int k = 0; k < 4; ++k) sk[co][k] = __fmaf_rn(w, xv[k] + b1c, sk[co][k]);
                }
            }
#pragma unroll
            for (int cb = 0; cb < CB; ++cb) {
                float *row = s_lo + ((size_t)cb * LR + rs) * ZPL + 1 + 4 * zq;
#pragma unroll
                for (int k = 0; k < 4; ++k) { t[cb][k] = elu1(t[cb][k] + b2a) + b2b; row[k] = t[cb][k]; }
                if (zq == 0) row[-1] = t[cb][0];
                if (zq == ZQ - 1) row[4] = t[cb][3];
            }
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                float *row = s_sk + ((size_t)co * LR + rs) * ZPL + 1 + 4 * zq;
#pragma unroll
                for (int k = 0; k < 4; ++k) row[k] = sk[co][k];
                if (zq == 0) row[-1] = sk[co][0];
                if (zq == ZQ - 1) row[4] = sk[co][3];
            }
        }
    }
    __syncthreads();

    const int q = tid % ZQo, slot = tid / ZQo, nslots = kUpThreads / ZQo;
    // ---- stage A2: trilinear x2 into the haloed hi-res tile ------------------------------------------------
    for (int hr = slot; hr < HR; hr += nslots) {
        const int qh = hr / IW, qw = hr - qh * IW;
        const int h0 = (int)s_tab[3 * qh], h1 = (int)s_tab[3 * qh + 1]; const float lh = s_tab[3 * qh + 2];
        const int w0 = (int)s_tab[3 * (IH + qw)], w1 = (int)s_tab[3 * (IH + qw) + 1]; const float lw = s_tab[3 * (IH + qw) + 2];
        const int r00 = h0 * LW + w0, r01 = h0 * LW + w1, r10 = h1 * LW + w0, r11 = h1 * LW + w1;
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
            const float *base = s_lo + (size_t)cb * LR * ZPL + 2 * q;
            float a00[4], a01[4], a10[4], a11[4];
            up_z4(base + (size_t)r00 * ZPL, a00); up_z4(base + (size_t)r01 * ZPL, a01);
            up_z4(base + (size_t)r10 * ZPL, a10); up_z4(base + (size_t)r11 * ZPL, a11);
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float c0 = a00[k] * (1.f - lw) + a01[k] * lw, c1 = a10[k] * (1.f - lw) + a11[k] * lw;
                o[k] = c0 * (1.f - lh) + c1 * lh;
            }
            float *row = s_hi + ((size_t)cb * HR + hr) * ZPH;
            *reinterpret_cast<float4 *>(row + 4 + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
            if (q == 0) row[Zo + 4] = o[0];              // circular padding on the hi-res grid
            if (q == ZQo - 1) row[3] = o[3];
        }
    }
    __syncthreads();

    // ---- stage B + C ---------------------------------------------------------------------------------------
    const float b3a = ld_scalar(p.b3a, 0.f), b3b = ld_scalar(p.b3b, 0.f), b4 = ld_scalar(p.b4, 0.f), sc = ld_scalar(p.scale, 1.f);
    const float b1d = ld_scalar(p.b1d, 0.f);
    auto up_epilogue = [&](float (&acc)[CB][4], const int lho, const int lwo, const int oh, const int ow) {
#pragma unroll
        for (int co = 0; co < CB; ++co)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[co][k] = elu1(acc[co][k] + b3a) + b3b;
        // skip: trilinear x2 of the low-res 1x1 skip at this hi row (tile rows lho+1, lwo+1 of the tap tables)
        const int qh = lho + 1, qw = lwo + 1;
        const int h0 = (int)s_tab[3 * qh], h1 = (int)s_tab[3 * qh + 1]; const float lh = s_tab[3 * qh + 2];
        const int w0 = (int)s_tab[3 * (IH + qw)], w1 = (int)s_tab[3 * (IH + qw) + 1]; const float lw = s_tab[3 * (IH + qw) + 2];
        const int r00 = h0 * LW + w0, r01 = h0 * LW + w1, r10 = h1 * LW + w0, r11 = h1 * LW + w1;
        const size_t off = ((size_t)oh * Wo + ow) * Zo + 4 * q;
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            float out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int cb = 0; cb < CB; ++cb) {
                const float w = s_w3[cb * COUT + c];
#pragma unroll
                for (int k = 0; k < 4; ++k) out[k] = __fmaf_rn(w, acc[cb][k], out[k]);
            }
            const float *base = s_sk + (size_t)c * LR * ZPL + 2 * q;
            float a00[4], a01[4], a10[4], a11[4];
            up_z4(base + (size_t)r00 * ZPL, a00); up_z4(base + (size_t)r01 * ZPL, a01);
            up_z4(base + (size_t)r10 * ZPL, a10); up_z4(base + (size_t)r11 * ZPL, a11);
            float yv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float c0 = a00[k] * (1.f - lw) + a01[k] * lw, c1 = a10[k] * (1.f - lw) + a11[k] * lw;
                yv[k] = __fmaf_rn(out[k], sc, b4) + (c0 * (1.f - lh) + c1 * lh + b1d);
            }
            *reinterpret_cast<float4 *>(p.y + ((size_t)b * COUT + c) * So + off) = make_float4(yv[0], yv[1], yv[2], yv[3]);
        }
    };
    if constexpr (CB % 2 == 0 && VQ3D_UP_NW2) {
        // two neighbouring hi-res output rows (w, w + 1) per thread, as in preact_row_kernel: windows and weights are
        // loaded once for both
        constexpr int CP = CB / 2;
        const int two2 = p.two >> 1;
        for (int ro = slot; ro < p.tho * two2; ro += nslots) {
            const int lho = ro / two2, lwo = 2 * (ro - lho * two2);
            const int oh = oh0 + lho, ow = ow0 + lwo;
            if (oh >= Ho || ow >= Wo) continue;
            float2 acc2[2][CP][4];
#pragma unroll
            for (int rw = 0; rw < 2; ++rw)
#pragma unroll
                for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc2[rw][cp][k] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int ci = 0; ci < CB; ++ci) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    float2 wprev[3][CP];
#pragma unroll
                    for (int kwi = 0; kwi < 4; ++kwi) {
                        const float *row = s_hi + ((size_t)ci * HR + (lho + kh) * IW + lwo + kwi) * ZPH + 4 * q;
                        const float4 m = *reinterpret_cast<const float4 *>(row + 4);
                        const float r[6] = {row[3], m.x, m.y, m.z, m.w, row[8]};
                        float2 wcur[3][CP];
                        if (kwi < 3) {
                            const float2 *wt = reinterpret_cast<const float2 *>(s_w2 + ((ci * 9 + kh * 3 + kwi) * 3) * CB);
#pragma unroll
                            for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                                for (int cp = 0; cp < CP; ++cp) wcur[kz][cp] = wt[kz * CP + cp];
                        }
#pragma unroll
                        for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                            for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (kwi < 3) acc2[0][cp][k] = ffma2_bcast(wcur[kz][cp], r[k + kz], acc2[0][cp][k]);
                                    if (kwi > 0) acc2[1][cp][k] = ffma2_bcast(wprev[kz][cp], r[k + kz], acc2[1][cp][k]);
                                }
                        if (kwi < 3) {
#pragma unroll
                            for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                                for (int cp = 0; cp < CP; ++cp) wprev[kz][cp] = wcur[kz][cp];
                        }
                    }
                }
            }
#pragma unroll
            for (int rw = 0; rw < 2; ++rw) {
                float acc[CB][4];
#pragma unroll
                for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) { acc[2 * cp][k] = acc2[rw][cp][k].x; acc[2 * cp + 1][k] = acc2[rw][cp][k].y; }
                up_epilogue(acc, lho, lwo + rw, oh, ow + rw);
            }
        }
    } else
    for (int ro = slot; ro < p.tho * p.two; ro += nslots) {
        const int lho = ro / p.two, lwo = ro - lho * p.two;
        const int oh = oh0 + lho, ow = ow0 + lwo;
        if (oh >= Ho || ow >= Wo) continue;
        float acc[CB][4];
        if constexpr (CB % 2 == 0) {
            constexpr int CP = CB / 2;      // two branch channels per FFMA2
            float2 acc2[CP][4];
#pragma unroll
            for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc2[cp][k] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int ci = 0; ci < CB; ++ci) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float *row = s_hi + ((size_t)ci * HR + (lho + kh) * IW + lwo + kw) * ZPH + 4 * q;
                        const float4 m = *reinterpret_cast<const float4 *>(row + 4);
                        const float r[6] = {row[3], m.x, m.y, m.z, m.w, row[8]};
                        const float2 *wt = reinterpret_cast<const float2 *>(s_w2 + ((ci * 9 + kh * 3 + kw) * 3) * CB);
#pragma unroll
                        for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                            for (int cp = 0; cp < CP; ++cp) {
                                const float2 w = wt[kz * CP + cp];
#pragma unroll
                                for (int k = 0; k < 4; ++k) acc2[cp][k] = ffma2_bcast(w, r[k + kz], acc2[cp][k]);
                            }
                    }
                }
            }
#pragma unroll
            for (int cp = 0; cp < CP; ++cp)
#pragma unroll
                for (int k = 0; k < 4; ++k) { acc[2 * cp][k] = acc2[cp][k].x; acc[2 * cp + 1][k] = acc2[cp][k].y; }
        } else {
#pragma unroll
        for (int co = 0; co < CB; ++co) acc[co][0] = acc[co][1] = acc[co][2] = acc[co][3] = 0.0f;
#pragma unroll
        for (int ci = 0; ci < CB; ++ci) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float *row = s_hi + ((size_t)ci * HR + (lho + kh) * IW + lwo + kw) * ZPH + 4 * q;
                    const float4 m = *reinterpret_cast<const float4 *>(row + 4);
                    const float r[6] = {row[3], m.x, m.y, m.z, m.w, row[8]};
                    const float *wt = s_w2 + ((ci * 9 + kh * 3 + kw) * 3) * CB;
#pragma unroll
                    for (int kz = 0; kz < 3; ++kz)
#pragma unroll
                        for (int co = 0; co < CB; ++co) {
                            const float w = wt[kz * CB + co];
#pragma unroll
                            for (int k = 0; k < 4; ++k) acc[co][k] = __fmaf_rn(w, r[k + kz], acc[co][k]);
                        }
                }
            }
        }
        }
        up_epilogue(acc, lho, lwo, oh, ow);
    }
}

template <int CIN, int CB, int COUT>
static int launch_up_row(const vq3d_preact_desc *d, void *stream) {
    using SM = UpSmem<CIN, CB, COUT>;
    UpParams p;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z;
    const int Ho = 2 * d->H, Wo = 2 * d->W;
    int tho = Ho < 4 ? Ho : 4, two = Wo < 4 ? Wo : 4;
    const size_t cap = 113 * 1024;
    while (SM::floats(tho, two, d->Z) * 4 > cap) {
        if (tho >= two && tho > 2) tho = 2;
        else if (two > 2) two = 2;
        else return fail(VQ3D_ERR_UNSUPPORTED, "preact_block(up row): tile does not fit shared memory");
    }
    p.tho = tho; p.two = two; p.ntho = (int)ceil_div(Ho, tho); p.ntwo = (int)ceil_div(Wo, two);
    p.x = d->x; p.w1 = d->w1; p.w2 = d->w2; p.w3 = d->w3; p.ws = d->wskip;
    p.b1a = d->b1a; p.b1b = d->b1b; p.b2a = d->b2a; p.b2b = d->b2b; p.b3a = d->b3a; p.b3b = d->b3b; p.b4 = d->b4; p.scale = d->scale;
    p.b1c = d->b1c; p.b1d = d->b1d; p.y = d->y;
    const int64_t grid = (int64_t)d->B * p.ntho * p.ntwo;
    if (grid > 0x7fffffff) return fail(VQ3D_ERR_INVALID, "preact_block(up row): grid too large");
    return launch("preact_up_row", preact_up_row_kernel<CIN, CB, COUT>, dim3((unsigned)grid), dim3(kUpThreads), SM::floats(tho, two, d->Z) * 4, stream, p);
}

// Z must be a multiple of 4 with Z/4 a power of two <= 32 (a row of Z/4 float4 lanes divides the CTA)
static bool row_shape_ok(const vq3d_preact_desc *d) {
    const int zq = d->Z / 4;
    return d->Z % 4 == 0 && zq >= 1 && zq <= 32 && (zq & (zq - 1)) == 0 &&
           (reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->out_w ? d->out_y : d->y) & 15) == 0;
}

int preact_row_dispatch(const vq3d_preact_desc *d, void *stream, bool *handled) {
    *handled = false;
    if (d->mode == 1 && d->wskip && !d->out_w) {
        // down: even extents, Z/8 output lanes per row and CB lanes per voxel must tile the 256-thread CTA
        const int zqo = d->Z / 8;
        const bool shape = d->Z % 8 == 0 && zqo >= 1 && (zqo & (zqo - 1)) == 0 && d->Z / 4 <= 32 && !((d->H | d->W) & 1) &&
                           (reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 15) == 0;
        int (*fn)(const vq3d_preact_desc *, void *) = nullptr;
        if (shape && d->Cin == 4 && d->Cb == 4 && d->Cout == 8 && zqo * 4 <= kDownThreads) fn = d->pre_w ? launch_down_row<4, 4, 8, true> : launch_down_row<4, 4, 8, false>;
        else if (shape && !d->pre_w && d->Cin == 8 && d->Cb == 8 && d->Cout == 16 && zqo * 8 <= kDownThreads) fn = launch_down_row<8, 8, 16, false>;
        if (fn) { *handled = true; return fn(d, stream); }
        return VQ3D_OK;
    }
    if (d->mode == 2 && d->wskip && !d->out_w && !d->pre_w) {
        // up: Z/4 low-res lanes and 2Z/4 hi-res lanes per row must tile the 256-thread CTA; low-res extents >= 2
        const int zq = d->Z / 4, zqo = d->Z / 2;
        const bool shape = d->Z % 4 == 0 && zq >= 1 && (zq & (zq - 1)) == 0 && zqo <= 64 && zqo <= kUpThreads && d->H >= 2 && d->W >= 2 &&
                           (reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 15) == 0;
        int (*fn)(const vq3d_preact_desc *, void *) = nullptr;
        if (shape && d->Cin == 8 && d->Cb == 4 && d->Cout == 4) fn = launch_up_row<8, 4, 4>;
        else if (shape && d->Cin == 4 && d->Cb == 2 && d->Cout == 2) fn = launch_up_row<4, 2, 2>;
        // (wider branches -- 16 -> 8 -> 8, 18 -> 9 -> 8 -- measured slower here than preact_fused_kernel: the fully unrolled
        //  conv2 needs > 200 registers; they stay on the generic fused kernel)
        if (fn) {
            const int rc = fn(d, stream);
            if (rc != VQ3D_ERR_UNSUPPORTED) { *handled = true; return rc; }
        }
        return VQ3D_OK;
    }
    if (d->mode != 0 || d->wskip || d->Cin != d->Cout || !row_shape_ok(d)) return VQ3D_OK;
    const bool outc = d->out_w != nullptr;
    int (*fn)(const vq3d_preact_desc *, void *) = nullptr;
    if (d->Cin == 2 && d->Cb == 1) fn = outc ? launch_row<2, 1, true> : launch_row<2, 1, false>;
    else if (d->Cin == 4 && d->Cb == 2) fn = outc ? launch_row<4, 2, true> : launch_row<4, 2, false>;
    else if (d->Cin == 8 && d->Cb == 4) fn = outc ? launch_row<8, 4, true> : launch_row<8, 4, false>;
    if (!fn) return VQ3D_OK;
    *handled = true;
    return fn(d, stream);
}

}  // namespace vq3d
