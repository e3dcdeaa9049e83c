// preact_kernels.cu -- one launch per PreActFixupResBlock (vqvae/layers.py:102-216).
//
//   o = conv1x1(ELU(x+b1a)+b1b); o = conv_k(ELU(o+b2a)+b2b); o = conv1x1(ELU(o+b3a)+b3b)
//   y = o*scale + b4 + (skip(x+b1c)+b1d | x)
//
// The reference runs ~16 ATen kernels per block and round-trips every intermediate (and a
// circularly padded copy, layers.py:109) through HBM.  Here a CTA owns a (th, tw, tz) tile of
// output voxels (tz along the contiguous depth axis) and does the whole block in three
// shared-memory stages, so HBM sees exactly: x once (+halo), y once.
//
//   stage A  t1 = ELU(conv1(ELU(x+b1a)+b1b)+b2a)+b2b on the tile + halo, wrapped coordinates
//            (circular padding costs nothing: the halo is just read from the other side).
//            mode down: the tile is the (2t+2)^3 input window of the k4 s2 conv.
//            mode up:   t1 and the 1x1 skip are evaluated on the low-res window, then
//                       trilinearly upsampled (align_corners=False) into the hi-res halo tile;
//                       conv1x1 and the upsample commute, so the skip costs 1/8 of the reference's.
//   stage B  k^3 conv from shared memory; a thread owns VX outputs along W x all CB channels and
//            slides a register window along W; weights are warp-broadcast 128-bit shared loads.
//   stage C  ELU, conv3, *scale + b4, skip / residual, coalesced store (threads run along Z).
//
// fp32 SIMT: this file serves the small channel counts of the model (branch widths 1..16),
// where a 16-wide tensor-core K/N would be mostly padding.
#include "vq3d_rt.h"

namespace vq3d {

int preact_row_dispatch(const vq3d_preact_desc *d, void *stream, bool *handled);    // preact_row_kernels.cu

struct PreactParams {
    int B, H, W, Z;            // input spatial
    int Ho, Wo, Zo;            // output spatial
    int th, tw, tz;            // output tile
    int nth, ntw, ntz;         // tiles per axis
    const float *x, *w1, *w2, *w3, *ws;
    const float *b1a, *b1b, *b2a, *b2b, *b3a, *b3b, *b4, *scale, *b1c, *b1d;
    float *y;
};

__host__ __device__ __forceinline__ int pmod(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}

__device__ __forceinline__ void up_taps_dev(int o, int n, int &i0, int &i1, float &l1) {
    float src = 0.5f * (float)o - 0.25f;
    if (src < 0.0f) src = 0.0f;
    i0 = (int)src;
    l1 = src - (float)i0;
    i1 = i0 + (i0 < n - 1 ? 1 : 0);
}

constexpr int round4(int v) { return (v + 3) & ~3; }

// MODE 0 same, 1 down, 2 up
template <int MODE> struct Geo {
    static constexpr int K = MODE == 1 ? 4 : 3;
    static constexpr int ST = MODE == 1 ? 2 : 1;
};

#ifndef VQ3D_FUSED_FFMA2
#define VQ3D_FUSED_FFMA2 1
#endif
template <int CIN, int CB, int COUT, int MODE, bool SKIP, int VX>
struct PreactSmem {
    static constexpr int K = Geo<MODE>::K, K3 = K * K * K;
    static constexpr int CBP = round4(CB), COUTP = round4(COUT);
    static constexpr int SKT = MODE == 1 ? 8 : 1;                 // skip taps (k2 s2 for down)
    static constexpr int w1 = 0;                                  // [CIN][CBP]
    static constexpr int w2 = w1 + CIN * CBP;                     // [CB][K3][CBP]
    static constexpr int w3 = w2 + CB * K3 * CBP;                 // [CB][COUTP]
    static constexpr int ws = w3 + CB * COUTP;                    // [CIN][SKT][COUTP]
    static constexpr int tiles = ws + (SKIP ? CIN * SKT * COUTP : 0);
    // input-tile extent of conv2 for an output tile of t
    __host__ __device__ static int in_ext(int t) { return (t - 1) * Geo<MODE>::ST + K; }
    __host__ __device__ static int lo_ext(int t) { return t / 2 + 2; }
    static size_t floats(int th, int tw, int tz) {
        size_t n = tiles + (size_t)CB * in_ext(th) * in_ext(tw) * in_ext(tz);
        if (MODE == 2) {
            n += (size_t)(CB + COUT) * lo_ext(th) * lo_ext(tw) * lo_ext(tz);   // low-res t1 + low-res skip
            n += 3 * (size_t)(in_ext(th) + in_ext(tw) + in_ext(tz));           // tap tables (i0, i1, lambda)
        }
        return n;
    }
};

template <int CIN, int CB, int COUT, int MODE, bool SKIP, int VX>
__global__ void __launch_bounds__(512)
preact_fused_kernel(PreactParams p) {
    using SM = PreactSmem<CIN, CB, COUT, MODE, SKIP, VX>;
    constexpr int K = SM::K, K3 = SM::K3, ST = Geo<MODE>::ST, CBP = SM::CBP, COUTP = SM::COUTP, SKT = SM::SKT;
    VQ3D_DYN_SMEM(float, smem);
    float *s_w1 = smem + SM::w1, *s_w2 = smem + SM::w2, *s_w3 = smem + SM::w3, *s_ws = smem + SM::ws;
    float *s_t1 = smem + SM::tiles;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int th = p.th, tw = p.tw, tz = p.tz;
    const int IH = SM::in_ext(th), IW = SM::in_ext(tw), IZ = SM::in_ext(tz);
    const int64_t S = (int64_t)p.H * p.W * p.Z, So = (int64_t)p.Ho * p.Wo * p.Zo;

    // tile coordinates
    int bid = blockIdx.x;
    const int tzi = bid % p.ntz; bid /= p.ntz;
    const int twi = bid % p.ntw; bid /= p.ntw;
    const int thi = bid % p.nth; bid /= p.nth;
    const int b = bid;
    const int oh0 = thi * th, ow0 = twi * tw, oz0 = tzi * tz;
    const float *xb = p.x + (size_t)b * CIN * S;

    // ---- weights -> shared (transposed so that the output channel is innermost) -------------
    for (int i = tid; i < CIN * CBP; i += nthr) {
        const int cb = i % CBP, ci = i / CBP;
        s_w1[i] = cb < CB ? p.w1[cb * CIN + ci] : 0.0f;
    }
    for (int i = tid; i < CB * K3 * CBP; i += nthr) {
        const int co = i % CBP, t = (i / CBP) % K3, ci = i / (CBP * K3);
        s_w2[i] = co < CB ? p.w2[((size_t)co * CB + ci) * K3 + t] : 0.0f;
    }
    for (int i = tid; i < CB * COUTP; i += nthr) {
        const int co = i % COUTP, cb = i / COUTP;
        s_w3[i] = co < COUT ? p.w3[co * CB + cb] : 0.0f;
    }
    if (SKIP) {
        for (int i = tid; i < CIN * SKT * COUTP; i += nthr) {
            const int co = i % COUTP, t = (i / COUTP) % SKT, ci = i / (COUTP * SKT);
            s_ws[i] = co < COUT ? p.ws[((size_t)co * CIN + ci) * SKT + t] : 0.0f;
        }
    }
    const float b1a = ld_scalar(p.b1a, 0.f), b1b = ld_scalar(p.b1b, 0.f), b2a = ld_scalar(p.b2a, 0.f), b2b = ld_scalar(p.b2b, 0.f);
    __syncthreads();

    // ---- stage A ------------------------------------------------------------------------
    if (MODE != 2) {
        const int ih0 = oh0 * ST - 1, iw0 = ow0 * ST - 1, iz0 = oz0 * ST - 1;     // pad = 1
        const int n_in = IH * IW * IZ;
        for (int i = tid; i < n_in; i += nthr) {
            const int lz = i % IZ, lw = (i / IZ) % IW, lh = i / (IZ * IW);
            const int gh = pmod(ih0 + lh, p.H), gw = pmod(iw0 + lw, p.W), gz = pmod(iz0 + lz, p.Z);
            const float *px = xb + ((size_t)gh * p.W + gw) * p.Z + gz;
            float acc[CB];
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[c] = 0.0f;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const float v = elu1(__ldg(px + (size_t)ci * S) + b1a) + b1b;
#pragma unroll
                for (int c = 0; c < CB; ++c) acc[c] = __fmaf_rn(s_w1[ci * CBP + c], v, acc[c]);
            }
#pragma unroll
            for (int c = 0; c < CB; ++c) s_t1[(size_t)c * n_in + i] = elu1(acc[c] + b2a) + b2b;
        }
    } else {
        // low-res window, then trilinear upsample into the hi-res halo tile
        const int LH = SM::lo_ext(th), LW = SM::lo_ext(tw), LZ = SM::lo_ext(tz);
        const int n_lo = LH * LW * LZ, n_in = IH * IW * IZ;
        float *s_lo = s_t1 + (size_t)CB * n_in;                  // [CB][n_lo]
        float *s_sk = s_lo + (size_t)CB * n_lo;                  // [COUT][n_lo]
        float *s_tab = s_sk + (size_t)COUT * n_lo;               // [3][IH+IW+IZ]: i0 | i1 | lambda
        const int lh0 = oh0 / 2 - 1, lw0 = ow0 / 2 - 1, lz0 = oz0 / 2 - 1;
        const float b1c = ld_scalar(p.b1c, 0.f);
        for (int i = tid; i < n_lo; i += nthr) {
            const int lz = i % LZ, lw = (i / LZ) % LW, lh = i / (LZ * LW);
            const int gh = pmod(lh0 + lh, p.H), gw = pmod(lw0 + lw, p.W), gz = pmod(lz0 + lz, p.Z);
            const float *px = xb + ((size_t)gh * p.W + gw) * p.Z + gz;
            float acc[CB], sk[COUT];
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[c] = 0.0f;
#pragma unroll
            for (int c = 0; c < COUT; ++c) sk[c] = 0.0f;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const float xv = __ldg(px + (size_t)ci * S);
                const float v = elu1(xv + b1a) + b1b;
#pragma unroll
                for (int c = 0; c < CB; ++c) acc[c] = __fmaf_rn(s_w1[ci * CBP + c], v, acc[c]);
                const float xs = xv + b1c;
#pragma unroll
                for (int c = 0; c < COUT; ++c) sk[c] = __fmaf_rn(s_ws[ci * COUTP + c], xs, sk[c]);
            }
#pragma unroll
            for (int c = 0; c < CB; ++c) s_lo[(size_t)c * n_lo + i] = elu1(acc[c] + b2a) + b2b;
#pragma unroll
            for (int c = 0; c < COUT; ++c) s_sk[(size_t)c * n_lo + i] = sk[c];
        }
        // tap tables for the hi-res halo positions of each axis (upsample clamps at the edges,
        // the circular wrap of the conv padding applies to the upsampled tensor)
        const int ntab = IH + IW + IZ;
        for (int i = tid; i < ntab; i += nthr) {
            int q, o0, n_hi, n_lo_ax, l0, ext;
            if (i < IH) { q = i; o0 = oh0; n_hi = p.Ho; n_lo_ax = p.H; l0 = lh0; ext = LH; }
            else if (i < IH + IW) { q = i - IH; o0 = ow0; n_hi = p.Wo; n_lo_ax = p.W; l0 = lw0; ext = LW; }
            else { q = i - IH - IW; o0 = oz0; n_hi = p.Zo; n_lo_ax = p.Z; l0 = lz0; ext = LZ; }
            const int pos = pmod(o0 - 1 + q, n_hi);
            int i0, i1; float l1;
            up_taps_dev(pos, n_lo_ax, i0, i1, l1);
            s_tab[i] = (float)min(pmod(i0 - l0, n_lo_ax), ext - 1);
            s_tab[ntab + i] = (float)min(pmod(i1 - l0, n_lo_ax), ext - 1);
            s_tab[2 * ntab + i] = l1;
        }
        __syncthreads();
        for (int i = tid; i < n_in; i += nthr) {
            const int qz = i % IZ, qw = (i / IZ) % IW, qh = i / (IZ * IW);
            const int h0 = (int)s_tab[qh], h1 = (int)s_tab[ntab + qh]; const float lh = s_tab[2 * ntab + qh];
            const int w0 = (int)s_tab[IH + qw], w1 = (int)s_tab[ntab + IH + qw]; const float lw = s_tab[2 * ntab + IH + qw];
            const int z0 = (int)s_tab[IH + IW + qz], z1 = (int)s_tab[ntab + IH + IW + qz]; const float lz = s_tab[2 * ntab + IH + IW + qz];
            const int o00 = (h0 * LW + w0) * LZ, o01 = (h0 * LW + w1) * LZ, o10 = (h1 * LW + w0) * LZ, o11 = (h1 * LW + w1) * LZ;
#pragma unroll
            for (int c = 0; c < CB; ++c) {
                const float *s = s_lo + (size_t)c * n_lo;
                const float a00 = s[o00 + z0] * (1.f - lz) + s[o00 + z1] * lz, a01 = s[o01 + z0] * (1.f - lz) + s[o01 + z1] * lz;
                const float a10 = s[o10 + z0] * (1.f - lz) + s[o10 + z1] * lz, a11 = s[o11 + z0] * (1.f - lz) + s[o11 + z1] * lz;
                const float c0 = a00 * (1.f - lw) + a01 * lw, c1 = a10 * (1.f - lw) + a11 * lw;
                s_t1[(size_t)c * n_in + i] = c0 * (1.f - lh) + c1 * lh;
            }
        }
    }
    __syncthreads();

    // ---- stage B + C ----------------------------------------------------------------------
    // thread -> (ly, lxv, lz), lz fastest: a warp walks the contiguous depth axis
    const int nxv = tw / VX;
    const int lz = tid % tz, lxv = (tid / tz) % nxv, ly = tid / (tz * nxv);
    const bool in_tile = ly < th;
    if (in_tile) {
        const int n_in = IH * IW * IZ;
        float acc[VX][CB];
        constexpr int NW = (VX - 1) * ST + K;       // register window along W
        const int bh = ly * ST, bw = lxv * VX * ST, bz = lz * ST;
        if constexpr (VQ3D_FUSED_FFMA2 && CB >= 4) {
            // two branch channels per packed FMA (weight pair x broadcast window value); an odd CB pads the last pair
            // with the (ignored) lane of the padded weight column CB < CBP
            constexpr int CP = (CB + 1) / 2;
            static_assert(2 * CP <= CBP, "padded weight column");
            float2 acc2[VX][CP];
#pragma unroll
            for (int v = 0; v < VX; ++v)
#pragma unroll
                for (int c = 0; c < CP; ++c) acc2[v][c] = make_float2(0.0f, 0.0f);
            for (int ci = 0; ci < CB; ++ci) {
                const float *t1c = s_t1 + (size_t)ci * n_in;
                const float *wc = s_w2 + (size_t)ci * K3 * CBP;
#pragma unroll
                for (int kh = 0; kh < K; ++kh) {
#pragma unroll
                    for (int kz = 0; kz < K; ++kz) {
                        float win[NW];
                        const float *row = t1c + ((size_t)(bh + kh) * IW + bw) * IZ + bz + kz;
#pragma unroll
                        for (int j = 0; j < NW; ++j) win[j] = row[(size_t)j * IZ];
#pragma unroll
                        for (int kw = 0; kw < K; ++kw) {
                            const float2 *wt = reinterpret_cast<const float2 *>(wc + ((kh * K + kw) * K + kz) * CBP);
#pragma unroll
                            for (int c = 0; c < CP; ++c) {
                                const float2 wv = wt[c];
#pragma unroll
                                for (int v = 0; v < VX; ++v) acc2[v][c] = ffma2_bcast(wv, win[v * ST + kw], acc2[v][c]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < VX; ++v)
#pragma unroll
                for (int c = 0; c < CB; ++c) acc[v][c] = (c & 1) ? acc2[v][c >> 1].y : acc2[v][c >> 1].x;
        } else {
#pragma unroll
        for (int v = 0; v < VX; ++v)
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[v][c] = 0.0f;
        for (int ci = 0; ci < CB; ++ci) {
            const float *t1c = s_t1 + (size_t)ci * n_in;
            const float *wc = s_w2 + (size_t)ci * K3 * CBP;
#pragma unroll
            for (int kh = 0; kh < K; ++kh) {
#pragma unroll
                for (int kz = 0; kz < K; ++kz) {
                    float win[NW];
                    const float *row = t1c + ((size_t)(bh + kh) * IW + bw) * IZ + bz + kz;
#pragma unroll
                    for (int j = 0; j < NW; ++j) win[j] = row[(size_t)j * IZ];
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
                        const float *wt = wc + ((kh * K + kw) * K + kz) * CBP;
#pragma unroll
                        for (int c = 0; c < CB; ++c) {
                            const float wv = wt[c];
#pragma unroll
                            for (int v = 0; v < VX; ++v) acc[v][c] = __fmaf_rn(wv, win[v * ST + kw], acc[v][c]);
                        }
                    }
                }
            }
        }
        }
        // stage C
        const float b3a = ld_scalar(p.b3a, 0.f), b3b = ld_scalar(p.b3b, 0.f), b4 = ld_scalar(p.b4, 0.f), sc = ld_scalar(p.scale, 1.f);
        const float b1c = ld_scalar(p.b1c, 0.f), b1d = ld_scalar(p.b1d, 0.f);
        const int oh = oh0 + ly, oz = oz0 + lz;
#pragma unroll
        for (int v = 0; v < VX; ++v) {
            const int ow = ow0 + lxv * VX + v;
            if (oh < p.Ho && ow < p.Wo && oz < p.Zo) {
                float out[COUT];
#pragma unroll
                for (int c = 0; c < COUT; ++c) out[c] = 0.0f;
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) {
                    const float t2 = elu1(acc[v][cb] + b3a) + b3b;
#pragma unroll
                    for (int c = 0; c < COUT; ++c) out[c] = __fmaf_rn(s_w3[cb * COUTP + c], t2, out[c]);
                }
                float res[COUT];
                if (!SKIP) {
                    const float *px = xb + ((size_t)oh * p.W + ow) * p.Z + oz;
#pragma unroll
                    for (int c = 0; c < COUT; ++c) res[c] = __ldg(px + (size_t)(c < CIN ? c : 0) * S);
                } else if (MODE == 0) {
                    const float *px = xb + ((size_t)oh * p.W + ow) * p.Z + oz;
#pragma unroll
                    for (int c = 0; c < COUT; ++c) res[c] = b1d;
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) {
                        const float xs = __ldg(px + (size_t)ci * S) + b1c;
#pragma unroll
                        for (int c = 0; c < COUT; ++c) res[c] = __fmaf_rn(s_ws[ci * COUTP + c], xs, res[c]);
                    }
                } else if (MODE == 1) {
#pragma unroll
                    for (int c = 0; c < COUT; ++c) res[c] = 0.0f;
                    for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int ih = 2 * oh + (t >> 2), iw = 2 * ow + ((t >> 1) & 1), iz = 2 * oz + (t & 1);
                            const float xs = __ldg(xb + (size_t)ci * S + ((size_t)ih * p.W + iw) * p.Z + iz) + b1c;
#pragma unroll
                            for (int c = 0; c < COUT; ++c) res[c] = __fmaf_rn(s_ws[(ci * 8 + t) * COUTP + c], xs, res[c]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < COUT; ++c) res[c] += b1d;
                } else {
                    const int LH = SM::lo_ext(th), LW = SM::lo_ext(tw), LZ = SM::lo_ext(tz);
                    const int n_lo = LH * LW * LZ, ntab = IH + IW + IZ;
                    const float *s_sk = s_t1 + (size_t)CB * n_in + (size_t)CB * n_lo;
                    const float *s_tab = s_sk + (size_t)COUT * n_lo;
                    const int qh = ly + 1, qw = lxv * VX + v + 1, qz = lz + 1;    // halo offset 1
                    const int h0 = (int)s_tab[qh], h1 = (int)s_tab[ntab + qh]; const float lh = s_tab[2 * ntab + qh];
                    const int w0 = (int)s_tab[IH + qw], w1 = (int)s_tab[ntab + IH + qw]; const float lw = s_tab[2 * ntab + IH + qw];
                    const int z0 = (int)s_tab[IH + IW + qz], z1 = (int)s_tab[ntab + IH + IW + qz]; const float lzz = s_tab[2 * ntab + IH + IW + qz];
                    const int o00 = (h0 * LW + w0) * LZ, o01 = (h0 * LW + w1) * LZ, o10 = (h1 * LW + w0) * LZ, o11 = (h1 * LW + w1) * LZ;
#pragma unroll
                    for (int c = 0; c < COUT; ++c) {
                        const float *s = s_sk + (size_t)c * n_lo;
                        const float a00 = s[o00 + z0] * (1.f - lzz) + s[o00 + z1] * lzz, a01 = s[o01 + z0] * (1.f - lzz) + s[o01 + z1] * lzz;
                        const float a10 = s[o10 + z0] * (1.f - lzz) + s[o10 + z1] * lzz, a11 = s[o11 + z0] * (1.f - lzz) + s[o11 + z1] * lzz;
                        const float c0 = a00 * (1.f - lw) + a01 * lw, c1 = a10 * (1.f - lw) + a11 * lw;
                        res[c] = c0 * (1.f - lh) + c1 * lh + b1d;
                    }
                }
                float *py = p.y + (size_t)b * COUT * So + ((size_t)oh * p.Wo + ow) * p.Zo + oz;
#pragma unroll
                for (int c = 0; c < COUT; ++c) py[(size_t)c * So] = __fmaf_rn(out[c], sc, b4) + res[c];
            }
        }
    }
}

template <int CIN, int CB, int COUT, int MODE, bool SKIP, int VX>
static int launch_fused(const vq3d_preact_desc *d, void *stream) {
    using SM = PreactSmem<CIN, CB, COUT, MODE, SKIP, VX>;
    PreactParams p;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Z = d->Z;
    if (MODE == 1) {
        if ((d->H | d->W | d->Z) & 1) return fail(VQ3D_ERR_UNSUPPORTED, "preact_block(down): odd input size");
        p.Ho = d->H / 2; p.Wo = d->W / 2; p.Zo = d->Z / 2;
    } else if (MODE == 2) {
        p.Ho = 2 * d->H; p.Wo = 2 * d->W; p.Zo = 2 * d->Z;
    } else {
        p.Ho = d->H; p.Wo = d->W; p.Zo = d->Z;
    }
    // tile: tz along depth (<= 32), tw multiple of VX, th; even for mode up; shrink to fit smem / 512 threads
    auto even_up = [](int v) { return MODE == 2 ? ((v + 1) & ~1) : v; };
    int tz = even_up(p.Zo < 32 ? p.Zo : 32);
    int tw = (int)ceil_div(p.Wo < 8 ? p.Wo : 8, VX) * VX;
    if (MODE == 2) tw = even_up(tw);
    int th = even_up(p.Ho < 8 ? p.Ho : 8);
    const size_t smem_cap = 200 * 1024;
    auto threads = [&]() { return th * (tw / VX) * tz; };
    while ((SM::floats(th, tw, tz) * 4 > smem_cap || threads() > 512)) {
        if (th > (MODE == 2 ? 2 : 1)) th = even_up((th + 1) / 2);
        else if (tw > VX && (tw / 2) % VX == 0 && !(MODE == 2 && ((tw / 2) & 1))) tw /= 2;
        else if (tz > (MODE == 2 ? 2 : 1)) tz = even_up((tz + 1) / 2);
        else return fail(VQ3D_ERR_UNSUPPORTED, "preact_block: tile does not fit shared memory");
    }
    // small problems: prefer more CTAs over big tiles (148 SMs)
    auto ntiles = [&]() { return (int64_t)p.B * ceil_div(p.Ho, th) * ceil_div(p.Wo, tw) * ceil_div(p.Zo, tz); };
    while (ntiles() < 2 * kNumSMs && th > (MODE == 2 ? 2 : 1) && threads() > 64) th = even_up((th + 1) / 2);
    p.th = th; p.tw = tw; p.tz = tz;
    p.nth = (int)ceil_div(p.Ho, th); p.ntw = (int)ceil_div(p.Wo, tw); p.ntz = (int)ceil_div(p.Zo, tz);
    p.x = d->x; p.w1 = d->w1; p.w2 = d->w2; p.w3 = d->w3; p.ws = d->wskip;
    p.b1a = d->b1a; p.b1b = d->b1b; p.b2a = d->b2a; p.b2b = d->b2b; p.b3a = d->b3a; p.b3b = d->b3b;
    p.b4 = d->b4; p.scale = d->scale; p.b1c = d->b1c; p.b1d = d->b1d; p.y = d->y;
    int nthr = threads();
    nthr = (nthr + 31) & ~31;
    const int64_t grid = ntiles();
    if (grid > 0x7fffffff) return fail(VQ3D_ERR_INVALID, "preact_block: grid too large");
    return launch("preact_fused", preact_fused_kernel<CIN, CB, COUT, MODE, SKIP, VX>, dim3((unsigned)grid), dim3((unsigned)nthr),
                  SM::floats(th, tw, tz) * 4, stream, p);
}

struct FusedEntry {
    int cin, cb, cout, mode, skip;
    int (*fn)(const vq3d_preact_desc *, void *);
};

#define VQ3D_FUSED(CIN, CB, COUT, MODE, SKIP, VX) {CIN, CB, COUT, MODE, SKIP, launch_fused<CIN, CB, COUT, MODE, (SKIP) != 0, VX>}
static const FusedEntry kFused[] = {
    // same (no skip): every stack / post-scale block of the Full and downscaled models with branch width <= 16
    VQ3D_FUSED(2, 1, 2, 0, 0, 4), VQ3D_FUSED(4, 2, 4, 0, 0, 4), VQ3D_FUSED(8, 4, 8, 0, 0, 4), VQ3D_FUSED(16, 8, 16, 0, 0, 4),
    VQ3D_FUSED(18, 9, 18, 0, 0, 4), VQ3D_FUSED(32, 16, 32, 0, 0, 2), VQ3D_FUSED(6, 3, 6, 0, 0, 4),
    // same with 1x1 skip: pre_q of level 0 (18 -> 2), test shapes
    VQ3D_FUSED(18, 9, 2, 0, 1, 4), VQ3D_FUSED(4, 2, 1, 0, 1, 4),
    // down (k4 s2 + k2 s2 skip)
    VQ3D_FUSED(4, 4, 8, 1, 1, 2), VQ3D_FUSED(8, 8, 16, 1, 1, 2), VQ3D_FUSED(16, 16, 32, 1, 1, 2),
    // up (trilinear x2 + k3, 1x1 skip)
    VQ3D_FUSED(8, 4, 4, 2, 1, 4), VQ3D_FUSED(4, 2, 2, 2, 1, 4), VQ3D_FUSED(16, 8, 8, 2, 1, 4), VQ3D_FUSED(18, 9, 8, 2, 1, 4),
    VQ3D_FUSED(32, 16, 16, 2, 1, 2),
};

static const FusedEntry *find_fused(const vq3d_preact_desc *d) {
    for (const FusedEntry &e : kFused)
        if (e.cin == d->Cin && e.cb == d->Cb && e.cout == d->Cout && e.mode == d->mode && e.skip == (d->wskip != nullptr)) return &e;
    return nullptr;
}

static int validate(const vq3d_preact_desc *d) {
    if (!d) return fail(VQ3D_ERR_INVALID, "preact_block: null descriptor");
    if (!d->x || !d->w1 || !d->w2 || !d->w3) return fail(VQ3D_ERR_INVALID, "preact_block: null x/weights");
    if (d->out_w ? !d->out_y : !d->y) return fail(VQ3D_ERR_INVALID, "preact_block: null output");
    if (d->B < 1 || d->H < 1 || d->W < 1 || d->Z < 1 || d->Cin < 1 || d->Cb < 1 || d->Cout < 1 || d->mode < 0 || d->mode > 2)
        return fail(VQ3D_ERR_INVALID, "preact_block: bad sizes");
    if (d->mode != 0 && !d->wskip) return fail(VQ3D_ERR_INVALID, "preact_block: mode down/up needs a skip conv");
    if (d->mode == 0 && !d->wskip && d->Cin != d->Cout) return fail(VQ3D_ERR_INVALID, "preact_block: Cin != Cout needs a skip conv");
    return VQ3D_OK;
}

}  // namespace vq3d

using namespace vq3d;

extern "C" int vq3d_preact_block(const vq3d_preact_desc *d, void *stream) {
    int rc = validate(d);
    if (rc) return rc;
    bool handled = false;
    rc = preact_row_dispatch(d, stream, &handled);
    if (rc || handled) return rc;
    if (d->out_w) return fail(VQ3D_ERR_UNSUPPORTED, "preact_block: no kernel fuses the trailing 1x1 convolution for this shape");
    if (d->pre_w) return fail(VQ3D_ERR_UNSUPPORTED, "preact_block: no kernel fuses the leading 1x1 convolution for this shape");
    const FusedEntry *e = find_fused(d);
    if (!e) return fail(VQ3D_ERR_UNSUPPORTED, "preact_block: no fused instantiation for Cin=%d Cb=%d Cout=%d mode=%d", d->Cin, d->Cb, d->Cout, d->mode);
    return e->fn(d, stream);
}

extern "C" int vq3d_preact_stack(const vq3d_preact_desc *blocks, int n, float *tmp, void *stream) {
    if (!blocks || n < 1 || !tmp) return fail(VQ3D_ERR_INVALID, "preact_stack: bad arguments");
    for (int i = 0; i < n; ++i) {
        int rc = validate(&blocks[i]);
        if (rc) return rc;
        if (blocks[i].mode != 0 || blocks[i].wskip || blocks[i].Cin != blocks[0].Cin || blocks[i].Cb != blocks[0].Cb)
            return fail(VQ3D_ERR_INVALID, "preact_stack: blocks must be equal-shape 'same' blocks without skip");
    }
    for (int i = 0; i + 1 < n; ++i)
        if (blocks[i].out_w) return fail(VQ3D_ERR_INVALID, "preact_stack: only the last block may carry a trailing 1x1 convolution");
    const FusedEntry *e = find_fused(&blocks[0]);
    // ping-pong between tmp and y so that block n-1 lands in blocks[n-1].y (or .out_y through the fused 1x1)
    const float *src = blocks[0].x;
    float *out = blocks[n - 1].y;
    for (int i = 0; i < n; ++i) {
        vq3d_preact_desc d = blocks[i];
        d.x = src;
        d.y = ((n - 1 - i) % 2 == 0) ? out : tmp;
        bool handled = false;
        int rc = preact_row_dispatch(&d, stream, &handled);
        if (rc) return rc;
        if (!handled) {
            if (d.out_w) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack: no kernel fuses the trailing 1x1 convolution for this shape");
            if (!e) return fail(VQ3D_ERR_UNSUPPORTED, "preact_stack: no fused instantiation for C=%d", blocks[0].Cin);
            rc = e->fn(&d, stream);
            if (rc) return rc;
        }
        src = d.y;
    }
    return VQ3D_OK;
}
