// Exact fp32 kernels for the layers of the reference's SHIPPED YAML that the tcgen05 conv stacks do not cover:
//   * general 2-d convolution on NCHW tensors — rectangular kernels, any stride / padding, no bias — in its three roles
//     (forward, input gradient, weight gradient).  nn.ConvTranspose2d is the same three kernels with the roles of input
//     and output swapped.  Used by SoundEncoder_v2 / SoundDecoder_v2 (encoder.py:661-721, observation_model.py:420-472:
//     kernels 3x9, 4x8, 3x4, 4x4, 7x7, Conv1d k1) and by the BatchNorm variants of ImageEncoder / ImageDecoder
//     (encoder.py:324-337, observation_model.py:75-86: Conv / ConvT without bias followed by BatchNorm2d + ReLU);
//   * nn.BatchNorm2d and nn.InstanceNorm2d / 1d (affine, running statistics) forward / backward, optional fused ReLU;
//   * nn.GLU(dim=1) forward / backward.
// The convolutions are tiled implicit GEMMs on the CUDA cores (64 x 64 x 16 tiles, 4 x 4 outputs per thread): exact fp32,
// the first correct version of these layers; they are not on the benchmarked path (BASELINE configs carry no sound / BatchNorm).
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int BK = 16, NT = 256;      // tiles: (BM, BN) = (64, 64), or (256, 16) when there are at most 16 columns (BM * BN = 16 * NT)

struct GC {
    int N, Cin, H, W, Cout, KH, KW, SH, SW, PH, PW, Ho, Wo;
    const float* x;      // [N,Cin,H,W]
    const float* w;      // [Cout,Cin,KH,KW]
    const float* dy;     // [N,Cout,Ho,Wo]
    float* out;
    long long M, K, kchunk;
    int Ncols;
};

// MODE 0: y = conv(x, w)          M = N*Ho*Wo, cols = Cout,          K = Cin*KH*KW
// MODE 1: dx = conv^T(dy, w)      cols = Cin; one grid.z slice per stride-parity class (rh, rw) of the input positions: only the
//         taps kh = rh + i*SH, kw = rw + j*SW reach a position with (h + PH) % SH == rh, so a class is a dense GEMM with
//         M = N*Hc*Wc rows and K = Cout*nkh*nkw (no wasted taps, the same weight tile for every row of a CTA)
// MODE 2: dw += dy^T im2col(x)    M = Cout,    cols = Cin*KH*KW,     K = N*Ho*Wo (split over grid.z, atomics)
template <int MODE, int BM, int BN>
__global__ void __launch_bounds__(NT) gconv_kernel(const GC g) {
    constexpr int RA = BM / 16, RB = BN / 16;          // A rows / B columns each thread loads per K step
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, kk = tid & 15, r0 = tid >> 4;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    long long kbeg = 0, kend = g.K;
    if (MODE == 2) {
        kbeg = (long long)blockIdx.z * g.kchunk;
        kend = kbeg + g.kchunk < g.K ? kbeg + g.kchunk : g.K;
    }
    const int KHW = g.KH * g.KW, HoWo = g.Ho * g.Wo, HW = g.H * g.W;
    // MODE 1: this CTA's parity class
    int rh = 0, rw = 0, h0 = 0, w0 = 0, Hc = 0, Wc = 0, nkh = 0, nkw = 0;
    long long Mrows = g.M;
    if (MODE == 1) {
        rh = blockIdx.z / g.SW; rw = blockIdx.z - rh * g.SW;
        h0 = ((rh - g.PH) % g.SH + g.SH) % g.SH; w0 = ((rw - g.PW) % g.SW + g.SW) % g.SW;
        Hc = h0 < g.H ? (g.H - h0 + g.SH - 1) / g.SH : 0; Wc = w0 < g.W ? (g.W - w0 + g.SW - 1) / g.SW : 0;
        nkh = rh < g.KH ? (g.KH - rh + g.SH - 1) / g.SH : 0; nkw = rw < g.KW ? (g.KW - rw + g.SW - 1) / g.SW : 0;
        Mrows = (long long)g.N * Hc * Wc;
        if (m0 >= Mrows) return;
        kend = (long long)g.Cout * nkh * nkw;
    }

    // per-thread decode of its 4 A rows and 4 B columns (fixed over the K loop)
    int a_n[RA], a_a[RA], a_b[RA];
    bool a_ok[RA];
    int b_c[RB], b_kh[RB], b_kw[RB];
    bool b_ok[RB];
#pragma unroll
    for (int j = 0; j < RA; ++j) {
        const long long m = m0 + r0 + 16 * j;
        a_ok[j] = m < Mrows;
        a_n[j] = a_a[j] = a_b[j] = 0;
        if (a_ok[j]) {
            if (MODE == 0) {
                const int n = (int)(m / HoWo), r = (int)(m - (long long)n * HoWo), ho = r / g.Wo, wo = r - ho * g.Wo;
                a_n[j] = n; a_a[j] = ho * g.SH - g.PH; a_b[j] = wo * g.SW - g.PW;
            } else if (MODE == 1) {
                const int HWc = Hc * Wc, n = (int)(m / HWc), r = (int)(m - (long long)n * HWc), hi = r / Wc, wi = r - hi * Wc;
                a_n[j] = n; a_a[j] = (h0 + hi * g.SH + g.PH - rh) / g.SH; a_b[j] = (w0 + wi * g.SW + g.PW - rw) / g.SW;   // ho, wo of tap (rh, rw)
            } else {
                a_n[j] = (int)m;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < RB; ++j) {
        const int c = n0 + r0 + 16 * j;
        b_ok[j] = c < g.Ncols;
        b_c[j] = c; b_kh[j] = b_kw[j] = 0;
        if (MODE == 2 && b_ok[j]) {
            const int ci = c / KHW, r = c - ci * KHW;
            b_c[j] = ci; b_kh[j] = r / g.KW; b_kw[j] = r - b_kh[j] * g.KW;
        }
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int ty = tid / (BN / 4), tx = tid % (BN / 4);

    for (long long k0 = kbeg; k0 < kend; k0 += BK) {
        const long long k = k0 + kk;
        const bool kok = k < kend;
        float av[RA], bv[RB];
#pragma unroll
        for (int j = 0; j < RA; ++j) av[j] = 0.f;
#pragma unroll
        for (int j = 0; j < RB; ++j) bv[j] = 0.f;
        if (kok) {
            if (MODE == 0) {
                const int ci = (int)(k / KHW), r = (int)(k - (long long)ci * KHW), kh = r / g.KW, kw = r - kh * g.KW;
#pragma unroll
                for (int j = 0; j < RA; ++j) {
                    const int h = a_a[j] + kh, w = a_b[j] + kw;
                    if (a_ok[j] && h >= 0 && h < g.H && w >= 0 && w < g.W)
                        av[j] = __ldg(g.x + ((long long)a_n[j] * g.Cin + ci) * HW + (long long)h * g.W + w);
                }
#pragma unroll
                for (int j = 0; j < RB; ++j)
                    if (b_ok[j]) bv[j] = __ldg(g.w + (long long)b_c[j] * g.K + k);
            } else if (MODE == 1) {
                const int nk = nkh * nkw, co = (int)(k / nk), r = (int)(k - (long long)co * nk), ti = r / nkw, tj = r - ti * nkw;
                const int wofs = (rh + ti * g.SH) * g.KW + rw + tj * g.SW;
#pragma unroll
                for (int j = 0; j < RA; ++j) {
                    const int ho = a_a[j] - ti, wo = a_b[j] - tj;
                    if (a_ok[j] && ho >= 0 && ho < g.Ho && wo >= 0 && wo < g.Wo)
                        av[j] = __ldg(g.dy + ((long long)a_n[j] * g.Cout + co) * HoWo + (long long)ho * g.Wo + wo);
                }
#pragma unroll
                for (int j = 0; j < RB; ++j)
                    if (b_ok[j]) bv[j] = __ldg(g.w + ((long long)co * g.Cin + b_c[j]) * KHW + wofs);
            } else {
                const int n = (int)(k / HoWo), r = (int)(k - (long long)n * HoWo), ho = r / g.Wo, wo = r - ho * g.Wo;
#pragma unroll
                for (int j = 0; j < RA; ++j)
                    if (a_ok[j]) av[j] = __ldg(g.dy + ((long long)n * g.Cout + a_n[j]) * HoWo + r);
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const int h = ho * g.SH - g.PH + b_kh[j], w = wo * g.SW - g.PW + b_kw[j];
                    if (b_ok[j] && h >= 0 && h < g.H && w >= 0 && w < g.W)
                        bv[j] = __ldg(g.x + ((long long)n * g.Cin + b_c[j]) * HW + (long long)h * g.W + w);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < RA; ++j) As[kk][r0 + 16 * j] = av[j];
#pragma unroll
        for (int j = 0; j < RB; ++j) Bs[kk][r0 + 16 * j] = bv[j];
        __syncthreads();
#pragma unroll
        for (int p = 0; p < BK; ++p) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[p][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[p][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= Mrows) continue;
        long long base = 0, cstride = 1;
        if (MODE == 0) {
            const int n = (int)(m / HoWo), r = (int)(m - (long long)n * HoWo);
            base = (long long)n * g.Cout * HoWo + r; cstride = HoWo;
        } else if (MODE == 1) {
            const int HWc = Hc * Wc, n = (int)(m / HWc), r = (int)(m - (long long)n * HWc), hi = r / Wc, wi = r - hi * Wc;
            base = (long long)n * g.Cin * HW + (long long)(h0 + hi * g.SH) * g.W + (w0 + wi * g.SW); cstride = HW;
        } else {
            base = m * g.Ncols;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c >= g.Ncols) continue;
            if (MODE == 2) atomicAdd(g.out + base + c, acc[i][j]);
            else g.out[base + (long long)c * cstride] = acc[i][j];
        }
    }
}

// ---- convolutions to very few output channels (SoundDecoder_v2.out: 64 -> 1, 7x7): the GEMM tile would be 15/16 padding ------------
// forward: one thread per output pixel, all (<= 4) output channels, weights [k][co] in shared memory
template <int CO>
__global__ void gconv_fwd_small_kernel(const GC g) {
    extern __shared__ float wsm[];
    const int KHW = g.KH * g.KW, K = g.Cin * KHW, HoWo = g.Ho * g.Wo, HW = g.H * g.W;
    for (int i = threadIdx.x; i < K * CO; i += blockDim.x) {
        const int k = i / CO, co = i - k * CO;
        wsm[i] = co < g.Cout ? g.w[(long long)co * K + k] : 0.f;
    }
    __syncthreads();
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < g.M; m += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(m / HoWo), r = (int)(m - (long long)n * HoWo), ho = r / g.Wo, wo = r - ho * g.Wo;
        const int hb = ho * g.SH - g.PH, wb = wo * g.SW - g.PW;
        float acc[CO];
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[c] = 0.f;
        for (int ci = 0; ci < g.Cin; ++ci) {
            const float* xp = g.x + ((long long)n * g.Cin + ci) * HW;
            const float* wp = wsm + (long long)ci * KHW * CO;
            for (int kh = 0; kh < g.KH; ++kh) {
                const int h = hb + kh;
                if (h < 0 || h >= g.H) continue;
                for (int kw = 0; kw < g.KW; ++kw) {
                    const int w = wb + kw;
                    if (w < 0 || w >= g.W) continue;
                    const float v = __ldg(xp + (long long)h * g.W + w);
#pragma unroll
                    for (int c = 0; c < CO; ++c) acc[c] = fmaf(v, wp[(kh * g.KW + kw) * CO + c], acc[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CO; ++c)
            if (c < g.Cout) g.out[((long long)n * g.Cout + c) * HoWo + r] = acc[c];
    }
}

// weight gradient: CTA = (input channel ci, output channel co, slice of the images); a thread walks output pixels with one register
// accumulator per tap (kernel size is a template parameter so that the taps are unrolled with constant indices), then the block
// reduces the taps and adds them to dw[co][ci][*]
template <int KH, int KW>
__global__ void __launch_bounds__(256) gconv_wgrad_small_kernel(const GC g, int n_slices) {
    constexpr int KHW = KH * KW;
    __shared__ float red[8][KHW];
    const int HoWo = g.Ho * g.Wo, HW = g.H * g.W;
    const int ci = blockIdx.x, co = blockIdx.y;
    float acc[KHW];
#pragma unroll
    for (int t = 0; t < KHW; ++t) acc[t] = 0.f;
    for (int n = blockIdx.z; n < g.N; n += n_slices) {
        const float* xp = g.x + ((long long)n * g.Cin + ci) * HW;
        const float* dp = g.dy + ((long long)n * g.Cout + co) * HoWo;
        for (int r = threadIdx.x; r < HoWo; r += blockDim.x) {
            const int ho = r / g.Wo, wo = r - ho * g.Wo, hb = ho * g.SH - g.PH, wb = wo * g.SW - g.PW;
            const float d = __ldg(dp + r);
#pragma unroll
            for (int kh = 0; kh < KH; ++kh) {
                const int h = hb + kh;
                if (h < 0 || h >= g.H) continue;
                const float* row = xp + (long long)h * g.W;
#pragma unroll
                for (int kw = 0; kw < KW; ++kw) {
                    const int w = wb + kw;
                    if (w >= 0 && w < g.W) acc[kh * KW + kw] = fmaf(d, __ldg(row + w), acc[kh * KW + kw]);
                }
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < KHW; ++t) {
        const float v = warp_sum(acc[t]);
        if (lane == 0) red[warp][t] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < KHW; t += blockDim.x) {
        float v = 0.f;
        for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) v += red[wq][t];
        atomicAdd(g.out + ((long long)co * g.Cin + ci) * KHW + t, v);
    }
}

int check_geom(const mrssm_gconv_args* a) {
    MRSSM_CHECK(a && a->N > 0 && a->Cin > 0 && a->Cout > 0 && a->H > 0 && a->W > 0 && a->KH > 0 && a->KW > 0 && a->SH > 0 && a->SW > 0 &&
                    a->PH >= 0 && a->PW >= 0, "gconv: bad geometry");
    MRSSM_CHECK(a->Ho == (a->H + 2 * a->PH - a->KH) / a->SH + 1 && a->Wo == (a->W + 2 * a->PW - a->KW) / a->SW + 1 && a->Ho > 0 && a->Wo > 0,
                "gconv: output %dx%d does not match input %dx%d, kernel %dx%d, stride %dx%d, padding %dx%d", a->Ho, a->Wo, a->H, a->W,
                a->KH, a->KW, a->SH, a->SW, a->PH, a->PW);
    return 0;
}

GC make(const mrssm_gconv_args* a) {
    GC g;
    g.N = a->N; g.Cin = a->Cin; g.H = a->H; g.W = a->W; g.Cout = a->Cout; g.KH = a->KH; g.KW = a->KW; g.SH = a->SH; g.SW = a->SW;
    g.PH = a->PH; g.PW = a->PW; g.Ho = a->Ho; g.Wo = a->Wo;
    g.x = a->x; g.w = a->w; g.dy = a->y; g.out = nullptr; g.M = 0; g.K = 0; g.kchunk = 0; g.Ncols = 0;
    return g;
}

// ---- normalisation ------------------------------------------------------------------------------------------------------
// A "group" is what one mean / variance is taken over: BatchNorm: channel c over (n, i); InstanceNorm: plane (n, c) over i.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    return t;
}

__global__ void norm_stats_kernel(const float* __restrict__ x, int N, int C, int HW, int instance, float* __restrict__ mean,
                                  float* __restrict__ var) {
    __shared__ double sh[32];
    const int gidx = blockIdx.x;
    double s = 0.0, q = 0.0;
    if (instance) {
        const float* p = x + (long long)gidx * HW;
        for (int i = threadIdx.x; i < HW; i += blockDim.x) { const double v = p[i]; s += v; q += v * v; }
    } else {
        if (HW >= (int)blockDim.x) {
            for (int n = 0; n < N; ++n) {
                const float* p = x + ((long long)n * C + gidx) * HW;
                for (int i = threadIdx.x; i < HW; i += blockDim.x) { const double v = p[i]; s += v; q += v * v; }
            }
        } else {
            const long long total = (long long)N * HW;
            for (long long e = threadIdx.x; e < total; e += blockDim.x) {
                const long long n = e / HW, i = e - n * HW;
                const double v = x[(n * C + gidx) * HW + i];
                s += v; q += v * v;
            }
        }
    }
    s = block_sum(s, sh);
    q = block_sum(q, sh);
    if (threadIdx.x == 0) {
        const double cnt = instance ? (double)HW : (double)N * HW, m = s / cnt;
        mean[gidx] = (float)m;
        var[gidx] = (float)fmax(q / cnt - m * m, 0.0);
    }
}

// running <- (1 - momentum) running + momentum * (batch mean, unbiased batch variance); InstanceNorm: averaged over the N planes
__global__ void norm_running_kernel(const float* __restrict__ mean, const float* __restrict__ var, int N, int C, int HW, int instance,
                                    float momentum, float* __restrict__ rmean, float* __restrict__ rvar) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float m, v;
    if (instance) {
        double sm = 0.0, sv = 0.0;
        for (int n = 0; n < N; ++n) { sm += mean[n * C + c]; sv += var[n * C + c]; }
        m = (float)(sm / N);
        v = (float)(sv / N) * ((float)HW / (float)(HW - 1));
    } else {
        const float cnt = (float)N * (float)HW;
        m = mean[c];
        v = var[c] * (cnt / (cnt - 1.f));
    }
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * m;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * v;
}

// y = [relu](gamma (x - mean_g) / sqrt(var_g + eps) + beta);  per_plane: statistics indexed by (n, c), else by c
__global__ void norm_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, long long total, int C, int HW,
                                  int per_plane, int relu, float eps, float* __restrict__ y) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long pl = e / HW;
        const int c = (int)(pl % C);
        const long long gi = per_plane ? pl : c;
        float v = (x[e] - mean[gi]) * rsqrtf(var[gi] + eps) * gamma[c] + beta[c];
        if (relu && v < 0.f) v = 0.f;
        y[e] = v;
    }
}

// the same with one CTA per (n, c) plane (HW >= 64): the per-plane scalars are read once, no index arithmetic per element
__global__ void norm_apply_plane_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, int C, int HW, int per_plane,
                                        int relu, float eps, float* __restrict__ y) {
    const long long pl = blockIdx.x;
    const int c = (int)(pl % C);
    const long long gi = per_plane ? pl : c;
    const float mu = mean[gi], sc = rsqrtf(var[gi] + eps) * gamma[c], bt = beta[c];
    const float* xp = x + pl * HW;
    float* yp = y + pl * HW;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        float v = (xp[i] - mu) * sc + bt;
        if (relu && v < 0.f) v = 0.f;
        yp[i] = v;
    }
}
__global__ void norm_bwd_apply_plane_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ x,
                                            const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ gamma,
                                            const float* __restrict__ sum_g, const float* __restrict__ sum_gx, int N, int C, int HW, int instance,
                                            int per_plane_stats, int batch_stats, int relu, float eps, float* __restrict__ dx) {
    const long long pl = blockIdx.x;
    const int c = (int)(pl % C);
    const long long gi = instance ? pl : c, si = (instance && !per_plane_stats) ? c : gi;
    const float inv_cnt = instance ? 1.f / (float)HW : 1.f / ((float)N * (float)HW);
    const float mu = mean[si], rs = rsqrtf(var[si] + eps), gm = gamma[c] * rs;
    const float mg = batch_stats ? sum_g[gi] * inv_cnt : 0.f, mgx = batch_stats ? sum_gx[gi] * inv_cnt : 0.f;
    const long long o = pl * HW;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        float gg = g[o + i];
        if (relu && !(y[o + i] > 0.f)) gg = 0.f;
        dx[o + i] = gm * (gg - mg - (x[o + i] - mu) * rs * mgx);
    }
}

// per group: sum of g^ = g * [y > 0 if relu] and of g^ * x^;  dgamma[c] += sum g^ x^, dbeta[c] += sum g^
__global__ void norm_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ x,
                                       const float* __restrict__ mean, const float* __restrict__ var, int N, int C, int HW, int instance,
                                       int per_plane_stats, int relu, float eps, float* __restrict__ sum_g, float* __restrict__ sum_gx,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ double sh[32];
    const int gidx = blockIdx.x;
    double s = 0.0, q = 0.0;
    if (instance) {
        const int c = gidx % C;
        const long long si = per_plane_stats ? gidx : c;
        const float mu = mean[si], rs = rsqrtf(var[si] + eps);
        const long long o = (long long)gidx * HW;
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {
            float gg = g[o + i];
            if (relu && !(y[o + i] > 0.f)) gg = 0.f;
            s += gg; q += (double)gg * ((x[o + i] - mu) * rs);
        }
    } else {
        const float mu = mean[gidx], rs = rsqrtf(var[gidx] + eps);
        if (HW >= (int)blockDim.x) {
            for (int n = 0; n < N; ++n) {
                const long long o0 = ((long long)n * C + gidx) * HW;
                for (int i = threadIdx.x; i < HW; i += blockDim.x) {
                    float gg = g[o0 + i];
                    if (relu && !(y[o0 + i] > 0.f)) gg = 0.f;
                    s += gg; q += (double)gg * ((x[o0 + i] - mu) * rs);
                }
            }
        } else {
            const long long total = (long long)N * HW;
            for (long long e = threadIdx.x; e < total; e += blockDim.x) {
                const long long n = e / HW, i = e - n * HW, o = (n * C + gidx) * HW + i;
                float gg = g[o];
                if (relu && !(y[o] > 0.f)) gg = 0.f;
                s += gg; q += (double)gg * ((x[o] - mu) * rs);
            }
        }
    }
    s = block_sum(s, sh);
    q = block_sum(q, sh);
    if (threadIdx.x == 0) {
        sum_g[gidx] = (float)s;
        sum_gx[gidx] = (float)q;
        const int c = instance ? gidx % C : gidx;
        if (dgamma) atomicAdd(dgamma + c, (float)q);
        if (dbeta) atomicAdd(dbeta + c, (float)s);
    }
}

// batch statistics: dx = gamma / sigma (g^ - mean(g^) - x^ mean(g^ x^));  fixed (running) statistics: dx = gamma / sigma g^
__global__ void norm_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ x,
                                      const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ gamma,
                                      const float* __restrict__ sum_g, const float* __restrict__ sum_gx, long long total, int N, int C, int HW,
                                      int instance, int per_plane_stats, int batch_stats, int relu, float eps, float* __restrict__ dx) {
    const float inv_cnt = instance ? 1.f / (float)HW : 1.f / ((float)N * (float)HW);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long pl = e / HW;
        const int c = (int)(pl % C);
        const long long gi = instance ? pl : c, si = (instance && !per_plane_stats) ? c : gi;
        const float rs = rsqrtf(var[si] + eps);
        float gg = g[e];
        if (relu && !(y[e] > 0.f)) gg = 0.f;
        float v = gg;
        if (batch_stats) v = gg - sum_g[gi] * inv_cnt - (x[e] - mean[si]) * rs * sum_gx[gi] * inv_cnt;
        dx[e] = gamma[c] * rs * v;
    }
}

// ---- GLU over dim 1: x [N, 2C, L] -> y [N, C, L] = a * sigmoid(b) ---------------------------------------------------------------
__global__ void glu_fwd_kernel(const float* __restrict__ x, long long total, int C, int L, float* __restrict__ y) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long cl = (long long)C * L, n = e / cl, r = e - n * cl;
        const float a = x[n * 2 * cl + r], b = x[n * 2 * cl + cl + r];
        y[e] = a * sigmoidf_(b);
    }
}
__global__ void glu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, long long total, int C, int L, float* __restrict__ dx) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long cl = (long long)C * L, n = e / cl, r = e - n * cl;
        const float a = x[n * 2 * cl + r], s = sigmoidf_(x[n * 2 * cl + cl + r]), gg = g[e];
        dx[n * 2 * cl + r] = gg * s;
        dx[n * 2 * cl + cl + r] = gg * a * s * (1.f - s);
    }
}

// y[n][c][i] = x[n][c][i] + bias[c] (the last ConvTranspose2d of the BatchNorm decoder keeps its bias) and the bias gradient
__global__ void chan_bias_fwd_kernel(const float* __restrict__ x, long long total, int C, int HW, const float* __restrict__ bias, int relu,
                                     float* __restrict__ y) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const float v = x[e] + bias[(int)((e / HW) % C)];
        y[e] = (relu && v < 0.f) ? 0.f : v;
    }
}
__global__ void chan_bias_bwd_kernel(const float* __restrict__ g, int N, int C, int HW, float* __restrict__ dbias) {
    __shared__ double sh[32];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        const float* p = g + ((long long)n * C + c) * HW;
        for (int i = threadIdx.x; i < HW; i += blockDim.x) s += p[i];
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) atomicAdd(dbias + c, (float)s);
}

// ---- bf16 tensor-core route of the general convolution (bf16 mode): NHWC staging, explicit im2col rows, dense tcgen05 GEMMs ----
// batched transpose in[n][A][B] -> out[n][B][A] with conversion (NCHW fp32 -> NHWC bf16: A = C, B = HW; NHWC fp32 -> NCHW fp32: A = HW, B = C)
template <typename Tin, typename Tout>
__global__ void btranspose_kernel(const Tin* __restrict__ in, int A, int B, Tout* __restrict__ out) {
    __shared__ float tile[32][33];
    const long long img = blockIdx.z;
    const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
    const Tin* ip = in + img * (long long)A * B;
    Tout* op = out + img * (long long)A * B;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int a = a0 + i, b = b0 + threadIdx.x;
        tile[i][threadIdx.x] = (a < A && b < B) ? (float)ip[(long long)a * B + b] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int b = b0 + i, a = a0 + threadIdx.x;
        if (a < A && b < B) op[(long long)b * A + a] = (Tout)tile[threadIdx.x][i];
    }
}

// col[m][tap * Cin + ci] = x_nhwc[n][ho*SH - PH + kh][wo*SW - PW + kw][ci]  (bf16, 8 channels = 16 bytes per thread)
__global__ void im2col_nhwc_kernel(const uint4* __restrict__ x, int N, int H, int W, int C8, int KH, int KW, int SH, int SW, int PH, int PW,
                                   int Ho, int Wo, uint4* __restrict__ col) {
    const long long total = (long long)N * Ho * Wo * KH * KW * C8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C8);
        long long r = e / C8;
        const int tap = (int)(r % (KH * KW));
        r /= KH * KW;
        const int wo = (int)(r % Wo);
        r /= Wo;
        const int ho = (int)(r % Ho), n = (int)(r / Ho);
        const int kh = tap / KW, kw = tap - kh * KW, h = ho * SH - PH + kh, w = wo * SW - PW + kw;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(x + (((long long)n * H + h) * W + w) * C8 + c);
        col[e] = v;
    }
}

// dx_nhwc[n][h][w][ci] = sum over the taps that reach (h, w) of dcol[(n, ho, wo)][tap * Cin + ci]   (fp32 accumulate / output)
__global__ void col2im_nhwc_kernel(const uint4* __restrict__ dcol, int N, int H, int W, int C8, int KH, int KW, int SH, int SW, int PH, int PW,
                                   int Ho, int Wo, float4* __restrict__ dx) {
    const long long total = (long long)N * H * W * C8;
    const long long rowv = (long long)KH * KW * C8;          // uint4 per im2col row
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C8);
        long long r = e / C8;
        const int w = (int)(r % W);
        r /= W;
        const int h = (int)(r % H), n = (int)(r / H);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int kh = (h + PH) % SH; kh < KH; kh += SH) {
            const int ho = (h + PH - kh) / SH;
            if (ho < 0 || ho >= Ho) continue;
            for (int kw = (w + PW) % SW; kw < KW; kw += SW) {
                const int wo = (w + PW - kw) / SW;
                if (wo < 0 || wo >= Wo) continue;
                const uint4 v = __ldg(dcol + (((long long)n * Ho + ho) * Wo + wo) * rowv + (long long)(kh * KW + kw) * C8 + c);
                const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __bfloat1622float2(p[i]);
                    acc[2 * i] += f.x; acc[2 * i + 1] += f.y;
                }
            }
        }
        dx[2 * e] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dx[2 * e + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// grad[co][ci][tap] += dw2[co][tap * Cin + ci]   (the GEMM's tap-major weight gradient back into the parameter's layout)
__global__ void wperm_add_kernel(const float* __restrict__ dw2, long long total, int Cin, int KHW, float* __restrict__ grad) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long K = (long long)Cin * KHW, co = e / K, k = e - co * K;
        const int tap = (int)(k / Cin), ci = (int)(k - (long long)tap * Cin);
        grad[co * K + (long long)ci * KHW + tap] += dw2[e];
    }
}
// w2[co][tap * Cin + ci] = w[co][ci][tap]
__global__ void wperm_kernel(const float* __restrict__ w, long long total, int Cin, int KHW, float* __restrict__ w2) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long K = (long long)Cin * KHW, co = e / K, k = e - co * K;
        const int tap = (int)(k / Cin), ci = (int)(k - (long long)tap * Cin);
        w2[e] = w[co * K + (long long)ci * KHW + tap];
    }
}

inline int ew_blocks(long long total) { return (int)std::max<long long>(1, std::min<long long>(148 * 16, (total + 255) / 256)); }

}  // namespace

extern "C" int mrssm_gconv_fwd(const mrssm_gconv_args* a, void* stream) {
    if (int e = check_geom(a)) return e;
    MRSSM_CHECK(a->x && a->w && a->y, "gconv_fwd: null tensor");
    GC g = make(a);
    g.out = a->y; g.M = (long long)a->N * a->Ho * a->Wo; g.Ncols = a->Cout; g.K = (long long)a->Cin * a->KH * a->KW;
    const size_t wsm = (size_t)g.K * 4 * sizeof(float);
    if (g.Ncols <= 4 && wsm <= 96 * 1024) {
        MRSSM_CUDA(cudaFuncSetAttribute(gconv_fwd_small_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
        gconv_fwd_small_kernel<4><<<(unsigned)std::min<long long>(148 * 16, ceil_div64(g.M, 256)), 256, wsm, (cudaStream_t)stream>>>(g);
    } else if (g.Ncols <= 16) {
        dim3 grid((unsigned)ceil_div64(g.M, 256), 1);
        gconv_kernel<0, 256, 16><<<grid, NT, 0, (cudaStream_t)stream>>>(g);
    } else {
        dim3 grid((unsigned)ceil_div64(g.M, 64), (unsigned)ceil_div64(g.Ncols, 64));
        gconv_kernel<0, 64, 64><<<grid, NT, 0, (cudaStream_t)stream>>>(g);
    }
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_gconv_dgrad(const mrssm_gconv_args* a, void* stream) {
    if (int e = check_geom(a)) return e;
    MRSSM_CHECK(a->dx && a->w && a->y, "gconv_dgrad: null tensor");
    GC g = make(a);
    g.out = a->dx; g.Ncols = a->Cin; g.K = (long long)a->Cout * a->KH * a->KW;
    g.M = (long long)a->N * ceil_div64(a->H, a->SH) * ceil_div64(a->W, a->SW);          // rows of the largest parity class
    if (g.Ncols <= 16) {
        dim3 grid((unsigned)ceil_div64(g.M, 256), 1, (unsigned)(a->SH * a->SW));
        gconv_kernel<1, 256, 16><<<grid, NT, 0, (cudaStream_t)stream>>>(g);
    } else {
        dim3 grid((unsigned)ceil_div64(g.M, 64), (unsigned)ceil_div64(g.Ncols, 64), (unsigned)(a->SH * a->SW));
        gconv_kernel<1, 64, 64><<<grid, NT, 0, (cudaStream_t)stream>>>(g);
    }
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_gconv_wgrad(const mrssm_gconv_args* a, void* stream) {
    if (int e = check_geom(a)) return e;
    MRSSM_CHECK(a->x && a->dw && a->y, "gconv_wgrad: null tensor");
    GC g = make(a);
    g.out = a->dw; g.M = a->Cout; g.Ncols = a->Cin * a->KH * a->KW; g.K = (long long)a->N * a->Ho * a->Wo;
    if (a->Cout <= 4 && a->KH == 7 && a->KW == 7) {          // SoundDecoder_v2.out: the 7x7 Conv2d to one channel
        const int n_slices = (int)std::max<long long>(1, std::min<long long>(a->N, ceil_div64(4 * 148, (long long)a->Cin * a->Cout)));
        dim3 grid((unsigned)a->Cin, (unsigned)a->Cout, (unsigned)n_slices);
        gconv_wgrad_small_kernel<7, 7><<<grid, 256, 0, (cudaStream_t)stream>>>(g, n_slices);
        MRSSM_LAUNCH_CHECK();
        return 0;
    }
    // few output channels: the transposed tile, 16 weight rows x 256 taps per CTA
    const bool narrow = g.M <= 16;
    const int BMw = narrow ? 16 : 64, BNw = narrow ? 256 : 64;
    const long long tiles = ceil_div64(g.M, BMw) * ceil_div64(g.Ncols, BNw);
    long long splits = std::max<long long>(1, std::min<long long>(ceil_div64(4 * 148, tiles), g.K / (8 * BK)));
    splits = std::min<long long>(splits, 65535);
    g.kchunk = ceil_div64(ceil_div64(g.K, splits), BK) * BK;
    splits = ceil_div64(g.K, g.kchunk);
    dim3 grid((unsigned)ceil_div64(g.M, BMw), (unsigned)ceil_div64(g.Ncols, BNw), (unsigned)splits);
    if (narrow) gconv_kernel<2, 16, 256><<<grid, NT, 0, (cudaStream_t)stream>>>(g);
    else gconv_kernel<2, 64, 64><<<grid, NT, 0, (cudaStream_t)stream>>>(g);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_norm_fwd(const mrssm_norm_args* a, void* stream) {
    MRSSM_CHECK(a && a->x && a->y && a->gamma && a->beta && a->mean && a->var && a->N > 0 && a->C > 0 && a->HW > 0, "norm_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int groups = a->instance ? a->N * a->C : a->C;
    const long long total = (long long)a->N * a->C * a->HW;
    if (a->batch_stats) {
        MRSSM_CHECK(a->instance ? a->HW > 1 || !a->running_mean : (long long)a->N * a->HW > 1, "norm_fwd: a single value per group has no variance");
        norm_stats_kernel<<<groups, 256, 0, st>>>(a->x, a->N, a->C, a->HW, a->instance, a->mean, a->var);
        MRSSM_LAUNCH_CHECK();
        if (a->running_mean && a->running_var) {
            norm_running_kernel<<<(a->C + 127) / 128, 128, 0, st>>>(a->mean, a->var, a->N, a->C, a->HW, a->instance, a->momentum,
                                                                   a->running_mean, a->running_var);
            MRSSM_LAUNCH_CHECK();
        }
        if (a->HW >= 64)
            norm_apply_plane_kernel<<<(unsigned)((long long)a->N * a->C), a->HW >= 512 ? 256 : 64, 0, st>>>(a->x, a->mean, a->var, a->gamma, a->beta,
                                                                                                        a->C, a->HW, a->instance, a->relu, a->eps, a->y);
        else
            norm_apply_kernel<<<ew_blocks(total), 256, 0, st>>>(a->x, a->mean, a->var, a->gamma, a->beta, total, a->C, a->HW, a->instance, a->relu,
                                                               a->eps, a->y);
    } else {
        MRSSM_CHECK(a->running_mean && a->running_var, "norm_fwd: fixed statistics requested without running buffers");
        if (a->HW >= 64)
            norm_apply_plane_kernel<<<(unsigned)((long long)a->N * a->C), a->HW >= 512 ? 256 : 64, 0, st>>>(a->x, a->running_mean, a->running_var,
                                                                                                        a->gamma, a->beta, a->C, a->HW, 0, a->relu,
                                                                                                        a->eps, a->y);
        else
            norm_apply_kernel<<<ew_blocks(total), 256, 0, st>>>(a->x, a->running_mean, a->running_var, a->gamma, a->beta, total, a->C, a->HW, 0,
                                                               a->relu, a->eps, a->y);
    }
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_norm_bwd(const mrssm_norm_args* a, const float* g, float* sum_g, float* sum_gx, float* dgamma, float* dbeta, float* dx,
                              void* stream) {
    MRSSM_CHECK(a && a->x && a->y && a->gamma && g && sum_g && sum_gx && dx && a->N > 0 && a->C > 0 && a->HW > 0, "norm_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int groups = a->instance ? a->N * a->C : a->C;
    const long long total = (long long)a->N * a->C * a->HW;
    const float* mean = a->batch_stats ? a->mean : a->running_mean;
    const float* var = a->batch_stats ? a->var : a->running_var;
    MRSSM_CHECK(mean && var, "norm_bwd: statistics missing");
    const int per_plane = a->instance && a->batch_stats;
    norm_bwd_reduce_kernel<<<groups, 256, 0, st>>>(g, a->y, a->x, mean, var, a->N, a->C, a->HW, a->instance, per_plane, a->relu, a->eps, sum_g,
                                                  sum_gx, dgamma, dbeta);
    MRSSM_LAUNCH_CHECK();
    if (a->HW >= 64)
        norm_bwd_apply_plane_kernel<<<(unsigned)((long long)a->N * a->C), a->HW >= 512 ? 256 : 64, 0, st>>>(g, a->y, a->x, mean, var, a->gamma, sum_g,
                                                                                                        sum_gx, a->N, a->C, a->HW, a->instance,
                                                                                                        per_plane, a->batch_stats, a->relu, a->eps, dx);
    else
        norm_bwd_apply_kernel<<<ew_blocks(total), 256, 0, st>>>(g, a->y, a->x, mean, var, a->gamma, sum_g, sum_gx, total, a->N, a->C, a->HW,
                                                               a->instance, per_plane, a->batch_stats, a->relu, a->eps, dx);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_glu_fwd(const float* x, int64_t N, int32_t C, int32_t L, float* y, void* stream) {
    MRSSM_CHECK(x && y && N > 0 && C > 0 && L > 0, "glu_fwd: bad arguments");
    const long long total = (long long)N * C * L;
    glu_fwd_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, total, C, L, y);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_glu_bwd(const float* x, const float* g, int64_t N, int32_t C, int32_t L, float* dx, void* stream) {
    MRSSM_CHECK(x && g && dx && N > 0 && C > 0 && L > 0, "glu_bwd: bad arguments");
    const long long total = (long long)N * C * L;
    glu_bwd_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, g, total, C, L, dx);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_chan_bias_fwd(const float* x, int64_t N, int32_t C, int32_t HW, const float* bias, int32_t relu, float* y, void* stream) {
    MRSSM_CHECK(x && y && bias && N > 0 && C > 0 && HW > 0, "chan_bias_fwd: bad arguments");
    const long long total = (long long)N * C * HW;
    chan_bias_fwd_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, total, C, HW, bias, relu, y);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_chan_bias_bwd(const float* g, int64_t N, int32_t C, int32_t HW, float* dbias, void* stream) {
    MRSSM_CHECK(g && dbias && N > 0 && C > 0 && HW > 0, "chan_bias_bwd: bad arguments");
    dim3 grid((unsigned)C, (unsigned)std::min<long long>(N, 64));
    chan_bias_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, (int)N, C, HW, dbias);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

// ---- staging kernels of the tensor-core route (mrssm_b200/ops.py: GConvTCFn) -----------------------------------------------------
extern "C" int mrssm_nchw_to_nhwc_bf16(const float* x, int64_t N, int32_t C, int32_t HW, void* out, void* stream) {
    MRSSM_CHECK(x && out && N > 0 && N <= 65535 && C > 0 && HW > 0, "nchw_to_nhwc_bf16: bad arguments");
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N), block(32, 8);
    btranspose_kernel<float, __nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(x, C, HW, (__nv_bfloat16*)out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_nhwc_to_nchw_f32(const float* x, int64_t N, int32_t C, int32_t HW, float* out, void* stream) {
    MRSSM_CHECK(x && out && N > 0 && N <= 65535 && C > 0 && HW > 0, "nhwc_to_nchw_f32: bad arguments");
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)((HW + 31) / 32), (unsigned)N), block(32, 8);
    btranspose_kernel<float, float><<<grid, block, 0, (cudaStream_t)stream>>>(x, HW, C, out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_im2col_nhwc(const mrssm_gconv_args* a, const void* x_nhwc, void* col, void* stream) {
    if (int e = check_geom(a)) return e;
    MRSSM_CHECK(x_nhwc && col && a->Cin % 8 == 0, "im2col_nhwc: needs bf16 NHWC input with channels in multiples of 8");
    const long long total = (long long)a->N * a->Ho * a->Wo * a->KH * a->KW * (a->Cin / 8);
    im2col_nhwc_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)x_nhwc, a->N, a->H, a->W, a->Cin / 8, a->KH, a->KW, a->SH,
                                                                          a->SW, a->PH, a->PW, a->Ho, a->Wo, (uint4*)col);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_col2im_nhwc(const mrssm_gconv_args* a, const void* dcol, float* dx_nhwc, void* stream) {
    if (int e = check_geom(a)) return e;
    MRSSM_CHECK(dcol && dx_nhwc && a->Cin % 8 == 0, "col2im_nhwc: needs channels in multiples of 8");
    const long long total = (long long)a->N * a->H * a->W * (a->Cin / 8);
    col2im_nhwc_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)dcol, a->N, a->H, a->W, a->Cin / 8, a->KH, a->KW, a->SH,
                                                                          a->SW, a->PH, a->PW, a->Ho, a->Wo, (float4*)dx_nhwc);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_gconv_weight_perm(const float* w, int64_t Cout, int32_t Cin, int32_t KHW, float* w2, void* stream) {
    MRSSM_CHECK(w && w2 && Cout > 0 && Cin > 0 && KHW > 0, "gconv_weight_perm: bad arguments");
    const long long total = (long long)Cout * Cin * KHW;
    wperm_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(w, total, Cin, KHW, w2);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_gconv_weight_perm_add(const float* dw2, int64_t Cout, int32_t Cin, int32_t KHW, float* grad, void* stream) {
    MRSSM_CHECK(dw2 && grad && Cout > 0 && Cin > 0 && KHW > 0, "gconv_weight_perm_add: bad arguments");
    const long long total = (long long)Cout * Cin * KHW;
    wperm_add_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(dw2, total, Cin, KHW, grad);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
