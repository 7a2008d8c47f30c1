// Generic CUDA-core implicit GEMM for the conv family (fp32 accumulate).
#include <algorithm>
//
// This is the exact-fp32 path ("fp32 mode", parity rtol 1e-3 and the bring-up reference for the
// tcgen05 kernels).  All three ops are one tile engine over a *separable gather*: an operand
// element is p[rowOff(row) + kOff(k)] (zero when a bounds predicate fails), so im2col is never
// materialised and NCHW / NHWC / strided views are all handled by strides.
//
//   DOWN : C[m=(img,hs,ws)][n=cs]      = sum_{k=(kh,kw,cl)}  large[img,2hs+kh,2ws+kw,cl] * W[cs,cl,kh,kw]
//   UP   : C[m=(img,a,b)|parity][n=cl] = sum_{k=(th,tw,cs)}  small[img,a-th,b-tw,cs]   * W[cs,cl,ph+2th,pw+2tw]
//   WGRAD: C[m=cs][n=(kh,kw,cl)]      += sum_{k=(img,hs,ws)} small[k,cs] * large[img,2hs+kh,2ws+kw,cl]
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;
enum { OP_DOWN = 0, OP_UP = 1, OP_WGRAD = 2 };

struct T4 {
    const void* p;
    long long sI, sH, sW, sC;
};

struct ConvK {
    int n_img, Hl, Wl, Cl, Hs, Ws, Cs, ksz;
    int act, mask_mode, accumulate;
    T4 large, small;
    float* weight;
    long long w_ss, w_sl;
    const float* bias;
    const void* mask;
    // derived
    int M, N, K;         // GEMM extents (per parity class for UP: M computed in-kernel)
    int nt;              // taps per dim for UP
    long long kchunk;    // WGRAD split-K chunk
};

template <typename T> __device__ __forceinline__ float ldf(const T* p, long long i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, long long i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, long long i) {
    return __bfloat162float(p[i]);
}
template <typename T> __device__ __forceinline__ void stf(T* p, long long i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, long long i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, long long i, float v) {
    p[i] = __float2bfloat16(v);
}

template <int OP, typename T>
__global__ void __launch_bounds__(NT) conv_simt_kernel(ConvK a) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    __shared__ long long rowOffA[BM], rowOut[BM], rowOffB[BN], colOut[BN];
    __shared__ int rowY[BM], rowX[BM];
    __shared__ long long kOffA[BK], kOffB[BK];
    __shared__ int kDy[BK], kDx[BK], kOk[BK];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    int ph = 0, pw = 0, Ha = 0, Wa = 0;
    int M = a.M, N = a.N;
    long long k_begin = 0, k_end = a.K;
    if (OP == OP_UP) {
        ph = blockIdx.z >> 1;
        pw = blockIdx.z & 1;
        Ha = (a.Hl - ph + 1) / 2;
        Wa = (a.Wl - pw + 1) / 2;
        M = a.n_img * Ha * Wa;
        if (m0 >= M) return;
    }
    if (OP == OP_WGRAD) {
        k_begin = (long long)blockIdx.z * a.kchunk;
        k_end = min((long long)a.K, k_begin + a.kchunk);
        if (k_begin >= k_end) return;
    }
    const T* Lp = (const T*)a.large.p;
    const T* Sp = (const T*)a.small.p;

    // ---- per-block row / column decode -------------------------------------------------------
    if (tid < BM) {
        int m = m0 + tid;
        long long offA = 0, offO = 0;
        int y = 0, x = 0;
        if (m < M) {
            if (OP == OP_DOWN) {
                int img = m / (a.Hs * a.Ws), r = m % (a.Hs * a.Ws), hs = r / a.Ws, ws = r % a.Ws;
                offA = img * a.large.sI + 2 * hs * a.large.sH + 2 * ws * a.large.sW;
                offO = img * a.small.sI + hs * a.small.sH + ws * a.small.sW;
            } else if (OP == OP_UP) {
                int img = m / (Ha * Wa), r = m % (Ha * Wa), aa = r / Wa, bb = r % Wa;
                offA = img * a.small.sI + aa * a.small.sH + bb * a.small.sW;
                offO = img * a.large.sI + (2 * aa + ph) * a.large.sH + (2 * bb + pw) * a.large.sW;
                y = aa;
                x = bb;
            } else {
                offA = m * a.small.sC;
                offO = m * a.w_ss;
            }
        }
        rowOffA[tid] = offA;
        rowOut[tid] = offO;
        rowY[tid] = y;
        rowX[tid] = x;
    } else if (tid < BM + BN) {
        int j = tid - BM, n = n0 + j;
        long long offB = 0, offO = 0;
        if (n < N) {
            if (OP == OP_DOWN) {
                offB = n * a.w_ss;
                offO = n * a.small.sC;
            } else if (OP == OP_UP) {
                offB = n * a.w_sl;
                offO = n * a.large.sC;
            } else {
                int tap = n / a.Cl, cl = n % a.Cl, kh = tap / a.ksz, kw = tap % a.ksz;
                offB = kh * a.large.sH + kw * a.large.sW + cl * a.large.sC;
                offO = cl * a.w_sl + tap;
            }
        }
        rowOffB[j] = offB;
        colOut[j] = offO;
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int ty = tid / 16, tx = tid % 16;

    for (long long kt = k_begin; kt < k_end; kt += BK) {
        __syncthreads();   // previous tile consumed; also orders the row decode on the first pass
        if (tid < BK) {
            long long k = kt + tid;
            long long oa = 0, ob = 0;
            int dy = 0, dx = 0, ok = 0;
            if (k < k_end) {
                ok = 1;
                if (OP == OP_DOWN) {
                    int tap = (int)(k / a.Cl), cl = (int)(k % a.Cl), kh = tap / a.ksz, kw = tap % a.ksz;
                    oa = kh * a.large.sH + kw * a.large.sW + cl * a.large.sC;
                    ob = cl * a.w_sl + tap;
                } else if (OP == OP_UP) {
                    int tap = (int)(k / a.Cs), cs = (int)(k % a.Cs), th = tap / a.nt, tw = tap % a.nt;
                    int kh = ph + 2 * th, kw = pw + 2 * tw;
                    ok = (kh < a.ksz) && (kw < a.ksz);
                    oa = -th * a.small.sH - tw * a.small.sW + cs * a.small.sC;
                    ob = cs * a.w_ss + (ok ? kh * a.ksz + kw : 0);
                    dy = th;
                    dx = tw;
                } else {
                    int img = (int)(k / (a.Hs * a.Ws)), r = (int)(k % (a.Hs * a.Ws)), hs = r / a.Ws, ws = r % a.Ws;
                    oa = img * a.small.sI + hs * a.small.sH + ws * a.small.sW;
                    ob = img * a.large.sI + 2 * hs * a.large.sH + 2 * ws * a.large.sW;
                }
            }
            kOffA[tid] = oa;
            kOffB[tid] = ob;
            kDy[tid] = dy;
            kDx[tid] = dx;
            kOk[tid] = ok;
        }
        __syncthreads();
        // ---- gather the two tiles --------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int r, kk;
            if (OP == OP_WGRAD) {
                r = tid % 64;
                kk = tid / 64 + 4 * i;
            } else {
                kk = tid % 16;
                r = tid / 16 + 16 * i;
            }
            float va = 0.f, vb = 0.f;
            if (kOk[kk]) {
                if (m0 + r < M) {
                    if (OP == OP_DOWN) {
                        va = ldf<T>(Lp, rowOffA[r] + kOffA[kk]);
                    } else if (OP == OP_UP) {
                        int yy = rowY[r] - kDy[kk], xx = rowX[r] - kDx[kk];
                        if ((unsigned)yy < (unsigned)a.Hs && (unsigned)xx < (unsigned)a.Ws)
                            va = ldf<T>(Sp, rowOffA[r] + kOffA[kk]);
                    } else {
                        va = ldf<T>(Sp, rowOffA[r] + kOffA[kk]);
                    }
                }
                if (n0 + r < N) {
                    if (OP == OP_WGRAD)
                        vb = ldf<T>(Lp, rowOffB[r] + kOffB[kk]);
                    else
                        vb = __ldg(a.weight + rowOffB[r] + kOffB[kk]);
                }
            }
            As[kk][r] = va;
            Bs[kk][r] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }

    // ---- epilogue ---------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int r = ty * 4 + i;
        if (m0 + r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = tx * 4 + j;
            if (n0 + c >= N) continue;
            long long off = rowOut[r] + colOut[c];
            float v = acc[i][j];
            if (OP == OP_WGRAD) {
                atomicAdd(a.weight + off, v);
            } else {
                T* outp = (T*)(OP == OP_DOWN ? a.small.p : a.large.p);
                if (a.accumulate) v += ldf<T>(outp, off);
                if (a.bias) v += __ldg(a.bias + n0 + c);
                v = act_apply(v, a.act);
                if (a.mask_mode) v *= act_grad_from_out(ldf<T>((const T*)a.mask, off), a.mask_mode);
                stf<T>(outp, off, v);
            }
        }
    }
}

// dbias[c] += sum over pixels of small[pix][c]
template <typename T>
__global__ void colsum_t4_kernel(T4 s, int n_img, int Hs, int Ws, int Cs, float* out, long long chunk) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cs) return;
    long long P = (long long)n_img * Hs * Ws;
    long long p0 = (long long)blockIdx.y * chunk, p1 = min(P, p0 + chunk);
    float acc = 0.f;
    for (long long p = p0; p < p1; ++p) {
        int img = (int)(p / (Hs * Ws)), r = (int)(p % (Hs * Ws)), hs = r / Ws, ws = r % Ws;
        acc += ldf<T>((const T*)s.p, img * s.sI + hs * s.sH + ws * s.sW + c * s.sC);
    }
    atomicAdd(out + c, acc);
}

int fill(const mrssm_conv_args* a, ConvK& k) {
    MRSSM_CHECK(a != nullptr, "conv: null args");
    MRSSM_CHECK(a->n_img > 0 && a->Cl > 0 && a->Cs > 0 && a->ksz > 0, "conv: bad geometry");
    // Conv2d floors: the last row/column of `large` may be unused (64 -> 31 -> 14), never the reverse
    MRSSM_CHECK(a->Hl >= 2 * (a->Hs - 1) + a->ksz && a->Wl >= 2 * (a->Ws - 1) + a->ksz && a->Hs > 0 && a->Ws > 0,
                "conv: Hl=%d Hs=%d ksz=%d inconsistent (need Hl >= 2*(Hs-1)+ksz)", a->Hl, a->Hs, a->ksz);
    MRSSM_CHECK(a->dtype == MRSSM_F32 || a->dtype == MRSSM_BF16, "conv: bad dtype %d", a->dtype);
    MRSSM_CHECK(a->large.ptr && a->small.ptr && a->weight, "conv: null tensor");
    k.n_img = a->n_img; k.Hl = a->Hl; k.Wl = a->Wl; k.Cl = a->Cl;
    k.Hs = a->Hs; k.Ws = a->Ws; k.Cs = a->Cs; k.ksz = a->ksz;
    k.act = a->act; k.mask_mode = a->mask ? a->mask_mode : 0; k.accumulate = a->accumulate;
    k.large = {a->large.ptr, a->large.sI, a->large.sH, a->large.sW, a->large.sC};
    k.small = {a->small.ptr, a->small.sI, a->small.sH, a->small.sW, a->small.sC};
    k.weight = a->weight; k.w_ss = a->w_ss; k.w_sl = a->w_sl;
    k.bias = a->bias; k.mask = a->mask;
    k.nt = (a->ksz + 1) / 2;
    k.kchunk = 0;
    return 0;
}

}  // namespace

int conv_down_simt(const mrssm_conv_args* a, cudaStream_t st) {
    ConvK k;
    if (int e = fill(a, k)) return e;
    k.M = a->n_img * a->Hs * a->Ws; k.N = a->Cs; k.K = a->ksz * a->ksz * a->Cl;
    dim3 grid((unsigned)ceil_div64(k.M, BM), (unsigned)ceil_div64(k.N, BN), 1);
    if (a->dtype == MRSSM_F32) conv_simt_kernel<OP_DOWN, float><<<grid, NT, 0, st>>>(k);
    else conv_simt_kernel<OP_DOWN, __nv_bfloat16><<<grid, NT, 0, st>>>(k);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

int conv_up_simt(const mrssm_conv_args* a, cudaStream_t st) {
    ConvK k;
    if (int e = fill(a, k)) return e;
    int Ha = (a->Hl + 1) / 2, Wa = (a->Wl + 1) / 2;
    k.M = a->n_img * Ha * Wa; k.N = a->Cl; k.K = k.nt * k.nt * a->Cs;
    dim3 grid((unsigned)ceil_div64(k.M, BM), (unsigned)ceil_div64(k.N, BN), 4);
    if (a->dtype == MRSSM_F32) conv_simt_kernel<OP_UP, float><<<grid, NT, 0, st>>>(k);
    else conv_simt_kernel<OP_UP, __nv_bfloat16><<<grid, NT, 0, st>>>(k);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

int conv_wgrad_simt(const mrssm_conv_args* a, cudaStream_t st) {
    ConvK k;
    if (int e = fill(a, k)) return e;
    k.M = a->Cs; k.N = a->ksz * a->ksz * a->Cl;
    long long P = (long long)a->n_img * a->Hs * a->Ws;
    MRSSM_CHECK(P < (1ll << 31), "wgrad: too many pixels");
    k.K = (int)P;
    long long tiles = ceil_div64(k.M, BM) * ceil_div64(k.N, BN);
    long long want = ceil_div64(4 * 148, tiles);                 // ~4 waves of CTAs
    long long splits = std::max<long long>(1, std::min<long long>(want, ceil_div64(P, 4 * BK)));
    splits = std::min<long long>(splits, 65535);
    k.kchunk = ceil_div64(ceil_div64(P, splits), BK) * BK;
    splits = ceil_div64(P, k.kchunk);
    dim3 grid((unsigned)ceil_div64(k.M, BM), (unsigned)ceil_div64(k.N, BN), (unsigned)splits);
    if (a->dtype == MRSSM_F32) conv_simt_kernel<OP_WGRAD, float><<<grid, NT, 0, st>>>(k);
    else conv_simt_kernel<OP_WGRAD, __nv_bfloat16><<<grid, NT, 0, st>>>(k);
    MRSSM_LAUNCH_CHECK();
    if (a->bias) {
        long long chunk = std::max<long long>(256, ceil_div64(P, 592));
        dim3 g2((unsigned)ceil_div64(a->Cs, 128), (unsigned)ceil_div64(P, chunk));
        if (a->dtype == MRSSM_F32)
            colsum_t4_kernel<float><<<g2, 128, 0, st>>>(k.small, a->n_img, a->Hs, a->Ws, a->Cs, a->bias, chunk);
        else
            colsum_t4_kernel<__nv_bfloat16><<<g2, 128, 0, st>>>(k.small, a->n_img, a->Hs, a->Ws, a->Cs, a->bias, chunk);
        MRSSM_LAUNCH_CHECK();
    }
    return 0;
}

// out[c] += sum_{img,h,w} large[img,h,w,c]   (bias gradient of a ConvTranspose2d output; a->bias is the output)
extern "C" int mrssm_colsum_t4(const mrssm_conv_args* a, void* stream) {
    MRSSM_CHECK(a && a->large.ptr && a->bias && a->n_img > 0 && a->Cl > 0, "colsum_t4: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    T4 t = {a->large.ptr, a->large.sI, a->large.sH, a->large.sW, a->large.sC};
    long long P = (long long)a->n_img * a->Hl * a->Wl;
    long long chunk = std::max<long long>(256, ceil_div64(P, 2368));
    dim3 g2((unsigned)ceil_div64(a->Cl, 32), (unsigned)ceil_div64(P, chunk));
    if (a->dtype == MRSSM_F32) colsum_t4_kernel<float><<<g2, 32, 0, st>>>(t, a->n_img, a->Hl, a->Wl, a->Cl, a->bias, chunk);
    else colsum_t4_kernel<__nv_bfloat16><<<g2, 32, 0, st>>>(t, a->n_img, a->Hl, a->Wl, a->Cl, a->bias, chunk);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
