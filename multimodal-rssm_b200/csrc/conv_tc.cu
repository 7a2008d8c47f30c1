// tcgen05 implicit-GEMM kernels for the conv family (bf16 operands, fp32 accumulators in TMEM).
//
//   fwd-type (DOWN / UP):  D[128 pixel rows][BN out-channels] += A[rows][64 k] * W[BN][64 k]
//       A rows are gathered straight from the NHWC bf16 activation (im2col never exists) by cp.async
//       (16-byte channel vectors, zero-fill for padding taps) into 128B-swizzled K-major tiles;
//       W comes pre-packed (bf16, K-major, zero padded).  One elected thread issues tcgen05.mma
//       (M=128, N=BN, K=16), tcgen05.commit releases smem stages through mbarriers, 4 epilogue warps
//       read the accumulator with tcgen05.ld and apply bias / activation / act'-mask, storing bf16
//       NHWC or fp32 with arbitrary strides (NCHW).
//   wgrad: D[128 cs][BN (tap,cl) columns] += sum over 64-pixel blocks of small[pix][cs] x large-gather[pix][col]
//       both operands MN-major (a smem row is one pixel = one contraction index), split-K over pixel
//       ranges, fp32 atomics straight into the PyTorch-layout gradient.
#include <algorithm>
#include <cuda.h>
#include <string.h>
#include "tc_common.cuh"

int mrssm_tma_map_2d_sw128(void* map, const void* base, long long inner, long long outer, long long row_bytes, int box_inner, int box_outer);

namespace {

using bf16 = __nv_bfloat16;
enum { OP_DOWN = 0, OP_UP = 1 };
constexpr int BM = 128, BK = 64, NLOAD = 256, NTHREADS = 288, LAG = 2;

struct T4 {
    const void* p;
    long long sI, sH, sW, sC;
};

struct FwdK {
    int n_img, Hl, Wl, Hs, Ws, ksz, nt;
    int Cin;            // channels (padded, multiple of 8) of the gathered tensor
    int M;              // rows (DOWN); UP: per parity class, computed in-kernel
    int BN;             // output channels handled per CTA (multiple of 16, <= 256)
    int N_total;        // padded output channels over all N tiles
    int n_valid;        // real output channels
    int bias_mod;       // bias index = n % bias_mod
    int Kpad;           // multiple of 64
    int act, mask_mode, out_f32;
    int stages, tmem_cols;
    T4 in, out, mask;
    const bf16* w;      // [classes][N_total][Kpad]
    const float* bias;
};

template <int OP>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_fwd_kernel(FwdK a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[4], bar_empty[4], bar_accum;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = BM * 128 + a.BN * 128;
    const int nkb = a.Kpad / BK;
    const int S = a.stages;

    int ph = 0, pw = 0, Ha = 0, Wa = 0, M = a.M;
    if (OP == OP_UP) {
        ph = blockIdx.z >> 1;
        pw = blockIdx.z & 1;
        Ha = (a.Hl - ph + 1) / 2;
        Wa = (a.Wl - pw + 1) / 2;
        M = a.n_img * Ha * Wa;
        if ((int)blockIdx.x * BM >= M) return;
    }
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * a.BN;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(tc::smem_u32(&bar_full[s]), NLOAD);
            tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1);
        }
        tc::mbar_init(tc::smem_u32(&bar_accum), 1);
        tc::fence_barrier_init();
    }
    if (warp == 8) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), a.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 8) {
        // ------------------------------ loaders ------------------------------
        const bf16* inp = (const bf16*)a.in.p;
        long long in_base = 0, out_off = 0, mask_off = 0;
        int ya = 0, xb = 0;
        bool row_ok = false;
        int brow0 = 0;
        if (warp < 4) {
            int m = m0 + tid;
            row_ok = m < M;
            if (row_ok) {
                if (OP == OP_DOWN) {
                    int img = m / (a.Hs * a.Ws), r = m % (a.Hs * a.Ws), hs = r / a.Ws, ws = r % a.Ws;
                    in_base = img * a.in.sI + 2 * hs * a.in.sH + 2 * ws * a.in.sW;
                    out_off = img * a.out.sI + hs * a.out.sH + ws * a.out.sW;
                    mask_off = img * a.mask.sI + hs * a.mask.sH + ws * a.mask.sW;
                } else {
                    int img = m / (Ha * Wa), r = m % (Ha * Wa);
                    ya = r / Wa;
                    xb = r % Wa;
                    in_base = img * a.in.sI + ya * a.in.sH + xb * a.in.sW;
                    out_off = img * a.out.sI + (2 * ya + ph) * a.out.sH + (2 * xb + pw) * a.out.sW;
                    mask_off = img * a.mask.sI + (2 * ya + ph) * a.mask.sH + (2 * xb + pw) * a.mask.sW;
                }
            }
        } else {
            brow0 = tid - 128;
        }
        const bf16* wbase = a.w + ((long long)(OP == OP_UP ? blockIdx.z : 0) * a.N_total + n0) * a.Kpad;
        const int ntap = (OP == OP_DOWN) ? a.ksz * a.ksz : a.nt * a.nt;

        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            if (kb >= S) tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ((kb / S) - 1) & 1);
            const uint32_t sA = smem0 + s * stage_bytes, sB = sA + BM * 128;
            if (warp < 4) {
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    int k0 = kb * BK + p * 8;
                    int tap = k0 / a.Cin, c = k0 - tap * a.Cin;
                    bool ok = row_ok && tap < ntap;
                    long long off = in_base + c;
                    if (OP == OP_DOWN) {
                        int kh = tap / a.ksz, kw = tap - kh * a.ksz;
                        off += kh * a.in.sH + kw * a.in.sW;
                    } else {
                        int th = tap / a.nt, tw = tap - th * a.nt;
                        int yy = ya - th, xx = xb - tw;
                        ok = ok && (ph + 2 * th < a.ksz) && (pw + 2 * tw < a.ksz) && (unsigned)yy < (unsigned)a.Hs &&
                             (unsigned)xx < (unsigned)a.Ws;
                        off -= th * a.in.sH + tw * a.in.sW;
                    }
                    tc::cp_async16(sA + tc::sw128_off(tid, p), ok ? (const void*)(inp + off) : (const void*)inp, ok);
                }
            } else {
                for (int r = brow0; r < a.BN; r += 128) {
                    const bf16* src = wbase + (long long)r * a.Kpad + kb * BK;
#pragma unroll
                    for (int p = 0; p < 8; ++p) tc::cp_async16(sB + tc::sw128_off(r, p), src + p * 8, true);
                }
            }
            tc::cp_async_commit();
            if (kb >= LAG) {
                tc::cp_async_wait<LAG>();
                tc::fence_proxy_async();
                tc::mbar_arrive(tc::smem_u32(&bar_full[(kb - LAG) % S]));
            }
        }
        // drain: the last LAG groups
        if (nkb >= 2) {
            tc::cp_async_wait<1>();
            tc::fence_proxy_async();
            tc::mbar_arrive(tc::smem_u32(&bar_full[(nkb - 2) % S]));
        }
        tc::cp_async_wait<0>();
        tc::fence_proxy_async();
        tc::mbar_arrive(tc::smem_u32(&bar_full[(nkb - 1) % S]));

        if (warp < 4) {
            // ------------------------------ epilogue ------------------------------
            tc::mbar_wait(tc::smem_u32(&bar_accum), 0);
            tc::tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < a.BN; c0 += 16) {
                float v[16];
                tc::tmem_ld16(trow + c0, v);
                if (!row_ok) continue;
                const int nb = n0 + c0;
                if (nb >= a.n_valid && a.out_f32) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    int n = nb + i;
                    float x = v[i];
                    if (a.bias && n < a.n_valid) x += __ldg(a.bias + (n % a.bias_mod));
                    x = act_apply(x, a.act);
                    if (a.mask_mode) {
                        float mk = n < a.n_valid ? __bfloat162float(((const bf16*)a.mask.p)[mask_off + n * a.mask.sC]) : 0.f;
                        x *= act_grad_from_out(mk, a.mask_mode);
                    }
                    v[i] = n < a.n_valid ? x : 0.f;
                }
                if (a.out_f32) {
                    float* op = (float*)a.out.p + out_off;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (nb + i < a.n_valid) op[(long long)(nb + i) * a.out.sC] = v[i];
                } else {
                    __align__(16) bf16 h[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) h[i] = __float2bfloat16(v[i]);
                    uint4* dst = reinterpret_cast<uint4*>((bf16*)a.out.p + out_off + nb);
                    dst[0] = reinterpret_cast<const uint4*>(h)[0];
                    dst[1] = reinterpret_cast<const uint4*>(h)[1];
                }
            }
            tc::tc_fence_before();
        }
    } else {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_bf16(BM, a.BN, 0, 0);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % S;
                tc::mbar_wait(tc::smem_u32(&bar_full[s]), (kb / S) & 1);
                tc::tc_fence_after();
                const uint32_t sA = smem0 + s * stage_bytes, sB = sA + BM * 128;
#pragma unroll
                for (int j = 0; j < BK / 16; ++j) {
                    uint64_t ad = tc::smem_desc_sw128(sA + j * 32, 16, 1024);
                    uint64_t bd = tc::smem_desc_sw128(sB + j * 32, 16, 1024);
                    tc::umma_bf16(tmem_base, ad, bd, idesc, (kb | j) != 0);
                }
                tc::umma_commit(tc::smem_u32(&bar_empty[s]));
            }
            tc::umma_commit(tc::smem_u32(&bar_accum));
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 8) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------------
struct WgK {
    int n_img, Hl, Wl, Hs, Ws, ksz;
    int Cs, Cl;             // padded channels of the bf16 tensors
    int Cs_valid, Cl_valid;
    int BN;                 // columns per CTA: multiple of 64, <= 256
    int Ncols;              // ksz*ksz*Cl
    long long P, pchunk;    // contraction length (pixels) and split size (multiple of 64)
    int stages, tmem_cols;
    T4 small, large;
    float* dw;
    long long w_ss, w_sl;
    int tma;                    // dense case (1x1 small map, taps contiguous in the large row): both operands arrive by TMA (128B swizzle)
};

__device__ __forceinline__ void wg_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wg_tma_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap mS, const __grid_constant__ CUtensorMap mL, const WgK a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[4], bar_empty[4], bar_accum;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = a.stages;
    const uint32_t a_bytes = 2 * 8192;                     // 2 slabs of [64 pix][64 cs]
    const uint32_t stage_bytes = a_bytes + (a.BN / 64) * 8192;
    const int m0 = blockIdx.x * BM;                        // cs tile
    const int n0 = blockIdx.y * a.BN;                      // column tile
    const long long p_begin = (long long)blockIdx.z * a.pchunk, p_end = min(a.P, p_begin + a.pchunk);
    if (p_begin >= p_end) return;
    const int nkb = (int)((p_end - p_begin + BK - 1) / BK);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(tc::smem_u32(&bar_full[s]), a.tma ? 1 : NLOAD);
            tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1);
        }
        tc::mbar_init(tc::smem_u32(&bar_accum), 1);
        tc::fence_barrier_init();
    }
    if (warp == 8) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), a.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 8) {
        const bf16* sp = (const bf16*)a.small.p;
        const bf16* lp = (const bf16*)a.large.p;
        // thread -> (pixel row r of the 64-pixel block, quarter q): r = tid & 63, q = tid >> 6 (0..3)
        const int r = tid & 63, q = tid >> 6;
        const int nslabB = a.BN / 64;
        if (a.tma) {
            // dense layers: a K block is 64 whole rows of both matrices; one thread issues 2 + BN/64 TMA boxes of [64 rows][128 B]
            // (rows past the end and feature columns past the padded width are zero-filled by the TMA unit)
            if (tid == 0) {
                for (int kb = 0; kb < nkb; ++kb) {
                    const int s = kb % S;
                    if (kb >= S) tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ((kb / S) - 1) & 1);
                    const uint32_t sA = smem0 + s * stage_bytes, sB = sA + a_bytes;
                    const uint32_t bar = tc::smem_u32(&bar_full[s]);
                    const int row = (int)(p_begin + (long long)kb * BK);
                    wg_expect_tx(bar, stage_bytes);
                    wg_tma_2d(sA, &mS, m0, row, bar);
                    wg_tma_2d(sA + 8192, &mS, m0 + 64, row, bar);
                    for (int i = 0; i < nslabB; ++i) wg_tma_2d(sB + i * 8192, &mL, n0 + 64 * i, row, bar);
                }
            }
        } else
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            if (kb >= S) tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ((kb / S) - 1) & 1);
            const uint32_t sA = smem0 + s * stage_bytes, sB = sA + a_bytes;
            const long long pix = p_begin + (long long)kb * BK + r;
            const bool pok = pix < p_end;
            long long soff = 0, loff = 0;
            if (pok) {
                int img = (int)(pix / (a.Hs * a.Ws)), rr = (int)(pix % (a.Hs * a.Ws)), hs = rr / a.Ws, ws = rr % a.Ws;
                soff = img * a.small.sI + hs * a.small.sH + ws * a.small.sW;
                loff = img * a.large.sI + 2 * hs * a.large.sH + 2 * ws * a.large.sW;
            }
            // A': 128 cs = 16 chunks per pixel; this thread copies chunks q*4 .. q*4+3
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int ch = q * 4 + i;                        // 0..15
                int cs = m0 + ch * 8;
                bool ok = pok && cs < a.Cs;
                tc::cp_async16(sA + (ch >> 3) * 8192 + tc::sw128_off(r, ch & 7), ok ? (const void*)(sp + soff + cs) : (const void*)sp, ok);
            }
            // B': BN columns = BN/8 chunks per pixel, split over the 4 quarters
            for (int ch = q; ch < a.BN / 8; ch += 4) {
                int col = n0 + ch * 8;
                int tap = col / a.Cl, cl = col - tap * a.Cl;
                int kh = tap / a.ksz, kw = tap - kh * a.ksz;
                bool ok = pok && col < a.Ncols;
                tc::cp_async16(sB + (ch >> 3) * 8192 + tc::sw128_off(r, ch & 7),
                               ok ? (const void*)(lp + loff + kh * a.large.sH + kw * a.large.sW + cl) : (const void*)lp, ok);
            }
            (void)nslabB;
            tc::cp_async_commit();
            if (kb >= LAG) {
                tc::cp_async_wait<LAG>();
                tc::fence_proxy_async();
                tc::mbar_arrive(tc::smem_u32(&bar_full[(kb - LAG) % S]));
            }
        }
        if (!a.tma) {
            if (nkb >= 2) {
                tc::cp_async_wait<1>();
                tc::fence_proxy_async();
                tc::mbar_arrive(tc::smem_u32(&bar_full[(nkb - 2) % S]));
            }
            tc::cp_async_wait<0>();
            tc::fence_proxy_async();
            tc::mbar_arrive(tc::smem_u32(&bar_full[(nkb - 1) % S]));
        }

        if (warp < 4) {
            tc::mbar_wait(tc::smem_u32(&bar_accum), 0);
            tc::tc_fence_after();
            const int cs = m0 + tid;
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < a.BN; c0 += 16) {
                float v[16];
                tc::tmem_ld16(trow + c0, v);
                if (cs >= a.Cs_valid) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    int col = n0 + c0 + i;
                    if (col >= a.Ncols) continue;
                    int tap = col / a.Cl, cl = col - tap * a.Cl;
                    if (cl < a.Cl_valid) atomicAdd(a.dw + cs * a.w_ss + cl * a.w_sl + tap, v[i]);
                }
            }
            tc::tc_fence_before();
        }
    } else {
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_bf16(BM, a.BN, 1, 1);
            const uint32_t a_bytes2 = 2 * 8192;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % S;
                tc::mbar_wait(tc::smem_u32(&bar_full[s]), (kb / S) & 1);
                tc::tc_fence_after();
                const uint32_t sA = smem0 + s * stage_bytes, sB = sA + a_bytes2;
#pragma unroll
                for (int j = 0; j < BK / 16; ++j) {
                    // 16 pixels (contraction) per MMA = 2 swizzle atoms of 8 rows: advance 2048 B
                    uint64_t ad = tc::smem_desc_sw128(sA + j * 2048, 8192, 1024);
                    uint64_t bd = tc::smem_desc_sw128(sB + j * 2048, 8192, 1024);
                    tc::umma_bf16(tmem_base, ad, bd, idesc, (kb | j) != 0);
                }
                tc::umma_commit(tc::smem_u32(&bar_empty[s]));
            }
            tc::umma_commit(tc::smem_u32(&bar_accum));
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 8) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------
// packing / conversion helpers
// ------------------------------------------------------------------------------------------------------
// mode 0 DOWN : out[n=cs][k=(tap,cl)]                       = W[cs][cl][tap]
// mode 1 UP   : out[class][n=cl][k=(th,tw,cs)]              = W[cs][cl][ph+2th][pw+2tw]  (0 if tap outside kernel)
// mode 2 UP1x1: out[n=(tap,cl)][k=cs]                       = W[cs][cl][tap]            (ConvT on a 1x1 input = dense)
__global__ void pack_weight_kernel(const float* __restrict__ w, long long w_ss, long long w_sl, int Cs_valid, int Cl_valid,
                                   int Cs_pad, int Cl_pad, int ksz, int mode, int Npad, int Kpad, bf16* __restrict__ out) {
    long long total = (long long)(mode == 1 ? 4 : 1) * Npad * Kpad;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    const int nt = (ksz + 1) / 2;
    for (; i < total; i += stride) {
        int k = (int)(i % Kpad);
        long long t = i / Kpad;
        int n = (int)(t % Npad), cls = (int)(t / Npad);
        float v = 0.f;
        if (mode == 0) {
            int tap = k / Cl_pad, cl = k % Cl_pad;
            if (n < Cs_valid && cl < Cl_valid && tap < ksz * ksz) v = w[n * w_ss + cl * w_sl + tap];
        } else if (mode == 1) {
            int ph = cls >> 1, pw = cls & 1;
            int tap = k / Cs_pad, cs = k % Cs_pad, th = tap / nt, tw = tap % nt;
            int kh = ph + 2 * th, kw = pw + 2 * tw;
            if (n < Cl_valid && cs < Cs_valid && tap < nt * nt && kh < ksz && kw < ksz) v = w[cs * w_ss + n * w_sl + kh * ksz + kw];
        } else {
            int tap = n / Cl_pad, cl = n % Cl_pad;
            if (k < Cs_valid && cl < Cl_valid && tap < ksz * ksz) v = w[k * w_ss + cl * w_sl + tap];
        }
        out[i] = __float2bfloat16(v);
    }
}

// fp32 strided [n_img,H,W,C] -> bf16 NHWC with C padded to Cpad (zeros); scale folds constants (e.g. dY scaling)
__global__ void to_bf16_nhwc_kernel(T4 src, int n_img, int H, int W, int Cc, int Cpad, float scale, bf16* __restrict__ dst) {
    long long total = (long long)n_img * H * W * Cpad;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    const float* sp = (const float*)src.p;
    for (; i < total; i += stride) {
        int c = (int)(i % Cpad);
        long long t = i / Cpad;
        int x = (int)(t % W);
        t /= W;
        int y = (int)(t % H);
        int img = (int)(t / H);
        float v = c < Cc ? sp[img * src.sI + y * src.sH + x * src.sW + c * src.sC] * scale : 0.f;
        dst[i] = __float2bfloat16(v);
    }
}

// bf16 NHWC (padded C) -> fp32 strided
__global__ void from_bf16_nhwc_kernel(const bf16* __restrict__ src, int n_img, int H, int W, int Cc, int Cpad, T4 dst) {
    long long total = (long long)n_img * H * W * Cc;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    float* dp = (float*)dst.p;
    for (; i < total; i += stride) {
        int c = (int)(i % Cc);
        long long t = i / Cc;
        int x = (int)(t % W);
        t /= W;
        int y = (int)(t % H);
        int img = (int)(t / H);
        dp[img * dst.sI + y * dst.sH + x * dst.sW + c * dst.sC] = __bfloat162float(src[((long long)(img * H + y) * W + x) * Cpad + c]);
    }
}

// dbias[c] += sum over rows of bf16 x[row][c]  (x contiguous [rows][Cpad])
__global__ void colsum_bf16_kernel(const bf16* __restrict__ x, long long rows, int Cpad, int Cvalid, float* out, long long chunk) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cvalid) return;
    long long r0 = (long long)blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) acc += __bfloat162float(x[r * Cpad + c]);
    atomicAdd(out + c, acc);
}

inline int pow2_cols(int n) {
    int c = 32;
    while (c < n) c <<= 1;
    return c;
}

inline T4 cvt(const mrssm_t4& t) { return T4{t.ptr, t.sI, t.sH, t.sW, t.sC}; }

}  // namespace

extern "C" int mrssm_tc_pack_weight(const float* w, int64_t w_ss, int64_t w_sl, int32_t Cs_valid, int32_t Cl_valid,
                                    int32_t Cs_pad, int32_t Cl_pad, int32_t ksz, int32_t mode, int32_t Npad, int32_t Kpad,
                                    void* out, void* stream) {
    MRSSM_CHECK(w && out && Npad % 16 == 0 && Kpad % 64 == 0 && mode >= 0 && mode <= 2, "tc_pack_weight: bad args");
    long long total = (long long)(mode == 1 ? 4 : 1) * Npad * Kpad;
    int blocks = (int)std::min<long long>(148 * 8, ceil_div64(total, 256));
    pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, w_ss, w_sl, Cs_valid, Cl_valid, Cs_pad, Cl_pad, ksz, mode,
                                                                 Npad, Kpad, (bf16*)out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_tc_to_bf16(const mrssm_t4* src, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t Cpad, float scale,
                                void* dst, void* stream) {
    MRSSM_CHECK(src && src->ptr && dst && Cpad >= C, "tc_to_bf16: bad args");
    long long total = (long long)n_img * H * W * Cpad;
    int blocks = (int)std::min<long long>(148 * 16, ceil_div64(total, 256));
    to_bf16_nhwc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cvt(*src), n_img, H, W, C, Cpad, scale, (bf16*)dst);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_tc_from_bf16(const void* src, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t Cpad,
                                  const mrssm_t4* dst, void* stream) {
    MRSSM_CHECK(src && dst && dst->ptr && Cpad >= C, "tc_from_bf16: bad args");
    long long total = (long long)n_img * H * W * C;
    int blocks = (int)std::min<long long>(148 * 16, ceil_div64(total, 256));
    from_bf16_nhwc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, n_img, H, W, C, Cpad, cvt(*dst));
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_tc_colsum(const void* x, int64_t rows, int32_t Cpad, int32_t Cvalid, float* out, void* stream) {
    MRSSM_CHECK(x && out && rows > 0 && Cvalid <= Cpad, "tc_colsum: bad args");
    long long chunk = std::max<long long>(64, ceil_div64(rows, 1184));
    dim3 grid((unsigned)ceil_div64(Cvalid, 64), (unsigned)ceil_div64(rows, chunk));
    colsum_bf16_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>((const bf16*)x, rows, Cpad, Cvalid, out, chunk);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

int dense_tc_launch(const void* A, long long lda, int M, int K, const void* wpacked, int Npad, int Kpad, int n_valid, const float* bias,
                    int bias_mod, int act, const void* mask, long long ldm, int mask_mode, void* out, long long ldc, long long cstride,
                    int out_f32, cudaStream_t st, const float* addend = nullptr, long long addend_ld = 0, int group_n = 0, int group_k = 0);

static int launch_fwd(const mrssm_tc_conv_args* a, int op, cudaStream_t st) {
    MRSSM_CHECK(a && a->large.ptr && a->small.ptr && a->wpacked, "tc_conv: null tensor");
    {   // dense layers (nn.Linear, ConvT / Conv whose window covers the whole map): the TMA-fed GEMM of dense_tc.cu
        const mrssm_t4& in = (op == OP_DOWN) ? a->large : a->small;
        const mrssm_t4& out = (op == OP_DOWN) ? a->small : a->large;
        const int Cin = (op == OP_DOWN) ? a->Cl : a->Cs;
        const bool full_window = op == OP_DOWN ? (a->Hs == 1 && a->Ws == 1 && a->Hl == a->ksz && a->Wl == a->ksz &&
                                                  (a->ksz == 1 || (in.sW == Cin && in.sH == (long long)a->Wl * Cin)))
                                               : (a->ksz == 1 && a->Hl == 1 && a->Wl == 1 && a->Hs == 1 && a->Ws == 1);
        if (full_window && in.sC == 1 && Cin % 8 == 0 && in.sI % 8 == 0 && a->n_out_pad % 16 == 0 &&
            (a->n_out_pad <= 4096 || !a->bias || (a->bias_mod > 0 && a->bias_mod % 8 == 0 && a->bias_mod <= 4096 && a->n_out_valid % a->bias_mod == 0)) &&
            (a->out_f32 || (out.sC == 1 && out.sI % 8 == 0)) && (!a->mask.ptr || (a->mask.sC == 1 && a->mask.sI % 8 == 0))) {
            const int K = (op == OP_DOWN) ? a->ksz * a->ksz * Cin : Cin;
            const int Kpad = (int)(ceil_div64(a->group_n > 0 ? a->group_k : K, BK) * BK);
            return dense_tc_launch(in.ptr, in.sI, a->n_img, K, a->wpacked, a->n_out_pad, Kpad, a->n_out_valid, a->bias, a->bias_mod, a->act,
                                   a->mask.ptr, a->mask.sI, a->mask_mode, out.ptr, out.sI, out.sC, a->out_f32, st, a->addend, a->addend_ld,
                                   a->group_n, a->group_k);
        }
    }
    MRSSM_CHECK(!a->addend && a->group_n == 0, "tc_conv: addend / grouping are supported by the dense (1x1 / full-window) layers only");
    FwdK k;
    k.n_img = a->n_img; k.Hl = a->Hl; k.Wl = a->Wl; k.Hs = a->Hs; k.Ws = a->Ws; k.ksz = a->ksz; k.nt = (a->ksz + 1) / 2;
    const mrssm_t4& in = (op == OP_DOWN) ? a->large : a->small;
    const mrssm_t4& out = (op == OP_DOWN) ? a->small : a->large;
    k.Cin = (op == OP_DOWN) ? a->Cl : a->Cs;
    MRSSM_CHECK(k.Cin % 8 == 0 && in.sC == 1 && in.sW % 8 == 0 && in.sH % 8 == 0 && in.sI % 8 == 0,
                "tc_conv: gathered tensor must be bf16 NHWC with channels padded to 8 (Cin=%d)", k.Cin);
    k.M = (op == OP_DOWN) ? a->n_img * a->Hs * a->Ws : a->n_img * ((a->Hl + 1) / 2) * ((a->Wl + 1) / 2);
    k.N_total = a->n_out_pad; k.n_valid = a->n_out_valid; k.bias_mod = a->bias_mod > 0 ? a->bias_mod : a->n_out_valid;
    MRSSM_CHECK(k.N_total % 16 == 0 && k.N_total > 0, "tc_conv: n_out_pad %d must be a multiple of 16", k.N_total);
    k.BN = k.N_total <= 256 ? k.N_total : (k.N_total % 256 == 0 ? 256 : 128);
    MRSSM_CHECK(k.N_total % k.BN == 0, "tc_conv: n_out_pad %d not tileable", k.N_total);
    int ktaps = (op == OP_DOWN) ? a->ksz * a->ksz : k.nt * k.nt;
    k.Kpad = (int)(ceil_div64((long long)ktaps * k.Cin, BK) * BK);
    k.act = a->act; k.mask_mode = a->mask.ptr ? a->mask_mode : 0; k.out_f32 = a->out_f32;
    MRSSM_CHECK(k.out_f32 || (out.sC == 1 && out.sW % 8 == 0), "tc_conv: bf16 output must be NHWC with padded channels");
    k.in = cvt(in); k.out = cvt(out); k.mask = cvt(a->mask);
    k.w = (const bf16*)a->wpacked; k.bias = a->bias;
    uint32_t stage = BM * 128 + k.BN * 128;
    k.stages = (int)std::max<uint32_t>(LAG + 1, std::min<uint32_t>(4, (200 * 1024) / stage));   // needs > LAG stages
    k.tmem_cols = pow2_cols(k.BN);
    size_t smem = (size_t)k.stages * stage + 1024;
    dim3 grid((unsigned)ceil_div64(k.M, BM), (unsigned)(k.N_total / k.BN), op == OP_UP ? 4 : 1);
    if (op == OP_DOWN) {
        MRSSM_CUDA(cudaFuncSetAttribute(conv_tc_fwd_kernel<OP_DOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_tc_fwd_kernel<OP_DOWN><<<grid, NTHREADS, smem, st>>>(k);
    } else {
        MRSSM_CUDA(cudaFuncSetAttribute(conv_tc_fwd_kernel<OP_UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_tc_fwd_kernel<OP_UP><<<grid, NTHREADS, smem, st>>>(k);
    }
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_tc_conv_down(const mrssm_tc_conv_args* a, void* stream) { return launch_fwd(a, OP_DOWN, (cudaStream_t)stream); }
extern "C" int mrssm_tc_conv_up(const mrssm_tc_conv_args* a, void* stream) { return launch_fwd(a, OP_UP, (cudaStream_t)stream); }

extern "C" int mrssm_tc_conv_wgrad(const mrssm_tc_conv_args* a, void* stream) {
    MRSSM_CHECK(a && a->large.ptr && a->small.ptr && a->dweight, "tc_wgrad: null tensor");
    WgK k;
    k.n_img = a->n_img; k.Hl = a->Hl; k.Wl = a->Wl; k.Hs = a->Hs; k.Ws = a->Ws; k.ksz = a->ksz;
    k.Cs = a->Cs; k.Cl = a->Cl; k.Cs_valid = a->cs_valid; k.Cl_valid = a->cl_valid;
    MRSSM_CHECK(k.Cs % 8 == 0 && k.Cl % 8 == 0 && a->small.sC == 1 && a->large.sC == 1, "tc_wgrad: tensors must be bf16 NHWC padded to 8");
    k.Ncols = a->ksz * a->ksz * k.Cl;
    k.BN = k.Ncols >= 256 ? 256 : (int)(ceil_div64(k.Ncols, 64) * 64);
    k.P = (long long)a->n_img * a->Hs * a->Ws;
    long long tiles = ceil_div64(k.Cs, BM) * ceil_div64(k.Ncols, k.BN);
    long long want = std::max<long long>(1, ceil_div64(2 * 148, tiles));
    long long splits = std::min<long long>(want, std::max<long long>(1, k.P / (8 * BK)));
    k.pchunk = ceil_div64(ceil_div64(k.P, splits), BK) * BK;
    splits = ceil_div64(k.P, k.pchunk);
    k.small = cvt(a->small); k.large = cvt(a->large);
    k.dw = a->dweight; k.w_ss = a->w_ss; k.w_sl = a->w_sl;
    uint32_t stage = 2 * 8192 + (k.BN / 64) * 8192;
    k.stages = (int)std::max<uint32_t>(LAG + 1, std::min<uint32_t>(4, (200 * 1024) / stage));   // needs > LAG stages
    k.tmem_cols = pow2_cols(k.BN);
    size_t smem = (size_t)k.stages * stage + 1024;
    dim3 grid((unsigned)ceil_div64(k.Cs, BM), (unsigned)ceil_div64(k.Ncols, k.BN), (unsigned)splits);
    CUtensorMap mS, mL;
    memset(&mS, 0, sizeof(mS));
    memset(&mL, 0, sizeof(mL));
    k.tma = 0;
    const bool dense = a->Hs == 1 && a->Ws == 1 && a->Hl == a->ksz && a->Wl == a->ksz &&
                       (a->ksz == 1 || (a->large.sW == k.Cl && a->large.sH == (long long)a->ksz * k.Cl)) && a->small.sI % 8 == 0 &&
                       a->large.sI % 8 == 0 && ((uintptr_t)a->small.ptr & 15) == 0 && ((uintptr_t)a->large.ptr & 15) == 0 &&
                       a->small.sI >= k.Cs && a->large.sI >= k.Ncols && a->n_img >= 64;
    if (dense) {
        if (int rc = mrssm_tma_map_2d_sw128(&mS, a->small.ptr, k.Cs, a->n_img, a->small.sI * 2, 64, 64)) return rc;
        if (int rc = mrssm_tma_map_2d_sw128(&mL, a->large.ptr, k.Ncols, a->n_img, a->large.sI * 2, 64, 64)) return rc;
        k.tma = 1;
    }
    MRSSM_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_wgrad_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(mS, mL, k);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
