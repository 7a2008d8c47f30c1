// Dense layers on tcgen05 with both operands fed by TMA: C[M][N] = act(A[M][K] * W[N][K]^T + bias) (* act'-mask).
// Used for nn.Linear forward / dgrad and for the ConvTranspose2d on the 1x1 map (a GEMM with N = k*k*Cout) of
// observation_model.py:99-102,65-66 and the hoisted embedding half of encoder.py:172-176.
// Persistent CTAs over (row tile, column tile) with the column tile fastest (the A tile is re-read from L2), a ring of
// K blocks of 64 (A: 128 rows x 128 B, W: BN rows x 128 B, 128B swizzle), one tcgen05.mma issuer, two accumulator sets
// in TMEM, 16 epilogue warps.
#include <cuda.h>
#include <algorithm>
#include "tc_common.cuh"

int mrssm_tma_map_2d_sw128(void* map, const void* base, long long inner, long long outer, long long row_bytes, int box_inner, int box_outer);

namespace {

using bf16 = __nv_bfloat16;
constexpr int DT = 704, D_MMA_WARP = 21, D_EPI = 16;

struct DenseP {
    int M, K, Npad, n_valid, BN, n_ntiles, n_mtiles, nkb, NS;
    int act, mask_mode, out_f32, bias_mod;
    int bias_period;                      // entries of the shared-memory bias table: column n reads entry n % bias_period
    long long ldc, cstride, ldm;          // output row / column stride (elements), mask row stride
    void* out;
    const bf16* mask;
    const float* bias;
    const float* addend;                  // fp32 [M][lda_add] added before the activation, or NULL
    long long lda_add;
    int group_n, group_k;                 // block-diagonal GEMM: column tile -> first A column it contracts; 0 = plain
    int vec4;                             // fp32 output rows allow 16-byte stores
};

__device__ __forceinline__ void d_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void d_tma_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void d_umma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void d_tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(DT, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mW, const DenseP P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full[8], empty[8], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[4096];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_bytes = 128u * 128u, b_bytes = (uint32_t)P.BN * 128u, stage = a_bytes + b_bytes;
    const int n_tiles = P.n_mtiles * P.n_ntiles;

    for (int c = tid; c < P.bias_period; c += DT) bias_s[c] = (P.bias && c < P.n_valid) ? P.bias[c % P.bias_mod] : 0.f;
    if (tid == 0) {
        for (int s = 0; s < 8; ++s) {
            tc::mbar_init(tc::smem_u32(&full[s]), 1);
            tc::mbar_init(tc::smem_u32(&empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(tc::smem_u32(&acc_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&acc_empty[s]), D_EPI);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int mt = tile / P.n_ntiles, nt = tile - mt * P.n_ntiles;
                const int koff = P.group_n ? (nt * P.BN) / P.group_n * P.group_k : 0;
                for (int kb = 0; kb < P.nkb; ++kb, ++cnt) {
                    const int s = cnt % P.NS;
                    tc::mbar_wait(tc::smem_u32(&empty[s]), ((cnt / P.NS) & 1) ^ 1);
                    const uint32_t bar = tc::smem_u32(&full[s]);
                    d_expect_tx(bar, stage);
                    d_tma_2d(smem0 + (uint32_t)s * stage, &mA, koff + kb * 64, mt * 128, bar);
                    d_tma_2d(smem0 + (uint32_t)s * stage + a_bytes, &mW, kb * 64, nt * P.BN, bar);
                }
            }
        }
        __syncwarp();
    } else if (warp == D_MMA_WARP) {
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_bf16(128, P.BN, 0, 0);
            const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024 B, version 1, 128B swizzle
            uint32_t cnt = 0, ccnt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ccnt) {
                const int set = ccnt & 1;
                tc::mbar_wait(tc::smem_u32(&acc_empty[set]), ((ccnt >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(set * 256);
                for (int kb = 0; kb < P.nkb; ++kb, ++cnt) {
                    const int s = cnt % P.NS;
                    tc::mbar_wait(tc::smem_u32(&full[s]), (cnt / P.NS) & 1);
                    tc::tc_fence_after();
                    const uint32_t sA = smem0 + (uint32_t)s * stage;
                    const uint32_t a_lo = ((sA >> 4) & 0x3FFFu) | (1u << 16), b_lo = (((sA + a_bytes) >> 4) & 0x3FFFu) | (1u << 16);
                    d_umma(d, a_lo, b_lo, hi, idesc, kb != 0);
                    d_umma(d, a_lo + 2, b_lo + 2, hi, idesc, 1);
                    d_umma(d, a_lo + 4, b_lo + 4, hi, idesc, 1);
                    d_umma(d, a_lo + 6, b_lo + 6, hi, idesc, 1);
                    tc::umma_commit(tc::smem_u32(&empty[s]));
                }
                tc::umma_commit(tc::smem_u32(&acc_full[set]));
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + D_EPI) {
        const int e = warp - 4, q = e & 3, part = e >> 2;
        const int ncp = (P.BN % 64 == 0) ? 4 : ((P.BN % 32 == 0) ? 2 : 1);
        const int ncols = P.BN / ncp;
        uint32_t ccnt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ccnt) {
            const int mt = tile / P.n_ntiles, nt = tile - mt * P.n_ntiles;
            const int set = ccnt & 1;
            tc::mbar_wait(tc::smem_u32(&acc_full[set]), (ccnt >> 1) & 1);
            tc::tc_fence_after();
            const int row = mt * 128 + q * 32 + lane;
            const bool row_ok = row < P.M;
            for (int cp = part; cp < ncp; cp += D_EPI / 4) {
                const int col0 = cp * ncols;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * 256 + col0);
                for (int c0 = 0; c0 < ncols; c0 += 16) {
                    float v[16];
                    d_tmem_ld16(taddr + c0, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (!row_ok) continue;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int n = nt * P.BN + col0 + c0 + 8 * h;
                        if (n >= P.Npad) continue;
                        const int nb = n % P.bias_period;
                        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + nb), b1 = *reinterpret_cast<const float4*>(bias_s + nb + 4);
                        float x[8] = {v[8 * h] + b0.x, v[8 * h + 1] + b0.y, v[8 * h + 2] + b0.z, v[8 * h + 3] + b0.w,
                                      v[8 * h + 4] + b1.x, v[8 * h + 5] + b1.y, v[8 * h + 6] + b1.z, v[8 * h + 7] + b1.w};
                        if (P.addend) {
                            const float* ap = P.addend + (long long)row * P.lda_add + n;
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (n + i < P.n_valid) x[i] += __ldg(ap + i);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) x[i] = act_apply(x[i], P.act);
                        if (P.mask_mode) {
                            const uint4 mk = __ldg(reinterpret_cast<const uint4*>(P.mask + (long long)row * P.ldm + n));
                            const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mk);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 m = __bfloat1622float2(mh[i]);
                                x[2 * i] *= act_grad_from_out(m.x, P.mask_mode);
                                x[2 * i + 1] *= act_grad_from_out(m.y, P.mask_mode);
                            }
                        }
                        if (P.out_f32) {
                            float* op = (float*)P.out + (long long)row * P.ldc;
                            if (P.vec4 && n + 8 <= P.n_valid) {
                                *reinterpret_cast<float4*>(op + n) = make_float4(x[0], x[1], x[2], x[3]);
                                *reinterpret_cast<float4*>(op + n + 4) = make_float4(x[4], x[5], x[6], x[7]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (n + i < P.n_valid) op[(long long)(n + i) * P.cstride] = x[i];
                            }
                        } else {
                            uint4 pk;
                            __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                ph[i] = __floats2bfloat162_rn(n + 2 * i < P.n_valid ? x[2 * i] : 0.f, n + 2 * i + 1 < P.n_valid ? x[2 * i + 1] : 0.f);
                            *reinterpret_cast<uint4*>((bf16*)P.out + (long long)row * P.ldc + n) = pk;
                        }
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[set]));
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

// A: bf16 [M][K] with row stride lda (elements, multiple of 8); wpacked: bf16 [Npad][Kpad] (Kpad multiple of 64, zero padded)
int dense_tc_launch(const void* A, long long lda, int M, int K, const void* wpacked, int Npad, int Kpad, int n_valid, const float* bias,
                    int bias_mod, int act, const void* mask, long long ldm, int mask_mode, void* out, long long ldc, long long cstride,
                    int out_f32, cudaStream_t st, const float* addend, long long addend_ld, int group_n, int group_k) {
    MRSSM_CHECK(A && wpacked && out && M > 0 && K > 0 && Npad % 16 == 0 && Kpad % 64 == 0 && lda % 8 == 0 && K % 8 == 0,
                "dense_tc: bad arguments (M %d K %d Npad %d Kpad %d lda %lld)", M, K, Npad, Kpad, lda);
    // the bias table holds 4096 entries: all columns, or one period of a periodic bias (ConvTranspose2d on the 1x1 map: n = tap * C + c)
    const bool periodic = bias_mod > 0 && bias_mod % 8 == 0 && bias_mod <= 4096 && n_valid % bias_mod == 0;
    MRSSM_CHECK(Npad <= 4096 || !bias || periodic, "dense_tc: %d output columns exceed the bias table", Npad);
    MRSSM_CHECK(out_f32 || (ldc % 8 == 0 && cstride == 1), "dense_tc: bf16 output rows must be 16-byte aligned");
    DenseP P;
    P.M = M; P.K = K; P.Npad = Npad; P.n_valid = n_valid;
    P.n_mtiles = (M + 127) / 128;
    P.BN = Npad <= 256 ? Npad : (Npad % 256 == 0 ? 256 : (Npad % 128 == 0 ? 128 : 64));
    // few row tiles (the per-step rollout GEMMs: M = batch): narrower column tiles put the K loop on more SMs
    while (P.BN > 64 && P.BN % 2 == 0 && (P.BN / 2) % 16 == 0 && Npad % (P.BN / 2) == 0 && P.n_mtiles * ((Npad + P.BN - 1) / P.BN) < 74) P.BN /= 2;
    P.group_n = group_n > 0 ? group_n : 0; P.group_k = group_k;
    if (P.group_n) {
        MRSSM_CHECK(group_k > 0 && group_n % 16 == 0 && Npad % group_n == 0, "dense_tc: bad grouping (%d columns over %d inputs)", group_n, group_k);
        while (group_n % P.BN != 0 && P.BN > 16) P.BN = (P.BN % 32 == 0) ? P.BN / 2 : 16;
        MRSSM_CHECK(group_n % P.BN == 0 && Npad % P.BN == 0, "dense_tc: group of %d columns not tileable", group_n);
    }
    MRSSM_CHECK(Npad % P.BN == 0 || Npad > 256, "dense_tc: %d columns not tileable", Npad);
    P.n_ntiles = (Npad + P.BN - 1) / P.BN;
    P.nkb = Kpad / 64;
    const int stage = 128 * 128 + P.BN * 128;
    P.NS = std::max(2, std::min(8, (200 * 1024) / stage));
    P.act = act; P.mask_mode = mask ? mask_mode : 0; P.out_f32 = out_f32; P.bias_mod = bias_mod > 0 ? bias_mod : std::max(1, n_valid);
    P.bias_period = (Npad > 4096 && periodic) ? bias_mod : std::min(Npad, 4096);
    P.ldc = ldc; P.cstride = cstride; P.ldm = ldm;
    P.out = out; P.mask = (const bf16*)mask; P.bias = bias;
    P.addend = addend; P.lda_add = addend_ld;
    P.vec4 = out_f32 && cstride == 1 && ldc % 4 == 0 && ((uintptr_t)out & 15) == 0;
    CUtensorMap mA, mW;
    if (int rc = mrssm_tma_map_2d_sw128(&mA, A, K, M, lda * 2, 64, 128)) return rc;
    if (int rc = mrssm_tma_map_2d_sw128(&mW, wpacked, Kpad, Npad, (long long)Kpad * 2, 64, P.BN)) return rc;
    const size_t smem = (size_t)P.NS * stage + 1024;
    const int grid = std::min(P.n_mtiles * P.n_ntiles, 148);
    // (set once: the call takes a driver lock, and the per-step rollout launches this kernel ~1000 times per train step)
    constexpr size_t kMaxSmem = 200 * 1024 + 1024;          // NS * stage <= 200 KB by construction of NS
    static bool attr_set = false;
    if (!attr_set) {
        MRSSM_CUDA(cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
        attr_set = true;
    }
    MRSSM_CHECK(smem <= kMaxSmem, "dense_tc: %zu bytes of shared memory", smem);
    dense_tc_kernel<<<grid, DT, std::max<size_t>(smem, 120 * 1024), st>>>(mA, mW, P);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
