// "Plane" convolution kernels for sm_100a: stride-2 Conv2d / ConvTranspose2d forward, dgrad and wgrad on
// tcgen05 tensor cores with the activation tile loaded ONCE into shared memory by TMA.
//
// Idea.  A stride-2 convolution over the space-to-depth (2x2 parity) view of its input is a stride-1
// correlation, and a stride-2 transposed convolution is, per output-parity class, a stride-1 correlation
// over the zero-padded input.  For a stride-1 correlation the im2col matrix of tap (a,b) is the activation
// itself shifted by a*pitch+b pixels.  So the activation tile is put in shared memory as channel-chunk
// planes  plane[chunk][pixel][8 channels]  (16 bytes per pixel, pixel index linear over the padded tile) —
// exactly the un-swizzled UMMA operand layout (core matrix = 8 pixels x 16 bytes, SBO = 128 B, LBO = plane
// stride) — and every tap is the SAME planes addressed through a descriptor whose start address is advanced
// by shift*16 bytes.  im2col is never materialised, not even in shared memory; TMA does the parity split
// (one tensor map per parity with doubled strides) and the zero padding (out-of-bounds fill).
//
//   fwd-type (down / up):  D[128 pixels][BN] += A(shifted planes, K-major) * W[BN][64]   (W streamed by TMA,
//       128B-swizzled ring; accumulators for several 128-pixel blocks live in TMEM, two sets ping-pong
//       between the MMA issuer and the 4 epilogue warps: bias, activation, act'-mask, bf16 NHWC or fp32 strided)
//   wgrad:  D[128 cs][2*Cl] += small^T (MN-major planes) * large parity planes shifted by the tap (MN-major);
//       one accumulator per (kh, kw/2) tap group, contraction over pixels, accumulated across all tiles of
//       a CTA in TMEM, then fp32 atomics into the PyTorch-layout master gradient.
#include <cuda.h>
#include <algorithm>
#include <string.h>
#include "tc_common.cuh"

namespace {

using bf16 = __nv_bfloat16;
enum { OP_DOWN = 0, OP_UP = 1 };
constexpr int NTHREADS = 256;
constexpr int SMEM_TOTAL = 227 * 1024 - 9216;   // dynamic budget: 227 KB minus the 1 KB alignment slack and the static barriers/tables

long long* g_prof = nullptr;    // optional device buffer for in-kernel clock64 timelines (bring-up / tuning)
// tuning aid (mrssm_pl_set_plan_override): force the tiling of the forward-type planner.  0 / -1 = let the planner choose.
struct PlanOvr {
    int BI, TH, NA, bres, NB;
};
int g_sm_budget = 148;       // SMs a persistent plane kernel may occupy (mrssm_pl_set_sm_budget: kernels sharing the GPU with the rollout)
PlanOvr g_ovr = {0, 0, 0, -1, 0};
int g_dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // bring-up switches: 0 barrier polling, 1 no fast epilogue, 2 no L2 prefetch, 3 no column-tap
                                            // replication (wgrad), 4 no grouped TMA boxes

struct T4 {
    const void* p;
    long long sI, sH, sW, sC;
};
inline T4 cvt(const mrssm_t4& t) { return T4{t.ptr, t.sI, t.sH, t.sW, t.sC}; }

// bf16 activation view with channels in chunks of 8 (mrssm_tv): linear (NHWC: sW = C, sK = 8; planar: sW = 8, sK = H*W*8)
// or parity-planar (par): pixel (y,x) lives in plane (y&1)*2+(x&1) at row y>>1, column x>>1.
struct TV {
    const void* p;
    long long sI, sH, sW, sK, sP;
    int par;
};
inline TV cvt(const mrssm_tv& t) { return TV{t.ptr, t.sI, t.sH, t.sW, t.sK, t.sP, t.par}; }
// element offset of channel 0 of pixel (y, x) of image img
__device__ __forceinline__ long long tv_pix(const TV& t, int img, int y, int x) {
    if (t.par) return img * t.sI + ((y & 1) * 2 + (x & 1)) * t.sP + (y >> 1) * t.sH + (x >> 1) * t.sW;
    return img * t.sI + y * t.sH + x * t.sW;
}

// ---------------------------------------------------------------------------------------------------
// TMA + descriptors
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// ---- converged-warp issue: the producer and the MMA warps run their loops with all 32 lanes (so every operand lives in the
// uniform datapath and a TMA / MMA is ONE warp-level instruction, not an ELECT retry loop per instruction); the asynchronous
// instruction itself is guarded by the predicate of the lane chosen once by elect.sync.
// Barrier wait of a converged warp.  mode 0: every lane polls; mode 1: only the elected lane polls and the result is spread with
// a vote (warp-uniform by construction, so the code behind it stays in the uniform datapath) — 32 pollers per warp load the
// barrier unit that also has to deliver the TMA / MMA completions.
__device__ __forceinline__ void mbar_wait_conv(uint32_t bar, uint32_t parity, uint32_t lead, int mode) {
    if (mode == 0) {
        tc::mbar_wait(bar, parity);
        return;
    }
    for (uint32_t it = 0;; ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p, e;\n\t"
            "setp.ne.b32 e, %3, 0;\n\t"
            "setp.ne.b32 p, 0, 0;\n\t"
            "@e mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(lead)
            : "memory");
        if (__any_sync(0xffffffffu, done != 0)) return;
        if (it > (1u << 24)) {
            if (lead) printf("mrssm plane: mbarrier wait timeout (block %d,%d warp %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x >> 5, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ uint32_t elect_one_lane() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mbar_expect_tx_if(uint32_t bar, uint32_t bytes, uint32_t lead) {
    asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %2, 0;\n\t@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes),
                 "r"(lead)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d_if(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t bar, uint32_t lead) {
    asm volatile(
        "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %7, 0;\n\t"
        "@e cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n\t}"
        ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar), "r"(lead)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_if(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar, uint32_t lead) {
    asm volatile(
        "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %5, 0;\n\t"
        "@e cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
        ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar), "r"(lead)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d_if(const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t lead) {
    asm volatile(
        "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %5, 0;\n\t"
        "@e cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];\n\t}"
        ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(lead)
        : "memory");
}
__device__ __forceinline__ void prefetch_l2_if(const void* p, uint32_t bytes, uint32_t lead) {
    asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %2, 0;\n\t@e cp.async.bulk.prefetch.L2.global [%0], %1;\n\t}" ::"l"(p), "r"(bytes), "r"(lead)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t bar, uint32_t lead) {
    asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %1, 0;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar),
                 "r"(lead)
                 : "memory");
}
// un-swizzled (INTERLEAVE) operand descriptor.
//   K-major : core matrix = 8 rows(M/N) x 16 B(K); SBO = byte stride between 8-row groups, LBO = between the two 8-element K chunks
//   MN-major: core matrix = 8 rows(K)   x 16 B(MN); LBO = byte stride between 8-row K groups, SBO = between 8-element MN chunks
__device__ __forceinline__ uint64_t smem_desc_plain(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// activation map, element form: bf16 [N][Y][X][C] with byte strides, box {8, bx, by, bi}, no swizzle, zero OOB fill
int make_act_map(CUtensorMap* m, const void* base, long long C, long long X, long long Y, long long N, long long sx, long long sy,
                 long long sn, int bx, int by, int bi) {
    EncodeTiledFn enc = get_encode();
    MRSSM_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sx, (cuuint64_t)sy, (cuuint64_t)sn};
    cuuint32_t box[4] = {8, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bi};
    cuuint32_t es[4] = {1, 1, 1, 1};
    MRSSM_CHECK(X >= 1 && Y >= 1 && N >= 1 && bx <= 256 && by <= 256 && bi <= 256 && sx % 16 == 0 && sy % 16 == 0 && sn % 16 == 0 &&
                    ((uintptr_t)base & 15) == 0,
                "plane conv: activation tensor not TMA-addressable (dims %lld %lld %lld %lld strides %lld %lld %lld box %d %d %d)", C, X, Y,
                N, sx, sy, sn, bx, by, bi);
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MRSSM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: %d (dims %lld %lld %lld %lld strides %lld %lld %lld box %d %d %d)",
                (int)r, C, X, Y, N, sx, sy, sn, bx, by, bi);
    return 0;
}
// activation map, merged form for x-contiguous planes: the 16-byte pixel-chunk is two 64-bit elements, so one box row is a
// whole run of bx pixels (bx*16 contiguous bytes).  dims {2X, Y, chunks, N}, box {2bx, by, 1, bi}.
int make_act_map_merged(CUtensorMap* m, const void* base, long long X, long long Y, long long chunks, long long N, long long sy,
                        long long sk, long long sn, int bx, int by, int bi) {
    EncodeTiledFn enc = get_encode();
    MRSSM_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)(2 * X), (cuuint64_t)Y, (cuuint64_t)chunks, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sy, (cuuint64_t)sk, (cuuint64_t)sn};
    cuuint32_t box[4] = {(cuuint32_t)(2 * bx), (cuuint32_t)by, 1, (cuuint32_t)bi};
    cuuint32_t es[4] = {1, 1, 1, 1};
    MRSSM_CHECK(X >= 1 && Y >= 1 && N >= 1 && chunks >= 1 && 2 * bx <= 256 && by <= 256 && bi <= 256 && sy % 16 == 0 && sk % 16 == 0 &&
                    sn % 16 == 0 && ((uintptr_t)base & 15) == 0,
                "plane conv: planar tensor not TMA-addressable (dims %lld %lld %lld %lld strides %lld %lld %lld box %d %d %d)", X, Y, chunks,
                N, sy, sk, sn, bx, by, bi);
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MRSSM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(planar) failed: %d (dims %lld %lld %lld %lld strides %lld %lld %lld box %d %d %d)",
                (int)r, X, Y, chunks, N, sy, sk, sn, bx, by, bi);
    return 0;
}
// grouped form for x-contiguous planes: ONE box covers every chunk plane of the map.  dims {2X, Y, N, chunks} — the image dimension
// before the chunk dimension, whatever their strides — box {2bx, by, bi, chunks}: the box lands in shared memory as
// [chunk][img][y][x], i.e. the chunk planes the MMA descriptors walk, back to back (plane stride = bi*by*bx*16 bytes).  A TMA box
// costs the TMA unit ~200 cycles whatever its size (profiles/micro/tma_rate.cu), so 4-64 plane boxes per tile were the
// bottleneck of most layers.
int make_act_map_grouped(CUtensorMap* m, const void* base, long long X, long long Y, long long chunks, long long N, long long sy,
                         long long sk, long long sn, int bx, int by, int bi, int bch) {
    EncodeTiledFn enc = get_encode();
    MRSSM_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)(2 * X), (cuuint64_t)Y, (cuuint64_t)N, (cuuint64_t)chunks};
    cuuint64_t strides[3] = {(cuuint64_t)sy, (cuuint64_t)sn, (cuuint64_t)sk};
    cuuint32_t box[4] = {(cuuint32_t)(2 * bx), (cuuint32_t)by, (cuuint32_t)bi, (cuuint32_t)bch};
    cuuint32_t es[4] = {1, 1, 1, 1};
    MRSSM_CHECK(X >= 1 && Y >= 1 && N >= 1 && chunks >= 1 && 2 * bx <= 256 && by <= 256 && bi <= 256 && bch <= 256 && sy % 16 == 0 && sk % 16 == 0 &&
                    sn % 16 == 0 && ((uintptr_t)base & 15) == 0,
                "plane conv: planar tensor not TMA-addressable (grouped; dims %lld %lld %lld %lld strides %lld %lld %lld box %d %d %d %d)", X, Y, N,
                chunks, sy, sn, sk, bx, by, bi, bch);
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MRSSM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(grouped planar) failed: %d (dims %lld %lld %lld %lld strides %lld %lld %lld box %d %d %d %d)",
                (int)r, X, Y, N, chunks, sy, sn, sk, bx, by, bi, bch);
    return 0;
}
// can the view be loaded with grouped boxes?  (x-contiguous planes: planar or parity-planar)
inline bool view_groupable(const mrssm_tv& t, bool parity_split) { return t.sW == 8 && (parity_split ? t.par != 0 : t.par == 0); }

// Tensor maps of a source view [N][H][W][Cp]: one map (whole tensor) or four (its 2x2 parity sub-grids).  *merged reports
// the coordinate convention the kernel must use: merged (2*x, y, chunk, img) or element (chunk*8, x, y, img).
int make_view_maps(CUtensorMap* maps, int* merged, const mrssm_tv& t, int H, int W, int Cp, int N, bool parity_split, int bx, int by, int bi,
                   int group_chunks = 0) {
    const char* base = (const char*)t.ptr;
    if (group_chunks > 0) {                 // *merged = 2: coordinates (2*x, y, img, first chunk)
        MRSSM_CHECK(view_groupable(t, parity_split), "plane conv: grouped boxes need x-contiguous planes");
        *merged = 2;
        if (!parity_split) {
            if (int rc = make_act_map_grouped(&maps[0], base, W, H, Cp / 8, N, 2 * t.sH, 2 * t.sK, 2 * t.sI, bx, by, bi, group_chunks)) return rc;
            maps[1] = maps[2] = maps[3] = maps[0];
            return 0;
        }
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                const long long X = std::max(1, (W - px + 1) / 2), Y = std::max(1, (H - py + 1) / 2);
                if (int rc = make_act_map_grouped(&maps[py * 2 + px], base + 2 * (py * 2 + px) * t.sP, X, Y, Cp / 8, N, 2 * t.sH, 2 * t.sK, 2 * t.sI,
                                                  bx, by, bi, group_chunks))
                    return rc;
            }
        return 0;
    }
    if (!parity_split) {
        MRSSM_CHECK(!t.par, "plane conv: this operand must be a linear (NHWC or planar) view, not parity-planar");
        if (t.sW == 8 && 2 * bx <= 256) {
            *merged = 1;
            if (int rc = make_act_map_merged(&maps[0], base, W, H, Cp / 8, N, 2 * t.sH, 2 * t.sK, 2 * t.sI, bx, by, bi)) return rc;
        } else {
            MRSSM_CHECK(t.sK == 8, "plane conv: a pixel-strided view must keep its channels contiguous (sK == 8)");
            *merged = 0;
            if (int rc = make_act_map(&maps[0], base, Cp, W, H, N, 2 * t.sW, 2 * t.sH, 2 * t.sI, bx, by, bi)) return rc;
        }
        maps[1] = maps[2] = maps[3] = maps[0];
        return 0;
    }
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const long long X = std::max(1, (W - px + 1) / 2), Y = std::max(1, (H - py + 1) / 2);
            if (t.par) {
                MRSSM_CHECK(t.sW == 8 && 2 * bx <= 256, "plane conv: parity-planar views must be x-contiguous (sW == 8) with boxes <= 128 pixels");
                *merged = 1;
                if (int rc = make_act_map_merged(&maps[py * 2 + px], base + 2 * (py * 2 + px) * t.sP, X, Y, Cp / 8, N, 2 * t.sH, 2 * t.sK, 2 * t.sI,
                                                 bx, by, bi))
                    return rc;
            } else {
                MRSSM_CHECK(t.sK == 8, "plane conv: a linear view split by parity must be NHWC (sK == 8); use a parity-planar view");
                *merged = 0;
                if (int rc = make_act_map(&maps[py * 2 + px], base + 2 * (py * t.sH + px * t.sW), Cp, X, Y, N, 4 * t.sW, 4 * t.sH, 2 * t.sI, bx,
                                          by, bi))
                    return rc;
            }
        }
    return 0;
}
// weight map: bf16 [N_total][K_total] row-major, box {64, BN}, 128B swizzle
int make_w_map(CUtensorMap* m, const void* base, long long K_total, long long N_total, int BN) {
    EncodeTiledFn enc = get_encode();
    MRSSM_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)K_total, (cuuint64_t)N_total};
    cuuint64_t strides[1] = {(cuuint64_t)K_total * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MRSSM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed: %d (K %lld N %lld BN %d)", (int)r, K_total, N_total, BN);
    return 0;
}

}  // namespace

// bf16 row-major [outer][inner] matrix, box {box_inner, box_outer}, 128B swizzle (used by dense_tc.cu as well)
int mrssm_tma_map_2d_sw128(void* map, const void* base, long long inner, long long outer, long long row_bytes, int box_inner, int box_outer) {
    EncodeTiledFn enc = get_encode();
    MRSSM_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t es[2] = {1, 1};
    MRSSM_CHECK(row_bytes % 16 == 0 && ((uintptr_t)base & 15) == 0 && box_inner * 2 == 128 && box_outer <= 256,
                "tma_map_2d: matrix not TMA-addressable (inner %lld outer %lld row bytes %lld box %d %d)", inner, outer, row_bytes, box_inner, box_outer);
    CUresult r = enc((CUtensorMap*)map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MRSSM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: %d (inner %lld outer %lld row bytes %lld box %d %d)", (int)r, inner, outer,
                row_bytes, box_inner, box_outer);
    return 0;
}

namespace {

// ---------------------------------------------------------------------------------------------------
// forward-type kernel (down / up)
// ---------------------------------------------------------------------------------------------------
struct FwdP {
    int op;
    int n_img, n_groups, n_bands, BI, BX, BY, TH;
    int Hv, Wv;                 // valid output extent in position space (down: Hs,Ws; up: ceil(Hl/2),ceil(Wl/2))
    int planes, ppm;            // A planes per tile, planes per tensor map
    int PS, plane_bytes;        // plane stride / bytes written by TMA per plane
    int x0, y0;                 // box origin (up: -(nt-1))
    int n_ksteps, nkb, nt, J;
    int BN, n_ntiles;
    int MB_total, MBs, n_passes, n_sets, set_cols;
    int NA, NB, b_res;          // b_res: the whole packed weight stays resident in shared memory (NB == nkb)
    int act, mask_mode, out_f32, n_valid, Cop, Ho, Wo;
    int a_stage_bytes;
    int mergedA;                // TMA coordinate convention of the activation maps (2: grouped boxes, one per map)
    int group;                  // one box per tensor map: every chunk plane of the map (planes are dense, PS = plane_bytes)
    int s2d;                    // down: the source is the space-to-depth (16-channel) form of a <= 4-channel image
    long long* prof;            // [CTA][16 tiles][8 slots] clock64 stamps or NULL
    TV out, mask;               // bf16 output / act'-mask views (indexed at the output pixel)
    long long mask_img_bytes;   // > 0: the mask view stores every image as one dense block of this many bytes (L2 prefetch per tile)
    long long tgt_img_bytes;    // > 0: same for the fp32 target images of the fused reconstruction loss
    uint32_t ks_off16[320];     // (plane*PS + shift*16) >> 4 per K-step: read by the MMA issuer from the constant bank (uniform regs)
    T4 out32;                   // fp32 output (F32 kernels)
    const float* bias;
    const float* scale_ptr;     // optional output scale: (*scale_ptr) * scale_mul
    float scale_mul;
    const float* mse_target;    // fused MSE (up, F32): residual vs target -> *mse_sum and bf16 space-to-depth residual in `out`
    float* mse_sum;
    float mse_scale;
    int mse_vec;                // target rows are x-contiguous and 8-byte aligned: float2 loads
    int mse_lean;               // 3 channels, 32 accumulator columns: the lean fused-loss kernel (NC = -1)
    int mse_on;                 // fused reconstruction loss (target: fp32 `mse_target` or the bf16 space-to-depth view `tgt`)
    TV tgt;                     // bf16 space-to-depth target (the encoder's own imported image): two 16-byte loads per position
    // fast epilogue (bf16 output, act none / ReLU, every item inside one parity class): nc columns per item (16, 32 or 64)
    int fast, ncp, nc;
    int poll;                   // barrier polling of the converged warps (mbar_wait_conv)
    uint8_t* bits_out;          // ReLU sign bits of the output, byte (pixel, 8-channel chunk) at [img][y][x][Cop/8]
    const uint8_t* bits_in;     // act'-mask of a dgrad in the same form (instead of the bf16 `mask` view)
    long long bits_img_bytes;   // bytes of one image of bits_in (L2 prefetch per tile)
};
#define PROF(slot) do { if (P.prof && lit < 16 && (threadIdx.x & 31) == 0) P.prof[((long long)blockIdx.x * 16 + lit) * 8 + (slot)] = clock64(); } while (0)

constexpr int FWD_THREADS = 704;      // warp 0 TMA, 2 TMEM alloc, 4..19 epilogue (four per TMEM lane quarter), 21 MMA issuer
constexpr int MMA_WARP = 21;          // the highest warp id of its scheduler: the issue arbiter favours high warp ids
constexpr int EPI_WARPS = 16;

__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_lohi_if(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                  uint32_t accumulate, uint32_t lead) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 e, %7, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(lead)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// bias + activation + act'-mask + store of 8 consecutive output channels of one pixel
template <bool F32>
__device__ __forceinline__ void epi_store8(const FwdP& P, const float* v, const float* bias_s, int cl0, long long o_off, const uint4 mk,
                                           float oscale) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias_s + cl0), b1 = *reinterpret_cast<const float4*>(bias_s + cl0 + 4);
    float x[8] = {v[0] + b0.x, v[1] + b0.y, v[2] + b0.z, v[3] + b0.w, v[4] + b1.x, v[5] + b1.y, v[6] + b1.z, v[7] + b1.w};
    if (P.act == MRSSM_ACT_RELU) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = fmaxf(x[e], 0.f);
    } else if (P.act == MRSSM_ACT_ELU) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = x[e] > 0.f ? x[e] : expm1f(x[e]);
    }
    if (P.mask_mode) {
        const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mk);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 m = __bfloat1622float2(mh[e]);
            x[2 * e] *= act_grad_from_out(m.x, P.mask_mode);
            x[2 * e + 1] *= act_grad_from_out(m.y, P.mask_mode);
        }
    }
    if (P.scale_ptr) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] *= oscale;
    }
    if (F32) {
        float* op = (float*)P.out32.p + o_off;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (cl0 + e < P.n_valid) op[(long long)(cl0 + e) * P.out32.sC] = x[e];
    } else {
        uint4 pk;
        __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) ph[e] = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
        *reinterpret_cast<uint4*>((bf16*)P.out.p + o_off + (cl0 >> 3) * P.out.sK) = pk;
    }
}

// Fused reconstruction loss for one output row of the last ConvTranspose2d (columns = 4 parity classes x 8, NV real
// channels): residual vs target, sum of squares, optional reconstruction store, bf16 space-to-depth residual store.
// NV is a template parameter so that every register-array index is static.
// target values of the four parity classes of position (y, x): tv[class][channel] (issued before the accumulators are waited for)
template <int NV>
__device__ __forceinline__ void mse_load(const FwdP& P, int img, int y, int x, float (&tv)[4][NV]) {
    const float* tg = P.mse_target ? P.mse_target + img * P.out32.sI : nullptr;
    if (P.tgt.p) {
        // space-to-depth target: channel (parity class, c) of position (y, x) — the residual's own layout
        const bf16* tp = (const bf16*)P.tgt.p + tv_pix(P.tgt, img, y, x);
        const uint4 t0 = __ldg(reinterpret_cast<const uint4*>(tp)), t1 = __ldg(reinterpret_cast<const uint4*>(tp + P.tgt.sK));
        float t16[16];
        const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&t0);
        const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&t1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = __bfloat1622float2(h0[e]), b = __bfloat1622float2(h1[e]);
            t16[2 * e] = a.x; t16[2 * e + 1] = a.y; t16[8 + 2 * e] = b.x; t16[8 + 2 * e + 1] = b.y;
        }
#pragma unroll
        for (int cls = 0; cls < 4; ++cls)
#pragma unroll
            for (int c = 0; c < NV; ++c) tv[cls][c] = t16[cls * NV + c];
    } else
#pragma unroll
    for (int py = 0; py < 2; ++py) {
        const int yy = 2 * y + py;
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            tv[py * 2][c] = tv[py * 2 + 1][c] = 0.f;
            if (yy < P.Ho) {
                const float* tp = tg + yy * P.out32.sH + (long long)c * P.out32.sC + 2 * x * P.out32.sW;
                if (P.mse_vec) {           // x-contiguous target, even width: one 8-byte load covers both column parities
                    const float2 t2 = __ldg(reinterpret_cast<const float2*>(tp));
                    tv[py * 2][c] = t2.x;
                    tv[py * 2 + 1][c] = t2.y;
                } else {
                    tv[py * 2][c] = __ldg(tp);
                    if (2 * x + 1 < P.Wo) tv[py * 2 + 1][c] = __ldg(tp + P.out32.sW);
                }
            }
        }
    }
}

template <int NV>
__device__ __forceinline__ void mse_apply(const FwdP& P, const float (&v)[32], const float (&tv)[4][NV], const float* bias_s, int img, int y, int x,
                                          float& mse_acc) {
    float* rc = P.out32.p ? (float*)P.out32.p + img * P.out32.sI : nullptr;
    float r16[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) r16[e] = 0.f;
#pragma unroll
    for (int cls = 0; cls < 4; ++cls) {
        const int yy = 2 * y + (cls >> 1), xx = 2 * x + (cls & 1);
        const bool ok = yy < P.Ho && xx < P.Wo;
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            float val = v[cls * 8 + c] + bias_s[c];
            if (P.act == MRSSM_ACT_RELU) val = fmaxf(val, 0.f);
            else if (P.act == MRSSM_ACT_ELU) val = val > 0.f ? val : expm1f(val);
            const float res = ok ? val - tv[cls][c] : 0.f;
            if (ok && rc) rc[yy * P.out32.sH + xx * P.out32.sW + (long long)c * P.out32.sC] = val;
            mse_acc = fmaf(res, res, mse_acc);
            r16[cls * NV + c] = res;
        }
    }
    bf16* dp = (bf16*)P.out.p + tv_pix(P.out, img, y, x);
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        uint4 pk;
        __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) ph[e] = __floats2bfloat162_rn(r16[ch * 8 + 2 * e], r16[ch * 8 + 2 * e + 1]);
        *reinterpret_cast<uint4*>(dp + ch * P.out.sK) = pk;
    }
}

template <int NV>
__device__ __forceinline__ void mse_row(const FwdP& P, const float (&v)[32], const float* bias_s, int img, int y, int x, float& mse_acc) {
    float tv[4][NV];
    mse_load<NV>(P, img, y, x, tv);
    mse_apply<NV>(P, v, tv, bias_s, img, y, x, mse_acc);
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// prmt with sign replication: every byte of the result is 0xFF / 0x00 from the msb of the selected source byte
__device__ __forceinline__ uint32_t half_masks_from_flags(uint32_t w) {      // bit 15 -> low half, bit 31 -> high half
    uint32_t d;
    asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(d) : "r"(w));
    return d;
}
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) { return (a & m) | (b & ~m); }

// ReLU sign byte of 8 channels held as four packed bf16x2 words (values >= +0): bit (3 - j) = channel 2j > 0, bit (7 - j) =
// channel 2j + 1 > 0.  (x + 0x7FFF per half sets the half's msb iff the half is non-zero.)
__device__ __forceinline__ uint32_t relu_byte(const uint32_t (&p)[4]) {
    uint32_t r = p[0] + 0x7FFF7FFFu;
    r = bitsel(r, (p[1] + 0x7FFF7FFFu) >> 1, 0x80008000u);
    r = bitsel(r, (p[2] + 0x7FFF7FFFu) >> 2, 0xC000C000u);
    r = bitsel(r, (p[3] + 0x7FFF7FFFu) >> 3, 0xE000E000u);
    return ((r >> 12) & 0xFu) | ((r >> 24) & 0xF0u);
}
// keep the halves of the four packed words whose flag bit is set in the sign byte b
__device__ __forceinline__ void apply_relu_byte(uint32_t (&p)[4], uint32_t b) {
    const uint32_t e = (b & 0xFu) | ((b & 0xF0u) << 12);
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] &= half_masks_from_flags(e << (12 + j));
}
// the same from a bf16 activation (act'-mask given as the forward output itself): half > 0
__device__ __forceinline__ void apply_relu_bf16(uint32_t (&p)[4], const uint4 m) {
    const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t nz = (mw[j] & 0x7FFF7FFFu) + 0x7FFF7FFFu;          // msb of a half = magnitude non-zero
        p[j] &= half_masks_from_flags(nz & ~mw[j]);                      // and not negative
    }
}

// One epilogue item of the fast path: 32 rows (this warp's TMEM lanes) x NC accumulator columns that lie inside one
// parity class -> bias, ReLU, act'-mask, scale, bf16 pack, one 16-byte store per 8 channels, optional sign bytes out.
template <int OP, int NC>
__device__ __forceinline__ void epi_fast_item(const FwdP& P, const float* bias_s, uint32_t taddr, bool row_ok, int img, int y, int x, int n0,
                                              float oscale) {
    int cls = 0, cl0 = n0;
    if (OP == OP_UP) {
        cls = n0 / P.Cop;
        cl0 = n0 - cls * P.Cop;
    }
    const int yy = OP == OP_UP ? 2 * y + (cls >> 1) : y, xx = OP == OP_UP ? 2 * x + (cls & 1) : x;
    const bool ok = row_ok && yy < P.Ho && xx < P.Wo;
    bf16* o_ptr = nullptr;
    const bf16* m_ptr = nullptr;
    long long boff = 0;
    if (ok) {
        o_ptr = (bf16*)P.out.p + tv_pix(P.out, img, yy, xx) + (cl0 >> 3) * P.out.sK;
        boff = ((long long)(img * P.Ho + yy) * P.Wo + xx) * (P.Cop >> 3) + (cl0 >> 3);
        if (!P.bits_in && P.mask_mode) m_ptr = (const bf16*)P.mask.p + tv_pix(P.mask, img, yy, xx) + (cl0 >> 3) * P.mask.sK;
    }
    constexpr int W = NC >= 32 ? 32 : 16;          // columns per tcgen05.ld (one sign word of W / 8 bytes each)
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += W) {
        float v[W];
        if (W == 32) tmem_ld32_nowait(taddr + c0, v);
        else tmem_ld16_nowait(taddr + c0, v);
        uint32_t mb = 0xFFFFFFFFu, ob = 0;
        uint4 mk[W / 8];
        if (ok && P.bits_in) {
            if (W == 32) mb = __ldg(reinterpret_cast<const unsigned int*>(P.bits_in + boff + (c0 >> 3)));
            else mb = __ldg(reinterpret_cast<const unsigned short*>(P.bits_in + boff + (c0 >> 3)));
        }
        if (m_ptr) {
#pragma unroll
            for (int j = 0; j < W / 8; ++j) mk[j] = __ldg(reinterpret_cast<const uint4*>(m_ptr + (long long)((c0 >> 3) + j) * P.mask.sK));
        }
        tmem_wait_ld();
        if (ok) {
#pragma unroll
            for (int j = 0; j < W / 8; ++j) {
                const float* vv = v + 8 * j;
                const int cl = cl0 + c0 + 8 * j;
                const float4 b0 = *reinterpret_cast<const float4*>(bias_s + cl), b1 = *reinterpret_cast<const float4*>(bias_s + cl + 4);
                float t[8] = {vv[0] + b0.x, vv[1] + b0.y, vv[2] + b0.z, vv[3] + b0.w, vv[4] + b1.x, vv[5] + b1.y, vv[6] + b1.z, vv[7] + b1.w};
                if (P.act == MRSSM_ACT_RELU) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) t[e] = fmaxf(t[e], 0.f);
                }
                if (P.scale_ptr) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) t[e] *= oscale;
                }
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(t[2 * e], t[2 * e + 1]);
                    pk[e] = *reinterpret_cast<const uint32_t*>(&h);
                }
                if (P.bits_in) apply_relu_byte(pk, (mb >> (8 * j)) & 0xFFu);
                else if (m_ptr) apply_relu_bf16(pk, mk[j]);
                if (P.bits_out) ob |= relu_byte(pk) << (8 * j);
                *reinterpret_cast<uint4*>(o_ptr + (long long)((c0 >> 3) + j) * P.out.sK) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            if (P.bits_out) {
                if (W == 32) *reinterpret_cast<unsigned int*>(P.bits_out + boff + (c0 >> 3)) = ob;
                else *reinterpret_cast<unsigned short*>(P.bits_out + boff + (c0 >> 3)) = (unsigned short)ob;
            }
        }
    }
}

// NC = 0: generic epilogue (fp32 outputs, fused loss, ELU, odd widths); NC = 16 / 32 / 64: the fast bf16 epilogue with NC
// accumulator columns per item (a kernel of its own, so neither path pays for the other's registers)
template <int OP, bool F32, int NC>
__global__ void __launch_bounds__(FWD_THREADS, 1)
plane_fwd_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1, const __grid_constant__ CUtensorMap mA2,
                 const __grid_constant__ CUtensorMap mA3, const __grid_constant__ CUtensorMap mB, const __grid_constant__ FwdP P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[2], a_empty[2], acc_full[2], acc_empty[2];
    __shared__ uint64_t b_full[80], b_empty[8];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[1024];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_stage = (uint32_t)P.BN * 128u;
    const uint32_t smemA = smem0 + (uint32_t)P.NB * b_stage;
    const int n_tiles = P.n_groups * P.n_bands;

    for (int c = tid; c < 1024; c += FWD_THREADS) bias_s[c] = (P.bias && c < P.n_valid) ? P.bias[c] : 0.f;
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(tc::smem_u32(&a_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&a_empty[s]), 1);
            tc::mbar_init(tc::smem_u32(&acc_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&acc_empty[s]), EPI_WARPS);
        }
        for (int s = 0; s < 80; ++s) tc::mbar_init(tc::smem_u32(&b_full[s]), 1);
        for (int s = 0; s < 8; ++s) tc::mbar_init(tc::smem_u32(&b_empty[s]), 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ------------------------------ TMA producer (converged warp, elected lane issues) ------------------------------
        {
            const uint32_t lead = elect_one_lane();
            uint32_t acnt = 0, bcnt = 0;
            auto plane_coords = [&](int tile, int q, const CUtensorMap*& m, int& c0, int& c1, int& c2, int& c3) {
                const int ig = tile / P.n_bands, band = tile - ig * P.n_bands;
                const int mi = q / P.ppm, ch = q - mi * P.ppm;
                m = mi == 0 ? &mA0 : (mi == 1 ? &mA1 : (mi == 2 ? &mA2 : &mA3));
                if (P.mergedA) {
                    c0 = 2 * P.x0; c1 = band * P.TH + P.y0; c2 = ch; c3 = ig * P.BI;
                } else {
                    c0 = ch * 8; c1 = P.x0; c2 = band * P.TH + P.y0; c3 = ig * P.BI;
                }
            };
            auto prefetch_range = [&](const char* base, long long lo_b, long long hi_b) {     // 16-byte granules, 64 KB pieces
                const uintptr_t lo = ((uintptr_t)base + (uintptr_t)lo_b) & ~(uintptr_t)15;
                const uintptr_t hi = ((uintptr_t)base + (uintptr_t)hi_b + 15) & ~(uintptr_t)15;
                const char* mp = (const char*)lo;
                long long left = (long long)(hi - lo);
                while (left > 0) {
                    const uint32_t sz = (uint32_t)(left < 65536 ? left : 65536);
                    prefetch_l2_if(mp, sz, lead);
                    mp += sz;
                    left -= sz;
                }
            };
            auto issue_A = [&](int tile) {
                const int sa = acnt % P.NA;
                mbar_wait_conv(tc::smem_u32(&a_empty[sa]), ((acnt / P.NA) & 1) ^ 1, lead, P.poll);
                const uint32_t bar = tc::smem_u32(&a_full[sa]);
                mbar_expect_tx_if(bar, (uint32_t)P.planes * (uint32_t)P.plane_bytes, lead);
                const int ig = tile / P.n_bands, band = tile - ig * P.n_bands;
                const uint32_t dst = smemA + (uint32_t)sa * (uint32_t)P.a_stage_bytes;
                if (P.group) {          // one box per map: all of its chunk planes, [chunk][img][y][x]
                    const uint32_t map_bytes = (uint32_t)P.ppm * (uint32_t)P.PS;
                    const int cy = band * P.TH + P.y0, ci = ig * P.BI;
                    tma_load_4d_if(dst, &mA0, 2 * P.x0, cy, ci, 0, bar, lead);
                    if (P.planes > P.ppm) {
                        tma_load_4d_if(dst + map_bytes, &mA1, 2 * P.x0, cy, ci, 0, bar, lead);
                        tma_load_4d_if(dst + 2 * map_bytes, &mA2, 2 * P.x0, cy, ci, 0, bar, lead);
                        tma_load_4d_if(dst + 3 * map_bytes, &mA3, 2 * P.x0, cy, ci, 0, bar, lead);
                    }
                } else {
                    for (int q = 0; q < P.planes; ++q) {
                        const CUtensorMap* m;
                        int c0, c1, c2, c3;
                        plane_coords(tile, q, m, c0, c1, c2, c3);
                        tma_load_4d_if(dst + (uint32_t)q * (uint32_t)P.PS, m, c0, c1, c2, c3, bar, lead);
                    }
                }
                if (band == 0) {
                    // what the epilogue of this tile reads straight from global memory (sign bytes, act'-mask, loss target): pull
                    // the tile's images into L2 now, a tile ahead of their use
                    const int i0 = ig * P.BI, ni = min(P.BI, P.n_img - i0);
                    if (P.bits_img_bytes > 0) prefetch_range((const char*)P.bits_in, (long long)i0 * P.bits_img_bytes, (long long)(i0 + ni) * P.bits_img_bytes);
                    if (P.mask_img_bytes > 0) prefetch_range((const char*)P.mask.p, (long long)i0 * P.mask_img_bytes, (long long)(i0 + ni) * P.mask_img_bytes);
                    if (P.tgt_img_bytes > 0) prefetch_range(P.tgt.p ? (const char*)P.tgt.p : (const char*)P.mse_target, (long long)i0 * P.tgt_img_bytes, (long long)(i0 + ni) * P.tgt_img_bytes);
                }
                ++acnt;
            };
            // single-buffered activation stage: the load of the next tile can only start when this tile's MMAs are done, so its
            // planes are pulled into L2 meanwhile (the TMA load then hits L2 instead of HBM)
            auto prefetch_A = [&](int tile) {
                if (P.group) {
                    const int ig = tile / P.n_bands, band = tile - ig * P.n_bands;
                    const int cy = band * P.TH + P.y0, ci = ig * P.BI;
                    tma_prefetch_4d_if(&mA0, 2 * P.x0, cy, ci, 0, lead);
                    if (P.planes > P.ppm) {
                        tma_prefetch_4d_if(&mA1, 2 * P.x0, cy, ci, 0, lead);
                        tma_prefetch_4d_if(&mA2, 2 * P.x0, cy, ci, 0, lead);
                        tma_prefetch_4d_if(&mA3, 2 * P.x0, cy, ci, 0, lead);
                    }
                    return;
                }
                for (int q = 0; q < P.planes; ++q) {
                    const CUtensorMap* m;
                    int c0, c1, c2, c3;
                    plane_coords(tile, q, m, c0, c1, c2, c3);
                    tma_prefetch_4d_if(m, c0, c1, c2, c3, lead);
                }
            };
            int tile = blockIdx.x;
            int lit = 0;
            PROF(0);
            if (P.b_res) {
                for (int kb = 0; kb < P.nkb; ++kb) {
                    const uint32_t bar = tc::smem_u32(&b_full[kb]);
                    mbar_expect_tx_if(bar, b_stage, lead);
                    tma_load_2d_if(smem0 + (uint32_t)kb * b_stage, &mB, kb * 64, 0, bar, lead);
                }
            }
            if (tile < n_tiles) issue_A(tile);
            for (; tile < n_tiles; tile += gridDim.x, ++lit) {
                const int next = tile + gridDim.x;
                PROF(1);
                if (P.NA > 1 && next < n_tiles) issue_A(next);
                if (P.NA == 1 && next < n_tiles) prefetch_A(next);
                if (!P.b_res) {
                    for (int it = 0; it < P.n_ntiles * P.n_passes; ++it) {
                        const int nti = it / P.n_passes;
                        for (int kb = 0; kb < P.nkb; ++kb) {
                            const int sb = bcnt % P.NB;
                            mbar_wait_conv(tc::smem_u32(&b_empty[sb]), ((bcnt / P.NB) & 1) ^ 1, lead, P.poll);
                            const uint32_t bar = tc::smem_u32(&b_full[sb]);
                            mbar_expect_tx_if(bar, b_stage, lead);
                            tma_load_2d_if(smem0 + (uint32_t)sb * b_stage, &mB, kb * 64, nti * P.BN, bar, lead);
                            ++bcnt;
                        }
                    }
                }
                if (P.NA == 1 && next < n_tiles) issue_A(next);
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // ------------------------------ MMA issuer (converged warp, elected lane issues) ------------------------------
        {
            const uint32_t lead = elect_one_lane();
            const uint32_t idesc = tc::idesc_bf16(128, P.BN, 0, 0);
            const uint32_t a_hi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1, no swizzle
            const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, 128B swizzle
            const uint32_t a_lbo = ((uint32_t)P.PS >> 4) << 16;
            const uint32_t BN = (uint32_t)P.BN;
            uint32_t acnt = 0, bcnt = 0, ccnt = 0;
            int lit = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lit) {
                const int sa = acnt % P.NA;
                mbar_wait_conv(tc::smem_u32(&a_full[sa]), (acnt / P.NA) & 1, lead, P.poll);
                tc::tc_fence_after();
                PROF(2);
                const uint32_t a_base = (((smemA + (uint32_t)sa * (uint32_t)P.a_stage_bytes) >> 4) & 0x3FFFu) | a_lbo;
                for (int it = 0; it < P.n_ntiles * P.n_passes; ++it) {
                    const int pass = it % P.n_passes;
                    const int set = ccnt % P.n_sets;
                    mbar_wait_conv(tc::smem_u32(&acc_empty[set]), ((ccnt / P.n_sets) & 1) ^ 1, lead, P.poll);
                    tc::tc_fence_after();
                    if (it == 0) PROF(6);
                    const int mb0 = pass * P.MBs, nmb = min(P.MBs, P.MB_total - mb0);
                    const uint32_t tacc = tmem_base + (uint32_t)(set * P.set_cols);
                    const uint32_t a_pass = a_base + (uint32_t)mb0 * 128u;
                    if (P.b_res) {
                        // whole weight resident: one accumulator at a time, all of its K steps back to back
                        if (lit == 0 && it == 0) {
                            for (int kb = 0; kb < P.nkb; ++kb) mbar_wait_conv(tc::smem_u32(&b_full[kb]), 0, lead, P.poll);
                            tc::tc_fence_after();
                        }
                        const uint32_t b_base = ((smem0 >> 4) & 0x3FFFu) | (1u << 16);
                        const uint32_t b_step = b_stage >> 4;
                        const int nfull = P.n_ksteps >> 2, ntail = P.n_ksteps & 3;
                        uint32_t a_row = a_pass, d = tacc;
                        for (int mb = 0; mb < nmb; ++mb, a_row += 128u, d += BN) {
                            uint32_t b_lo = b_base;
                            for (int kb = 0; kb < nfull; ++kb, b_lo += b_step) {
                                const uint4 ko = make_uint4(P.ks_off16[kb * 4], P.ks_off16[kb * 4 + 1], P.ks_off16[kb * 4 + 2], P.ks_off16[kb * 4 + 3]);
                                umma_bf16_lohi_if(d, a_row + ko.x, a_hi, b_lo, b_hi, idesc, kb != 0, lead);
                                umma_bf16_lohi_if(d, a_row + ko.y, a_hi, b_lo + 2, b_hi, idesc, 1, lead);
                                umma_bf16_lohi_if(d, a_row + ko.z, a_hi, b_lo + 4, b_hi, idesc, 1, lead);
                                umma_bf16_lohi_if(d, a_row + ko.w, a_hi, b_lo + 6, b_hi, idesc, 1, lead);
                            }
                            if (ntail) {
                                const uint4 ko = make_uint4(P.ks_off16[nfull * 4], P.ks_off16[nfull * 4 + 1], P.ks_off16[nfull * 4 + 2], P.ks_off16[nfull * 4 + 3]);
                                umma_bf16_lohi_if(d, a_row + ko.x, a_hi, b_lo, b_hi, idesc, nfull != 0, lead);
                                if (ntail > 1) umma_bf16_lohi_if(d, a_row + ko.y, a_hi, b_lo + 2, b_hi, idesc, 1, lead);
                                if (ntail > 2) umma_bf16_lohi_if(d, a_row + ko.z, a_hi, b_lo + 4, b_hi, idesc, 1, lead);
                            }
                        }
                    } else {
                        for (int kb = 0; kb < P.nkb; ++kb) {
                            const int sb = bcnt % P.NB;
                            mbar_wait_conv(tc::smem_u32(&b_full[sb]), (bcnt / P.NB) & 1, lead, P.poll);
                            tc::tc_fence_after();
                            const uint32_t b_lo = (((smem0 + (uint32_t)sb * b_stage) >> 4) & 0x3FFFu) | (1u << 16);
                            const uint4 ko = make_uint4(P.ks_off16[kb * 4], P.ks_off16[kb * 4 + 1], P.ks_off16[kb * 4 + 2], P.ks_off16[kb * 4 + 3]);
                            const int nk = min(4, P.n_ksteps - kb * 4);
                            uint32_t a_row = a_pass, d = tacc;
                            if (nk == 4) {
                                for (int mb = 0; mb < nmb; ++mb, a_row += 128u, d += BN) {
                                    umma_bf16_lohi_if(d, a_row + ko.x, a_hi, b_lo, b_hi, idesc, kb != 0, lead);
                                    umma_bf16_lohi_if(d, a_row + ko.y, a_hi, b_lo + 2, b_hi, idesc, 1, lead);
                                    umma_bf16_lohi_if(d, a_row + ko.z, a_hi, b_lo + 4, b_hi, idesc, 1, lead);
                                    umma_bf16_lohi_if(d, a_row + ko.w, a_hi, b_lo + 6, b_hi, idesc, 1, lead);
                                }
                            } else {
                                for (int mb = 0; mb < nmb; ++mb, a_row += 128u, d += BN) {
                                    umma_bf16_lohi_if(d, a_row + ko.x, a_hi, b_lo, b_hi, idesc, kb != 0, lead);
                                    if (nk > 1) umma_bf16_lohi_if(d, a_row + ko.y, a_hi, b_lo + 2, b_hi, idesc, 1, lead);
                                    if (nk > 2) umma_bf16_lohi_if(d, a_row + ko.z, a_hi, b_lo + 4, b_hi, idesc, 1, lead);
                                }
                            }
                            umma_commit_if(tc::smem_u32(&b_empty[sb]), lead);
                            ++bcnt;
                        }
                    }
                    if (it == 0) PROF(7);
                    umma_commit_if(tc::smem_u32(&acc_full[set]), lead);
                    ++ccnt;
                }
                umma_commit_if(tc::smem_u32(&a_empty[sa]), lead);
                PROF(3);
                ++acnt;
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + EPI_WARPS) {
        // ------------------------------ epilogue ------------------------------
        // Four warps share each TMEM lane quarter.  A pass is cut into items (128-row block, column part); the
        // warps of a quarter take items round-robin.
        const int e = warp - 4, q = e & 3, part = e >> 2;
        const bool mse = (NC == 0 || NC == -1) && F32 && OP == OP_UP && P.mse_on;       // fused reconstruction loss: whole rows per warp
        // column parts per block: narrow outputs stay whole (the row decode is amortised over more columns)
        const int ncp = P.ncp;
        const int ncols = P.BN / ncp;
        const int IP = P.BY * P.BX;
        const float oscale = P.scale_ptr ? __ldg(P.scale_ptr) * P.scale_mul : 1.f;
        float mse_acc = 0.f;
        const float inv_IP = 1.f / (float)IP, inv_BX = 1.f / (float)P.BX;
        uint32_t ccnt = 0;
        int lit = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lit) {
            const int ig = tile / P.n_bands, band = tile - ig * P.n_bands;
            for (int it = 0; it < P.n_ntiles * P.n_passes; ++it) {
                const int nti = it / P.n_passes, pass = it % P.n_passes;
                const int set = ccnt % P.n_sets;
                tc::mbar_wait(tc::smem_u32(&acc_full[set]), (ccnt / P.n_sets) & 1);
                tc::tc_fence_after();
                if (it == 0 && tid == 128) PROF(4);
                const int mb0 = pass * P.MBs, nmb = min(P.MBs, P.MB_total - mb0);
                for (int item = part; item < nmb * ncp; item += EPI_WARPS / 4) {
                    const int mb = item / ncp, col0 = (item - mb * ncp) * ncols;
                    const int n0 = nti * P.BN + col0;
                    const int p = (mb0 + mb) * 128 + q * 32 + lane;
                    // p < 2^16: (p + 0.5) / d is at least 0.5/d away from an integer, so the float quotient truncates exactly
                    const int i = __float2int_rz(((float)p + 0.5f) * inv_IP), r = p - i * IP;
                    const int yr = __float2int_rz(((float)r + 0.5f) * inv_BX), x = r - yr * P.BX;
                    const int img = ig * P.BI + i, y = band * P.TH + yr;
                    const bool row_ok = i < P.BI && img < P.n_img && yr < P.TH && y < P.Hv && x < P.Wv;
                    int cls = 0, cl0 = n0;
                    if (OP == OP_UP) {
                        cls = n0 / P.Cop;
                        cl0 = n0 - cls * P.Cop;
                    }
                    long long o_base = 0, m_base = 0;
                    if (OP == OP_DOWN && row_ok) {
                        o_base = F32 ? img * P.out32.sI + y * P.out32.sH + x * P.out32.sW : tv_pix(P.out, img, y, x);
                        if (P.mask_mode) m_base = tv_pix(P.mask, img, y, x);
                    }
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * P.set_cols + mb * P.BN + col0);
                    if (NC > 0) {
                        epi_fast_item<OP, NC <= 0 ? 16 : NC>(P, bias_s, taddr, row_ok, img, y, x, n0, oscale);
                        continue;
                    }
                    if (NC == -1) {
                        // fused reconstruction loss of a 3-channel image, its own kernel: target loads first, then the accumulators
                        float tv3[4][3];
                        if (row_ok) mse_load<3>(P, img, y, x, tv3);
                        float v[32];
                        tmem_ld32_nowait(taddr, v);
                        tmem_wait_ld();
                        if (row_ok) mse_apply<3>(P, v, tv3, bias_s, img, y, x, mse_acc);
                        continue;
                    }
                    if (mse) {
                        // columns = (parity class, channel): exactly the space-to-depth pixel (y, x) of the residual
                        float v[32];
                        tmem_ld16_nowait(taddr, v);
                        if (P.BN > 16) tmem_ld16_nowait(taddr + 16, v + 16);
                        tmem_wait_ld();
                        if (row_ok) {
                            switch (P.n_valid) {
                                case 1: mse_row<1>(P, v, bias_s, img, y, x, mse_acc); break;
                                case 2: mse_row<2>(P, v, bias_s, img, y, x, mse_acc); break;
                                case 3: mse_row<3>(P, v, bias_s, img, y, x, mse_acc); break;
                                default: mse_row<4>(P, v, bias_s, img, y, x, mse_acc); break;
                            }
                        }
                        continue;
                    }
                    // 16 columns per iteration (small loop body: the 16 epilogue warps, the issuer and the producer share one
                    // instruction cache)
                    int cur_cls = -1;                 // up: pixel offsets are recomputed only when the parity class changes
                    bool ok_c = row_ok;
                    long long o_pix = o_base, m_pix = m_base;
                    // act'-mask vectors: when the item stays inside one parity class (always, for the masked layers) they are loaded
                    // one 16-column iteration ahead of their use (an L2 hit is still ~400 cycles)
                    const bool m_ahead = P.mask_mode && (OP == OP_DOWN || cl0 + ncols <= P.Cop);
                    const bf16* m_ptr = nullptr;
                    uint4 mk0 = make_uint4(0u, 0u, 0u, 0u), mk1 = mk0;
                    if (m_ahead) {
                        bool ok_m = row_ok;
                        long long mp = m_base;
                        if (OP == OP_UP) {
                            const int yy = 2 * y + (cls >> 1), xx = 2 * x + (cls & 1);
                            ok_m = row_ok && yy < P.Ho && xx < P.Wo;
                            if (ok_m) mp = tv_pix(P.mask, img, yy, xx);
                        }
                        if (ok_m) {
                            m_ptr = (const bf16*)P.mask.p + mp + (cl0 >> 3) * P.mask.sK;
                            mk0 = __ldg(reinterpret_cast<const uint4*>(m_ptr));
                            if (ncols > 8) mk1 = __ldg(reinterpret_cast<const uint4*>(m_ptr + P.mask.sK));
                        }
                    }
#pragma unroll 1
                    for (int c0 = 0; c0 < ncols; c0 += 16) {
                        float v[16];
                        tmem_ld16_nowait(taddr + c0, v);
                        uint4 mc[2] = {mk0, mk1};
                        if (m_ahead && m_ptr && c0 + 16 < ncols) {
                            mk0 = __ldg(reinterpret_cast<const uint4*>(m_ptr + (long long)((c0 + 16) >> 3) * P.mask.sK));
                            mk1 = __ldg(reinterpret_cast<const uint4*>(m_ptr + (long long)((c0 + 24) >> 3) * P.mask.sK));
                        }
                        tmem_wait_ld();
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (OP == OP_UP && cls != cur_cls) {
                                cur_cls = cls;
                                const int yy = 2 * y + (cls >> 1), xx = 2 * x + (cls & 1);
                                ok_c = row_ok && yy < P.Ho && xx < P.Wo;
                                if (ok_c) {
                                    o_pix = F32 ? img * P.out32.sI + yy * P.out32.sH + xx * P.out32.sW : tv_pix(P.out, img, yy, xx);
                                    if (P.mask_mode && !m_ahead) m_pix = tv_pix(P.mask, img, yy, xx);
                                }
                            }
                            if (ok_c) {
                                if (P.mask_mode && !m_ahead) mc[h] = __ldg(reinterpret_cast<const uint4*>((const bf16*)P.mask.p + m_pix + (cl0 >> 3) * P.mask.sK));
                                epi_store8<F32>(P, v + 8 * h, bias_s, cl0, o_pix, mc[h], oscale);
                            }
                            cl0 += 8;
                            if (OP == OP_UP && cl0 == P.Cop) {
                                cl0 = 0;
                                ++cls;
                            }
                        }
                    }
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[set]));
                ++ccnt;
            }
            if (tid == 128) PROF(5);
        }
        if (mse) {
            const float sacc = warp_sum(mse_acc);
            if (lane == 0) atomicAdd(P.mse_sum, sacc * P.mse_scale);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------
// planner for the forward-type kernel
// ---------------------------------------------------------------------------------------------------
// Tilings measured on B200 for the layer geometries of the shipped 64x64 stacks at large frame counts (profiles/sweep_plans.py);
// anything else takes the planner's heuristic.  Key: op, space-to-depth source, (Hl, Cl, Hs, Cs, k) with padded channels.
struct TunedPlan {
    int op, s2d, Hl, Cl, Hs, Cs, k;
    int BI, TH, NA, bres, NB;
};
const TunedPlan kTuned[] = {
    // op s2d  Hl  Cl  Hs   Cs  k   BI  TH NA res NB     measured at n = 50 176 (planner's own choice -> tuned), round 2
    {OP_DOWN, 1, 64, 16, 31, 32, 4, 4, 31, 1, 0, 0},     // E1 forward           1.228 -> 1.147 ms
    {OP_UP, 0, 14, 64, 6, 128, 4, 4, 7, 1, 0, 4},        // E3 dgrad             1.108 -> 0.930 ms
    {OP_UP, 0, 31, 32, 14, 64, 4, 4, 16, 1, 1, 0},       // E2 dgrad             1.795 -> 1.363 ms
    {OP_UP, 0, 13, 64, 5, 128, 5, 3, 7, 1, 0, 3},        // D2 forward           2.050 -> 1.874 ms
    {OP_DOWN, 1, 64, 16, 30, 32, 6, 3, 30, 2, 1, 0},     // D4 dgrad             1.329 -> 1.261 ms
};
const TunedPlan* find_tuned(const mrssm_pl_conv_args* a, int op) {
    if (a->n_img < 1024) return nullptr;             // small launches: the heuristic (few tiles, fill matters more than overlap)
    for (const TunedPlan& t : kTuned)
        if (t.op == op && t.s2d == (a->s2d_cq > 0) && t.Hl == a->Hl && t.Cl == a->Cl && t.Hs == a->Hs && t.Cs == a->Cs && t.k == a->ksz &&
            a->Hl == a->Wl && a->Hs == a->Ws)
            return &t;
    return nullptr;
}

int plan_fwd(const mrssm_pl_conv_args* a, int op, FwdP& P, size_t& smem_bytes) {
    memset(&P, 0, sizeof(P));
    const int k = a->ksz, nt = (k + 1) / 2;
    P.op = op;
    P.n_img = a->n_img;
    P.nt = nt;
    int Cp, N_total;
    if (op == OP_DOWN) {
        Cp = a->Cl;
        P.s2d = a->s2d_cq > 0;
        MRSSM_CHECK(Cp % 8 == 0, "plane down: source channels %d must be padded to 8", Cp);
        MRSSM_CHECK(!P.s2d || (Cp % 16 == 0 && 4 * a->s2d_cq <= Cp), "plane down: a space-to-depth source needs 4*cq <= Cl, Cl %% 16 == 0");
        P.Hv = a->Hs; P.Wv = a->Ws;
        P.x0 = P.y0 = 0;
        if (P.s2d) {
            P.planes = Cp / 8; P.ppm = P.planes;
            P.J = Cp / 16;
            P.n_ksteps = nt * nt * P.J;
        } else {
            P.planes = 4 * (Cp / 8); P.ppm = Cp / 8;
            P.J = Cp / 8;
            P.n_ksteps = k * nt * P.J;
        }
        N_total = a->n_out_pad;
        P.Cop = N_total;
        P.Ho = a->Hs; P.Wo = a->Ws;
    } else {
        Cp = a->Cs;
        MRSSM_CHECK(Cp % 16 == 0, "plane up: source channels %d must be padded to 16", Cp);
        P.Hv = (a->Hl + 1) / 2; P.Wv = (a->Wl + 1) / 2;
        P.planes = Cp / 8; P.ppm = P.planes;
        P.x0 = P.y0 = -(nt - 1);
        P.J = Cp / 16;
        P.n_ksteps = nt * nt * P.J;
        MRSSM_CHECK(a->n_out_pad % 8 == 0, "plane up: n_out_pad %d must be a multiple of 8", a->n_out_pad);
        N_total = 4 * a->n_out_pad;
        P.Cop = a->n_out_pad;
        P.Ho = a->Hl; P.Wo = a->Wl;
    }
    MRSSM_CHECK(P.n_ksteps <= 320, "plane conv: %d K-steps exceed the table", P.n_ksteps);
    MRSSM_CHECK(N_total <= 1024 || op == OP_UP, "plane conv: %d output channels exceed the bias table", N_total);
    MRSSM_CHECK(P.Cop <= 1024, "plane conv: %d output channels exceed the bias table", P.Cop);
    MRSSM_CHECK(N_total % 16 == 0, "plane conv: %d output columns not a multiple of 16", N_total);
    P.nkb = (P.n_ksteps + 3) / 4;
    P.BN = N_total <= 256 ? N_total : (N_total % 256 == 0 ? 256 : (N_total % 128 == 0 ? 128 : 64));
    MRSSM_CHECK(N_total % P.BN == 0, "plane conv: %d output columns not tileable", N_total);
    P.n_ntiles = N_total / P.BN;
    P.n_sets = 2; P.set_cols = 256;
    P.MBs = std::max(1, 256 / P.BN);
    P.BX = P.Wv + nt - 1;
    // grouped boxes (one TMA box per tensor map) put the chunk planes back to back: plane stride = the plane's own bytes.  The
    // rows an MMA block reads past its plane's end (M round-up, tap shifts) then alias the next plane — they only feed rows
    // whose outputs are discarded — and the last plane needs that much slack behind it.  Every map's first plane must start
    // 128-byte aligned: the row pitch is widened until ppm * BX is a multiple of 8 pixels (whatever BI and BY turn out to be).
    const mrssm_tv& src_view = (op == OP_DOWN) ? a->large : a->small;
    const bool can_group = !g_dbg[4] && src_view.ptr && view_groupable(src_view, op == OP_DOWN && !P.s2d) && P.ppm <= 256;
    if (can_group)
        while ((P.ppm * P.BX) % 8 != 0) ++P.BX;
    const int maxshift = (nt - 1) * P.BX + nt - 1;
    const long long min_a = (long long)P.planes * ((128 + maxshift) * 16 + 128);    // one 128-row block of the smallest tile
    // the whole packed weight stays resident when it is small; otherwise it streams through a ring of 64-wide K blocks
    const bool res_fits = P.n_ntiles == 1 && P.nkb <= 80 && (long long)P.nkb * P.BN * 128 + std::max<long long>(min_a, 36 * 1024) <= SMEM_TOTAL;
    P.b_res = res_fits && (long long)P.nkb * P.BN * 128 <= 152 * 1024;
    const TunedPlan* tuned = find_tuned(a, op);
    PlanOvr ovr = g_ovr;
    if (tuned && !ovr.BI && !ovr.TH && !ovr.NA && ovr.bres < 0) ovr = PlanOvr{tuned->BI, tuned->TH, tuned->NA, tuned->bres, tuned->NB};
    if (ovr.bres >= 0) {
        MRSSM_CHECK(!ovr.bres || res_fits, "plane conv: plan override asks for resident weights that do not fit");
        P.b_res = ovr.bres;
    }
    if (P.b_res) {
        P.NB = P.nkb;
    } else {
        P.NB = ovr.NB > 0 ? ovr.NB : std::min(8, std::max(2, (64 * 1024) / (P.BN * 128)));
        while (P.NB > 2 && min_a > SMEM_TOTAL - (long long)P.NB * P.BN * 128) --P.NB;
    }
    const long long bring = (long long)P.NB * P.BN * 128;
    const long long avail = SMEM_TOTAL - bring;
    bool grouped = false;
    auto stage_bytes = [&](int BI, int BY, int TH, int& MB_total, int& PS) {
        long long Lout = ((long long)(BI - 1) * BY + (TH - 1)) * P.BX + P.Wv;
        MB_total = (int)((Lout + 127) / 128);
        const long long dense = (long long)BI * BY * P.BX * 16, need = ((long long)MB_total * 128 + maxshift) * 16;
        grouped = can_group && 2 * P.BX <= 256 && BI <= 256 && BY <= 256 && ((long long)P.ppm * dense) % 128 == 0 && dense < 262144;
        if (grouped) {
            PS = (int)dense;
            return ((long long)P.planes * dense + std::max<long long>(0, need - dense) + 1023) / 1024 * 1024;
        }
        long long ps = std::max<long long>(dense, need);
        PS = (int)((ps + 127) / 128 * 128);
        return (long long)P.planes * PS;
    };
    int MBt, PS;
    const int BYfull = P.Hv + nt - 1;
    long long one = stage_bytes(1, BYfull, P.Hv, MBt, PS);
    if (ovr.BI > 0 || ovr.TH > 0 || ovr.NA > 0) {
        // forced tiling (tuned table / tuning sweep): whole images (TH == Hv, BI images per tile) or row bands of one image
        const int TH = ovr.TH > 0 ? std::min(ovr.TH, P.Hv) : P.Hv;
        P.n_bands = (P.Hv + TH - 1) / TH;
        P.BI = P.n_bands > 1 ? 1 : std::max(1, std::min(ovr.BI, a->n_img));
        P.TH = TH;
        P.BY = TH + nt - 1;
        P.NA = ovr.NA > 0 ? std::min(ovr.NA, 2) : 2;
        const long long sb = stage_bytes(P.BI, P.BY, P.TH, MBt, PS);
        MRSSM_CHECK(sb * P.NA <= avail && PS < 262144, "plane conv: forced plan (BI %d TH %d NA %d) needs %lld bytes of %lld", P.BI, P.TH, P.NA,
                    sb * P.NA, avail);
    } else if (one <= avail) {
        // images per tile and A stages: maximise (valid rows / MMA rows) x (accumulator-set fill when the weights stream),
        // taking the smallest image count within 2 % of the best; single buffering must win by 50 % to be chosen
        auto best_for = [&](int NA, int& bestBI) -> double {
            const long long budget = avail / NA;
            double best = -1.0;
            bestBI = 0;
            for (int BI = 1; BI <= std::min(a->n_img, 256); ++BI) {
                long long sb = stage_bytes(BI, BYfull, P.Hv, MBt, PS);
                if (sb > budget || PS >= 262144 || (MBt > 32 && BI > 1)) break;
                double sc = (double)BI * P.Hv * P.Wv / ((double)MBt * 128);
                if (!P.b_res) sc *= (double)MBt / (double)(((MBt + P.MBs - 1) / P.MBs) * P.MBs);
                if (sc > best + 0.02) {
                    best = sc;
                    bestBI = BI;
                }
            }
            return best;
        };
        int bi2 = 0, bi1 = 0;
        const double s2 = best_for(2, bi2), s1 = best_for(1, bi1);
        if (bi2 > 0 && s2 * 1.5 >= s1) {
            P.NA = 2; P.BI = bi2;
        } else {
            P.NA = 1; P.BI = std::max(1, bi1);
        }
        P.BY = BYfull; P.TH = P.Hv; P.n_bands = 1;
    } else {
        P.NA = 2;
        const long long budget = avail / 2;
        int TH = P.Hv;
        while (TH > 1 && stage_bytes(1, TH + nt - 1, TH, MBt, PS) > budget) --TH;
        MRSSM_CHECK(stage_bytes(1, TH + nt - 1, TH, MBt, PS) <= budget, "plane conv: a single row band does not fit shared memory");
        P.n_bands = (P.Hv + TH - 1) / TH;
        TH = (P.Hv + P.n_bands - 1) / P.n_bands;
        P.BI = 1; P.TH = TH; P.BY = TH + nt - 1;
    }
    P.a_stage_bytes = (int)stage_bytes(P.BI, P.BY, P.TH, P.MB_total, P.PS);
    P.group = grouped ? 1 : 0;
    MRSSM_CHECK(P.PS / 16 < 16384, "plane conv: plane stride %d too large for the descriptor", P.PS);
    P.plane_bytes = P.BI * P.BY * P.BX * 16;
    P.n_groups = (a->n_img + P.BI - 1) / P.BI;
    P.n_passes = (P.MB_total + P.MBs - 1) / P.MBs;
    // equal passes: the two TMEM sets ping-pong between passes, so an 8 + 1 split would serialise the big pass's MMAs with its own epilogue
    P.MBs = (P.MB_total + P.n_passes - 1) / P.n_passes;
    smem_bytes = (size_t)bring + (size_t)P.NA * P.a_stage_bytes + 1024;
    MRSSM_CHECK(smem_bytes <= (size_t)SMEM_TOTAL + 1024, "plane conv: shared memory plan %zu too large", smem_bytes);
    // epilogue
    P.act = a->act; P.mask_mode = a->mask.ptr ? a->mask_mode : 0; P.out_f32 = a->out_f32;
    P.n_valid = a->n_out_valid;
    const mrssm_tv& out = (op == OP_DOWN) ? a->small : a->large;
    P.out = cvt(out); P.mask = cvt(a->mask); P.out32 = cvt(a->out32);
    P.mask_img_bytes = 0;
    if (P.mask_mode && !g_dbg[2]) {
        // dense per-image blocks: parity-planar = 4 parity planes of (Cp/8) chunk planes; planar = (Cp/8) chunk planes; NHWC = H*W*Cp
        const mrssm_tv& mv = a->mask;
        const long long img_bytes = 2 * mv.sI;
        if (img_bytes > 0 && img_bytes % 16 == 0 && ((uintptr_t)mv.ptr & 15) == 0 && img_bytes <= (1 << 20)) P.mask_img_bytes = img_bytes;
    }
    P.bias = a->bias;
    P.prof = g_prof;
    P.poll = g_dbg[0];
    P.scale_ptr = a->scale_ptr; P.scale_mul = a->scale_mul;
    P.mse_target = a->mse_target; P.mse_sum = a->mse_sum; P.mse_scale = a->mse_scale;
    P.tgt = cvt(a->mse_target_s2d);
    if (P.tgt.p) P.mse_target = nullptr;
    P.mse_on = (P.mse_target || P.tgt.p) ? 1 : 0;
    if (P.mse_on) {
        MRSSM_CHECK(op == OP_UP && P.out_f32 && P.mse_sum && a->n_out_valid <= 4 && a->n_out_pad == 8 && a->large.ptr && !a->large.par,
                    "plane up: fused MSE needs out_f32, <= 4 channels (n_out_pad 8), a sum buffer and a linear space-to-depth residual view");
        P.out = cvt(a->large);
        P.mse_vec = a->out32.sW == 1 && (P.Wo & 1) == 0 && a->out32.sH % 2 == 0 && a->out32.sC % 2 == 0 && a->out32.sI % 2 == 0 &&
                    a->mse_target && ((uintptr_t)a->mse_target & 7) == 0;
        const long long tb = P.tgt.p ? 2 * P.tgt.sI : 4 * a->out32.sI;
        const void* tbase = P.tgt.p ? P.tgt.p : (const void*)a->mse_target;
        P.tgt_img_bytes = (!g_dbg[2] && tb > 0 && tb % 16 == 0 && ((uintptr_t)tbase & 15) == 0 && tb <= (1 << 20)) ? tb : 0;
        MRSSM_CHECK(!P.tgt.p || (!P.tgt.par && P.tgt.sK % 8 == 0 && P.tgt.sW % 8 == 0 && P.tgt.sH % 8 == 0 && P.tgt.sI % 8 == 0 && ((uintptr_t)P.tgt.p & 15) == 0),
                    "plane up: the space-to-depth loss target must be a linear view with 16-byte aligned chunks");
    }
    // epilogue shape: columns per item.  Fast path: bf16 output, act none / ReLU, ReLU (or no) act'-mask, and items that stay
    // inside one parity class; otherwise the generic epilogue (fp32 outputs, fused loss, ELU, odd widths).
    P.bits_out = a->relu_bits_out; P.bits_in = a->relu_bits_in;
    {
        const bool mse_k = P.mse_on != 0;
        P.mse_lean = (mse_k && op == OP_UP && P.n_valid == 3 && P.BN == 32 && P.act == 0 && !g_dbg[1]) ? 1 : 0;
        int nc = 0;
        for (int c : {64, 32, 16}) {
            if (P.BN % c == 0 && (op == OP_DOWN || P.Cop % c == 0) && (c == 16 || P.BN / c <= 16)) {
                nc = c;
                break;
            }
        }
        // narrow tiles: keep >= 4 items per 128-row block per lane quarter busy only when it costs nothing
        if (nc == 64 && P.BN < 256) nc = 32;
        const bool act_ok = P.act == 0 || P.act == MRSSM_ACT_RELU;
        const bool mask_ok = P.mask_mode == 0 || P.mask_mode == MRSSM_ACT_RELU;
        P.fast = (!P.out_f32 && !mse_k && nc > 0 && act_ok && mask_ok && !g_dbg[1]) ? 1 : 0;
        if (P.bits_in) {
            MRSSM_CHECK(a->mask_mode == MRSSM_ACT_RELU && !a->mask.ptr, "plane conv: relu_bits_in replaces the mask view (mask_mode must be ReLU)");
            P.mask_mode = MRSSM_ACT_RELU;
        }
        MRSSM_CHECK(!(P.bits_in || P.bits_out) || P.fast, "plane conv: ReLU sign bits need the bf16 fast epilogue (bf16 output, ReLU, %d columns)", P.BN);
        MRSSM_CHECK(!P.bits_out || P.act == MRSSM_ACT_RELU, "plane conv: relu_bits_out needs act = ReLU");
        if (P.fast) {
            P.nc = nc; P.ncp = P.BN / nc;
        } else {
            P.ncp = mse_k ? 1 : ((P.BN % 64 == 0 && P.BN >= 128) ? 4 : ((P.BN % 32 == 0 && P.BN >= 64) ? 2 : 1));
            P.nc = P.BN / P.ncp;
        }
        P.bits_img_bytes = (P.bits_in && !g_dbg[2]) ? (long long)P.Ho * P.Wo * (P.Cop >> 3) : 0;
        if (P.bits_in) P.mask_img_bytes = 0;
    }
    auto vec_ok = [](const mrssm_tv& t) {
        return t.sW % 8 == 0 && t.sH % 8 == 0 && t.sI % 8 == 0 && t.sK % 8 == 0 && t.sP % 8 == 0 && ((uintptr_t)t.ptr & 15) == 0;
    };
    MRSSM_CHECK(P.out_f32 ? (a->out32.ptr != nullptr || P.mse_on) : vec_ok(out),
                "plane conv: bf16 output view must keep 8-channel chunks 16-byte aligned");
    MRSSM_CHECK(!P.mask_mode || vec_ok(a->mask), "plane conv: mask view must keep 8-channel chunks 16-byte aligned");
    return 0;
}

int launch_fwd(const mrssm_pl_conv_args* a, int op, cudaStream_t st) {
    MRSSM_CHECK(a && a->wpacked, "plane conv: null weight");
    MRSSM_CHECK(a->ksz >= 2, "plane conv: kernel size %d (dense layers use mrssm_tc_conv_*)", a->ksz);
    FwdP P;
    size_t smem;
    if (int rc = plan_fwd(a, op, P, smem)) return rc;
    CUtensorMap mA[4], mB;
    memset(mA, 0, sizeof(mA));
    if (op == OP_DOWN) {
        MRSSM_CHECK(a->large.ptr, "plane down: null source");
        if (P.s2d) {
            if (int rc = make_view_maps(mA, &P.mergedA, a->large, (a->Hl + 1) / 2, (a->Wl + 1) / 2, a->Cl, a->n_img, false, P.BX, P.BY, P.BI, P.group ? P.ppm : 0))
                return rc;
        } else if (int rc = make_view_maps(mA, &P.mergedA, a->large, a->Hl, a->Wl, a->Cl, a->n_img, true, P.BX, P.BY, P.BI, P.group ? P.ppm : 0)) {
            return rc;
        }
    } else {
        MRSSM_CHECK(a->small.ptr, "plane up: null source");
        if (int rc = make_view_maps(mA, &P.mergedA, a->small, a->Hs, a->Ws, a->Cs, a->n_img, false, P.BX, P.BY, P.BI, P.group ? P.ppm : 0)) return rc;
    }
    const long long K_total = (long long)P.nkb * 64;
    const long long N_total = (long long)P.BN * P.n_ntiles;
    if (int rc = make_w_map(&mB, a->wpacked, K_total, N_total, P.BN)) return rc;
    // K-step table: which chunk plane and which pixel shift every K step (tap, 16-channel slab) reads
    for (int ks = 0; ks < 320; ++ks) {
        uint32_t off = 0;
        if (ks < P.n_ksteps) {
            const int j = ks % P.J, t = ks / P.J, b = t % P.nt, r = t / P.nt;
            int plane, shift;
            if (op == OP_DOWN && P.s2d) {   // r = a: all four input parities are channels of the same pixel
                plane = 2 * j;
                shift = r * P.BX + b;
            } else if (op == OP_DOWN) {     // r = kh
                plane = (r & 1) * 2 * P.J + 2 * j;
                shift = (r >> 1) * P.BX + b;
            } else {                        // r = a
                plane = 2 * j;
                shift = (P.nt - 1 - r) * P.BX + (P.nt - 1 - b);
            }
            off = ((uint32_t)plane * (uint32_t)P.PS + (uint32_t)shift * 16u) >> 4;
        }
        P.ks_off16[ks] = off;
    }
    const int n_tiles = P.n_groups * P.n_bands;
    int grid = std::min(n_tiles, g_sm_budget);
    smem = std::max<size_t>(smem, 120 * 1024);       // > half an SM: one CTA per SM (each allocates all 512 TMEM columns)
#define PL_LAUNCH(OPV, F32V, NCV)                                                                                                   \
    do {                                                                                                                             \
        MRSSM_CUDA(cudaFuncSetAttribute(plane_fwd_kernel<OPV, F32V, NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        plane_fwd_kernel<OPV, F32V, NCV><<<grid, FWD_THREADS, smem, st>>>(mA[0], mA[1], mA[2], mA[3], mB, P);                        \
    } while (0)
#define PL_LAUNCH_OP(OPV)                                        \
    do {                                                         \
        if (P.out_f32 && P.mse_lean) PL_LAUNCH(OP_UP, true, -1); \
        else if (P.out_f32) PL_LAUNCH(OPV, true, 0);             \
        else if (!P.fast) PL_LAUNCH(OPV, false, 0);              \
        else if (P.nc == 64) PL_LAUNCH(OPV, false, 64);          \
        else if (P.nc == 32) PL_LAUNCH(OPV, false, 32);          \
        else PL_LAUNCH(OPV, false, 16);                          \
    } while (0)
    if (op == OP_DOWN) PL_LAUNCH_OP(OP_DOWN); else PL_LAUNCH_OP(OP_UP);
#undef PL_LAUNCH_OP
#undef PL_LAUNCH
    MRSSM_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// weight packing for the forward-type kernel:  out[n][k], k = kstep*16 + e  (zero padded to K_total)
// ---------------------------------------------------------------------------------------------------
__global__ void pack_plane_kernel(const float* __restrict__ w, long long w_ss, long long w_sl, int Cs_valid, int Cl_valid, int Csp, int Clp,
                                  int ksz, int op, int cq, int N_total, int K_total, bf16* __restrict__ out) {
    const int nt = (ksz + 1) / 2;
    const long long total = (long long)N_total * K_total;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % K_total), n = (int)(i / K_total);
        const int ks = kk >> 4, e = kk & 15;
        float v = 0.f;
        if (op == 2) {                   // down over a space-to-depth source: channel = parity * cq + c
            const int J = Clp / 16;
            const int j = ks % J, t = ks / J, b = t % nt, aa = t / nt;
            const int ch = (2 * j + (e >> 3)) * 8 + (e & 7);
            const int par = ch / cq, c = ch - par * cq;
            const int kh = 2 * aa + (par >> 1), kw = 2 * b + (par & 1);
            if (aa < nt && par < 4 && n < Cs_valid && c < Cl_valid && kh < ksz && kw < ksz) v = w[n * w_ss + c * w_sl + kh * ksz + kw];
        } else if (op == OP_DOWN) {
            const int J = Clp / 8;
            const int j = ks % J, t = ks / J, b = t % nt, kh = t / nt;
            const int pl = 2 * j + (e >> 3);
            const int px = pl / J, chunk = pl - px * J;
            const int cl = chunk * 8 + (e & 7), kw = 2 * b + px;
            if (n < Cs_valid && cl < Cl_valid && kh < ksz && kw < ksz) v = w[n * w_ss + cl * w_sl + kh * ksz + kw];
        } else {
            const int J = Csp / 16;
            const int j = ks % J, t = ks / J, b = t % nt, aa = t / nt;
            const int cls = n / Clp, cl = n - cls * Clp;
            const int cs = (2 * j + (e >> 3)) * 8 + (e & 7);
            const int kh = (cls >> 1) + 2 * aa, kw = (cls & 1) + 2 * b;
            if (aa < nt && cls < 4 && cl < Cl_valid && cs < Cs_valid && kh < ksz && kw < ksz) v = w[cs * w_ss + cl * w_sl + kh * ksz + kw];
        }
        out[i] = __float2bfloat16(v);
    }
}

// ---------------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------------
struct WgP {
    int n_img, n_groups, n_bands, BI, BX, BY, TH, SBY;
    int nS, nL, cpl;            // small planes per m-half, large planes, chunks per large parity plane group (Clp/8)
    int PS_s, PS_l, small_bytes, large_bytes, offL, stage_bytes;
    int nksteps;                // K steps (16 pixels) per tile
    int ksz, nt, Clp, Csp, N;
    int NG, gpp, n_cpass, n_mhalf;
    int NA, zero_bytes, mergedS, mergedL, s2d_cq;
    int group;                  // one TMA box per tensor map (all chunk planes of the map, planes dense)
    int psplit;                 // the tap groups of even / odd kh read disjoint large planes (row parity): grid.y passes are split
                                // by row parity and each CTA loads only its parity's half of the large tile
    int nL_cta;                 // large planes a CTA loads (nL, or half of it with psplit)
    int mrep;                   // Cs <= 64: the small tile is loaded twice, the copy one row lower, so the 128 MMA rows are
                                // [cs | cs of the next row tap of the same parity]: one MMA accumulates row taps kh and kh + 2
    int rep, Ntot;              // rep: the large tile is loaded once per COLUMN tap b with the box origin shifted by b pixels (plane
                                // group = b), so one MMA of N = Ntot = nt * N covers the nt column taps of a row tap a; row taps stay
                                // descriptor shifts of a * BX pixels (space-to-depth sources: N = 16 per tap otherwise)
    const float* scale_ptr;
    float scale_mul;
    int cs_valid, cl_valid;
    float* dw;
    long long w_ss, w_sl;
    // bias gradient from the resident operand tile (warps 4..7 while the MMAs run): db_from 1 = small planes, 2 = large planes
    int poll;
    float* db;
    int db_from, db_nch, db_npar;     // 8-channel chunks to sum, plane groups per chunk (large, parity split: 4)
    int db_valid, db_fold;            // real channels; fold > 0: channel c of the tile is channel c % fold for c < 4 * fold
};

constexpr int WG_SUM_WARPS = 4;       // warps 4..7

__global__ void __launch_bounds__(NTHREADS, 1)
plane_wgrad_kernel(const __grid_constant__ CUtensorMap mS, const __grid_constant__ CUtensorMap mL0, const __grid_constant__ CUtensorMap mL1,
                   const __grid_constant__ CUtensorMap mL2, const __grid_constant__ CUtensorMap mL3, const WgP P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[2], a_empty[2], acc_full;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem0 - tc::smem_u32(smem_raw));
    const int cpass = blockIdx.y % P.n_cpass, mhalf = blockIdx.y / P.n_cpass;
    // tap groups of this CTA: [g0, g0 + ng) of all groups, or — row-parity split — of the groups of row parity `par`
    const int par = P.psplit ? (cpass & 1) : 0, sub = P.psplit ? (cpass >> 1) : cpass;
    const int n_kh_par = P.mrep ? (P.ksz - par + 3) / 4 : (P.ksz - par + 1) / 2;        // kh = par, par + 2 (or + 4), .. < ksz
    const int NGc = P.psplit ? n_kh_par * P.nt : P.NG;
    const int g0 = sub * P.gpp, ng = max(0, min(P.gpp, NGc - g0));
    const int n_tiles = ng > 0 ? P.n_groups * P.n_bands : 0;       // (a pass without groups — odd kernels, odd parity — has no work)

    // zero the operand stages once: pixels the TMA boxes never write (row tails, M/K round-up) must read as 0
    for (int i = tid * 16; i < P.zero_bytes; i += NTHREADS * 16) *reinterpret_cast<uint4*>(smem_al + i) = make_uint4(0, 0, 0, 0);
    tc::fence_proxy_async();
    if (tid == 0) {
        // a stage is released by the MMA commit and, when this CTA also sums the bias gradient, by the summing warps
        const bool do_sum = P.db_from != 0 && sub == 0 && (P.db_from == 1 ? par == 0 : mhalf == 0);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(tc::smem_u32(&a_full[s]), 1);
            tc::mbar_init(tc::smem_u32(&a_empty[s]), do_sum ? 1 + WG_SUM_WARPS : 1);
        }
        tc::mbar_init(tc::smem_u32(&acc_full), 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        {       // converged warp, elected lane issues (see the forward-type kernel)
            const uint32_t lead = elect_one_lane();
            uint32_t acnt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int sa = acnt % P.NA;
                mbar_wait_conv(tc::smem_u32(&a_empty[sa]), ((acnt / P.NA) & 1) ^ 1, lead, P.poll);
                const uint32_t bar = tc::smem_u32(&a_full[sa]);
                mbar_expect_tx_if(bar, (uint32_t)(P.mrep ? 2 : 1) * (uint32_t)P.nS * (uint32_t)P.small_bytes + (uint32_t)P.nL_cta * (uint32_t)P.large_bytes, lead);
                const int ig = tile / P.n_bands, band = tile - ig * P.n_bands;
                const uint32_t dst = smem0 + (uint32_t)sa * (uint32_t)P.stage_bytes;
                if (P.group) {
                    // one box for the small planes of this Cs half, one per large tensor map (parity, or column tap when replicated)
                    const int cy = band * P.TH, ci = ig * P.BI;
                    const uint32_t dl = dst + (uint32_t)P.offL, map_bytes = (uint32_t)P.cpl * (uint32_t)P.PS_l;
                    tma_load_4d_if(dst, &mS, 0, cy, ci, mhalf * 16, bar, lead);
                    if (P.mrep) tma_load_4d_if(dst + (uint32_t)P.nS * (uint32_t)P.PS_s, &mS, 0, cy - 1, ci, mhalf * 16, bar, lead);   // one row lower
                    if (P.rep) {
                        for (int mi = 0; mi < P.nt; ++mi) tma_load_4d_if(dl + (uint32_t)mi * map_bytes, &mL0, 2 * mi, cy, ci, 0, bar, lead);
                    } else if (P.psplit) {          // the two column-parity maps of this CTA's row parity
                        tma_load_4d_if(dl, par ? &mL2 : &mL0, 0, cy, ci, 0, bar, lead);
                        tma_load_4d_if(dl + map_bytes, par ? &mL3 : &mL1, 0, cy, ci, 0, bar, lead);
                    } else {
                        tma_load_4d_if(dl, &mL0, 0, cy, ci, 0, bar, lead);
                        if (P.nL > P.cpl) {
                            tma_load_4d_if(dl + map_bytes, &mL1, 0, cy, ci, 0, bar, lead);
                            tma_load_4d_if(dl + 2 * map_bytes, &mL2, 0, cy, ci, 0, bar, lead);
                            tma_load_4d_if(dl + 3 * map_bytes, &mL3, 0, cy, ci, 0, bar, lead);
                        }
                    }
                    if (P.NA == 1 && tile + (int)gridDim.x < n_tiles) {
                        // single stage: the next tile's boxes go to L2 now, its loads (after this tile's MMAs) then hit L2
                        const int nx = tile + gridDim.x, nig = nx / P.n_bands, ny = (nx - nig * P.n_bands) * P.TH, ni = nig * P.BI;
                        tma_prefetch_4d_if(&mS, 0, ny, ni, mhalf * 16, lead);
                        if (P.rep) {
                            for (int mi = 0; mi < P.nt; ++mi) tma_prefetch_4d_if(&mL0, 2 * mi, ny, ni, 0, lead);
                        } else if (P.psplit) {
                            tma_prefetch_4d_if(par ? &mL2 : &mL0, 0, ny, ni, 0, lead);
                            tma_prefetch_4d_if(par ? &mL3 : &mL1, 0, ny, ni, 0, lead);
                        } else {
                            tma_prefetch_4d_if(&mL0, 0, ny, ni, 0, lead);
                            if (P.nL > P.cpl) {
                                tma_prefetch_4d_if(&mL1, 0, ny, ni, 0, lead);
                                tma_prefetch_4d_if(&mL2, 0, ny, ni, 0, lead);
                                tma_prefetch_4d_if(&mL3, 0, ny, ni, 0, lead);
                            }
                        }
                    }
                    ++acnt;
                    continue;
                }
                for (int q = 0; q < P.nS; ++q) {
                    if (P.mergedS)
                        tma_load_4d_if(dst + (uint32_t)q * (uint32_t)P.PS_s, &mS, 0, band * P.TH, mhalf * 16 + q, ig * P.BI, bar, lead);
                    else
                        tma_load_4d_if(dst + (uint32_t)q * (uint32_t)P.PS_s, &mS, (mhalf * 16 + q) * 8, 0, band * P.TH, ig * P.BI, bar, lead);
                }
                for (int q = 0; q < P.nL; ++q) {
                    const int mi = q / P.cpl, ch = q - mi * P.cpl;
                    if (P.rep) {            // plane group mi = column tap b: the same box, origin shifted by b pixels
                        if (P.mergedL)
                            tma_load_4d_if(dst + (uint32_t)P.offL + (uint32_t)q * (uint32_t)P.PS_l, &mL0, 2 * mi, band * P.TH, ch, ig * P.BI, bar, lead);
                        else
                            tma_load_4d_if(dst + (uint32_t)P.offL + (uint32_t)q * (uint32_t)P.PS_l, &mL0, ch * 8, mi, band * P.TH, ig * P.BI, bar, lead);
                        continue;
                    }
                    const CUtensorMap* m = mi == 0 ? &mL0 : (mi == 1 ? &mL1 : (mi == 2 ? &mL2 : &mL3));
                    if (P.mergedL)
                        tma_load_4d_if(dst + (uint32_t)P.offL + (uint32_t)q * (uint32_t)P.PS_l, m, 0, band * P.TH, ch, ig * P.BI, bar, lead);
                    else
                        tma_load_4d_if(dst + (uint32_t)P.offL + (uint32_t)q * (uint32_t)P.PS_l, m, ch * 8, 0, band * P.TH, ig * P.BI, bar, lead);
                }
                ++acnt;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // (two issuing warps do not help: profiles/micro/umma_rate.cu pattern 6 — the ~46-cycle floor per MMA is not per thread)
        {
            const uint32_t lead = elect_one_lane();
            const uint32_t idesc = tc::idesc_bf16(128, P.N, 1, 1);
            // MN-major un-swizzled descriptors: LBO = 128 B between 8-pixel (K) groups, SBO = plane stride between 8-channel chunks
            const uint32_t s_hi = ((uint32_t)P.PS_s >> 4) | (1u << 14), l_hi = ((uint32_t)P.PS_l >> 4) | (1u << 14);
            const uint32_t lbo = (128u >> 4) << 16;
            uint32_t acnt = 0;
            uint32_t accum = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int sa = acnt % P.NA;
                mbar_wait_conv(tc::smem_u32(&a_full[sa]), (acnt / P.NA) & 1, lead, P.poll);
                tc::tc_fence_after();
                const uint32_t sS = smem0 + (uint32_t)sa * (uint32_t)P.stage_bytes, sL = sS + (uint32_t)P.offL;
                const uint32_t a0 = ((sS >> 4) & 0x3FFFu) | lbo;
                if (P.rep) {                // the nt column taps of row tap a in one MMA: B walks the nt * cpl x-shifted plane copies
                    const uint32_t idesc_all = tc::idesc_bf16(128, P.Ntot, 1, 1);
                    for (int ta = 0; ta < P.nt; ++ta) {
                        uint32_t a_lo = a0, b_lo = (((sL + (uint32_t)(ta * P.BX) * 16u) >> 4) & 0x3FFFu) | lbo;
                        const uint32_t d = tmem_base + (uint32_t)(ta * P.Ntot);
                        umma_bf16_lohi_if(d, a_lo, s_hi, b_lo, l_hi, idesc_all, accum, lead);
                        for (int ks = 1; ks < P.nksteps; ++ks) {
                            a_lo += 16u;
                            b_lo += 16u;
                            umma_bf16_lohi_if(d, a_lo, s_hi, b_lo, l_hi, idesc_all, 1, lead);
                        }
                    }
                    umma_commit_if(tc::smem_u32(&a_empty[sa]), lead);
                    accum = 1;
                    ++acnt;
                    continue;
                }
                for (int gi = 0; gi < ng; ++gi) {
                    const int g = g0 + gi, gk = g / P.nt, b = g - gk * P.nt;      // space-to-depth source: kh is the row tap a
                    int kh;
                    if (P.psplit) kh = par + (P.mrep ? 4 : 2) * gk;               // the taps of this CTA's row parity
                    else kh = P.mrep ? 4 * (gk >> 1) + (gk & 1) : gk;             // stacked: base taps 0, 1, 4, 5, ..
                    const uint32_t prow = P.psplit ? 0u : (uint32_t)((kh & 1) * 2 * P.cpl);      // split: only this parity's planes are loaded
                    const uint32_t boff = P.s2d_cq ? (uint32_t)(kh * P.BX + b) * 16u
                                                   : prow * (uint32_t)P.PS_l + (uint32_t)((kh >> 1) * P.BX + b) * 16u;
                    uint32_t a_lo = a0, b_lo = (((sL + boff) >> 4) & 0x3FFFu) | lbo;
                    const uint32_t d = tmem_base + (uint32_t)(gi * P.N);
                    umma_bf16_lohi_if(d, a_lo, s_hi, b_lo, l_hi, idesc, accum, lead);
                    for (int ks = 1; ks < P.nksteps; ++ks) {
                        a_lo += 16u;            // 16 pixels = 256 B
                        b_lo += 16u;
                        umma_bf16_lohi_if(d, a_lo, s_hi, b_lo, l_hi, idesc, 1, lead);
                    }
                }
                umma_commit_if(tc::smem_u32(&a_empty[sa]), lead);
                accum = 1;
                ++acnt;
            }
            umma_commit_if(tc::smem_u32(&acc_full), lead);
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp - 4;
        if (P.db_from != 0 && sub == 0 && (P.db_from == 1 ? par == 0 : mhalf == 0)) {      // `large` is shared by the Cs halves: one of them sums it
            // ---- bias gradient: per-channel sums over the pixels of the gradient operand, read from the tile the TMA already
            // put in shared memory for the MMAs (planes [chunk][pixel][8 ch]: a warp sweeps one plane with 16-byte loads).
            // Chunk c of this CTA is summed by warp (c % 4) [nch >= 4] or by warps {c, c + nch, ..} [nch < 4].
            const int nch = P.db_nch;
            const int cpw = nch >= WG_SUM_WARPS ? nch / WG_SUM_WARPS : 1;        // chunks per warp (1, 2 or 4)
            const int wpc = nch >= WG_SUM_WARPS ? 1 : WG_SUM_WARPS / nch;        // warps per chunk
            const int sub = nch >= WG_SUM_WARPS ? 0 : q / nch;                   // which share of the pixels (wpc > 1)
            const int c_first = nch >= WG_SUM_WARPS ? q : q % nch;
            const bool active = nch >= WG_SUM_WARPS || q < nch * wpc;
            float acc[4][8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
            const uint32_t ps = P.db_from == 1 ? (uint32_t)P.PS_s : (uint32_t)P.PS_l;
            const uint32_t off0 = P.db_from == 1 ? 0u : (uint32_t)P.offL;
            const int grp_stride = P.cpl;                                        // large: plane index = group * cpl + chunk
            uint32_t acnt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int sa = acnt % P.NA;
                tc::mbar_wait(tc::smem_u32(&a_full[sa]), (acnt / P.NA) & 1);
                const int band = tile % P.n_bands;
                // entries to count: the whole plane, or — row bands share their halo rows — the first TH rows (all of them in the last band)
                int cnt;
                if (P.db_from == 1) cnt = P.BI * P.SBY * P.BX;
                else cnt = (P.n_bands > 1 && band < P.n_bands - 1) ? P.TH * P.BX : P.BI * P.BY * P.BX;
                if (active) {
                    const uint8_t* stage = smem_al + (size_t)sa * P.stage_bytes + off0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        if (a < cpw) {
                            const int c = c_first + a * WG_SUM_WARPS;
                            for (int g = 0; g < P.db_npar; ++g) {
                                const uint8_t* pl = stage + (size_t)(g * grp_stride + c) * ps;
                                for (int e = sub * 32 + lane; e < cnt; e += 32 * wpc) {
                                    const uint4 pk = *reinterpret_cast<const uint4*>(pl + (size_t)e * 16);
                                    const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
                                    for (int h = 0; h < 4; ++h) {
                                        const float2 f = __bfloat1622float2(ph[h]);
                                        acc[a][2 * h] += f.x;
                                        acc[a][2 * h + 1] += f.y;
                                    }
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(&a_empty[sa]));
                ++acnt;
            }
            if (active) {
                const float sc = P.scale_ptr ? __ldg(P.scale_ptr) * P.scale_mul : 1.f;
                const int ch0 = P.db_from == 1 ? mhalf * 16 : 0;                  // small: this CTA holds chunks [16 mhalf, 16 mhalf + nS)
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    if (a < cpw) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float sum = warp_sum(acc[a][e]);
                            int c = (ch0 + c_first + a * WG_SUM_WARPS) * 8 + e;
                            bool ok = true;
                            if (P.db_fold > 0) {
                                ok = c < 4 * P.db_fold;
                                c = c % P.db_fold;
                            }
                            if (lane == 0 && ok && c < P.db_valid) atomicAdd(P.db + c, sum * sc);
                        }
                    }
                }
            }
        }
        tc::mbar_wait(tc::smem_u32(&acc_full), 0);
        tc::tc_fence_after();
        int cs = mhalf * 128 + q * 32 + lane;
        int kh_add = 0;
        if (P.mrep) {               // rows 64..127: the copy one row lower = the row tap two further on
            kh_add = cs >= 64 ? 2 : 0;
            cs &= 63;
        }
        const float oscale = P.scale_ptr ? __ldg(P.scale_ptr) * P.scale_mul : 1.f;
        for (int gi = 0; gi < ng; ++gi) {
            const int g = g0 + gi, gk = g / P.nt, b = g - gk * P.nt;
            const int kh = (P.psplit ? par + (P.mrep ? 4 : 2) * gk : (P.mrep ? 4 * (gk >> 1) + (gk & 1) : gk)) + kh_add;
            for (int c0 = 0; c0 < P.N; c0 += 16) {
                float v[16];
                tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(gi * P.N + c0), v);
                if (cs >= P.cs_valid) continue;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int c = c0 + e;
                    int cl, kh2, kw;
                    bool ok;
                    if (P.s2d_cq) {
                        const int par = c / P.s2d_cq;
                        cl = c - par * P.s2d_cq;
                        kh2 = 2 * kh + (par >> 1);
                        kw = 2 * b + (par & 1);
                        ok = par < 4 && kh2 < P.ksz;
                    } else {
                        const int px = c / P.Clp;
                        cl = c - px * P.Clp;
                        kh2 = kh;
                        kw = 2 * b + px;
                        ok = kh < P.ksz;
                    }
                    if (ok && cl < P.cl_valid && kw < P.ksz) atomicAdd(P.dw + cs * P.w_ss + cl * P.w_sl + kh2 * P.ksz + kw, v[e] * oscale);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

int plan_wgrad(const mrssm_pl_conv_args* a, WgP& P, size_t& smem_bytes, int& splits) {
    memset(&P, 0, sizeof(P));
    const int k = a->ksz, nt = (k + 1) / 2;
    P.n_img = a->n_img; P.ksz = k; P.nt = nt;
    P.Clp = a->Cl; P.Csp = a->Cs;
    MRSSM_CHECK(P.Clp % 8 == 0 && P.Csp % 8 == 0, "plane wgrad: channels must be padded to 8 (Cl %d Cs %d)", P.Clp, P.Csp);
    P.s2d_cq = a->s2d_cq;
    MRSSM_CHECK(!P.s2d_cq || (P.Clp % 16 == 0 && 4 * P.s2d_cq <= P.Clp), "plane wgrad: a space-to-depth source needs 4*cq <= Cl, Cl %% 16 == 0");
    P.N = P.s2d_cq ? P.Clp : 2 * P.Clp;
    MRSSM_CHECK(P.N % 16 == 0 && P.N <= 256, "plane wgrad: N = %d must be a multiple of 16, <= 256", P.N);
    P.cpl = P.Clp / 8;
    P.nL = P.s2d_cq ? P.cpl : 4 * P.cpl;
    P.n_mhalf = (P.Csp + 127) / 128;
    P.nS = std::min(16, P.Csp / 8);
    P.rep = (P.s2d_cq && nt * P.N <= 256 && nt * nt * P.N <= 512 && !g_dbg[3]) ? 1 : 0;
    P.Ntot = nt * P.N;
    if (P.rep) P.nL = nt * P.cpl;
    // row-tap stacking (mrep): with at most 64 small channels half of the 128 MMA rows would be empty; a second copy of the small
    // tile, one row lower, fills them with the contributions of the row tap two further on (whole-image tiles, even kernels)
    const bool can_mrep = !g_dbg[5] && !P.rep && !P.s2d_cq && P.nS == 8 && P.n_mhalf == 1 && k % 2 == 0;
    const int BXmin = a->Ws + nt - 1;
    // grouped TMA boxes (one per tensor map: every chunk plane of the map, planes back to back): both operands must be
    // x-contiguous planes.  The contraction runs over the pixels of a plane in steps of 16, so with nothing between the planes
    // the pixel count of a small plane must itself be a multiple of 16: the row pitch BX is widened until it is (the extra
    // columns are out-of-bounds zeros of the boxes).
    const bool can_group = !g_dbg[4] && view_groupable(a->small, false) && view_groupable(a->large, !P.s2d_cq);
    bool want_mrep = can_mrep, want_psplit = false;
    const bool can_psplit = !g_dbg[6] && !P.rep && !P.s2d_cq;
    auto plan = [&](int BI, int TH, bool banded, int NA) -> long long {
        const int BY = TH + nt - 1;
        const int SBY = banded ? TH : BY;
        int BX = BXmin;
        bool grp = can_group && BI <= 256 && BY <= 256 && P.nS <= 256 && P.cpl <= 256;
        if (grp) {
            while (((long long)BI * SBY * BX) % 16 != 0 && BX < BXmin + 16) ++BX;
            grp = ((long long)BI * SBY * BX) % 16 == 0 && 2 * BX <= 256;
            if (!grp) BX = BXmin;
        }
        if (grp) {       // every box destination 128-byte aligned
            const long long sb = (long long)BI * SBY * BX * 16, lb = (long long)BI * BY * BX * 16;
            grp = ((long long)P.nS * sb) % 128 == 0 && ((long long)P.cpl * lb) % 128 == 0 && sb < 262144 && lb < 262144;
            if (!grp) BX = BXmin;
        }
        P.BX = BX; P.group = grp ? 1 : 0;
        P.mrep = (want_mrep && grp && !banded) ? 1 : 0;
        P.psplit = (want_psplit && grp) ? 1 : 0;
        P.nL_cta = P.psplit ? P.nL / 2 : P.nL;
        const int scopies = P.mrep ? 2 : 1;
        const int maxshift = (nt - 1) * BX + nt - 1;
        const long long L = banded ? (long long)TH * BX : (long long)BI * BY * BX;
        const long long L16 = (L + 15) / 16 * 16;
        P.BI = BI; P.TH = TH; P.BY = BY; P.SBY = SBY; P.NA = NA;
        P.small_bytes = BI * SBY * BX * 16;
        P.large_bytes = BI * BY * BX * 16;
        if (grp) {
            P.PS_s = P.small_bytes;
            P.PS_l = P.large_bytes;
            P.offL = scopies * P.nS * P.PS_s;
            // behind the last large plane: what the last K steps and the tap shifts read past it (stays zero: no box writes it)
            P.stage_bytes = (int)(((long long)P.offL + (long long)P.nL_cta * P.PS_l + (maxshift + 16) * 16 + 1023) / 1024 * 1024);
        } else {
            P.PS_s = (int)((L16 * 16 + 127) / 128 * 128);
            P.PS_l = (int)(((L16 + maxshift) * 16 + 127) / 128 * 128);
            P.offL = P.nS * P.PS_s;
            P.stage_bytes = P.offL + P.nL * P.PS_l;
        }
        P.nksteps = (int)(L16 / 16);
        const long long span = 16LL * P.PS_s + 0;    // an M=128 descriptor walks 16 chunk planes from the stage start
        const long long last = std::max<long long>(P.stage_bytes, span);
        P.zero_bytes = (int)((NA - 1) * (long long)P.stage_bytes + last);
        return P.zero_bytes;
    };
    const long long avail = SMEM_TOTAL;
    const int Hs = a->Hs;
    bool done = false;
    // row-tap stacking doubles the small tile: when that only fits single-buffered it is still the better plan (fewer MMAs and
    // TMEM passes; the next tile is pulled into L2 while this one computes) — otherwise plan without it
    const int NG_plain = P.s2d_cq ? nt * nt : k * nt, NG_stack = 2 * ((k + 3) / 4) * nt;
    if (can_mrep && can_psplit && NG_stack * P.N > 512 && (NG_stack / 2) * P.N <= 512) {
        // stacked row taps AND the passes split by row parity: each parity's groups fit one TMEM pass and its CTA loads the
        // doubled small tile plus only half of the large tile — double-buffered (D3: 3 passes of 18 groups -> 2 of 6)
        want_psplit = true;
        if (plan(1, Hs, false, 2) <= avail && P.mrep && P.psplit) {
            // keep; the whole-image search below refines the image count
        } else {
            want_psplit = false;
        }
    }
    if (!want_psplit && can_psplit && !(can_mrep && NG_stack * P.N <= 512) && NG_plain * P.N > 512) {
        want_mrep = false;              // several passes anyway: split them by row parity (each loads half of the large tile)
        want_psplit = true;
        if (!(plan(1, Hs, false, 2) <= avail && P.psplit)) want_psplit = false, want_mrep = can_mrep;
    }
    if (can_mrep && !want_psplit) {
        if (plan(1, Hs, false, 2) <= avail && P.mrep) {
            // fits double-buffered: the general search below keeps it
        } else if (2 * ((k + 3) / 4) * nt * P.N <= 512 && plan(1, Hs, false, 1) <= avail && P.mrep) {
            // (only when every tap group then fits one TMEM pass: measured E2 1.57 -> 1.30 ms, but D3 — still two passes,
            // each with its tile load exposed — 3.53 -> 3.67 ms)
            P.n_bands = 1;
            done = true;
        } else {
            want_mrep = false;
        }
    }
    // whole images, double buffered, as many images per tile as fit (capped so a tile stays a few thousand pixels)
    if (!done && plan(1, Hs, false, 2) <= avail) {
        int BI = 1;
        while (BI < std::min(a->n_img, 256) && (long long)(BI + 1) * (Hs + nt - 1) * BXmin <= 4096 && plan(BI + 1, Hs, false, 2) <= avail) ++BI;
        plan(BI, Hs, false, 2);
        P.n_bands = 1;
        done = true;
    }
    if (!done) {
        int TH = Hs;
        while (TH > 1 && plan(1, TH, true, 2) > avail) --TH;
        MRSSM_CHECK(plan(1, TH, true, 2) <= avail, "plane wgrad: a single row band does not fit shared memory");
        P.n_bands = (Hs + TH - 1) / TH;
        TH = (Hs + P.n_bands - 1) / P.n_bands;
        plan(1, TH, true, 2);
    }
    MRSSM_CHECK(P.PS_s / 16 < 16384 && P.PS_l / 16 < 16384, "plane wgrad: plane stride too large");
    // tap groups: (kh, kw / 2) — with row-tap stacking only the base taps kh = 4 j + parity (each MMA also carries kh + 2)
    P.NG = P.s2d_cq ? nt * nt : (P.mrep ? 2 * ((k + 3) / 4) * nt : k * nt);
    if (P.psplit) {             // per row parity: kh = par, par + 2 (+ 4 stacked), ..; the even parity has the most taps
        const int ng_par = (P.mrep ? (k + 3) / 4 : (k + 1) / 2) * nt;
        P.gpp = std::min(ng_par, 512 / P.N);
        const int n_sub = (ng_par + P.gpp - 1) / P.gpp;
        P.gpp = (ng_par + n_sub - 1) / n_sub;
        P.n_cpass = 2 * n_sub;
    } else {
        P.gpp = std::min(P.NG, 512 / P.N);
        P.n_cpass = (P.NG + P.gpp - 1) / P.gpp;
        P.gpp = (P.NG + P.n_cpass - 1) / P.n_cpass;
    }
    P.n_groups = (a->n_img + P.BI - 1) / P.BI;
    smem_bytes = (size_t)P.zero_bytes + 1024;
    const int n_tiles = P.n_groups * P.n_bands;
    const int ypass = P.n_cpass * P.n_mhalf;
    splits = std::max(1, std::min(n_tiles, g_sm_budget / ypass));      // the whole grid is one wave (one CTA per SM)
    P.cs_valid = a->cs_valid; P.cl_valid = a->cl_valid;
    P.dw = a->dweight; P.w_ss = a->w_ss; P.w_sl = a->w_sl;
    P.scale_ptr = a->scale_ptr; P.scale_mul = a->scale_mul;
    P.poll = g_dbg[0];
    P.db = a->dbias; P.db_from = a->dbias ? a->dbias_from : 0;
    if (P.db_from) {
        MRSSM_CHECK(P.db_from == 1 || P.db_from == 2, "plane wgrad: dbias_from %d (1 = small, 2 = large)", P.db_from);
        if (P.db_from == 1) {
            P.db_nch = P.nS; P.db_npar = 1; P.db_valid = a->cs_valid; P.db_fold = 0;
        } else {
            P.db_nch = P.cpl; P.db_npar = (P.s2d_cq || P.rep) ? 1 : (P.psplit ? 2 : 4); P.db_valid = a->cl_valid; P.db_fold = P.s2d_cq;
        }
        MRSSM_CHECK(P.db_nch == 1 || P.db_nch == 2 || P.db_nch == 4 || P.db_nch == 8 || P.db_nch == 16,
                    "plane wgrad: bias-gradient sums need 1, 2, 4, 8 or 16 channel chunks (got %d)", P.db_nch);
        MRSSM_CHECK(P.n_bands == 1 || P.BI == 1, "plane wgrad: banded tiles hold one image");
    }
    return 0;
}

int launch_wgrad(const mrssm_pl_conv_args* a, cudaStream_t st) {
    MRSSM_CHECK(a && a->large.ptr && a->small.ptr && a->dweight, "plane wgrad: null tensor");
    MRSSM_CHECK(a->ksz >= 2, "plane wgrad: kernel size %d (dense layers use mrssm_tc_conv_wgrad)", a->ksz);
    WgP P;
    size_t smem;
    int splits;
    if (int rc = plan_wgrad(a, P, smem, splits)) return rc;
    CUtensorMap mS[4], mL[4];
    if (int rc = make_view_maps(mS, &P.mergedS, a->small, a->Hs, a->Ws, a->Cs, a->n_img, false, P.BX, P.SBY, P.BI, P.group ? P.nS : 0)) return rc;
    if (P.s2d_cq) {
        if (int rc = make_view_maps(mL, &P.mergedL, a->large, (a->Hl + 1) / 2, (a->Wl + 1) / 2, a->Cl, a->n_img, false, P.BX, P.BY, P.BI, P.group ? P.cpl : 0)) return rc;
    } else if (int rc = make_view_maps(mL, &P.mergedL, a->large, a->Hl, a->Wl, a->Cl, a->n_img, true, P.BX, P.BY, P.BI, P.group ? P.cpl : 0)) {
        return rc;
    }
    dim3 grid((unsigned)splits, (unsigned)(P.n_cpass * P.n_mhalf));
    smem = std::max<size_t>(smem, 120 * 1024);
    MRSSM_CUDA(cudaFuncSetAttribute(plane_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    plane_wgrad_kernel<<<grid, NTHREADS, smem, st>>>(mS[0], mL[0], mL[1], mL[2], mL[3], P);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int mrssm_pl_conv_down(const mrssm_pl_conv_args* a, void* stream) { return launch_fwd(a, OP_DOWN, (cudaStream_t)stream); }
extern "C" int mrssm_pl_conv_up(const mrssm_pl_conv_args* a, void* stream) { return launch_fwd(a, OP_UP, (cudaStream_t)stream); }
extern "C" int mrssm_pl_conv_wgrad(const mrssm_pl_conv_args* a, void* stream) { return launch_wgrad(a, (cudaStream_t)stream); }

extern "C" int mrssm_pl_packed_shape(int32_t op, int32_t Cs_pad, int32_t Cl_pad, int32_t ksz, int32_t* N_total, int32_t* K_total) {
    const int nt = (ksz + 1) / 2;
    MRSSM_CHECK(N_total && K_total && ksz >= 2 && op >= 0 && op <= 2, "pl_packed_shape: bad args");
    int n_ksteps;
    if (op == 2) {
        MRSSM_CHECK(Cl_pad % 16 == 0 && Cs_pad % 16 == 0, "pl_packed_shape(down, space-to-depth): Cl_pad %% 16, Cs_pad %% 16");
        n_ksteps = nt * nt * (Cl_pad / 16);
        *N_total = Cs_pad;
    } else if (op == OP_DOWN) {
        MRSSM_CHECK(Cl_pad % 8 == 0 && Cs_pad % 16 == 0, "pl_packed_shape(down): Cl_pad %% 8, Cs_pad %% 16");
        n_ksteps = ksz * nt * (Cl_pad / 8);
        *N_total = Cs_pad;
    } else {
        MRSSM_CHECK(Cs_pad % 16 == 0 && Cl_pad % 8 == 0, "pl_packed_shape(up): Cs_pad %% 16, Cl_pad %% 8");
        n_ksteps = nt * nt * (Cs_pad / 16);
        *N_total = 4 * Cl_pad;
    }
    *K_total = (n_ksteps + 3) / 4 * 64;
    return 0;
}

extern "C" int mrssm_pl_pack_weight(const float* w, int64_t w_ss, int64_t w_sl, int32_t Cs_valid, int32_t Cl_valid, int32_t Cs_pad,
                                    int32_t Cl_pad, int32_t ksz, int32_t op, int32_t s2d_cq, void* out, void* stream) {
    MRSSM_CHECK(op != 2 || (s2d_cq > 0 && 4 * s2d_cq <= Cl_pad), "pl_pack_weight: op 2 needs 0 < 4*s2d_cq <= Cl_pad");
    int32_t N_total, K_total;
    if (int rc = mrssm_pl_packed_shape(op, Cs_pad, Cl_pad, ksz, &N_total, &K_total)) return rc;
    MRSSM_CHECK(w && out, "pl_pack_weight: null pointer");
    const long long total = (long long)N_total * K_total;
    const int blocks = (int)std::min<long long>(148 * 8, ceil_div64(total, 256));
    pack_plane_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, w_ss, w_sl, Cs_valid, Cl_valid, Cs_pad, Cl_pad, ksz, op, s2d_cq, N_total, K_total,
                                                               (bf16*)out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// layout helpers: fp32 strided -> bf16 view (import), per-channel sums over a view (bias gradients)
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void import_view_kernel(T4 src, int n, int H, int W, int Cc, int nchunk, float scale, TV dst) {
    const long long total = (long long)n * H * W * nchunk;
    const float* sp0 = (const float*)src.p;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        long long t = i / W;
        const int y = (int)(t % H);
        t /= H;
        const int ch = (int)(t % nchunk), img = (int)(t / nchunk);
        const float* sp = sp0 + img * src.sI + y * src.sH + x * src.sW;
        uint4 pk;
        __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = ch * 8 + 2 * e;
            const float a = c < Cc ? sp[(long long)c * src.sC] * scale : 0.f;
            const float b = c + 1 < Cc ? sp[(long long)(c + 1) * src.sC] * scale : 0.f;
            ph[e] = __floats2bfloat162_rn(a, b);
        }
        *reinterpret_cast<uint4*>((bf16*)dst.p + tv_pix(dst, img, y, x) + ch * dst.sK) = pk;
    }
}

// bf16 view -> bf16 view (same logical [n,H,W,C] tensor, different layout): 16 bytes per (pixel, chunk)
__global__ void copy_view_kernel(TV src, int n, int H, int W, int nchunk, TV dst) {
    const long long total = (long long)n * H * W * nchunk;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % nchunk);
        long long t = i / nchunk;
        const int x = (int)(t % W);
        t /= W;
        const int y = (int)(t % H), img = (int)(t / H);
        const uint4 pk = __ldg(reinterpret_cast<const uint4*>((const bf16*)src.p + tv_pix(src, img, y, x) + ch * src.sK));
        *reinterpret_cast<uint4*>((bf16*)dst.p + tv_pix(dst, img, y, x) + ch * dst.sK) = pk;
    }
}

// fp32 strided [n,H,W,C<=4] -> space-to-depth bf16 view [n,ceil(H/2),ceil(W/2),16]: channel = (py*2+px)*C + c, rest zero
__global__ void import_s2d_kernel(T4 src, int n, int H, int W, int Cc, float scale, TV dst) {
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const long long total = (long long)n * H2 * W2 * 2;
    const float* sp0 = (const float*)src.p;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x2 = (int)(i % W2);
        long long t = i / W2;
        const int y2 = (int)(t % H2);
        t /= H2;
        const int ch = (int)(t & 1), img = (int)(t >> 1);
        const float* sp = sp0 + img * src.sI;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int cidx = ch * 8 + e, par = cidx / Cc, c = cidx - par * Cc;
            const int y = 2 * y2 + (par >> 1), x = 2 * x2 + (par & 1);
            f[e] = (par < 4 && y < H && x < W) ? sp[y * src.sH + x * src.sW + (long long)c * src.sC] * scale : 0.f;
        }
        uint4 pk;
        __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) ph[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
        *reinterpret_cast<uint4*>((bf16*)dst.p + tv_pix(dst, img, y2, x2) + ch * dst.sK) = pk;
    }
}

// same, for x-contiguous sources (src.sW == 1, W even, 3 channels... any C <= 4): one thread per s2d pixel reads the 2x2
// patch of every channel with float2 loads (a warp reads 256 contiguous bytes per row) and writes both 8-channel chunks
__global__ void import_s2d_rows_kernel(T4 src, int n, int H, int W, int Cc, float scale, TV dst) {
    const int H2 = (H + 1) / 2, W2 = W / 2;
    const long long total = (long long)n * H2 * W2;
    const float* sp0 = (const float*)src.p;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x2 = (int)(i % W2);
        const long long t = i / W2;
        const int y2 = (int)(t % H2), img = (int)(t / H2);
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = 0.f;
        const float* sp = sp0 + img * src.sI + 2 * x2;
#pragma unroll
        for (int py = 0; py < 2; ++py) {
            const int y = 2 * y2 + py;
            if (y < H) {
                for (int c = 0; c < Cc; ++c) {
                    const float2 v = *reinterpret_cast<const float2*>(sp + y * src.sH + (long long)c * src.sC);
                    f[(py * 2) * Cc + c] = v.x * scale;
                    f[(py * 2 + 1) * Cc + c] = v.y * scale;
                }
            }
        }
        bf16* dp = (bf16*)dst.p + tv_pix(dst, img, y2, x2);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint4 pk;
            __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) ph[e] = __floats2bfloat162_rn(f[ch * 8 + 2 * e], f[ch * 8 + 2 * e + 1]);
            *reinterpret_cast<uint4*>(dp + ch * dst.sK) = pk;
        }
    }
}

// flat variant for tiny spatial extents (H*W < 64): one thread per (image, pixel), generic index arithmetic
__global__ void colsum_view_flat_kernel(TV t, int n, int H, int W, int Cvalid, int fold, const float* scale_ptr, float scale_mul,
                                        float* __restrict__ out) {
    const int ch = blockIdx.y;
    const long long total = (long long)n * H * W;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const long long r = i / W;
        const int y = (int)(r % H), img = (int)(r / H);
        const uint4 pk = __ldg(reinterpret_cast<const uint4*>((const bf16*)t.p + tv_pix(t, img, y, x) + ch * t.sK));
        const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(ph[e]);
            acc[2 * e] += f.x;
            acc[2 * e + 1] += f.y;
        }
    }
    __shared__ float red[8][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float s = warp_sum(acc[e]);
        if (lane == 0) red[warp][e] = s;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
        if (scale_ptr) s *= __ldg(scale_ptr) * scale_mul;
        const int c = ch * 8 + threadIdx.x;
        if (fold > 0) {
            if (c < 4 * fold && c % fold < Cvalid) atomicAdd(out + c % fold, s);
        } else if (c < Cvalid) {
            atomicAdd(out + c, s);
        }
    }
}

// Per-channel sums over a view.  blockIdx.y = 8-channel chunk; blocks stride over (image, parity plane) pairs and walk
// each plane row by row (a row is a contiguous run of W*16 bytes in the planar layouts): no per-element index arithmetic.
// fold > 0: channel c of the view is channel (c % fold) of the logical tensor for c < 4*fold (space-to-depth views)
__global__ void colsum_view_kernel(TV t, int n, int H, int W, int Cvalid, int fold, const float* scale_ptr, float scale_mul,
                                   float* __restrict__ out) {
    const int ch = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int npar = t.par ? 4 : 1;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // one (image, parity plane) per warp at a time; lanes sweep the plane's valid pixels (rows are contiguous runs)
    for (long long pl = (long long)blockIdx.x * nwarps + warp; pl < (long long)n * npar; pl += (long long)gridDim.x * nwarps) {
        const int img = (int)(pl / npar), par = (int)(pl - (long long)img * npar);
        int Hv = H, Wv = W;
        if (t.par) {
            Hv = (H - (par >> 1) + 1) / 2;
            Wv = (W - (par & 1) + 1) / 2;
        }
        const float invW = 1.f / (float)Wv;
        const bf16* base = (const bf16*)t.p + img * t.sI + par * t.sP + ch * t.sK;
        const int cnt = Hv * Wv;
#pragma unroll 4
        for (int i = lane; i < cnt; i += 32) {
            const int y = __float2int_rz(((float)i + 0.5f) * invW), x = i - y * Wv;
            const uint4 pk = __ldg(reinterpret_cast<const uint4*>(base + y * t.sH + x * t.sW));
            const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(ph[e]);
                acc[2 * e] += f.x;
                acc[2 * e + 1] += f.y;
            }
        }
    }
    __shared__ float red[8][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float s = warp_sum(acc[e]);
        if (lane == 0) red[warp][e] = s;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
        if (scale_ptr) s *= __ldg(scale_ptr) * scale_mul;
        const int c = ch * 8 + threadIdx.x;
        if (fold > 0) {
            if (c < 4 * fold && c % fold < Cvalid) atomicAdd(out + c % fold, s);
        } else if (c < Cvalid) {
            atomicAdd(out + c, s);
        }
    }
}
}  // namespace

extern "C" int mrssm_pl_import(const mrssm_t4* src, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t Cpad, float scale,
                               const mrssm_tv* dst, void* stream) {
    MRSSM_CHECK(src && src->ptr && dst && dst->ptr && Cpad % 8 == 0 && Cpad >= C, "pl_import: bad args");
    const long long total = (long long)n_img * H * W * (Cpad / 8);
    const int blocks = (int)std::min<long long>(148 * 16, ceil_div64(total, 256));
    import_view_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cvt(*src), n_img, H, W, C, Cpad / 8, scale, cvt(*dst));
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_pl_copy(const mrssm_tv* src, int32_t n_img, int32_t H, int32_t W, int32_t Cpad, const mrssm_tv* dst, void* stream) {
    MRSSM_CHECK(src && src->ptr && dst && dst->ptr && Cpad % 8 == 0 && Cpad > 0, "pl_copy: bad args");
    const long long total = (long long)n_img * H * W * (Cpad / 8);
    const int blocks = (int)std::min<long long>(148 * 16, ceil_div64(total, 256));
    copy_view_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cvt(*src), n_img, H, W, Cpad / 8, cvt(*dst));
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_pl_import_s2d(const mrssm_t4* src, int32_t n_img, int32_t H, int32_t W, int32_t C, float scale, const mrssm_tv* dst,
                                   void* stream) {
    MRSSM_CHECK(src && src->ptr && dst && dst->ptr && C >= 1 && C <= 4, "pl_import_s2d: bad args (C must be <= 4)");
    const long long total = (long long)n_img * ((H + 1) / 2) * ((W + 1) / 2) * 2;
    const int blocks = (int)std::min<long long>(148 * 16, ceil_div64(total, 256));
    if (src->sW == 1 && W % 2 == 0 && src->sH % 2 == 0 && src->sC % 2 == 0 && src->sI % 2 == 0 && ((uintptr_t)src->ptr & 7) == 0) {
        const long long tot2 = (long long)n_img * ((H + 1) / 2) * (W / 2);
        const int b2 = (int)std::min<long long>(148 * 16, ceil_div64(tot2, 256));
        import_s2d_rows_kernel<<<b2, 256, 0, (cudaStream_t)stream>>>(cvt(*src), n_img, H, W, C, scale, cvt(*dst));
    } else {
        import_s2d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cvt(*src), n_img, H, W, C, scale, cvt(*dst));
    }
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_pl_colsum(const mrssm_tv* x, int32_t n_img, int32_t H, int32_t W, int32_t Cpad, int32_t Cvalid, int32_t fold,
                               const float* scale_ptr, float scale_mul, float* out, void* stream) {
    MRSSM_CHECK(x && x->ptr && out && Cpad % 8 == 0 && Cvalid <= Cpad && (fold == 0 || 4 * fold <= Cpad), "pl_colsum: bad args");
    const int nchunk = fold > 0 ? (4 * fold + 7) / 8 : (Cvalid + 7) / 8;
    const long long planes = (long long)n_img * (x->par ? 4 : 1);
    const int bx = (int)std::max<long long>(1, std::min<long long>(ceil_div64(planes, 8), std::max(1, 1184 / nchunk)));
    dim3 grid((unsigned)bx, (unsigned)nchunk);
    if (H * W < 64) {
        const long long total = (long long)n_img * H * W;
        dim3 g2((unsigned)std::max<long long>(1, std::min<long long>(ceil_div64(total, 256 * 8), std::max(1, 1184 / nchunk))), (unsigned)nchunk);
        colsum_view_flat_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(cvt(*x), n_img, H, W, Cvalid, fold, scale_ptr, scale_mul, out);
    } else {
        colsum_view_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cvt(*x), n_img, H, W, Cvalid, fold, scale_ptr, scale_mul, out);
    }
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_pl_set_debug(int32_t key, int32_t value) {
    MRSSM_CHECK(key >= 0 && key < 8, "pl_set_debug: bad key");
    g_dbg[key] = value;
    return 0;
}

extern "C" int mrssm_pl_set_sm_budget(int32_t n_sm) {
    MRSSM_CHECK(n_sm >= 8 && n_sm <= 148, "pl_set_sm_budget: %d SMs (8 .. 148)", n_sm);
    g_sm_budget = n_sm;
    return 0;
}

extern "C" int mrssm_pl_set_plan_override(int32_t BI, int32_t TH, int32_t NA, int32_t bres, int32_t NB) {
    g_ovr = PlanOvr{BI, TH, NA, bres, NB};
    return 0;
}

extern "C" int mrssm_pl_set_profile_buffer(void* dev_buf) {
    g_prof = (long long*)dev_buf;
    return 0;
}

// Host-only: describe the tiling plan of a layer (no GPU needed; used by the CPU tests and for tuning).
extern "C" int mrssm_pl_describe(const mrssm_pl_conv_args* a, int32_t op, char* buf, int32_t buflen) {
    MRSSM_CHECK(a && buf && buflen > 0, "pl_describe: bad args");
    if (op == 2) {
        WgP P;
        size_t smem;
        int splits;
        if (int rc = plan_wgrad(a, P, smem, splits)) return rc;
        snprintf(buf, buflen,
                 "wgrad BI=%d bands=%d TH=%d BX=%d BY=%d nS=%d nL=%d PS_s=%d PS_l=%d stage=%d NA=%d ksteps/tile=%d N=%d groups=%d gpp=%d cpass=%d mhalf=%d "
                 "splits=%d smem=%zu tiles=%d rep=%d group=%d mrep=%d psplit=%d",
                 P.BI, P.n_bands, P.TH, P.BX, P.BY, P.nS, P.nL, P.PS_s, P.PS_l, P.stage_bytes, P.NA, P.nksteps, P.N, P.NG, P.gpp, P.n_cpass,
                 P.n_mhalf, splits, smem, P.n_groups * P.n_bands, P.rep, P.group, P.mrep, P.psplit);
    } else {
        FwdP P;
        size_t smem;
        if (int rc = plan_fwd(a, op, P, smem)) return rc;
        double eff = (double)P.BI * std::min(P.TH, P.Hv) * P.Wv / ((double)P.MB_total * 128);
        snprintf(buf, buflen,
                 "%s BI=%d bands=%d TH=%d BX=%d BY=%d planes=%d PS=%d stage=%d NA=%d NB=%d bres=%d BN=%d ntiles=%d ksteps=%d MB_total=%d MBs=%d passes=%d "
                 "smem=%zu tiles=%d group=%d row_eff=%.2f",
                 op == OP_DOWN ? "down" : "up", P.BI, P.n_bands, P.TH, P.BX, P.BY, P.planes, P.PS, P.a_stage_bytes, P.NA, P.NB, P.b_res, P.BN, P.n_ntiles,
                 P.n_ksteps, P.MB_total, P.MBs, P.n_passes, smem, P.n_groups * P.n_bands, P.group, eff);
    }
    return 0;
}
