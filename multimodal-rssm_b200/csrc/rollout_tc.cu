// Fused RSSM rollout, forward, on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 operands, fp32 accumulate and
// fp32 recurrent state.  Replaces the same reference code as rollout_simt.cu — utils/models/transition_model.py:
// 200-285 (+ :50-114), nn.GRUCell (:160,235), the Gaussian heads of encoder.py:126-190, poe / MoPoE fusion
// (encoder.py:50-124) and the rsample calls — for the sizes of the shipped configs (D, H <= 208, S <= 32,
// S + A <= 48, <= 4 heads); mrssm_rollout_tc_eligible() says whether it applies.
//
// One CTA owns ROWS = 64 sequences for all T steps (tcgen05 M = 64: accumulator row i lives in TMEM lane
// 32*(i/16) + i%16, profiles/micro/tmem_layout.cu).  Per step the chain
//   xin=[s*mask,a] -> fc_embed -> GRU gates -> (1+E) heads fc1 -> fc2 -> fusion / rsample -> next xin
// is a fixed PROGRAM of tcgen05.mma instructions built on the host (mrssm_rollout_tc_plan): every MMA names its
// A operand (an activation buffer in shared memory, un-swizzled K-major chunk planes [chunk][row][8]), its B operand
// (a block of the packed bf16 weight image), its TMEM columns and its N.  The packed weight image (one contiguous
// byte stream per step, identical for every step, ~1.1 MB) is streamed through a shared-memory ring by cp.async.bulk
// — the producer runs ahead across phase and step boundaries.  Roles: warp 0 = weight producer, warp 1 = MMA issuer,
// warp 2 = TMEM allocator, warps 4..19 = epilogue (tcgen05.ld.16x256b so all 32 lanes carry data: thread t holds
// rows t/4 and t/4+8, columns 2(t%4), 2(t%4)+1 of every 8-column group).  The epilogues do the gate math, softplus,
// PoE / MoPoE fusion and rsample in fp32, write the API outputs and the BPTT stash to HBM, and write the NEXT
// operand (bf16) straight into the activation buffer the following MMAs read.  h is kept in fp32 in shared memory;
// only its bf16 copy feeds the tensor cores.
#include <algorithm>
#include <string.h>
#include <vector>
#include "tc_common.cuh"

namespace {

using bf16 = __nv_bfloat16;

constexpr int ROWS = 64;                  // sequences per CTA = MMA M
constexpr int CH_BYTES = ROWS * 16;       // one 8-feature chunk plane of an A operand
constexpr int MAXC = 26;                  // chunk planes per activation buffer (D, H <= 208)
constexpr int XIN_CH = 6;                 // S + A <= 48
constexpr int SLOT_BYTES = 14336;         // ring slot: 4 N=112 blocks, 2 N=208 blocks, 7 N=64 blocks
constexpr int MAX_SLOTS = 8;
constexpr int TC_THREADS = 608;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 16;
constexpr int ACC_STRIDE = 112;           // TMEM columns per gate accumulator / per fc1 set
constexpr int F2_COL = 2 * ACC_STRIDE;    // fc2 accumulators: 64 columns per head
constexpr int TC_MAX_HEADS = 4;
constexpr int MAX_TILES = 112, MAX_OPS = 336, MAX_RUNS = 192;
constexpr int TC_MAGIC = 0x52544331;      // "RTC1"
constexpr int PACK_TRANSPOSED = 0x100;

enum { BUF_XIN = 0, BUF_XU = 1, BUF_HPREV = 2, BUF_HNEW = 3 };
enum { EV_XIN = 0, EV_X, EV_HA, EV_HB, EV_U0, N_EV = EV_U0 + 2 * TC_MAX_HEADS };
enum { CM_X = 0, CM_GA, CM_GB, CM_F1, CM_F2 = CM_F1 + 2 * TC_MAX_HEADS, N_CM = CM_F2 + TC_MAX_HEADS };
enum { SRC_WSA = 0, SRC_WIH, SRC_WHH, SRC_W1, SRC_W2 = SRC_W1 + TC_MAX_HEADS, N_SRC = SRC_W2 + TC_MAX_HEADS };

// dynamic shared memory map (bytes from the 1024-aligned base)
constexpr int OFF_XIN = 0;
constexpr int OFF_XU = OFF_XIN + XIN_CH * CH_BYTES;
constexpr int OFF_H0 = OFF_XU + MAXC * CH_BYTES;
constexpr int OFF_H1 = OFF_H0 + MAXC * CH_BYTES;
constexpr int OFF_HF = OFF_H1 + MAXC * CH_BYTES;          // fp32 h: [chunk][row][8]
constexpr int OFF_RING = OFF_HF + MAXC * ROWS * 32;
constexpr int ZERO_BYTES = OFF_HF;                         // operand buffers zeroed at start (padding chunks stay zero)

struct TcTile {
    uint32_t src_off, bytes;              // byte range of the packed image
    uint16_t op_begin, op_end;
    int8_t wait_ev, commit;               // event to wait for before the first MMA / commit barrier after the last (-1: none)
    uint8_t run_begin, run_end;           // the same MMAs as runs (what the kernel issues)
};
struct TcOp {
    uint16_t a_off16;                     // A operand: byte offset >> 4 inside its buffer
    uint8_t a_buf, acc;
    uint16_t b_off16, d_col;              // B block offset >> 4 inside the ring slot; TMEM column
    uint32_t idesc;
};
struct TcRun {                            // `count` MMAs into one accumulator: A advances 2 chunk planes, B one block per MMA
    uint16_t a_off16;
    uint8_t a_buf, acc_first;
    uint16_t b_off16, d_col;
    uint32_t idesc;
    uint32_t count;
};
struct TcPack {                           // how one B block [2][N][8] is gathered from an fp32 PyTorch-layout weight
    int32_t src_id, k0, kvalid, N;
    int32_t seg_n[2], seg_src[2], seg_cnt[2];
    uint32_t dst_off;
    int32_t pad;
};
struct TcHeader {
    int32_t magic, D, S, H, A, NH, n_tiles, n_ops, n_pack, cA, nD8, cAH, nH8;
    uint32_t packed_bytes, tile_off, op_off, pack_off, total_bytes;
    uint32_t run_off;
    int32_t n_runs;
};
// What the kernel itself reads of the plan, passed BY VALUE (kernel parameters live in the constant bank: the MMA issuer and
// the producer index it with warp-uniform indices, so the loads and the descriptor arithmetic stay in uniform registers).
constexpr int PROG_RUNS = 128, PROG_TILES = 96;
struct TcProg {
    uint4 runs[PROG_RUNS];                // {A desc lo without smem base, B desc lo slot-relative, idesc, d_col | acc<<10 | flip<<11 | count<<13}
    uint32_t tiles[PROG_TILES];           // bytes/16 | n_runs<<10 | (wait_ev+1)<<13 | (commit+1)<<18
    int32_t n_tiles, cA, nD8, cAH, nH8;
};
struct PackSrc {
    const float* p[N_SRC];
    int32_t ld[N_SRC];
};

__host__ __device__ inline int ceil16(int v) { return (v + 15) / 16 * 16; }

// ---------------------------------------------------------------------------------------------------------------
// host: the per-step MMA program
// ---------------------------------------------------------------------------------------------------------------
struct PlanBuilder {
    std::vector<TcTile> tiles;
    std::vector<TcOp> ops;
    std::vector<TcPack> packs;
    std::vector<TcRun> runs;
    uint32_t packed = 0;
    int cur_wait = -1;
    bool first = true, open = false;

    void unit_begin(int wait_ev) { cur_wait = wait_ev; first = true; open = false; }
    void unit_end(int commit) { tiles.back().commit = (int8_t)commit; open = false; }
    // one MMA: D[d_col .. d_col+N) (+)= A(buf, chunk planes a_chunk, a_chunk+1) * Bblock^T, Bblock [N][16] gathered as `p` says
    void mma2(int a_buf, int a_chunk, int d_col, int N, int acc, TcPack p) {
        const uint32_t bytes = 32u * (uint32_t)N;
        if (!open || tiles.back().bytes + bytes > (uint32_t)SLOT_BYTES) {
            TcTile t;
            t.src_off = packed; t.bytes = 0;
            t.op_begin = t.op_end = (uint16_t)ops.size();
            t.wait_ev = (int8_t)(first ? cur_wait : -1);
            t.commit = -1;
            t.run_begin = t.run_end = (uint8_t)runs.size();
            tiles.push_back(t);
            first = false; open = true;
        }
        TcOp o;
        o.a_off16 = (uint16_t)(a_chunk * CH_BYTES / 16);
        o.a_buf = (uint8_t)a_buf; o.acc = (uint8_t)acc;
        o.b_off16 = (uint16_t)(tiles.back().bytes / 16);
        o.d_col = (uint16_t)d_col;
        o.idesc = tc::idesc_bf16(ROWS, N, 0, 0);
        ops.push_back(o);
        {
            TcTile& tl = tiles.back();
            bool ext = false;
            if (tl.run_end > tl.run_begin) {
                TcRun& r = runs.back();
                ext = acc && r.a_buf == o.a_buf && r.d_col == o.d_col && r.idesc == o.idesc && r.count < 7 &&
                      o.a_off16 == r.a_off16 + r.count * (2 * CH_BYTES / 16) && o.b_off16 == r.b_off16 + r.count * 2 * (uint32_t)N;
                if (ext) ++r.count;
            }
            if (!ext) {
                TcRun r;
                r.a_off16 = o.a_off16; r.a_buf = o.a_buf; r.acc_first = o.acc; r.b_off16 = o.b_off16; r.d_col = o.d_col;
                r.idesc = o.idesc; r.count = 1;
                runs.push_back(r);
                tl.run_end = (uint8_t)runs.size();
            }
        }
        p.N = N;
        p.dst_off = packed;
        packs.push_back(p);
        packed += bytes;
        tiles.back().bytes += bytes;
        tiles.back().op_end = (uint16_t)ops.size();
    }
    // forward-type block: Bblock[n][k] = W[row(n)][16 k + k'], rows through up to two segments of n
    void mma(int a_buf, int k, int d_col, int N, int acc, int src_id, int kin, int seg0_n, int seg0_src, int seg0_cnt, int seg1_n = 0,
             int seg1_src = 0, int seg1_cnt = 0) {
        TcPack p;
        memset(&p, 0, sizeof(p));
        p.src_id = src_id; p.k0 = 16 * k; p.kvalid = std::max(0, std::min(16, kin - 16 * k));
        p.seg_n[0] = seg0_n; p.seg_src[0] = seg0_src; p.seg_cnt[0] = seg0_cnt;
        p.seg_n[1] = seg1_n; p.seg_src[1] = seg1_src; p.seg_cnt[1] = seg1_cnt;
        mma2(a_buf, 2 * k, d_col, N, acc, p);
    }
    // transposed (dgrad-type) block: Bblock[n][k'] = W[row(16 k + k')][n_off + n] for n < n_cnt, rows through up to two
    // segments of the contraction index
    void mma_t(int a_chunk0, int k, int d_col, int N, int acc, int src_id, int n_off, int n_cnt, int seg0_k, int seg0_src, int seg0_cnt,
               int seg1_k = 0, int seg1_src = 0, int seg1_cnt = 0) {
        TcPack p;
        memset(&p, 0, sizeof(p));
        p.src_id = src_id | PACK_TRANSPOSED; p.k0 = 16 * k; p.kvalid = n_cnt; p.pad = n_off;
        p.seg_n[0] = seg0_k; p.seg_src[0] = seg0_src; p.seg_cnt[0] = seg0_cnt;
        p.seg_n[1] = seg1_k; p.seg_src[1] = seg1_src; p.seg_cnt[1] = seg1_cnt;
        mma2(0, a_chunk0 + 2 * k, d_col, N, acc, p);
    }
};

bool tc_eligible(int D, int S, int H, int A, int NH) {
    return D % 8 == 0 && H % 8 == 0 && D >= 16 && H >= 16 && D <= 8 * MAXC && H <= 8 * MAXC && S >= 1 && S <= 32 && A >= 0 &&
           S + A <= 8 * XIN_CH && NH >= 1 && NH <= TC_MAX_HEADS;
}

void build_plan(int D, int S, int H, int A, int NH, PlanBuilder& pb, TcHeader& h) {
    const int nD8 = D / 8, nH8 = H / 8, cA = (nD8 + 1) / 2, cAH = (nH8 + 1) / 2;
    const int nkD = (D + 15) / 16, nkH = (H + 15) / 16, nkX = (S + A + 15) / 16;
    // 1. fc_embed_state_action (transition_model.py:232-233)
    pb.unit_begin(EV_XIN);
    for (int k = 0; k < nkX; ++k) pb.mma(BUF_XIN, k, 0, ceil16(D), k != 0, SRC_WSA, S + A, 0, 0, D);
    pb.unit_end(CM_X);
    // 2. GRUCell gates in two column halves (r, z, i_n, h_n of the same features share TMEM)
    for (int hf = 0; hf < 2; ++hf) {
        const int c0 = hf ? cA : 0, nc = hf ? nD8 - cA : cA, n0 = 8 * c0, cnt = 8 * nc, N = ceil16(cnt);
        pb.unit_begin(hf ? EV_HA : EV_X);
        // r and z in ONE MMA of N = ACC_STRIDE + N (their accumulators are adjacent TMEM columns; the B block stacks the r rows
        // and the z rows): x and h contributions into the same accumulators
        for (int k = 0; k < nkD; ++k) pb.mma(BUF_XU, k, 0, ACC_STRIDE + N, k != 0, SRC_WIH, D, 0, n0, cnt, ACC_STRIDE, D + n0, cnt);
        for (int k = 0; k < nkD; ++k) pb.mma(BUF_HPREV, k, 0, ACC_STRIDE + N, 1, SRC_WHH, D, 0, n0, cnt, ACC_STRIDE, D + n0, cnt);
        for (int k = 0; k < nkD; ++k) pb.mma(BUF_XU, k, 2 * ACC_STRIDE, N, k != 0, SRC_WIH, D, 0, 2 * D + n0, cnt);
        for (int k = 0; k < nkD; ++k) pb.mma(BUF_HPREV, k, 3 * ACC_STRIDE, N, k != 0, SRC_WHH, D, 0, 2 * D + n0, cnt);
        pb.unit_end(hf ? CM_GB : CM_GA);
    }
    // 3./4. heads: fc1 (belief columns; the embedding columns are hoisted) in half-head units on two TMEM sets, fc2 per head
    auto fc1 = [&](int hd, int hf, int wait_ev) {
        const int c0 = hf ? cAH : 0, nc = hf ? nH8 - cAH : cAH, n0 = 8 * c0, cnt = 8 * nc, N = ceil16(cnt);
        pb.unit_begin(wait_ev);
        for (int k = 0; k < nkD; ++k) pb.mma(BUF_HNEW, k, hf * ACC_STRIDE, N, k != 0, SRC_W1 + hd, D, 0, n0, cnt);
        pb.unit_end(CM_F1 + 2 * hd + hf);
    };
    auto fc2 = [&](int hd) {
        pb.unit_begin(EV_U0 + 2 * hd + 1);
        for (int k = 0; k < nkH; ++k) pb.mma(BUF_XU, k, F2_COL + 64 * hd, 64, k != 0, SRC_W2 + hd, H, 0, 0, S, 32, S, S);
        pb.unit_end(CM_F2 + hd);
    };
    fc1(0, 0, EV_HB);
    fc1(0, 1, -1);
    for (int hd = 1; hd < NH; ++hd) {
        fc1(hd, 0, EV_U0 + 2 * hd - 2);
        fc2(hd - 1);
        fc1(hd, 1, -1);
    }
    fc2(NH - 1);

    memset(&h, 0, sizeof(h));
    h.magic = TC_MAGIC; h.D = D; h.S = S; h.H = H; h.A = A; h.NH = NH;
    h.n_tiles = (int)pb.tiles.size(); h.n_ops = (int)pb.ops.size(); h.n_pack = (int)pb.packs.size();
    h.cA = cA; h.nD8 = nD8; h.cAH = cAH; h.nH8 = nH8;
    h.packed_bytes = pb.packed;
    h.tile_off = 128;
    h.op_off = h.tile_off + (uint32_t)(pb.tiles.size() * sizeof(TcTile));
    h.pack_off = (h.op_off + (uint32_t)(pb.ops.size() * sizeof(TcOp)) + 15u) & ~15u;
    h.run_off = h.pack_off + (uint32_t)(pb.packs.size() * sizeof(TcPack));
    h.n_runs = (int)pb.runs.size();
    h.total_bytes = h.run_off + (uint32_t)(pb.runs.size() * sizeof(TcRun));
}

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void r_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void r_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void r_umma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 TMEM lanes x 8 columns: thread t gets (row t/4, cols 2(t%4), +1) in v[0], v[1] and (row t/4 + 8, same cols) in v[2], v[3]
__device__ __forceinline__ void ld_frag(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void st_shared_bf16(uint32_t addr, float v) {
    const unsigned short u = __bfloat16_as_ushort(__float2bfloat16(v));
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(u) : "memory");
}
__device__ __forceinline__ float2 ld_shared_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_f2(uint32_t addr, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_f2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

// rows x width fp32 block of a time-major [T, B, width] tensor -> L2 (one bulk prefetch; skipped when not 16-byte addressable)
__device__ __forceinline__ void prefetch_rows(const float* p, long long row0, int rows, int width) {
    if (!p || rows <= 0) return;
    const char* a = (const char*)(p + row0 * width);
    const uint32_t bytes = (uint32_t)rows * (uint32_t)width * 4u;
    if ((((uintptr_t)a) | bytes) & 15u) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes) : "memory");
}

// the epilogue warps signal "operand written / accumulator drained": one arrival per warp
__device__ __forceinline__ void epi_signal(uint64_t* bar, int lane) {
    tc::tc_fence_before();
    tc::fence_proxy_async();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(tc::smem_u32(bar));
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing: fp32 PyTorch-layout weights -> the bf16 B-operand stream of one step
// ---------------------------------------------------------------------------------------------------------------
__global__ void rollout_tc_pack_kernel(const uint8_t* __restrict__ plan, PackSrc src, bf16* __restrict__ out) {
    const TcHeader* h = reinterpret_cast<const TcHeader*>(plan);
    const TcPack pk = reinterpret_cast<const TcPack*>(plan + h->pack_off)[blockIdx.x];
    const float* w = src.p[pk.src_id & 0xff];
    const long long ld = src.ld[pk.src_id & 0xff];
    bf16* dst = out + pk.dst_off / 2;
    for (int i = threadIdx.x; i < pk.N * 16; i += blockDim.x) {
        const int ch = i / (pk.N * 8), n = (i >> 3) % pk.N, k = ch * 8 + (i & 7);
        float v = 0.f;
        if (pk.src_id & PACK_TRANSPOSED) {
            const int kk = pk.k0 + k;
            if (n < pk.kvalid) {
#pragma unroll
                for (int s = 0; s < 2; ++s)
                    if (kk >= pk.seg_n[s] && kk < pk.seg_n[s] + pk.seg_cnt[s]) v = w[(long long)(pk.seg_src[s] + kk - pk.seg_n[s]) * ld + pk.pad + n];
            }
        } else if (k < pk.kvalid) {
#pragma unroll
            for (int s = 0; s < 2; ++s)
                if (n >= pk.seg_n[s] && n < pk.seg_n[s] + pk.seg_cnt[s]) v = w[(long long)(pk.seg_src[s] + n - pk.seg_n[s]) * ld + pk.k0 + k];
        }
        dst[i] = __float2bfloat16(v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
// tuning aid: clock64 stamps of time step PROF_STEP of CTA 0 (role 0 producer, 1 MMA issuer, 2 first epilogue warp)
constexpr int PROF_STEP = 3, PROF_SLOTS = 512;   // prof buffer: 4 roles x PROF_SLOTS
#define STAMP(role) do { if (prof && blockIdx.x == 0 && t == PROF_STEP && pi < PROF_SLOTS) prof[(role) * PROF_SLOTS + pi++] = clock64(); } while (0)

// bf16-mode arithmetic: MUFU-based, absolute error ~1e-7 (the GEMM operands are rounded to bf16 anyway)
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950409f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - 2.f * rcp_approx(1.f + ex2_approx(2.8853900818f * x)); }
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : 0.6931471806f * lg2_approx(1.f + ex2_approx(1.4426950409f * x)); }
__device__ __forceinline__ float act_fast(float v, int act) {
    if (act == MRSSM_ACT_RELU) return fmaxf(v, 0.f);
    if (act == MRSSM_ACT_ELU) return v > 0.f ? v : ex2_approx(1.4426950409f * v) - 1.f;
    return v;
}
// waits of the many (epilogue) and the patient (producer): back off so the spinning warps leave the issue slots to the
// MMA warp and to the epilogue warps that have work
__device__ __forceinline__ void wait_backoff(uint64_t* bar, uint32_t parity) {
    const uint32_t b = tc::smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t it = 0; !done; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(b), "r"(parity)
            : "memory");
        if (!done) {
            __nanosleep(20);
            if (it > (1u << 22)) {
                printf("mrssm rollout_tc: barrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
                __trap();
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Statically unrolled MMA issue for the shipped model sizes: the same program build_plan / build_plan_bwd emit, walked at
// compile time so every descriptor offset, TMEM column, N and tile boundary is an immediate (one thread issues ~10
// instructions per MMA instead of ~40 through the run tables; the MMA issuer is the serial bottleneck of a step).  The
// tile split below MUST mirror PlanBuilder::mma2 (greedy fill of SLOT_BYTES per unit) — tests/test_rollout_tc_plan.py and the
// GPU parity tests cover both paths.
// ---------------------------------------------------------------------------------------------------------------
struct Issuer {
    uint32_t slot, phase, NS, ring16, tmem, par, sb;
    int it;
    uint64_t *full, *empty, *ev, *cm;
    int tile_bytes;
    bool open;
    __device__ __forceinline__ void unit_begin(int wev) {
        if (wev >= 0) {
            if (wev == 0) {
                if (it > 0) tc::mbar_wait(tc::smem_u32(&ev[0]), par ^ 1u);
            } else {
                tc::mbar_wait(tc::smem_u32(&ev[wev]), par);
            }
            tc::tc_fence_after();
        }
        open = false;
    }
    __device__ __forceinline__ void close_tile() {
        tc::umma_commit(tc::smem_u32(&empty[slot]));
        if (++slot == NS) { slot = 0; phase ^= 1u; }
    }
    __device__ __forceinline__ void unit_end(int cmi) {
        tc::umma_commit(tc::smem_u32(&empty[slot]));
        tc::umma_commit(tc::smem_u32(&cm[cmi]));
        if (++slot == NS) { slot = 0; phase ^= 1u; }
        open = false;
    }
    __device__ __forceinline__ void emit(uint32_t a16, int d_col, int N, int acc) {
        const int bytes = 32 * N;
        if (!open || tile_bytes + bytes > SLOT_BYTES) {
            if (open) close_tile();
            tc::mbar_wait(tc::smem_u32(&full[slot]), phase);
            tc::tc_fence_after();
            sb = ring16 + slot * (uint32_t)(SLOT_BYTES >> 4);
            tile_bytes = 0;
            open = true;
        }
        const uint32_t hi = (128u >> 4) | (1u << 14);
        r_umma(tmem + (uint32_t)d_col, a16 | (((uint32_t)CH_BYTES >> 4) << 16), (sb + (uint32_t)(tile_bytes >> 4)) | ((uint32_t)N << 16), hi,
               tc::idesc_bf16(ROWS, N, 0, 0), (uint32_t)acc);
        tile_bytes += bytes;
    }
};
constexpr uint32_t CH16 = CH_BYTES >> 4;
// opaque identity: keeps the compiler from hoisting every (buffer base + chunk offset) sum of the unrolled program out of the
// step loop; the add is redone next to its MMA (one uniform add) instead of living in a register across the whole step
__device__ __forceinline__ uint32_t opaque(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}

template <int D, int S, int H, int A, int NH>
__device__ __forceinline__ void static_step_fwd(Issuer& I, uint32_t xin16, uint32_t xu16, uint32_t hprev16, uint32_t hnew16) {
    constexpr int nD8 = D / 8, nH8 = H / 8, cA = (nD8 + 1) / 2, cAH = (nH8 + 1) / 2;
    constexpr int nkD = (D + 15) / 16, nkH = (H + 15) / 16, nkX = (S + A + 15) / 16, ND = (D + 15) / 16 * 16;
    I.unit_begin(EV_XIN);
#pragma unroll
    for (int k = 0; k < nkX; ++k) I.emit(opaque(xin16) + 2 * k * CH16, 0, ND, k != 0);
    I.unit_end(CM_X);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const int nc = hf ? nD8 - cA : cA, N = (8 * nc + 15) / 16 * 16;
        I.unit_begin(hf ? EV_HA : EV_X);
#pragma unroll
        for (int k = 0; k < nkD; ++k) I.emit(opaque(xu16) + 2 * k * CH16, 0, ACC_STRIDE + N, k != 0);
#pragma unroll
        for (int k = 0; k < nkD; ++k) I.emit(opaque(hprev16) + 2 * k * CH16, 0, ACC_STRIDE + N, 1);
#pragma unroll
        for (int k = 0; k < nkD; ++k) I.emit(opaque(xu16) + 2 * k * CH16, 2 * ACC_STRIDE, N, k != 0);
#pragma unroll
        for (int k = 0; k < nkD; ++k) I.emit(opaque(hprev16) + 2 * k * CH16, 3 * ACC_STRIDE, N, k != 0);
        I.unit_end(hf ? CM_GB : CM_GA);
    }
    auto fc1 = [&](int hd, int hf, int wev) {
        const int nc = hf ? nH8 - cAH : cAH, N = (8 * nc + 15) / 16 * 16;
        I.unit_begin(wev);
#pragma unroll
        for (int k = 0; k < nkD; ++k) I.emit(opaque(hnew16) + 2 * k * CH16, hf * ACC_STRIDE, N, k != 0);
        I.unit_end(CM_F1 + 2 * hd + hf);
    };
    auto fc2 = [&](int hd) {
        I.unit_begin(EV_U0 + 2 * hd + 1);
#pragma unroll
        for (int k = 0; k < nkH; ++k) I.emit(opaque(xu16) + 2 * k * CH16, F2_COL + 64 * hd, 64, k != 0);
        I.unit_end(CM_F2 + hd);
    };
    fc1(0, 0, EV_HB);
    fc1(0, 1, -1);
#pragma unroll
    for (int hd = 1; hd < NH; ++hd) {
        fc1(hd, 0, EV_U0 + 2 * hd - 2);
        fc2(hd - 1);
        fc1(hd, 1, -1);
    }
    fc2(NH - 1);
}

// Warp roles.  The MMA issuer is the highest warp id of its scheduler (the issue arbiter favours high warp ids).
constexpr int PROD_WARP = 16, ALLOC_WARP = 17, MMA_WARP = 18;

// NHS 0: table-driven MMA issue (any eligible size); > 0: statically unrolled issue for D=H=200, S=30, A=3, NHS heads.
// TWO: 64 sequences per CTA (both rows of every 16x256b fragment carry a sequence) instead of 32 (row t/4 only) — a template
// parameter so that the second row's registers and arithmetic do not exist in the 32-sequence kernels.
template <int NHS, bool TWO>
__global__ void __launch_bounds__(TC_THREADS, 1)
rollout_tc_fwd_kernel(const __grid_constant__ mrssm_rollout_args a, const __grid_constant__ TcProg prog, const uint8_t* __restrict__ packed,
                      const int NS, long long* __restrict__ prof) {
    constexpr int RPG = TWO ? 16 : 8;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full[MAX_SLOTS], empty[MAX_SLOTS], ev[N_EV], cm[N_CM];
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_x[8 * MAXC], bias_g[4][8 * MAXC], bias_1[TC_MAX_HEADS][8 * MAXC], bias_2[TC_MAX_HEADS][64];
    __shared__ uint32_t smask_s[32];                           // fusion subset (expert bitmask) of every state dim
    __shared__ const float* emb_pre_s[TC_MAX_HEADS];
    __shared__ float *st_u_s[TC_MAX_HEADS], *exp_means_s[TC_MAX_HEADS], *exp_stds_s[TC_MAX_HEADS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem = smem_raw + (smem0 - tc::smem_u32(smem_raw));
    const int D = a.D, S = a.S, H = a.H, A = a.A, E = a.n_experts, NH = 1 + E, B = a.B, T = a.T;
    const int n_tiles = prog.n_tiles, cA = prog.cA, nD8 = prog.nD8, cAH = prog.cAH, nH8 = prog.nH8;
    // MMA row r (0..63) <-> sequence b0 + (r/16)*RPG + r%16, valid when r%16 < RPG: with RPG = 8 a CTA owns 32 sequences and
    // every epilogue thread carries one valid row (half the elementwise work per CTA, twice the CTAs)
    const int b0 = blockIdx.x * 4 * RPG;

    // ---- prologue ---------------------------------------------------------------------------------------
    {
        for (int j = tid; j < 8 * MAXC; j += TC_THREADS) {
            const bool ok = j < D;
            bias_x[j] = ok ? a.b_sa[j] : 0.f;
            bias_g[0][j] = ok ? a.b_ih[j] + a.b_hh[j] : 0.f;
            bias_g[1][j] = ok ? a.b_ih[D + j] + a.b_hh[D + j] : 0.f;
            bias_g[2][j] = ok ? a.b_ih[2 * D + j] : 0.f;
            bias_g[3][j] = ok ? a.b_hh[2 * D + j] : 0.f;
            for (int h = 0; h < TC_MAX_HEADS; ++h) bias_1[h][j] = (h < NH && j < H && a.b1[h]) ? a.b1[h][j] : 0.f;
        }
        for (int j = tid; j < TC_MAX_HEADS * 64; j += TC_THREADS) {
            const int h = j >> 6, c = j & 63, s = c & 31;
            bias_2[h][c] = (h < NH && s < S) ? a.b2[h][(c >> 5) * S + s] : 0.f;
        }
        if (tid < 32) smask_s[tid] = (tid < S && a.n_subsets > 0) ? a.subset_mask[a.dim_subset[tid]] : 0u;
        if (tid < TC_MAX_HEADS) {
            emb_pre_s[tid] = tid < NH ? a.emb_pre[tid] : nullptr;
            st_u_s[tid] = tid < NH ? a.st_u[tid] : nullptr;
            exp_means_s[tid] = tid < NH ? a.exp_means[tid] : nullptr;
            exp_stds_s[tid] = tid < NH ? a.exp_stds[tid] : nullptr;
        }
        for (int i = tid; i < ZERO_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int s = 0; s < MAX_SLOTS; ++s) {
            tc::mbar_init(tc::smem_u32(&full[s]), 1);
            tc::mbar_init(tc::smem_u32(&empty[s]), 1);
        }
        for (int i = 0; i < N_EV; ++i) tc::mbar_init(tc::smem_u32(&ev[i]), EPI_WARPS);
        for (int i = 0; i < N_CM; ++i) tc::mbar_init(tc::smem_u32(&cm[i]), 1);
        tc::fence_barrier_init();
    }
    if (warp == ALLOC_WARP) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 512);
    __syncthreads();
    // h_{-1} (fp32 + bf16 operand) and xin of step 0
    for (int i = tid; i < ROWS * D; i += TC_THREADS) {
        const int r = i / D, j = i - r * D, seq = b0 + (r >> 4) * RPG + (r & 15);
        const float v = ((r & 15) < RPG && seq < B) ? a.prev_belief[(long long)seq * D + j] : 0.f;
        *reinterpret_cast<float*>(smem + OFF_HF + (j >> 3) * (ROWS * 32) + r * 32 + (j & 7) * 4) = v;
        *reinterpret_cast<bf16*>(smem + OFF_H0 + (j >> 3) * CH_BYTES + r * 16 + (j & 7) * 2) = __float2bfloat16(v);
    }
    for (int i = tid; i < ROWS * (S + A); i += TC_THREADS) {
        const int r = i / (S + A), j = i - r * (S + A), seq = b0 + (r >> 4) * RPG + (r & 15);
        float v = 0.f;
        if ((r & 15) < RPG && seq < B) {
            if (j < S) v = a.prev_state[(long long)seq * S + j] * (a.nonterminals ? a.nonterminals[seq] : 1.f);
            else v = a.actions[(long long)seq * A + (j - S)];
        }
        *reinterpret_cast<bf16*>(smem + OFF_XIN + (j >> 3) * CH_BYTES + r * 16 + (j & 7) * 2) = __float2bfloat16(v);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == PROD_WARP) {
        // ---- weight producer ------------------------------------------------------------------------------
        if (lane == 0) {
            int pi = 0;
            uint32_t slot = 0, phase = 0;
            const int prows = min(4 * RPG, B - b0);
            for (int t = 0; t < T; ++t) {
                // the next step's inputs of this CTA (hoisted embedding rows, noise) -> L2 while this step streams its weights
                for (int tt = (t == 0 ? 0 : t + 1); tt <= t + 1 && tt < T; ++tt) {
                    const long long r0 = (long long)tt * B + b0;
                    for (int h = 0; h < NH; ++h) prefetch_rows(emb_pre_s[h], r0, prows, H);
                    if (!a.det) {
                        prefetch_rows(a.eps_prior, r0, prows, S);
                        if (E > 0) prefetch_rows(a.eps_post, r0, prows, S);
                    }
                }
                uint32_t src = 0;
                for (int ti = 0; ti < n_tiles; ++ti) {
                    wait_backoff(&empty[slot], phase ^ 1u);
                    STAMP(0);
                    const uint32_t bar = tc::smem_u32(&full[slot]);
                    const uint32_t bytes = (prog.tiles[ti] & 1023u) << 4;
                    r_expect_tx(bar, bytes);
                    r_bulk_g2s(smem0 + OFF_RING + slot * SLOT_BYTES, packed + src, bytes, bar);
                    src += bytes;
                    if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // ---- MMA issuer ------------------------------------------------------------------------------------
        if (NHS > 0) {
            if (lane == 0) {
                Issuer I;
                I.slot = 0; I.phase = 0; I.NS = (uint32_t)NS; I.ring16 = (smem0 + OFF_RING) >> 4; I.tmem = tmem_base;
                I.full = full; I.empty = empty; I.ev = ev; I.cm = cm; I.tile_bytes = 0; I.open = false; I.sb = 0;
                for (int t = 0; t < T; ++t) {
                    I.it = t; I.par = (uint32_t)(t & 1);
                    static_step_fwd<200, 30, 200, 3, (NHS > 0 ? NHS : 1)>(I, (smem0 + OFF_XIN) >> 4, (smem0 + OFF_XU) >> 4,
                                                                        (smem0 + ((t & 1) ? OFF_H1 : OFF_H0)) >> 4, (smem0 + ((t & 1) ? OFF_H0 : OFF_H1)) >> 4);
                }
            }
        } else if (lane == 0) {
            const uint32_t hi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1, no swizzle
            const uint32_t hdelta = (uint32_t)(OFF_H1 - OFF_H0) >> 4;
            int pi = 0;
            uint32_t slot = 0, phase = 0;
            const uint32_t s16 = smem0 >> 4;
            for (int t = 0; t < T; ++t) {
                const uint32_t par = (uint32_t)(t & 1);
                const uint32_t f1 = s16 + (par ? hdelta : 0u), f2 = s16 - (par ? hdelta : 0u);
                int ri = 0;
                for (int ti = 0; ti < n_tiles; ++ti) {
                    const uint32_t tw = prog.tiles[ti];
                    const int wev = (int)((tw >> 13) & 31u) - 1, cmi = (int)((tw >> 18) & 31u) - 1;
                    if (wev >= 0) {
                        if (wev == EV_XIN) {
                            if (t > 0) tc::mbar_wait(tc::smem_u32(&ev[EV_XIN]), par ^ 1u);
                        } else {
                            tc::mbar_wait(tc::smem_u32(&ev[wev]), par);
                        }
                        tc::tc_fence_after();
                    }
                    STAMP(1);
                    tc::mbar_wait(tc::smem_u32(&full[slot]), phase);
                    tc::tc_fence_after();
                    STAMP(1);
                    const uint32_t sb = (smem0 + OFF_RING + slot * SLOT_BYTES) >> 4;
                    const int rend = ri + (int)((tw >> 10) & 7u);
                    for (; ri < rend; ++ri) {
                        const uint4 r = prog.runs[ri];
                        const uint32_t fl = (r.w >> 11) & 3u, n2 = ((r.z >> 17) & 63u) << 4, cntm = (r.w >> 13) & 7u;
                        const uint32_t a_lo = r.x + (fl == 1 ? f1 : (fl == 2 ? f2 : s16)), b_lo = r.y + sb;
                        const uint32_t d = tmem_base + (r.w & 1023u), id = r.z;
                        r_umma(d, a_lo, b_lo, hi, id, (r.w >> 10) & 1u);
                        if (cntm > 1) r_umma(d, a_lo + 128u, b_lo + n2, hi, id, 1);
                        if (cntm > 2) r_umma(d, a_lo + 256u, b_lo + 2 * n2, hi, id, 1);
                        if (cntm > 3) r_umma(d, a_lo + 384u, b_lo + 3 * n2, hi, id, 1);
                        if (cntm > 4) r_umma(d, a_lo + 512u, b_lo + 4 * n2, hi, id, 1);
                        if (cntm > 5) r_umma(d, a_lo + 640u, b_lo + 5 * n2, hi, id, 1);
                        if (cntm > 6) r_umma(d, a_lo + 768u, b_lo + 6 * n2, hi, id, 1);
                    }
                    tc::umma_commit(tc::smem_u32(&empty[slot]));
                    if (cmi >= 0) tc::umma_commit(tc::smem_u32(&cm[cmi]));
                    STAMP(1);
                    if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp < EPI_WARPS) {
        // ---- epilogue warps ------------------------------------------------------------------------------
        const int ew = warp, q = ew & 3, p = ew >> 2, etid = tid;
        const int r0 = 16 * q + (lane >> 2), r1 = r0 + 8, cj = 2 * (lane & 3);
        const int seq0 = b0 + q * RPG + (lane >> 2), seq1 = seq0 + 8;
        constexpr bool two = TWO;                                         // the second fragment row (r1) carries a sequence
        const bool ok0 = seq0 < B, ok1 = two && seq1 < B;
        constexpr int ne = two ? 4 : 2;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        const uint32_t opnd0 = (uint32_t)(r0 * 16 + cj * 2), opnd1 = (uint32_t)(r1 * 16 + cj * 2);      // inside a bf16 chunk plane
        const uint32_t hf0 = smem0 + OFF_HF + (uint32_t)(r0 * 32 + cj * 4), hf1 = smem0 + OFF_HF + (uint32_t)(r1 * 32 + cj * 4);
        const int act = a.act;
        const float min_std = a.min_std;
        const bool det = a.det != 0;
        int pi = 0;
#define ESTAMP() do { if (ew == 0 && lane == 0) STAMP(2); } while (0)
        int pj = 0;
#define DSTAMP() do { if (prof && ew == 0 && lane == 0 && blockIdx.x == 0 && t == PROF_STEP && pj < PROF_SLOTS) prof[3 * PROF_SLOTS + pj++] = clock64(); } while (0)
        for (int t = 0; t < T; ++t) {
            const uint32_t par = (uint32_t)(t & 1);
            const long long row0 = (long long)t * B + seq0, row1 = row0 + 8;
            // ---- E1: x = act(W_sa xin + b) -> XU, stash x ------------------------------------------------
            ESTAMP();
            wait_backoff(&cm[CM_X], par);
            tc::tc_fence_after();
            ESTAMP();
            {
                const int nc = ceil16(D) >> 3;
                for (int c = p; c < nc; c += 4) {
                    float v[4];
                    ld_frag(tlane + (uint32_t)(8 * c), v);
                    wait_ld();
                    const int j = 8 * c + cj;
                    const float bA = bias_x[j], bB = bias_x[j + 1];
                    const float x00 = act_fast(v[0] + bA, act), x01 = act_fast(v[1] + bB, act);
                    const uint32_t ch = smem0 + OFF_XU + (uint32_t)c * CH_BYTES;
                    st_shared_u32(ch + opnd0, pack_bf16x2(x00, x01));
                    if (a.st_x && j < D && ok0) st_f2(a.st_x + row0 * D + j, x00, x01);
                    if (two) {
                        const float x10 = act_fast(v[2] + bA, act), x11 = act_fast(v[3] + bB, act);
                        st_shared_u32(ch + opnd1, pack_bf16x2(x10, x11));
                        if (a.st_x && j < D && ok1) st_f2(a.st_x + row1 * D + j, x10, x11);
                    }
                }
            }
            epi_signal(&ev[EV_X], lane);
            ESTAMP();
            // ---- E2: GRU gate math per column half -> h (fp32 + bf16 operand), beliefs, stash ----------------
            const uint32_t hnew = smem0 + ((t & 1) ? OFF_H0 : OFF_H1);
            for (int hf = 0; hf < 2; ++hf) {
                wait_backoff(&cm[CM_GA + hf], par);
                tc::tc_fence_after();
                ESTAMP();
                const int c_lo = hf ? cA : 0, c_hi = hf ? nD8 : cA;
                for (int c = c_lo + p; c < c_hi; c += 4) {
                    const uint32_t col = (uint32_t)(8 * (c - c_lo));
                    float vr[4], vz[4], vi[4], vh[4];
                    ld_frag(tlane + col, vr);
                    ld_frag(tlane + ACC_STRIDE + col, vz);
                    ld_frag(tlane + 2 * ACC_STRIDE + col, vi);
                    DSTAMP();
                    ld_frag(tlane + 3 * ACC_STRIDE + col, vh);
                    wait_ld();
                    DSTAMP();
                    const int j = 8 * c + cj;
                    const float2 hp0 = ld_shared_f2(hf0 + (uint32_t)c * (ROWS * 32));
                    float2 hp1 = make_float2(0.f, 0.f);
                    if (two) hp1 = ld_shared_f2(hf1 + (uint32_t)c * (ROWS * 32));
                    const float hp[4] = {hp0.x, hp0.y, hp1.x, hp1.y};
                    const float br[2] = {bias_g[0][j], bias_g[0][j + 1]}, bz[2] = {bias_g[1][j], bias_g[1][j + 1]};
                    const float bi[2] = {bias_g[2][j], bias_g[2][j + 1]}, bh[2] = {bias_g[3][j], bias_g[3][j + 1]};
                    float rr[4], zz[4], nn[4], gg[4], hn[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (e < ne) {
                            rr[e] = sigmoid_fast(vr[e] + br[e & 1]);
                            zz[e] = sigmoid_fast(vz[e] + bz[e & 1]);
                            gg[e] = vh[e] + bh[e & 1];
                            nn[e] = tanh_fast(vi[e] + bi[e & 1] + rr[e] * gg[e]);
                            hn[e] = (1.f - zz[e]) * nn[e] + zz[e] * hp[e];
                        }
                    }
                    const uint32_t ch = hnew + (uint32_t)c * CH_BYTES;
                    DSTAMP();
                    st_shared_f2(hf0 + (uint32_t)c * (ROWS * 32), hn[0], hn[1]);
                    st_shared_u32(ch + opnd0, pack_bf16x2(hn[0], hn[1]));
                    if (ok0) {
                        const long long off = row0 * D + j;
                        st_f2(a.beliefs + off, hn[0], hn[1]);
                        if (a.st_r) {
                            st_f2(a.st_r + off, rr[0], rr[1]); st_f2(a.st_z + off, zz[0], zz[1]);
                            st_f2(a.st_n + off, nn[0], nn[1]); st_f2(a.st_ghn + off, gg[0], gg[1]);
                        }
                    }
                    if (two) {
                        st_shared_f2(hf1 + (uint32_t)c * (ROWS * 32), hn[2], hn[3]);
                        st_shared_u32(ch + opnd1, pack_bf16x2(hn[2], hn[3]));
                        if (ok1) {
                            const long long off = row1 * D + j;
                            st_f2(a.beliefs + off, hn[2], hn[3]);
                            if (a.st_r) {
                                st_f2(a.st_r + off, rr[2], rr[3]); st_f2(a.st_z + off, zz[2], zz[3]);
                                st_f2(a.st_n + off, nn[2], nn[3]); st_f2(a.st_ghn + off, gg[2], gg[3]);
                            }
                        }
                    }
                }
                DSTAMP();
                tc::tc_fence_before();
                tc::fence_proxy_async();
                DSTAMP();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(&ev[EV_HA + hf]));
                DSTAMP();
                ESTAMP();
            }
            // ---- E3: heads fc1 (+ hoisted embedding half) + act -> XU (as U), stash u ------------------------
            for (int hd = 0; hd < NH; ++hd) {
                const float* ep = emb_pre_s[hd];
                float* su = st_u_s[hd];
                for (int hf = 0; hf < 2; ++hf) {
                    const int c_lo = hf ? cAH : 0, c_hi = hf ? nH8 : cAH;
                    float pre[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c_lo + p + 4 * i;
                        pre[i][0] = pre[i][1] = pre[i][2] = pre[i][3] = 0.f;
                        if (ep && c < c_hi) {
                            const int j = 8 * c + cj;
                            if (ok0) { const float2 e0 = __ldg(reinterpret_cast<const float2*>(ep + row0 * H + j)); pre[i][0] = e0.x; pre[i][1] = e0.y; }
                            if (ok1) { const float2 e1 = __ldg(reinterpret_cast<const float2*>(ep + row1 * H + j)); pre[i][2] = e1.x; pre[i][3] = e1.y; }
                        }
                    }
                    // (hd, half 0) of hd >= 1 also needs fc2 of head hd-1 to be done reading U: that commit was issued later
                    const int cmi = (hf == 0 && hd > 0) ? CM_F2 + hd - 1 : CM_F1 + 2 * hd + hf;
                    ESTAMP();
                    wait_backoff(&cm[cmi], par);
                    tc::tc_fence_after();
                    ESTAMP();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c_lo + p + 4 * i;
                        if (c < c_hi) {
                            float v[4];
                            ld_frag(tlane + (uint32_t)(hf * ACC_STRIDE + 8 * (c - c_lo)), v);
                            wait_ld();
                            const int j = 8 * c + cj;
                            const float bA = bias_1[hd][j], bB = bias_1[hd][j + 1];
                            const float u00 = act_fast(v[0] + bA + pre[i][0], act), u01 = act_fast(v[1] + bB + pre[i][1], act);
                            const uint32_t ch = smem0 + OFF_XU + (uint32_t)c * CH_BYTES;
                            st_shared_u32(ch + opnd0, pack_bf16x2(u00, u01));
                            if (su && ok0) st_f2(su + row0 * H + j, u00, u01);
                            if (two) {
                                const float u10 = act_fast(v[2] + bA + pre[i][2], act), u11 = act_fast(v[3] + bB + pre[i][3], act);
                                st_shared_u32(ch + opnd1, pack_bf16x2(u10, u11));
                                if (su && ok1) st_f2(su + row1 * H + j, u10, u11);
                            }
                        }
                    }
                    epi_signal(&ev[EV_U0 + 2 * hd + hf], lane);
                    ESTAMP();
                }
            }
            // ---- E4: heads fc2 -> (mean, softplus + min_std), prior sample, fusion, posterior sample, next xin --
            const bool more = t + 1 < T;
            if (more) {                                       // action columns of xin(t+1): the buffer is idle since E1
                for (int i = etid; i < ROWS * A; i += EPI_WARPS * 32) {
                    const int r = i / A, k = i - r * A, col = S + k, seq = b0 + (r >> 4) * RPG + (r & 15);
                    const float v = ((r & 15) < RPG && seq < B) ? a.actions[((long long)(t + 1) * B + seq) * A + k] : 0.f;
                    st_shared_bf16(smem0 + OFF_XIN + (uint32_t)((col >> 3) * CH_BYTES + r * 16 + (col & 7) * 2), v);
                }
            }
            const int s0 = 8 * p + cj;                        // this thread: state dims s0, s0+1 of rows r0, r1
            const bool live = 8 * p < S;
            const bool pairok = s0 + 1 < S && (S & 1) == 0;   // float2 stores / loads of (s0, s0+1)
            float epr[4] = {0.f, 0.f, 0.f, 0.f}, epo[4] = {0.f, 0.f, 0.f, 0.f}, mk[2] = {1.f, 1.f};
            if (live) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int s = s0 + (e & 1);
                    const bool okr = (e < 2) ? ok0 : ok1;
                    if (okr && s < S && !det) {
                        const long long off = ((e < 2) ? row0 : row1) * S + s;
                        epr[e] = __ldg(a.eps_prior + off);
                        if (E > 0) epo[e] = __ldg(a.eps_post + off);
                    }
                }
                if (more && a.nonterminals) {
                    if (ok0) mk[0] = __ldg(a.nonterminals + row0 + B);
                    if (ok1) mk[1] = __ldg(a.nonterminals + row1 + B);
                }
            }
            ESTAMP();
            wait_backoff(&cm[CM_F2 + NH - 1], par);
            tc::tc_fence_after();
            ESTAMP();
            if (live) {
                float mu[TC_MAX_HEADS][4], sg[TC_MAX_HEADS][4];
#pragma unroll
                for (int h = 0; h < TC_MAX_HEADS; ++h) {
                    if (h < NH) {
                        ld_frag(tlane + (uint32_t)(F2_COL + 64 * h + 8 * p), mu[h]);
                        ld_frag(tlane + (uint32_t)(F2_COL + 64 * h + 32 + 8 * p), sg[h]);
                    }
                }
                wait_ld();
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {                  // one fragment row at a time (register pressure)
                    if (rr == 1 && !two) continue;
                    const bool okr = rr ? ok1 : ok0;
                    float o_pm[2], o_ps[2], o_pst[2], o_qm[2], o_qs[2], o_qst[2], o_em[TC_MAX_HEADS][2], o_es[TC_MAX_HEADS][2];
#pragma unroll
                    for (int ee = 0; ee < 2; ++ee) {
                        const int e = 2 * rr + ee;
                        const int s = min(s0 + ee, 31);
#pragma unroll
                        for (int h = 0; h < TC_MAX_HEADS; ++h) {
                            if (h < NH) {
                                o_em[h][ee] = mu[h][e] + bias_2[h][s];
                                o_es[h][ee] = softplus_fast(sg[h][e] + bias_2[h][32 + s]) + min_std;
                            }
                        }
                        o_pm[ee] = o_em[0][ee];
                        o_ps[ee] = o_es[0][ee];
                        o_pst[ee] = det ? o_pm[ee] : fmaf(o_ps[ee], epr[e], o_pm[ee]);
                        float nxt = o_pst[ee];
                        if (E > 0) {
                            if (a.n_subsets == 0) {
                                o_qm[ee] = o_em[1][ee];
                                o_qs[ee] = o_es[1][ee];
                            } else {
                                const unsigned mask = smask_s[s];
                                float sumT = 0.f, sumMT = 0.f;
#pragma unroll
                                for (int h = 1; h < TC_MAX_HEADS; ++h) {
                                    if (h < NH && (mask & (1u << (h - 1)))) {
                                        const float tt = rcp_approx(o_es[h][ee]);
                                        sumT += tt;
                                        sumMT = fmaf(o_em[h][ee], tt, sumMT);
                                    }
                                }
                                o_qs[ee] = rcp_approx(sumT);
                                o_qm[ee] = sumMT * o_qs[ee];
                            }
                            o_qst[ee] = det ? o_qm[ee] : fmaf(o_qs[ee], epo[e], o_qm[ee]);
                            nxt = o_qst[ee];
                        }
                        if (more && s0 + ee < S)
                            st_shared_bf16(smem0 + OFF_XIN + (uint32_t)((s >> 3) * CH_BYTES + (rr ? r1 : r0) * 16 + (s & 7) * 2),
                                           okr ? nxt * mk[rr] : 0.f);
                    }
                    if (!okr) continue;
                    // outputs: (s0, s0+1) of one row are adjacent in memory
                    const long long off = (rr ? row1 : row0) * S + s0;
                    auto put = [&](float* base, float v0, float v1) {
                        if (pairok) st_f2(base + off, v0, v1);
                        else {
                            if (s0 < S) base[off] = v0;
                            if (s0 + 1 < S) base[off + 1] = v1;
                        }
                    };
                    put(a.prior_means, o_pm[0], o_pm[1]);
                    put(a.prior_stds, o_ps[0], o_ps[1]);
                    put(a.prior_states, o_pst[0], o_pst[1]);
                    if (E > 0) {
                        put(a.post_means, o_qm[0], o_qm[1]);
                        put(a.post_stds, o_qs[0], o_qs[1]);
                        put(a.post_states, o_qst[0], o_qst[1]);
#pragma unroll
                        for (int h = 1; h < TC_MAX_HEADS; ++h) {
                            if (h < NH) {
                                put(exp_means_s[h], o_em[h][0], o_em[h][1]);
                                put(exp_stds_s[h], o_es[h][0], o_es[h][1]);
                            }
                        }
                    }
                }
            }
            epi_signal(&ev[EV_XIN], lane);
            ESTAMP();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == ALLOC_WARP) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// backward (BPTT), same machinery: per-step MMA program over the TRANSPOSED weights, activations = gradient operands.
// Per step t (descending):  go (head-output grads, elementwise from the upstream grads and the carried state grad)
//   -> gu_h = go_h W2_h, du_h = gu_h * act'(u_h)            (per half head, two TMEM sets)
//   -> G = g_beliefs + carry + sum_h du_h W1_h[:, :D]        (one accumulator)
//   -> GRU gate backward (elementwise)  -> gx = dgi W_ih, carry' = G z + dgh W_hh   (two accumulators)
//   -> dxpre = gx * act'(x)  -> g_xin = dxpre W_sa  -> state carry (masked), action grads
// The pre-activation gradients (d_o, d_u, d_gi, d_gh, d_xpre) and xin go to HBM for the deferred weight-gradient GEMMs.
// ---------------------------------------------------------------------------------------------------------------
enum { EVB_DO = 0, EVB_DU0, EVB_GRU = EVB_DU0 + 2 * TC_MAX_HEADS, EVB_DX, N_EVB };
enum { CMB_GB0 = 0, CMB_GCH0 = CMB_GB0 + 2 * TC_MAX_HEADS, CMB_GE = CMB_GCH0 + TC_MAX_HEADS, CMB_GF, N_CMB };
static_assert(N_EVB <= N_EV && N_CMB <= N_CM, "barrier arrays are shared with the forward kernel");
// operand region (chunk planes of CH_BYTES): go of head h at 8h; du at 32; later dr, dz, dn, dn*r at 0, 26, 52, 78; later dxpre at 0
constexpr int BCH_DO = 0, BCH_DU = 8 * TC_MAX_HEADS, BCH_DR = 0, BCH_DZ = MAXC, BCH_DN = 2 * MAXC, BCH_DNR = 3 * MAXC, BCH_DX = 0;
constexpr int BOFF_CG = 4 * MAXC * CH_BYTES;                  // fp32 carried grad wrt h: [chunk][row][8]
constexpr int BOFF_RING = BOFF_CG + MAXC * ROWS * 32;
constexpr int BCOL_SET = ACC_STRIDE, BCOL_GH = 2 * ACC_STRIDE, BCOL_GX = 0, BCOL_GHH = 208, BCOL_XIN = 448;

void build_plan_bwd(int D, int S, int H, int A, int NH, PlanBuilder& pb, TcHeader& h) {
    const int nD8 = D / 8, nH8 = H / 8, cA = (nD8 + 1) / 2, cAH = (nH8 + 1) / 2;
    const int nkD = (D + 15) / 16, nkH = (H + 15) / 16;
    const int ND = ceil16(D);
    auto gb = [&](int hd, int hf, int wait_ev) {      // gu_h[:, half] = go_h W2_h[:, half]   (K = [mean 0..S) | std 32..32+S))
        const int c0 = hf ? cAH : 0, nc = hf ? nH8 - cAH : cAH, n0 = 8 * c0, cnt = 8 * nc, N = ceil16(cnt);
        pb.unit_begin(wait_ev);
        for (int k = 0; k < 4; ++k) pb.mma_t(BCH_DO + 8 * hd, k, hf * BCOL_SET, N, k != 0, SRC_W2 + hd, n0, cnt, 0, 0, S, 32, S, S);
        pb.unit_end(CMB_GB0 + 2 * hd + hf);
    };
    auto gc = [&](int hd) {                           // G += du_h W1_h[:, :D]
        pb.unit_begin(EVB_DU0 + 2 * hd + 1);
        for (int k = 0; k < nkH; ++k) pb.mma_t(BCH_DU, k, BCOL_GH, ND, !(hd == 0 && k == 0), SRC_W1 + hd, 0, D, 0, 0, H);
        pb.unit_end(CMB_GCH0 + hd);
    };
    gb(0, 0, EVB_DO);
    gb(0, 1, -1);
    for (int hd = 0; hd < NH; ++hd) {
        if (hd + 1 < NH) gb(hd + 1, 0, EVB_DU0 + 2 * hd);
        gc(hd);
        if (hd + 1 < NH) gb(hd + 1, 1, -1);
    }
    // gx = [dr dz dn] W_ih ; gh = [dr dz dn*r] W_hh
    pb.unit_begin(EVB_GRU);
    for (int g = 0; g < 3; ++g)
        for (int k = 0; k < nkD; ++k) pb.mma_t(g == 0 ? BCH_DR : (g == 1 ? BCH_DZ : BCH_DN), k, BCOL_GX, ND, !(g == 0 && k == 0), SRC_WIH, 0, D, 0, g * D, D);
    for (int g = 0; g < 3; ++g)
        for (int k = 0; k < nkD; ++k) pb.mma_t(g == 0 ? BCH_DR : (g == 1 ? BCH_DZ : BCH_DNR), k, BCOL_GHH, ND, !(g == 0 && k == 0), SRC_WHH, 0, D, 0, g * D, D);
    pb.unit_end(CMB_GE);
    // g_xin = dxpre W_sa
    pb.unit_begin(EVB_DX);
    for (int k = 0; k < nkD; ++k) pb.mma_t(BCH_DX, k, BCOL_XIN, 8 * XIN_CH, k != 0, SRC_WSA, 0, S + A, 0, 0, D);
    pb.unit_end(CMB_GF);

    memset(&h, 0, sizeof(h));
    h.magic = TC_MAGIC + 1; h.D = D; h.S = S; h.H = H; h.A = A; h.NH = NH;
    h.n_tiles = (int)pb.tiles.size(); h.n_ops = (int)pb.ops.size(); h.n_pack = (int)pb.packs.size();
    h.cA = cA; h.nD8 = nD8; h.cAH = cAH; h.nH8 = nH8;
    h.packed_bytes = pb.packed;
    h.tile_off = 128;
    h.op_off = h.tile_off + (uint32_t)(pb.tiles.size() * sizeof(TcTile));
    h.pack_off = (h.op_off + (uint32_t)(pb.ops.size() * sizeof(TcOp)) + 15u) & ~15u;
    h.run_off = h.pack_off + (uint32_t)(pb.packs.size() * sizeof(TcPack));
    h.n_runs = (int)pb.runs.size();
    h.total_bytes = h.run_off + (uint32_t)(pb.runs.size() * sizeof(TcRun));
}

// the weight producer and the MMA issuer of one CTA (shared by both directions); `it` counts processed time steps
template <class OnStep>
__device__ __forceinline__ void producer_role(const TcProg& prog, const uint8_t* __restrict__ packed, uint32_t ring, int NS, int n_steps,
                                              uint64_t* full, uint64_t* empty, OnStep on_step) {
    uint32_t slot = 0, phase = 0;
    const int n_tiles = prog.n_tiles;
    for (int it = 0; it < n_steps; ++it) {
        on_step(it);
        uint32_t src = 0;
        for (int ti = 0; ti < n_tiles; ++ti) {
            wait_backoff(&empty[slot], phase ^ 1u);
            const uint32_t bar = tc::smem_u32(&full[slot]);
            const uint32_t bytes = (prog.tiles[ti] & 1023u) << 4;
            r_expect_tx(bar, bytes);
            r_bulk_g2s(ring + slot * SLOT_BYTES, packed + src, bytes, bar);
            src += bytes;
            if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1u; }
        }
    }
}
__device__ __forceinline__ void mma_role(const TcProg& prog, uint32_t abase, uint32_t ring, int NS, int n_steps, uint32_t tmem_base,
                                         uint64_t* full, uint64_t* empty, uint64_t* ev, uint64_t* cm) {
    const uint32_t hi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1, no swizzle
    const uint32_t s16 = abase >> 4;
    const int n_tiles = prog.n_tiles;
    uint32_t slot = 0, phase = 0;
    for (int it = 0; it < n_steps; ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        int ri = 0;
        for (int ti = 0; ti < n_tiles; ++ti) {
            const uint32_t tw = prog.tiles[ti];
            const int wev = (int)((tw >> 13) & 31u) - 1, cmi = (int)((tw >> 18) & 31u) - 1;
            if (wev >= 0) {
                if (wev == 0) {                                   // the step's first operand comes from the previous step's last epilogue
                    if (it > 0) tc::mbar_wait(tc::smem_u32(&ev[0]), par ^ 1u);
                } else {
                    tc::mbar_wait(tc::smem_u32(&ev[wev]), par);
                }
                tc::tc_fence_after();
            }
            tc::mbar_wait(tc::smem_u32(&full[slot]), phase);
            tc::tc_fence_after();
            const uint32_t sb = (ring + slot * SLOT_BYTES) >> 4;
            const int rend = ri + (int)((tw >> 10) & 7u);
            for (; ri < rend; ++ri) {
                const uint4 r = prog.runs[ri];
                const uint32_t n2 = ((r.z >> 17) & 63u) << 4, cntm = (r.w >> 13) & 7u;
                const uint32_t a_lo = r.x + s16, b_lo = r.y + sb;
                const uint32_t d = tmem_base + (r.w & 1023u), id = r.z;
                r_umma(d, a_lo, b_lo, hi, id, (r.w >> 10) & 1u);
                if (cntm > 1) r_umma(d, a_lo + 128u, b_lo + n2, hi, id, 1);
                if (cntm > 2) r_umma(d, a_lo + 256u, b_lo + 2 * n2, hi, id, 1);
                if (cntm > 3) r_umma(d, a_lo + 384u, b_lo + 3 * n2, hi, id, 1);
                if (cntm > 4) r_umma(d, a_lo + 512u, b_lo + 4 * n2, hi, id, 1);
                if (cntm > 5) r_umma(d, a_lo + 640u, b_lo + 5 * n2, hi, id, 1);
                if (cntm > 6) r_umma(d, a_lo + 768u, b_lo + 6 * n2, hi, id, 1);
            }
            tc::umma_commit(tc::smem_u32(&empty[slot]));
            if (cmi >= 0) tc::umma_commit(tc::smem_u32(&cm[cmi]));
            if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1u; }
        }
    }
}


template <int D, int S, int H, int A, int NH>
__device__ __forceinline__ void static_step_bwd(Issuer& I, uint32_t base16) {
    constexpr int nH8 = H / 8, cAH = (nH8 + 1) / 2;
    constexpr int nkD = (D + 15) / 16, nkH = (H + 15) / 16, ND = (D + 15) / 16 * 16;
    auto gb = [&](int hd, int hf, int wev) {
        const int nc = hf ? nH8 - cAH : cAH, N = (8 * nc + 15) / 16 * 16;
        I.unit_begin(wev);
#pragma unroll
        for (int k = 0; k < 4; ++k) I.emit(opaque(base16) + (BCH_DO + 8 * hd + 2 * k) * CH16, hf * BCOL_SET, N, k != 0);
        I.unit_end(CMB_GB0 + 2 * hd + hf);
    };
    auto gc = [&](int hd) {
        I.unit_begin(EVB_DU0 + 2 * hd + 1);
#pragma unroll
        for (int k = 0; k < nkH; ++k) I.emit(opaque(base16) + (BCH_DU + 2 * k) * CH16, BCOL_GH, ND, !(hd == 0 && k == 0));
        I.unit_end(CMB_GCH0 + hd);
    };
    gb(0, 0, EVB_DO);
    gb(0, 1, -1);
#pragma unroll
    for (int hd = 0; hd < NH; ++hd) {
        if (hd + 1 < NH) gb(hd + 1, 0, EVB_DU0 + 2 * hd);
        gc(hd);
        if (hd + 1 < NH) gb(hd + 1, 1, -1);
    }
    I.unit_begin(EVB_GRU);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int k = 0; k < nkD; ++k) I.emit(opaque(base16) + ((g == 0 ? BCH_DR : (g == 1 ? BCH_DZ : BCH_DN)) + 2 * k) * CH16, BCOL_GX, ND, !(g == 0 && k == 0));
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int k = 0; k < nkD; ++k) I.emit(opaque(base16) + ((g == 0 ? BCH_DR : (g == 1 ? BCH_DZ : BCH_DNR)) + 2 * k) * CH16, BCOL_GHH, ND, !(g == 0 && k == 0));
    }
    I.unit_end(CMB_GE);
    I.unit_begin(EVB_DX);
#pragma unroll
    for (int k = 0; k < nkD; ++k) I.emit(opaque(base16) + (BCH_DX + 2 * k) * CH16, BCOL_XIN, 8 * XIN_CH, k != 0);
    I.unit_end(CMB_GF);
}

struct BwdPtrs {          // per-head pointers the epilogue indexes dynamically (kept in shared memory)
    const float *st_u[TC_MAX_HEADS], *exp_means[TC_MAX_HEADS], *exp_stds[TC_MAX_HEADS], *g_exp_means[TC_MAX_HEADS], *g_exp_stds[TC_MAX_HEADS];
    float *d_u[TC_MAX_HEADS], *d_o[TC_MAX_HEADS];
};

template <int NHS, bool TWO>
__global__ void __launch_bounds__(TC_THREADS, 1)
rollout_tc_bwd_kernel(const __grid_constant__ mrssm_rollout_bwd_args g, const __grid_constant__ TcProg prog, const uint8_t* __restrict__ packed,
                      const int NS, long long* __restrict__ prof) {
    constexpr int RPG = TWO ? 16 : 8;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full[MAX_SLOTS], empty[MAX_SLOTS], ev[N_EV], cm[N_CM];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t smask_s[32];
    __shared__ BwdPtrs P;

    const mrssm_rollout_args& a = g.f;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem = smem_raw + (smem0 - tc::smem_u32(smem_raw));
    const int D = a.D, S = a.S, H = a.H, A = a.A, E = a.n_experts, NH = 1 + E, B = a.B, T = a.T;
    const int nD8 = prog.nD8, cAH = prog.cAH, nH8 = prog.nH8;
    const int b0 = blockIdx.x * 4 * RPG;

    if (tid < 32) smask_s[tid] = (tid < S) ? (a.n_subsets > 0 ? a.subset_mask[a.dim_subset[tid]] : 1u) : 0u;
    if (tid < TC_MAX_HEADS) {
        const bool ok = tid < NH;
        P.st_u[tid] = ok ? a.st_u[tid] : nullptr;
        P.exp_means[tid] = ok ? a.exp_means[tid] : nullptr;
        P.exp_stds[tid] = ok ? a.exp_stds[tid] : nullptr;
        P.g_exp_means[tid] = ok ? g.g_exp_means[tid] : nullptr;
        P.g_exp_stds[tid] = ok ? g.g_exp_stds[tid] : nullptr;
        P.d_u[tid] = ok ? g.d_u[tid] : nullptr;
        P.d_o[tid] = ok ? g.d_o[tid] : nullptr;
    }
    for (int i = tid; i < BOFF_RING / 16; i += TC_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < MAX_SLOTS; ++s) {
            tc::mbar_init(tc::smem_u32(&full[s]), 1);
            tc::mbar_init(tc::smem_u32(&empty[s]), 1);
        }
        for (int i = 0; i < N_EV; ++i) tc::mbar_init(tc::smem_u32(&ev[i]), EPI_WARPS);
        for (int i = 0; i < N_CM; ++i) tc::mbar_init(tc::smem_u32(&cm[i]), 1);
        tc::fence_barrier_init();
    }
    if (warp == ALLOC_WARP) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 512);
    __syncthreads();

    // epilogue-thread coordinates (all warps compute them; only warps < EPI_WARPS use them)
    const int q = warp & 3, p = warp >> 2;
    const int r0 = 16 * q + (lane >> 2), r1 = r0 + 8, cj = 2 * (lane & 3);
    const int seq0 = b0 + q * RPG + (lane >> 2);
    constexpr bool two = TWO;
    const bool ok0 = seq0 < B, ok1 = two && seq0 + 8 < B;
    const int act = a.act;
    const float min_std = a.min_std;
    const bool det = a.det != 0;
    const int s0 = 8 * p + cj;                                   // state dims s0, s0+1 (rows r0, r1) for the fusion backward
    const bool live = 8 * p < S;
    const bool pairS = s0 + 1 < S && (S & 1) == 0;
    float cgs[4] = {0.f, 0.f, 0.f, 0.f};                         // carried grad wrt the fed-back state at (r0|r1, s0|s0+1)

    // go of step t from the upstream grads and the state carry -> DO operands (bf16) + d_o (HBM)
    auto step_a = [&](int t) {
        if (!live) return;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            if (rr == 1 && !two) continue;
            const bool okr = rr ? ok1 : ok0;
            const int r = rr ? r1 : r0;
            const long long row = (long long)t * B + seq0 + 8 * rr;
            float gm[TC_MAX_HEADS][2], gs[TC_MAX_HEADS][2];
#pragma unroll
            for (int ee = 0; ee < 2; ++ee) {
                const int s = s0 + ee;
                const bool lv = okr && s < S;
                const long long off = lv ? row * S + s : 0;
                const float carry = cgs[2 * rr + ee];
                float gps = (g.g_prior_states && lv) ? __ldg(g.g_prior_states + off) : 0.f;
                if (E == 0) gps += carry;
                const float gpm = gps + ((g.g_prior_means && lv) ? __ldg(g.g_prior_means + off) : 0.f);
                float gpsd = (g.g_prior_stds && lv) ? __ldg(g.g_prior_stds + off) : 0.f;
                if (!det && lv) gpsd = fmaf(gps, __ldg(a.eps_prior + off), gpsd);
                const float psd = lv ? __ldg(a.prior_stds + off) : 1.f;
                gm[0][ee] = lv ? gpm : 0.f;
                gs[0][ee] = lv ? gpsd * (1.f - ex2_approx(-1.4426950409f * (psd - min_std))) : 0.f;
                if (E > 0) {
                    const float gq = ((g.g_post_states && lv) ? __ldg(g.g_post_states + off) : 0.f) + carry;
                    const float gqm = gq + ((g.g_post_means && lv) ? __ldg(g.g_post_means + off) : 0.f);
                    float gqs = (g.g_post_stds && lv) ? __ldg(g.g_post_stds + off) : 0.f;
                    if (!det && lv) gqs = fmaf(gq, __ldg(a.eps_post + off), gqs);
                    const unsigned mask = smask_s[min(s, 31)];
                    float Pp = 1.f, qm = 0.f;
                    if (a.n_subsets && lv) {
                        Pp = __ldg(a.post_stds + off);             // = 1 / sum of precisions
                        qm = __ldg(a.post_means + off);
                    }
#pragma unroll
                    for (int e = 1; e < TC_MAX_HEADS; ++e) {
                        if (e < NH) {
                            float vm = (P.g_exp_means[e] && lv) ? __ldg(P.g_exp_means[e] + off) : 0.f;
                            float vs = (P.g_exp_stds[e] && lv) ? __ldg(P.g_exp_stds[e] + off) : 0.f;
                            const float sd = lv ? __ldg(P.exp_stds[e] + off) : 1.f;
                            if (lv && (mask & (1u << (e - 1)))) {
                                if (a.n_subsets == 0) {
                                    vm += gqm;
                                    vs += gqs;
                                } else {
                                    const float te = rcp_approx(sd);
                                    const float mu = __ldg(P.exp_means[e] + off);
                                    vm = fmaf(gqm, te * Pp, vm);
                                    const float gT = gqm * (mu - qm) * Pp - gqs * (Pp * Pp);
                                    vs = fmaf(-gT, te * te, vs);
                                }
                            }
                            gm[e][ee] = lv ? vm : 0.f;
                            gs[e][ee] = lv ? vs * (1.f - ex2_approx(-1.4426950409f * (sd - min_std))) : 0.f;
                        }
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < TC_MAX_HEADS; ++e) {
                if (e < NH) {
                    const uint32_t base = smem0 + (uint32_t)((BCH_DO + 8 * e) * CH_BYTES + r * 16 + cj * 2);
                    st_shared_u32(base + (uint32_t)p * CH_BYTES, pack_bf16x2(gm[e][0], gm[e][1]));
                    st_shared_u32(base + (uint32_t)(4 + p) * CH_BYTES, pack_bf16x2(gs[e][0], gs[e][1]));
                    if (okr) {
                        float* dm = P.d_o[e] + row * 2 * S + s0;
                        if (pairS) { st_f2(dm, gm[e][0], gm[e][1]); st_f2(dm + S, gs[e][0], gs[e][1]); }
                        else {
                            if (s0 < S) { dm[0] = gm[e][0]; dm[S] = gs[e][0]; }
                            if (s0 + 1 < S) { dm[1] = gm[e][1]; dm[S + 1] = gs[e][1]; }
                        }
                    }
                }
            }
        }
    };
    if (warp < EPI_WARPS) step_a(T - 1);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == PROD_WARP) {
        if (lane == 0) {
            // The stash, the outputs and the upstream gradients were written a whole decoder pass ago (DRAM): while step `it`
            // streams its weights, pull the blocks the epilogues of the NEXT step read (rows of this CTA are contiguous in every
            // time-major tensor) into L2 with a few bulk prefetches.
            const int rows = min(4 * RPG, B - b0);
            auto prefetch_step = [&](int t, bool wide, bool narrow) {
                if (t < 0) return;
                const long long r0 = (long long)t * B + b0;
                if (wide) {
                    prefetch_rows(a.st_r, r0, rows, D); prefetch_rows(a.st_z, r0, rows, D); prefetch_rows(a.st_n, r0, rows, D);
                    prefetch_rows(a.st_ghn, r0, rows, D); prefetch_rows(a.st_x, r0, rows, D); prefetch_rows(g.g_beliefs, r0, rows, D);
                    if (t > 0) prefetch_rows(a.beliefs, r0 - B, rows, D);
                    for (int h = 0; h < NH; ++h) prefetch_rows(P.st_u[h], r0, rows, H);
                }
                if (narrow) {
                    prefetch_rows(g.g_prior_states, r0, rows, S); prefetch_rows(g.g_prior_means, r0, rows, S);
                    prefetch_rows(g.g_prior_stds, r0, rows, S); prefetch_rows(a.prior_stds, r0, rows, S);
                    prefetch_rows(a.eps_prior, r0, rows, S);
                    if (E > 0) {
                        prefetch_rows(g.g_post_states, r0, rows, S); prefetch_rows(g.g_post_means, r0, rows, S);
                        prefetch_rows(g.g_post_stds, r0, rows, S); prefetch_rows(a.eps_post, r0, rows, S);
                        prefetch_rows(a.post_stds, r0, rows, S); prefetch_rows(a.post_means, r0, rows, S);
                        for (int e = 1; e < NH; ++e) {
                            prefetch_rows(P.g_exp_means[e], r0, rows, S); prefetch_rows(P.g_exp_stds[e], r0, rows, S);
                            prefetch_rows(P.exp_stds[e], r0, rows, S); prefetch_rows(P.exp_means[e], r0, rows, S);
                        }
                    }
                }
            };
            producer_role(prog, packed, smem0 + BOFF_RING, NS, T, full, empty, [&](int it) {
                const int t = T - 1 - it;
                if (it == 0) prefetch_step(t, true, false);
                prefetch_step(t - 1, true, true);
            });
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        if (NHS > 0) {
            if (lane == 0) {
                Issuer I;
                I.slot = 0; I.phase = 0; I.NS = (uint32_t)NS; I.ring16 = (smem0 + BOFF_RING) >> 4; I.tmem = tmem_base;
                I.full = full; I.empty = empty; I.ev = ev; I.cm = cm; I.tile_bytes = 0; I.open = false; I.sb = 0;
                for (int it = 0; it < T; ++it) {
                    I.it = it; I.par = (uint32_t)(it & 1);
                    static_step_bwd<200, 30, 200, 3, (NHS > 0 ? NHS : 1)>(I, smem0 >> 4);
                }
            }
        } else if (lane == 0) {
            mma_role(prog, smem0, smem0 + BOFF_RING, NS, T, tmem_base, full, empty, ev, cm);
        }
        __syncwarp();
    } else if (warp < EPI_WARPS) {
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        const uint32_t opnd0 = (uint32_t)(r0 * 16 + cj * 2), opnd1 = (uint32_t)(r1 * 16 + cj * 2);
        const uint32_t cg0 = smem0 + BOFF_CG + (uint32_t)(r0 * 32 + cj * 4), cg1 = smem0 + BOFF_CG + (uint32_t)(r1 * 32 + cj * 4);
        int pi = 0;
#define BSTAMP() do { if (prof && warp == 0 && lane == 0 && blockIdx.x == 0 && it == PROF_STEP && pi < PROF_SLOTS) prof[2 * PROF_SLOTS + pi++] = clock64(); } while (0)
        int pj = 0;
#define BDSTAMP() do { if (prof && warp == 0 && lane == 0 && blockIdx.x == 0 && it == PROF_STEP && pj < PROF_SLOTS) prof[3 * PROF_SLOTS + pj++] = clock64(); } while (0)
        for (int it = 0; it < T; ++it) {
            const int t = T - 1 - it;
            const uint32_t par = (uint32_t)(it & 1);
            const long long row0 = (long long)t * B + seq0, row1 = row0 + 8;
            BSTAMP();
            // ---- Eb: du_h = gu_h * act'(u_h) -> DU operand, d_u ------------------------------------------------
            for (int hd = 0; hd < NH; ++hd) {
                const float* su = P.st_u[hd];
                float* du = P.d_u[hd];
                for (int hf = 0; hf < 2; ++hf) {
                    const int c_lo = hf ? cAH : 0, c_hi = hf ? nH8 : cAH;
                    float uu[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c_lo + p + 4 * i;
                        uu[i][0] = uu[i][1] = uu[i][2] = uu[i][3] = 0.f;
                        if (c < c_hi) {
                            const int j = 8 * c + cj;
                            if (ok0) { const float2 e0 = __ldg(reinterpret_cast<const float2*>(su + row0 * H + j)); uu[i][0] = e0.x; uu[i][1] = e0.y; }
                            if (ok1) { const float2 e1 = __ldg(reinterpret_cast<const float2*>(su + row1 * H + j)); uu[i][2] = e1.x; uu[i][3] = e1.y; }
                        }
                    }
                    // half 0 of head hd >= 1 overwrites DU: G += du_{hd-1} W1 must have read it (that commit was issued later)
                    const int cmi = (hf == 0 && hd > 0) ? CMB_GCH0 + hd - 1 : CMB_GB0 + 2 * hd + hf;
                    wait_backoff(&cm[cmi], par);
                    tc::tc_fence_after();
                    BSTAMP();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c_lo + p + 4 * i;
                        if (c < c_hi) {
                            float v[4];
                            ld_frag(tlane + (uint32_t)(hf * BCOL_SET + 8 * (c - c_lo)), v);
                            wait_ld();
                            const int j = 8 * c + cj;
                            const float d00 = ok0 ? v[0] * act_grad_from_out(uu[i][0], act) : 0.f, d01 = ok0 ? v[1] * act_grad_from_out(uu[i][1], act) : 0.f;
                            const uint32_t ch = smem0 + (uint32_t)(BCH_DU + c) * CH_BYTES;
                            st_shared_u32(ch + opnd0, pack_bf16x2(d00, d01));
                            if (ok0) st_f2(du + row0 * H + j, d00, d01);
                            if (two) {
                                const float d10 = ok1 ? v[2] * act_grad_from_out(uu[i][2], act) : 0.f, d11 = ok1 ? v[3] * act_grad_from_out(uu[i][3], act) : 0.f;
                                st_shared_u32(ch + opnd1, pack_bf16x2(d10, d11));
                                if (ok1) st_f2(du + row1 * H + j, d10, d11);
                            }
                        }
                    }
                    epi_signal(&ev[EVB_DU0 + 2 * hd + hf], lane);
                    BSTAMP();
                }
            }
            // ---- Ed: G = g_beliefs + carry + acc ; GRU gate backward -> dr, dz, dn, dn*r operands, d_gi, d_gh; carry = G z ------
            {
                // per chunk: stash (r, z, n, W_hn h), h_{t-1} and the upstream belief grad, software-pipelined one chunk ahead
                struct GruIn { float2 r, z, n, h, p, gb; };
                auto load_in = [&](int c, int rr) {
                    GruIn x;
                    x.r = x.z = x.n = x.h = x.p = x.gb = make_float2(0.f, 0.f);
                    if (c < nD8 && (rr ? ok1 : ok0)) {
                        const int j = 8 * c + cj;
                        const long long off = (rr ? row1 : row0) * D + j;
                        x.r = __ldg(reinterpret_cast<const float2*>(a.st_r + off));
                        x.z = __ldg(reinterpret_cast<const float2*>(a.st_z + off));
                        x.n = __ldg(reinterpret_cast<const float2*>(a.st_n + off));
                        x.h = __ldg(reinterpret_cast<const float2*>(a.st_ghn + off));
                        x.p = (t > 0) ? __ldg(reinterpret_cast<const float2*>(a.beliefs + off - (long long)B * D))
                                      : __ldg(reinterpret_cast<const float2*>(a.prev_belief + (long long)(seq0 + 8 * rr) * D + j));
                        if (g.g_beliefs) x.gb = __ldg(reinterpret_cast<const float2*>(g.g_beliefs + off));
                    }
                    return x;
                };
                // two chunks ahead when only one fragment row is live (the loads are L2 hits at ~600 cycles, an iteration is shorter)
                GruIn nx0 = load_in(p, 0), nx1 = load_in(two ? p : nD8, 1), nxx = load_in(two ? nD8 : p + 4, 0);
                wait_backoff(&cm[CMB_GCH0 + NH - 1], par);
                tc::tc_fence_after();
                BSTAMP();
                for (int c = p; c < nD8; c += 4) {
                    BDSTAMP();
                    const GruIn in0 = nx0, in1 = nx1;
                    if (two) {
                        nx0 = load_in(c + 4, 0);
                        nx1 = load_in(c + 4, 1);
                    } else {
                        nx0 = nxx;
                        nxx = load_in(c + 8, 0);
                    }
                    float v[4];
                    ld_frag(tlane + (uint32_t)(BCOL_GH + 8 * c), v);
                    wait_ld();
                    BDSTAMP();
                    const int j = 8 * c + cj;
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        if (rr == 1 && !two) continue;
                        const bool okr = rr ? ok1 : ok0;
                        const GruIn& x = rr ? in1 : in0;
                        const uint32_t cga = (rr ? cg1 : cg0) + (uint32_t)c * (ROWS * 32);
                        float gr[2] = {0.f, 0.f}, gz[2] = {0.f, 0.f}, gn[2] = {0.f, 0.f}, gnr[2] = {0.f, 0.f}, dir[2] = {0.f, 0.f};
                        if (okr) {
                            const float2 cg = ld_shared_f2(cga);
                            const float G[2] = {v[2 * rr] + cg.x + x.gb.x, v[2 * rr + 1] + cg.y + x.gb.y};
                            const float r_[2] = {x.r.x, x.r.y}, z_[2] = {x.z.x, x.z.y}, n_[2] = {x.n.x, x.n.y}, h_[2] = {x.h.x, x.h.y}, p_[2] = {x.p.x, x.p.y};
#pragma unroll
                            for (int ee = 0; ee < 2; ++ee) {
                                const float gnn = G[ee] * (1.f - z_[ee]), gzz = G[ee] * (p_[ee] - n_[ee]);
                                dir[ee] = G[ee] * z_[ee];
                                gn[ee] = gnn * (1.f - n_[ee] * n_[ee]);
                                gr[ee] = gn[ee] * h_[ee] * r_[ee] * (1.f - r_[ee]);
                                gz[ee] = gzz * z_[ee] * (1.f - z_[ee]);
                                gnr[ee] = gn[ee] * r_[ee];
                            }
                            if (rr == 0) BDSTAMP();
                            const long long o3 = (rr ? row1 : row0) * 3 * D + j;
                            st_f2(g.d_gi + o3, gr[0], gr[1]); st_f2(g.d_gi + o3 + D, gz[0], gz[1]); st_f2(g.d_gi + o3 + 2 * D, gn[0], gn[1]);
                            st_f2(g.d_gh + o3, gr[0], gr[1]); st_f2(g.d_gh + o3 + D, gz[0], gz[1]); st_f2(g.d_gh + o3 + 2 * D, gnr[0], gnr[1]);
                        }
                        if (rr == 0) BDSTAMP();
                        st_shared_f2(cga, dir[0], dir[1]);
                        const uint32_t o = (rr ? opnd1 : opnd0) + (uint32_t)c * CH_BYTES;
                        st_shared_u32(smem0 + BCH_DR * CH_BYTES + o, pack_bf16x2(gr[0], gr[1]));
                        st_shared_u32(smem0 + BCH_DZ * CH_BYTES + o, pack_bf16x2(gz[0], gz[1]));
                        st_shared_u32(smem0 + BCH_DN * CH_BYTES + o, pack_bf16x2(gn[0], gn[1]));
                        st_shared_u32(smem0 + BCH_DNR * CH_BYTES + o, pack_bf16x2(gnr[0], gnr[1]));
                    }
                }
            }
            // the padding chunk planes of the four gate operands alias go / du planes of this step: clear them
            if ((D & 15) != 0) {
                for (int i = tid; i < 4 * (CH_BYTES / 4); i += EPI_WARPS * 32) {
                    const int b = i / (CH_BYTES / 4), w = i - b * (CH_BYTES / 4);
                    st_shared_u32(smem0 + (uint32_t)((b * MAXC + nD8) * CH_BYTES + 4 * w), 0u);
                }
            }
            epi_signal(&ev[EVB_GRU], lane);
            BSTAMP();
            // ---- Ee: dxpre = gx * act'(x) -> DX operand, d_xpre ; carry += gh -----------------------------------
            float2 xn0 = make_float2(0.f, 0.f), xn1 = make_float2(0.f, 0.f);
            if (p < nD8) {
                if (ok0) xn0 = __ldg(reinterpret_cast<const float2*>(a.st_x + row0 * D + 8 * p + cj));
                if (ok1) xn1 = __ldg(reinterpret_cast<const float2*>(a.st_x + row1 * D + 8 * p + cj));
            }
            wait_backoff(&cm[CMB_GE], par);
            tc::tc_fence_after();
            BSTAMP();
            for (int c = p; c < nD8; c += 4) {
                const float2 xa = xn0, xb = xn1;
                if (c + 4 < nD8) {
                    if (ok0) xn0 = __ldg(reinterpret_cast<const float2*>(a.st_x + row0 * D + 8 * (c + 4) + cj));
                    if (ok1) xn1 = __ldg(reinterpret_cast<const float2*>(a.st_x + row1 * D + 8 * (c + 4) + cj));
                }
                float vx[4], vh[4];
                ld_frag(tlane + (uint32_t)(BCOL_GX + 8 * c), vx);
                ld_frag(tlane + (uint32_t)(BCOL_GHH + 8 * c), vh);
                wait_ld();
                const int j = 8 * c + cj;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    if (rr == 1 && !two) continue;
                    const bool okr = rr ? ok1 : ok0;
                    const long long off = (rr ? row1 : row0) * D + j;
                    const uint32_t cga = (rr ? cg1 : cg0) + (uint32_t)c * (ROWS * 32);
                    float dx0 = 0.f, dx1 = 0.f;
                    if (okr) {
                        const float2 xv = rr ? xb : xa;
                        dx0 = vx[2 * rr] * act_grad_from_out(xv.x, act);
                        dx1 = vx[2 * rr + 1] * act_grad_from_out(xv.y, act);
                        st_f2(g.d_xpre + off, dx0, dx1);
                        const float2 cg = ld_shared_f2(cga);
                        st_shared_f2(cga, cg.x + vh[2 * rr], cg.y + vh[2 * rr + 1]);
                    }
                    st_shared_u32(smem0 + (uint32_t)(BCH_DX + c) * CH_BYTES + (rr ? opnd1 : opnd0), pack_bf16x2(dx0, dx1));
                }
            }
            if ((D & 15) != 0) {
                for (int i = tid; i < CH_BYTES / 4; i += EPI_WARPS * 32) st_shared_u32(smem0 + (uint32_t)((BCH_DX + nD8) * CH_BYTES + 4 * i), 0u);
            }
            epi_signal(&ev[EVB_DX], lane);
            BSTAMP();
            // ---- Ef: g_xin -> state carry (masked), action grads, xin ; then go of step t-1 -------------------------
            wait_backoff(&cm[CMB_GF], par);
            tc::tc_fence_after();
            BSTAMP();
            for (int grp = p; grp * 8 < S + A; grp += 4) {
                float v[4];
                ld_frag(tlane + (uint32_t)(BCOL_XIN + 8 * grp), v);
                wait_ld();
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    if (rr == 1 && !two) continue;
                    const bool okr = rr ? ok1 : ok0;
                    const long long row = rr ? row1 : row0;
                    const float m = (okr && a.nonterminals) ? __ldg(a.nonterminals + row) : 1.f;
#pragma unroll
                    for (int ee = 0; ee < 2; ++ee) {
                        const int i = 8 * grp + cj + ee;
                        const float acc = v[2 * rr + ee];
                        if (i < S) {
                            cgs[2 * rr + ee] = okr ? acc * m : 0.f;           // grp == p here (S <= 32)
                            if (okr && g.xin) {
                                const float* src = (E > 0) ? a.post_states : a.prior_states;
                                const float sp = (t > 0) ? __ldg(src + (row - B) * S + i) : __ldg(a.prev_state + (long long)(seq0 + 8 * rr) * S + i);
                                g.xin[row * (S + A) + i] = sp * m;
                            }
                        } else if (i < S + A && okr) {
                            if (g.g_actions) g.g_actions[row * A + (i - S)] = acc;
                            if (g.xin) g.xin[row * (S + A) + i] = __ldg(a.actions + row * A + (i - S));
                        }
                    }
                }
            }
            BSTAMP();
            if (t > 0) step_a(t - 1);
            epi_signal(&ev[EVB_DO], lane);
            BSTAMP();
        }
        // ---- carries out -------------------------------------------------------------------------------------
        if (g.g_prev_belief) {
            for (int c = p; c < nD8; c += 4) {
                const int j = 8 * c + cj;
                if (ok0) { const float2 cg = ld_shared_f2(cg0 + (uint32_t)c * (ROWS * 32)); st_f2(g.g_prev_belief + (long long)seq0 * D + j, cg.x, cg.y); }
                if (ok1) { const float2 cg = ld_shared_f2(cg1 + (uint32_t)c * (ROWS * 32)); st_f2(g.g_prev_belief + (long long)(seq0 + 8) * D + j, cg.x, cg.y); }
            }
        }
        if (g.g_prev_state && live) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = s0 + (e & 1);
                if (s < S && ((e < 2) ? ok0 : ok1)) g.g_prev_state[(long long)(seq0 + 8 * (e >> 1)) * S + s] = cgs[e];
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == ALLOC_WARP) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" int mrssm_rollout_tc_eligible(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts) {
    return tc_eligible(D, S, H, A, 1 + n_experts) ? 1 : 0;
}

extern "C" int mrssm_rollout_tc_plan_bytes(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts, int64_t* plan_bytes,
                                           int64_t* packed_bytes) {
    MRSSM_CHECK(tc_eligible(D, S, H, A, 1 + n_experts), "rollout_tc: sizes not eligible (D %d S %d H %d A %d heads %d)", D, S, H, A, 1 + n_experts);
    PlanBuilder pb;
    TcHeader h;
    build_plan(D, S, H, A, 1 + n_experts, pb, h);
    MRSSM_CHECK(h.n_tiles <= MAX_TILES && h.n_ops <= MAX_OPS && h.n_runs < MAX_RUNS, "rollout_tc: program too long (%d tiles, %d MMAs)", h.n_tiles, h.n_ops);
    if (plan_bytes) *plan_bytes = h.total_bytes;
    if (packed_bytes) *packed_bytes = h.packed_bytes;
    return 0;
}

extern "C" int mrssm_rollout_tc_plan(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts, void* host_buf, int64_t buflen) {
    MRSSM_CHECK(host_buf && tc_eligible(D, S, H, A, 1 + n_experts), "rollout_tc_plan: bad arguments");
    PlanBuilder pb;
    TcHeader h;
    build_plan(D, S, H, A, 1 + n_experts, pb, h);
    MRSSM_CHECK(h.n_tiles <= MAX_TILES && h.n_ops <= MAX_OPS && h.n_runs < MAX_RUNS, "rollout_tc: program too long (%d tiles, %d MMAs)", h.n_tiles, h.n_ops);
    MRSSM_CHECK(buflen >= (int64_t)h.total_bytes, "rollout_tc_plan: buffer too small (%lld < %u)", (long long)buflen, h.total_bytes);
    uint8_t* o = (uint8_t*)host_buf;
    memset(o, 0, h.total_bytes);
    memcpy(o, &h, sizeof(h));
    memcpy(o + h.tile_off, pb.tiles.data(), pb.tiles.size() * sizeof(TcTile));
    memcpy(o + h.op_off, pb.ops.data(), pb.ops.size() * sizeof(TcOp));
    memcpy(o + h.pack_off, pb.packs.data(), pb.packs.size() * sizeof(TcPack));
    memcpy(o + h.run_off, pb.runs.data(), pb.runs.size() * sizeof(TcRun));
    return 0;
}

// a: weights in PyTorch layout ([out][in] fp32; ld1[h] = row stride of w1[h]); plan_dev: device copy of the plan;
// n_pack: the plan header's n_pack (host knows it from its own copy)
extern "C" int mrssm_rollout_tc_pack(const mrssm_rollout_args* a, const void* plan_dev, int32_t n_pack, void* packed_dev, void* stream) {
    MRSSM_CHECK(a && plan_dev && packed_dev && n_pack > 0, "rollout_tc_pack: bad arguments");
    const int NH = 1 + a->n_experts;
    MRSSM_CHECK(tc_eligible(a->D, a->S, a->H, a->A, NH), "rollout_tc_pack: sizes not eligible");
    PackSrc src;
    memset(&src, 0, sizeof(src));
    src.p[SRC_WSA] = a->w_sa; src.ld[SRC_WSA] = a->S + a->A;
    src.p[SRC_WIH] = a->w_ih; src.ld[SRC_WIH] = a->D;
    src.p[SRC_WHH] = a->w_hh; src.ld[SRC_WHH] = a->D;
    for (int h = 0; h < NH; ++h) {
        MRSSM_CHECK(a->w1[h] && a->w2[h] && a->ld1[h] >= a->D, "rollout_tc_pack: head %d weights missing", h);
        src.p[SRC_W1 + h] = a->w1[h]; src.ld[SRC_W1 + h] = (int32_t)a->ld1[h];
        src.p[SRC_W2 + h] = a->w2[h]; src.ld[SRC_W2 + h] = a->H;
    }
    MRSSM_CHECK(a->w_sa && a->w_ih && a->w_hh, "rollout_tc_pack: weights missing");
    rollout_tc_pack_kernel<<<n_pack, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)plan_dev, src, (bf16*)packed_dev);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

static long long* g_tc_prof = nullptr;
static int g_tc_rpg = 0;
static int g_tc_static = 1;
/* tuning / test aid: 0 forces the table-driven MMA issue even for the sizes that have a statically unrolled program */
extern "C" int mrssm_rollout_tc_set_static(int32_t on) {
    g_tc_static = on ? 1 : 0;
    return 0;
}
/* tuning aid: sequences per 16-row group of a CTA (16 -> 64 sequences per CTA, 8 -> 32; 0 = automatic) */
extern "C" int mrssm_rollout_tc_set_rows(int32_t rows_per_group) {
    MRSSM_CHECK(rows_per_group == 0 || rows_per_group == 8 || rows_per_group == 16, "rollout_tc_set_rows: 0, 8 or 16");
    g_tc_rpg = rows_per_group;
    return 0;
}
/* tuning aid: device buffer of 3*512 int64 receiving clock64 stamps of one time step (NULL = off) */
extern "C" int mrssm_rollout_tc_set_profile_buffer(void* dev_buf) {
    g_tc_prof = (long long*)dev_buf;
    return 0;
}

extern "C" int mrssm_rollout_tc_fwd(const mrssm_rollout_args* a, const void* plan_dev, const void* packed_dev, void* stream) {
    MRSSM_CHECK(a && plan_dev && packed_dev, "rollout_tc_fwd: bad arguments");
    MRSSM_CHECK(tc_eligible(a->D, a->S, a->H, a->A, 1 + a->n_experts), "rollout_tc_fwd: sizes not eligible");
    MRSSM_CHECK(a->T > 0 && a->B > 0 && a->prev_state && a->prev_belief && a->actions && a->beliefs && a->prior_states && a->prior_means &&
                    a->prior_stds && a->b_sa && a->b_ih && a->b_hh,
                "rollout_tc_fwd: missing tensors");
    for (int h = 0; h <= a->n_experts; ++h) MRSSM_CHECK(a->b2[h] && (a->b1[h] || a->emb_pre[h]), "rollout_tc_fwd: head %d bias missing", h);
    if (a->n_experts > 0) {
        MRSSM_CHECK(a->post_states && a->post_means && a->post_stds, "rollout_tc_fwd: posterior outputs missing");
        for (int h = 1; h <= a->n_experts; ++h) MRSSM_CHECK(a->exp_means[h] && a->exp_stds[h], "rollout_tc_fwd: expert outputs missing");
        MRSSM_CHECK(a->det || a->eps_post, "rollout_tc_fwd: eps_post missing");
    }
    MRSSM_CHECK(a->det || a->eps_prior, "rollout_tc_fwd: eps_prior missing");
    const int NH = 1 + a->n_experts;
    const bool stat = g_tc_static && a->D == 200 && a->S == 30 && a->H == 200 && a->A == 3 && (NH == 1 || NH == 2 || NH == 4);
    // 64 sequences per CTA when that already fills the machine, else 32 (one valid row per epilogue thread)
    const int RPG = g_tc_rpg ? g_tc_rpg : ((a->B + 63) / 64 >= 74 ? 16 : 8);
    auto kern = rollout_tc_fwd_kernel<0, false>;
    if (RPG == 16) kern = (stat && NH == 4) ? rollout_tc_fwd_kernel<4, true> : rollout_tc_fwd_kernel<0, true>;
    else if (stat) kern = NH == 1 ? rollout_tc_fwd_kernel<1, false> : (NH == 2 ? rollout_tc_fwd_kernel<2, false> : rollout_tc_fwd_kernel<4, false>);
    cudaFuncAttributes fa;
    MRSSM_CUDA(cudaFuncGetAttributes(&fa, kern));
    static thread_local TcProg prog;
    static thread_local int prog_key[5] = {-1, -1, -1, -1, -1};
    if (prog_key[0] != a->D || prog_key[1] != a->S || prog_key[2] != a->H || prog_key[3] != a->A || prog_key[4] != a->n_experts) {
        PlanBuilder pb;
        TcHeader h;
        build_plan(a->D, a->S, a->H, a->A, 1 + a->n_experts, pb, h);
        MRSSM_CHECK(h.n_tiles <= PROG_TILES && h.n_runs <= PROG_RUNS, "rollout_tc: program too long (%d tiles, %d runs)", h.n_tiles, h.n_runs);
        memset(&prog, 0, sizeof(prog));
        for (int i = 0; i < h.n_runs; ++i) {
            const TcRun& o = pb.runs[i];
            const uint32_t n = ((o.idesc >> 17) & 63u) << 3;
            uint32_t boff = OFF_XIN, flip = 0;
            if (o.a_buf == BUF_XU) boff = OFF_XU;
            if (o.a_buf == BUF_HPREV) { boff = OFF_H0; flip = 1; }      // odd steps: + (OFF_H1 - OFF_H0)
            if (o.a_buf == BUF_HNEW) { boff = OFF_H1; flip = 2; }       // odd steps: - (OFF_H1 - OFF_H0)
            prog.runs[i].x = ((boff >> 4) + o.a_off16) | (((uint32_t)CH_BYTES >> 4) << 16);
            prog.runs[i].y = (uint32_t)o.b_off16 | (n << 16);           // LBO = N * 16 bytes
            prog.runs[i].z = o.idesc;
            prog.runs[i].w = (uint32_t)o.d_col | ((uint32_t)o.acc_first << 10) | (flip << 11) | (o.count << 13);
        }
        for (int i = 0; i < h.n_tiles; ++i) {
            const TcTile& t = pb.tiles[i];
            prog.tiles[i] = (t.bytes >> 4) | ((uint32_t)(t.run_end - t.run_begin) << 10) | ((uint32_t)(t.wait_ev + 1) << 13) | ((uint32_t)(t.commit + 1) << 18);
        }
        prog.n_tiles = h.n_tiles; prog.cA = h.cA; prog.nD8 = h.nD8; prog.cAH = h.cAH; prog.nH8 = h.nH8;
        prog_key[0] = a->D; prog_key[1] = a->S; prog_key[2] = a->H; prog_key[3] = a->A; prog_key[4] = a->n_experts;
    }
    const int avail = 232448 - (int)fa.sharedSizeBytes - 1024 - OFF_RING;
    const int NS = std::min(MAX_SLOTS, avail / SLOT_BYTES);
    MRSSM_CHECK(NS >= 2, "rollout_tc_fwd: no room for the weight ring (%d bytes left)", avail);
    const size_t dyn = (size_t)OFF_RING + (size_t)NS * SLOT_BYTES + 1024;
    MRSSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<(a->B + 4 * RPG - 1) / (4 * RPG), TC_THREADS, dyn, (cudaStream_t)stream>>>(*a, prog, (const uint8_t*)packed_dev, NS, g_tc_prof);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

// ---- backward ----------------------------------------------------------------------------------------------------
extern "C" int mrssm_rollout_tc_bwd_plan_bytes(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts, int64_t* plan_bytes,
                                               int64_t* packed_bytes) {
    MRSSM_CHECK(tc_eligible(D, S, H, A, 1 + n_experts), "rollout_tc: sizes not eligible (D %d S %d H %d A %d heads %d)", D, S, H, A, 1 + n_experts);
    PlanBuilder pb;
    TcHeader h;
    build_plan_bwd(D, S, H, A, 1 + n_experts, pb, h);
    MRSSM_CHECK(h.n_tiles <= PROG_TILES && h.n_runs <= PROG_RUNS, "rollout_tc_bwd: program too long (%d tiles, %d runs)", h.n_tiles, h.n_runs);
    if (plan_bytes) *plan_bytes = h.total_bytes;
    if (packed_bytes) *packed_bytes = h.packed_bytes;
    return 0;
}

extern "C" int mrssm_rollout_tc_bwd_plan(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts, void* host_buf, int64_t buflen) {
    MRSSM_CHECK(host_buf && tc_eligible(D, S, H, A, 1 + n_experts), "rollout_tc_bwd_plan: bad arguments");
    PlanBuilder pb;
    TcHeader h;
    build_plan_bwd(D, S, H, A, 1 + n_experts, pb, h);
    MRSSM_CHECK(h.n_tiles <= PROG_TILES && h.n_runs <= PROG_RUNS, "rollout_tc_bwd: program too long (%d tiles, %d runs)", h.n_tiles, h.n_runs);
    MRSSM_CHECK(buflen >= (int64_t)h.total_bytes, "rollout_tc_bwd_plan: buffer too small (%lld < %u)", (long long)buflen, h.total_bytes);
    uint8_t* o = (uint8_t*)host_buf;
    memset(o, 0, h.total_bytes);
    memcpy(o, &h, sizeof(h));
    memcpy(o + h.tile_off, pb.tiles.data(), pb.tiles.size() * sizeof(TcTile));
    memcpy(o + h.op_off, pb.ops.data(), pb.ops.size() * sizeof(TcOp));
    memcpy(o + h.pack_off, pb.packs.data(), pb.packs.size() * sizeof(TcPack));
    memcpy(o + h.run_off, pb.runs.data(), pb.runs.size() * sizeof(TcRun));
    return 0;
}

static void fill_prog(const PlanBuilder& pb, const TcHeader& h, TcProg& prog) {
    memset(&prog, 0, sizeof(prog));
    for (int i = 0; i < h.n_runs; ++i) {
        const TcRun& o = pb.runs[i];
        const uint32_t n = ((o.idesc >> 17) & 63u) << 3;
        prog.runs[i].x = (uint32_t)o.a_off16 | (((uint32_t)CH_BYTES >> 4) << 16);
        prog.runs[i].y = (uint32_t)o.b_off16 | (n << 16);
        prog.runs[i].z = o.idesc;
        prog.runs[i].w = (uint32_t)o.d_col | ((uint32_t)o.acc_first << 10) | (o.count << 13);
    }
    for (int i = 0; i < h.n_tiles; ++i) {
        const TcTile& t = pb.tiles[i];
        prog.tiles[i] = (t.bytes >> 4) | ((uint32_t)(t.run_end - t.run_begin) << 10) | ((uint32_t)(t.wait_ev + 1) << 13) | ((uint32_t)(t.commit + 1) << 18);
    }
    prog.n_tiles = h.n_tiles; prog.cA = h.cA; prog.nD8 = h.nD8; prog.cAH = h.cAH; prog.nH8 = h.nH8;
}

int rollout_bwd_check(const mrssm_rollout_bwd_args* g);

// g: as mrssm_rollout_bwd (the weights are NOT read through the struct: packed_dev is the stream packed by mrssm_rollout_tc_pack
// from the plan of mrssm_rollout_tc_bwd_plan)
extern "C" int mrssm_rollout_tc_bwd(const mrssm_rollout_bwd_args* g, const void* packed_dev, void* stream) {
    MRSSM_CHECK(g && packed_dev, "rollout_tc_bwd: bad arguments");
    const mrssm_rollout_args* a = &g->f;
    MRSSM_CHECK(tc_eligible(a->D, a->S, a->H, a->A, 1 + a->n_experts), "rollout_tc_bwd: sizes not eligible");
    MRSSM_CHECK(a->T > 0 && a->B > 0 && a->prev_state && a->prev_belief && a->actions && a->beliefs && a->prior_states && a->prior_stds &&
                    a->st_x && a->st_r && a->st_z && a->st_n && a->st_ghn && g->d_xpre && g->d_gi && g->d_gh,
                "rollout_tc_bwd: missing tensors");
    for (int h = 0; h <= a->n_experts; ++h) MRSSM_CHECK(a->st_u[h] && g->d_u[h] && g->d_o[h], "rollout_tc_bwd: head %d buffers missing", h);
    if (a->n_experts > 0) {
        MRSSM_CHECK(a->post_states && a->post_means && a->post_stds, "rollout_tc_bwd: posterior tensors missing");
        for (int h = 1; h <= a->n_experts; ++h) MRSSM_CHECK(a->exp_means[h] && a->exp_stds[h], "rollout_tc_bwd: expert tensors missing");
        MRSSM_CHECK(a->det || a->eps_post, "rollout_tc_bwd: eps_post missing");
    }
    MRSSM_CHECK(a->det || a->eps_prior, "rollout_tc_bwd: eps_prior missing");
    static thread_local TcProg prog;
    static thread_local int key[5] = {-1, -1, -1, -1, -1};
    if (key[0] != a->D || key[1] != a->S || key[2] != a->H || key[3] != a->A || key[4] != a->n_experts) {
        PlanBuilder pb;
        TcHeader h;
        build_plan_bwd(a->D, a->S, a->H, a->A, 1 + a->n_experts, pb, h);
        MRSSM_CHECK(h.n_tiles <= PROG_TILES && h.n_runs <= PROG_RUNS, "rollout_tc_bwd: program too long (%d tiles, %d runs)", h.n_tiles, h.n_runs);
        fill_prog(pb, h, prog);
        key[0] = a->D; key[1] = a->S; key[2] = a->H; key[3] = a->A; key[4] = a->n_experts;
    }
    const int NH = 1 + a->n_experts;
    const bool stat = g_tc_static && a->D == 200 && a->S == 30 && a->H == 200 && a->A == 3 && (NH == 1 || NH == 2 || NH == 4);
    const int RPG = g_tc_rpg ? g_tc_rpg : ((a->B + 63) / 64 >= 74 ? 16 : 8);
    auto kern = rollout_tc_bwd_kernel<0, false>;
    if (RPG == 16) kern = (stat && NH == 4) ? rollout_tc_bwd_kernel<4, true> : rollout_tc_bwd_kernel<0, true>;
    else if (stat) kern = NH == 1 ? rollout_tc_bwd_kernel<1, false> : (NH == 2 ? rollout_tc_bwd_kernel<2, false> : rollout_tc_bwd_kernel<4, false>);
    cudaFuncAttributes fa;
    MRSSM_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int avail = 232448 - (int)fa.sharedSizeBytes - 1024 - BOFF_RING;
    const int NS = std::min(MAX_SLOTS, avail / SLOT_BYTES);
    MRSSM_CHECK(NS >= 2, "rollout_tc_bwd: no room for the weight ring (%d bytes left)", avail);
    const size_t dyn = (size_t)BOFF_RING + (size_t)NS * SLOT_BYTES + 1024;
    MRSSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<(a->B + 4 * RPG - 1) / (4 * RPG), TC_THREADS, dyn, (cudaStream_t)stream>>>(*g, prog, (const uint8_t*)packed_dev, NS, g_tc_prof);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
