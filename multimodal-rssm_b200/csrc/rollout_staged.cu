// Fused RSSM rollout (forward and BPTT backward), fp32 CUDA-core version with the weights STREAMED through shared
// memory: same arithmetic, same summation order and same outputs as rollout_simt.cu, but no thread ever waits on
// an L2 round trip for a weight.  All weight matrices of one time step are cut into row tiles (<= 38 KB); one
// thread keeps two tiles in flight ahead of the consumers with cp.async.bulk + mbarrier (a 3-slot ring), and
// every thread reads its weight column from shared memory, conflict free.  The tile schedule is identical for
// every step, so the prefetch runs across stage and step boundaries.  Used when one output feature per thread
// covers D and H (D, H <= 256); otherwise rollout_simt.cu runs.
// Reference: utils/models/transition_model.py:200-285 (+ :50-114), encoder.py:50-155 and their autograd.
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int RING_FWD = 4, RING_BWD = 3;   // ring slots (RING-1 tiles in flight): as many as shared memory allows
constexpr int TILE_FLOATS = 9600;          // 38 400 B per ring slot
constexpr int MAX_TILES = 96;
constexpr int FC1_PER_THREAD = 4;          // ceil(NH*H / NT) <= 4

struct Seg {
    const float* src;
    uint32_t bytes, dst_off;               // dst_off in floats
};
struct Tile {
    Seg seg[MRSSM_MAX_HEADS];
    int nseg;
    uint32_t total;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; !done; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (it > (1u << 24)) {
            printf("mrssm rollout: weight tile wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

template <int BT>
__device__ __forceinline__ void load_bt(float (&v)[BT], const float* p) {
#pragma unroll
    for (int b = 0; b < BT; b += 4) {
        float4 q = *reinterpret_cast<const float4*>(p + b);
        v[b] = q.x; v[b + 1] = q.y; v[b + 2] = q.z; v[b + 3] = q.w;
    }
}

// The weight streamer shared by both kernels.
template <int RING>
struct Streamer {
    Tile* tiles;
    uint64_t* full;
    float* ring;
    int ntiles;                 // tiles per time step
    long long total;            // tiles over the whole launch
    long long g;                // next tile to consume

    __device__ __forceinline__ void issue(long long gi) const {
        const Tile& t = tiles[gi % ntiles];
        const int slot = (int)(gi % RING);
        const uint32_t bar = smem_u32(&full[slot]);
        mbar_expect_tx(bar, t.total);
        const uint32_t dst = smem_u32(ring + (size_t)slot * TILE_FLOATS);
        for (int i = 0; i < t.nseg; ++i) bulk_g2s(dst + t.seg[i].dst_off * 4u, t.seg[i].src, t.seg[i].bytes, bar);
    }
    // every thread: returns the slot holding tile g (and keeps two more in flight)
    __device__ __forceinline__ const float* acquire() {
        if (threadIdx.x == 0 && g + RING - 1 < total) issue(g + RING - 1);      // that slot was released by the barrier after tile g-1
        const int slot = (int)(g % RING);
        mbar_wait(smem_u32(&full[slot]), (uint32_t)((g / RING) & 1));
        return ring + (size_t)slot * TILE_FLOATS;
    }
    __device__ __forceinline__ void release() {
        __syncthreads();
        ++g;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(NT) rollout_fwd_staged_kernel(mrssm_rollout_args a, int KC2, int KC3, int KC4) {
    constexpr int RING = RING_FWD;
    extern __shared__ __align__(128) float smem[];
    __shared__ Tile tiles[MAX_TILES];
    __shared__ uint64_t full[RING];
    const int tid = threadIdx.x;
    const int D = a.D, S = a.S, H = a.H, A = a.A, E = a.n_experts, NH = 1 + E, B = a.B;
    const int b0 = blockIdx.x * BT;
    float* ring = smem;                                  // [RING][TILE_FLOATS]
    float* xin = ring + RING * TILE_FLOATS;              // [S+A][BT]
    float* x = xin + (S + A) * BT;                       // [D][BT]
    float* hA = x + D * BT;
    float* hB = hA + D * BT;
    float* u = hB + D * BT;                              // [NH*H][BT]
    float* o = u + NH * H * BT;                          // [NH*2S][BT]
    float* sprev = o + NH * 2 * S * BT;                  // [S][BT]
    float* hprev = hA;
    float* hcur = hB;

    Streamer<RING> W;
    W.tiles = tiles; W.full = full; W.ring = ring; W.g = 0;
    if (tid == 0) {
        int n = 0;
        {   // stage 1: W_sa^T [S+A][D], one tile
            Tile& t = tiles[n++];
            t.nseg = 1;
            t.seg[0] = Seg{a.w_sa, (uint32_t)((S + A) * D * 4), 0u};
            t.total = t.seg[0].bytes;
        }
        for (int k0 = 0; k0 < D; k0 += KC2) {   // stage 2: rows of W_ih^T and W_hh^T ([D][3D])
            const int kc = min(KC2, D - k0);
            Tile& t = tiles[n++];
            t.nseg = 2;
            t.seg[0] = Seg{a.w_ih + (size_t)k0 * 3 * D, (uint32_t)(kc * 3 * D * 4), 0u};
            t.seg[1] = Seg{a.w_hh + (size_t)k0 * 3 * D, (uint32_t)(kc * 3 * D * 4), (uint32_t)(KC2 * 3 * D)};
            t.total = t.seg[0].bytes + t.seg[1].bytes;
        }
        for (int k0 = 0; k0 < D; k0 += KC3) {   // stage 3: rows of every head's fc1^T ([D][H], belief part)
            const int kc = min(KC3, D - k0);
            Tile& t = tiles[n++];
            t.nseg = NH;
            t.total = 0;
            for (int hd = 0; hd < NH; ++hd) {
                t.seg[hd] = Seg{a.w1[hd] + (size_t)k0 * H, (uint32_t)(kc * H * 4), (uint32_t)(hd * KC3 * H)};
                t.total += t.seg[hd].bytes;
            }
        }
        for (int k0 = 0; k0 < H; k0 += KC4) {   // stage 4: rows of every head's fc2^T ([H][2S])
            const int kc = min(KC4, H - k0);
            Tile& t = tiles[n++];
            t.nseg = NH;
            t.total = 0;
            for (int hd = 0; hd < NH; ++hd) {
                t.seg[hd] = Seg{a.w2[hd] + (size_t)k0 * 2 * S, (uint32_t)(kc * 2 * S * 4), (uint32_t)(hd * KC4 * 2 * S)};
                t.total += t.seg[hd].bytes;
            }
        }
        tiles[MAX_TILES - 1].nseg = n;       // publish the count through shared memory
        for (int s = 0; s < RING; ++s) mbar_init(smem_u32(&full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < D * BT; i += NT) {
        int k = i / BT, b = i % BT;
        hprev[i] = (b0 + b < B) ? a.prev_belief[(long long)(b0 + b) * D + k] : 0.f;
    }
    for (int i = tid; i < S * BT; i += NT) {
        int k = i / BT, b = i % BT;
        sprev[i] = (b0 + b < B) ? a.prev_state[(long long)(b0 + b) * S + k] : 0.f;
    }
    __syncthreads();
    W.ntiles = tiles[MAX_TILES - 1].nseg;
    W.total = (long long)W.ntiles * a.T;
    if (tid == 0) {
        for (int i = 0; i < RING - 1 && i < W.total; ++i) W.issue(i);
    }

    for (int t = 0; t < a.T; ++t) {
        const long long tb = (long long)t * B + b0;
        // ---- 0. masked [state, action] ----------------------------------------------------------
        for (int i = tid; i < (S + A) * BT; i += NT) {
            int k = i / BT, b = i % BT;
            float v = 0.f;
            if (b0 + b < B) {
                if (k < S) {
                    float m = a.nonterminals ? a.nonterminals[tb + b] : 1.f;
                    v = sprev[i] * m;
                } else {
                    v = a.actions[(tb + b) * A + (k - S)];
                }
            }
            xin[i] = v;
        }
        __syncthreads();
        // ---- 1. x = act(W_sa xin + b) -----------------------------------------------------------
        {
            const float* ws = W.acquire();
            const int j = tid;
            if (j < D) {
                float acc[BT];
                const float bj = a.b_sa[j];
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = bj;
                for (int k = 0; k < S + A; ++k) {
                    const float w = ws[k * D + j];
                    float v[BT];
                    load_bt<BT>(v, xin + k * BT);
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] = fmaf(w, v[b], acc[b]);
                }
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    const float xv = act_apply(acc[b], a.act);
                    x[j * BT + b] = xv;
                    if (a.st_x && b0 + b < B) a.st_x[(tb + b) * D + j] = xv;
                }
            }
            W.release();
        }
        // ---- 2. GRUCell --------------------------------------------------------------------------
        {
            const int j = tid;
            float ar[BT], az[BT], ain[BT], ahn[BT];
            if (j < D) {
                const float br = a.b_ih[j] + a.b_hh[j], bz = a.b_ih[D + j] + a.b_hh[D + j];
                const float bin = a.b_ih[2 * D + j], bhn = a.b_hh[2 * D + j];
#pragma unroll
                for (int b = 0; b < BT; ++b) { ar[b] = br; az[b] = bz; ain[b] = bin; ahn[b] = bhn; }
            }
            for (int k0 = 0; k0 < D; k0 += KC2) {
                const float* ws = W.acquire();
                const int kc = min(KC2, D - k0);
                if (j < D) {
                    const float* wi = ws + j;
                    const float* wh = ws + KC2 * 3 * D + j;
#pragma unroll 2
                    for (int kk = 0; kk < kc; ++kk) {
                        const float wir = wi[0], wiz = wi[D], win = wi[2 * D];
                        const float whr = wh[0], whz = wh[D], whn = wh[2 * D];
                        wi += 3 * D;
                        wh += 3 * D;
                        float xv[BT], hv[BT];
                        load_bt<BT>(xv, x + (k0 + kk) * BT);
                        load_bt<BT>(hv, hprev + (k0 + kk) * BT);
#pragma unroll
                        for (int b = 0; b < BT; ++b) {
                            ar[b] = fmaf(wir, xv[b], fmaf(whr, hv[b], ar[b]));
                            az[b] = fmaf(wiz, xv[b], fmaf(whz, hv[b], az[b]));
                            ain[b] = fmaf(win, xv[b], ain[b]);
                            ahn[b] = fmaf(whn, hv[b], ahn[b]);
                        }
                    }
                }
                if (k0 + KC2 >= D && j < D) {      // last tile: gate nonlinearities (hcur is not read by anyone yet)
#pragma unroll
                    for (int b = 0; b < BT; ++b) {
                        const float r = sigmoidf_(ar[b]), z = sigmoidf_(az[b]);
                        const float n = tanhf(ain[b] + r * ahn[b]);
                        const float hp = hprev[j * BT + b];
                        const float hn = (1.f - z) * n + z * hp;
                        hcur[j * BT + b] = hn;
                        if (b0 + b < B) {
                            const long long off = (tb + b) * D + j;
                            a.beliefs[off] = hn;
                            if (a.st_r) { a.st_r[off] = r; a.st_z[off] = z; a.st_n[off] = n; a.st_ghn[off] = ahn[b]; }
                        }
                    }
                }
                W.release();
            }
        }
        // ---- 3. heads fc1 (+ hoisted embedding half) + act ---------------------------------------
        {
            float acc[FC1_PER_THREAD][BT];
            int woff[FC1_PER_THREAD], hdm[FC1_PER_THREAD], cm[FC1_PER_THREAD];
#pragma unroll
            for (int m = 0; m < FC1_PER_THREAD; ++m) {
                const int idx = tid + m * NT;
                woff[m] = -1;
                hdm[m] = cm[m] = 0;
                if (idx < NH * H) {
                    const int hd = idx / H, c = idx - hd * H;
                    hdm[m] = hd;
                    cm[m] = c;
                    woff[m] = hd * KC3 * H + c;
                    const float bj = a.b1[hd] ? a.b1[hd][c] : 0.f;
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[m][b] = bj;
                    if (a.emb_pre[hd]) {
#pragma unroll
                        for (int b = 0; b < BT; ++b)
                            if (b0 + b < B) acc[m][b] += a.emb_pre[hd][(tb + b) * H + c];
                    }
                }
            }
            for (int k0 = 0; k0 < D; k0 += KC3) {
                const float* ws = W.acquire();
                const int kc = min(KC3, D - k0);
#pragma unroll 2
                for (int kk = 0; kk < kc; ++kk) {
                    float hv[BT];
                    load_bt<BT>(hv, hcur + (k0 + kk) * BT);
#pragma unroll
                    for (int m = 0; m < FC1_PER_THREAD; ++m) {
                        if (woff[m] >= 0) {
                            const float wv = ws[woff[m] + kk * H];
#pragma unroll
                            for (int b = 0; b < BT; ++b) acc[m][b] = fmaf(wv, hv[b], acc[m][b]);
                        }
                    }
                }
                if (k0 + KC3 >= D) {
#pragma unroll
                    for (int m = 0; m < FC1_PER_THREAD; ++m) {
                        if (woff[m] >= 0) {
                            const int idx = tid + m * NT, hd = hdm[m], c = cm[m];
#pragma unroll
                            for (int b = 0; b < BT; ++b) {
                                const float uv = act_apply(acc[m][b], a.act);
                                u[idx * BT + b] = uv;
                                if (a.st_u[hd] && b0 + b < B) a.st_u[hd][(tb + b) * H + c] = uv;
                            }
                        }
                    }
                }
                W.release();
            }
        }
        // ---- 4. heads fc2 -> (mean, softplus+min_std) --------------------------------------------
        {
            const int idx = tid;
            const bool live = idx < NH * 2 * S;
            const int hd = live ? idx / (2 * S) : 0, c = live ? idx - hd * 2 * S : 0;
            float acc[BT];
            const float bj = live ? a.b2[hd][c] : 0.f;
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] = bj;
            const float* uh = u + hd * H * BT;
            for (int k0 = 0; k0 < H; k0 += KC4) {
                const float* ws = W.acquire();
                const int kc = min(KC4, H - k0);
                if (live) {
                    const float* w = ws + hd * KC4 * 2 * S + c;
#pragma unroll 4
                    for (int kk = 0; kk < kc; ++kk) {
                        const float wv = w[kk * 2 * S];
                        float uv[BT];
                        load_bt<BT>(uv, uh + (k0 + kk) * BT);
#pragma unroll
                        for (int b = 0; b < BT; ++b) acc[b] = fmaf(wv, uv[b], acc[b]);
                    }
                    if (k0 + KC4 >= H) {
#pragma unroll
                        for (int b = 0; b < BT; ++b) o[idx * BT + b] = (c < S) ? acc[b] : softplusf_(acc[b]) + a.min_std;
                    }
                }
                W.release();
            }
        }
        // ---- 5. prior sample, expert outputs, fusion, posterior sample ---------------------------
        for (int idx = tid; idx < S * BT; idx += NT) {
            int b = idx / S, s = idx % S;
            if (b0 + b >= B) continue;
            long long off = (tb + b) * S + s;
            float pm = o[s * BT + b], ps = o[(S + s) * BT + b];
            float pst = a.det ? pm : fmaf(ps, a.eps_prior[off], pm);
            a.prior_means[off] = pm;
            a.prior_stds[off] = ps;
            a.prior_states[off] = pst;
            float nxt = pst;
            if (E > 0) {
                float qm, qs;
                if (a.n_subsets == 0) {
                    qm = o[(2 * S + s) * BT + b];
                    qs = o[(2 * S + S + s) * BT + b];
                } else {
                    unsigned mask = a.subset_mask[a.dim_subset[s]];
                    float sumT = 0.f, sumMT = 0.f;
                    for (int e = 1; e <= E; ++e) {
                        if (mask & (1u << (e - 1))) {
                            float tt = 1.f / o[(e * 2 * S + S + s) * BT + b];
                            sumT += tt;
                            sumMT = fmaf(o[(e * 2 * S + s) * BT + b], tt, sumMT);
                        }
                    }
                    qm = sumMT / sumT;
                    qs = 1.f / sumT;
                }
                for (int e = 1; e <= E; ++e) {
                    a.exp_means[e][off] = o[(e * 2 * S + s) * BT + b];
                    a.exp_stds[e][off] = o[(e * 2 * S + S + s) * BT + b];
                }
                float qst = a.det ? qm : fmaf(qs, a.eps_post[off], qm);
                a.post_means[off] = qm;
                a.post_stds[off] = qs;
                a.post_states[off] = qst;
                nxt = qst;
            }
            sprev[s * BT + b] = nxt;
        }
        __syncthreads();
        float* tmp = hprev; hprev = hcur; hcur = tmp;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward.  Weights in PyTorch layout; w1c[hd] are CONTIGUOUS [H][D] copies of the belief columns of fc1.
// ------------------------------------------------------------------------------------------------------------------
struct BwdExtra {
    const float* w1c[MRSSM_MAX_HEADS];
    int KCb, KCc, KCe;
};

template <int BT>
__global__ void __launch_bounds__(NT) rollout_bwd_staged_kernel(mrssm_rollout_bwd_args g, BwdExtra X) {
    constexpr int RING = RING_BWD;
    extern __shared__ __align__(128) float smem[];
    __shared__ Tile tiles[MAX_TILES];
    __shared__ uint64_t full[RING];
    const mrssm_rollout_args& a = g.f;
    const int tid = threadIdx.x;
    const int D = a.D, S = a.S, H = a.H, A = a.A, E = a.n_experts, NH = 1 + E, B = a.B;
    const int b0 = blockIdx.x * BT;
    const int KCb = X.KCb, KCc = X.KCc, KCe = X.KCe;
    float* ring = smem;
    float* cgh = ring + RING * TILE_FLOATS;   // [D][BT]   carried grad wrt h_t
    float* cgs = cgh + D * BT;                // [S][BT]   carried grad wrt the fed-back state
    float* go = cgs + S * BT;                 // [NH*2S][BT]
    float* gu = go + NH * 2 * S * BT;         // [NH*H][BT]
    float* Gh = gu + NH * H * BT;             // [D][BT]
    float* dgi = Gh + D * BT;                 // [3D][BT]
    float* dghn = dgi + 3 * D * BT;           // [D][BT]
    float* dxp = dghn + D * BT;               // [D][BT]

    Streamer<RING> W;
    W.tiles = tiles; W.full = full; W.ring = ring; W.g = 0;
    if (tid == 0) {
        int n = 0;
        for (int j0 = 0; j0 < 2 * S; j0 += KCb) {   // stage b: rows of every head's fc2 ([2S][H])
            const int kc = min(KCb, 2 * S - j0);
            Tile& t = tiles[n++];
            t.nseg = NH;
            t.total = 0;
            for (int hd = 0; hd < NH; ++hd) {
                t.seg[hd] = Seg{a.w2[hd] + (size_t)j0 * H, (uint32_t)(kc * H * 4), (uint32_t)(hd * KCb * H)};
                t.total += t.seg[hd].bytes;
            }
        }
        for (int c0 = 0; c0 < H; c0 += KCc) {       // stage c: rows of every head's fc1 belief part ([H][D], contiguous copy)
            const int kc = min(KCc, H - c0);
            Tile& t = tiles[n++];
            t.nseg = NH;
            t.total = 0;
            for (int hd = 0; hd < NH; ++hd) {
                t.seg[hd] = Seg{X.w1c[hd] + (size_t)c0 * D, (uint32_t)(kc * D * 4), (uint32_t)(hd * KCc * D)};
                t.total += t.seg[hd].bytes;
            }
        }
        for (int j0 = 0; j0 < 3 * D; j0 += KCe) {   // stage e: rows of W_ih and W_hh ([3D][D])
            const int kc = min(KCe, 3 * D - j0);
            Tile& t = tiles[n++];
            t.nseg = 2;
            t.seg[0] = Seg{a.w_ih + (size_t)j0 * D, (uint32_t)(kc * D * 4), 0u};
            t.seg[1] = Seg{a.w_hh + (size_t)j0 * D, (uint32_t)(kc * D * 4), (uint32_t)(KCe * D)};
            t.total = t.seg[0].bytes + t.seg[1].bytes;
        }
        {   // stage f: W_sa [D][S+A], one tile
            Tile& t = tiles[n++];
            t.nseg = 1;
            t.seg[0] = Seg{a.w_sa, (uint32_t)(D * (S + A) * 4), 0u};
            t.total = t.seg[0].bytes;
        }
        tiles[MAX_TILES - 1].nseg = n;
        for (int s = 0; s < RING; ++s) mbar_init(smem_u32(&full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < D * BT; i += NT) cgh[i] = 0.f;
    for (int i = tid; i < S * BT; i += NT) cgs[i] = 0.f;
    __syncthreads();
    W.ntiles = tiles[MAX_TILES - 1].nseg;
    W.total = (long long)W.ntiles * a.T;
    if (tid == 0) {
        for (int i = 0; i < RING - 1 && i < W.total; ++i) W.issue(i);
    }

    for (int t = a.T - 1; t >= 0; --t) {
        const long long tb = (long long)t * B + b0;
        // ---- a. sample / fusion / softplus backward -> go ----------------------------------------
        for (int idx = tid; idx < S * BT; idx += NT) {
            int b = idx / S, s = idx % S;
            bool live = b0 + b < B;
            long long off = live ? (tb + b) * S + s : 0;
            float carry = cgs[s * BT + b];
            float gps = (g.g_prior_states && live) ? g.g_prior_states[off] : 0.f;
            if (E == 0) gps += carry;
            float gpm = gps + ((g.g_prior_means && live) ? g.g_prior_means[off] : 0.f);
            float gpsd = ((g.g_prior_stds && live) ? g.g_prior_stds[off] : 0.f);
            if (!a.det && live) gpsd = fmaf(gps, a.eps_prior[off], gpsd);
            float psd = live ? a.prior_stds[off] : 1.f;
            go[s * BT + b] = live ? gpm : 0.f;
            go[(S + s) * BT + b] = live ? gpsd * (1.f - expf(-(psd - a.min_std))) : 0.f;
            if (E > 0) {
                float gq = ((g.g_post_states && live) ? g.g_post_states[off] : 0.f) + carry;
                float gqm = gq + ((g.g_post_means && live) ? g.g_post_means[off] : 0.f);
                float gqs = ((g.g_post_stds && live) ? g.g_post_stds[off] : 0.f);
                if (!a.det && live) gqs = fmaf(gq, a.eps_post[off], gqs);
                unsigned mask = a.n_subsets ? a.subset_mask[a.dim_subset[s]] : 1u;
                float P = 0.f, qm = 0.f;
                if (a.n_subsets && live) {
                    P = 1.f / a.post_stds[off];
                    qm = a.post_means[off];
                }
                for (int e = 1; e <= E; ++e) {
                    float gm = (g.g_exp_means[e] && live) ? g.g_exp_means[e][off] : 0.f;
                    float gs = (g.g_exp_stds[e] && live) ? g.g_exp_stds[e][off] : 0.f;
                    float sd = live ? a.exp_stds[e][off] : 1.f;
                    if (live && (mask & (1u << (e - 1)))) {
                        if (a.n_subsets == 0) {
                            gm += gqm;
                            gs += gqs;
                        } else {
                            float te = 1.f / sd;
                            float mu = a.exp_means[e][off];
                            gm = fmaf(gqm, te / P, gm);
                            float gT = gqm * (mu - qm) / P - gqs / (P * P);
                            gs = fmaf(-gT, te * te, gs);
                        }
                    }
                    go[(e * 2 * S + s) * BT + b] = live ? gm : 0.f;
                    go[(e * 2 * S + S + s) * BT + b] = live ? gs * (1.f - expf(-(sd - a.min_std))) : 0.f;
                }
            }
        }
        __syncthreads();
        for (int idx = tid; idx < NH * 2 * S * BT; idx += NT) {     // d_o -> HBM, [T,B,2S] per head
            int b = idx / (NH * 2 * S), r = idx % (NH * 2 * S), hd = r / (2 * S), c = r % (2 * S);
            if (b0 + b < B) g.d_o[hd][(tb + b) * 2 * S + c] = go[r * BT + b];
        }
        // ---- b. gu = (W2^T go) * act'(u) ---------------------------------------------------------
        {
            float acc[FC1_PER_THREAD][BT];
#pragma unroll
            for (int m = 0; m < FC1_PER_THREAD; ++m)
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[m][b] = 0.f;
            for (int j0 = 0; j0 < 2 * S; j0 += KCb) {
                const float* ws = W.acquire();
                const int kc = min(KCb, 2 * S - j0);
#pragma unroll
                for (int m = 0; m < FC1_PER_THREAD; ++m) {
                    const int idx = tid + m * NT;
                    if (idx < NH * H) {
                        const int hd = idx / H, c = idx - hd * H;
                        const float* w = ws + hd * KCb * H + c;
                        const float* gh = go + (hd * 2 * S + j0) * BT;
                        for (int jj = 0; jj < kc; ++jj) {
                            const float wv = w[jj * H];
                            float v[BT];
                            load_bt<BT>(v, gh + jj * BT);
#pragma unroll
                            for (int b = 0; b < BT; ++b) acc[m][b] = fmaf(wv, v[b], acc[m][b]);
                        }
                    }
                }
                if (j0 + KCb >= 2 * S) {
#pragma unroll
                    for (int m = 0; m < FC1_PER_THREAD; ++m) {
                        const int idx = tid + m * NT;
                        if (idx < NH * H) {
                            const int hd = idx / H, c = idx - hd * H;
#pragma unroll
                            for (int b = 0; b < BT; ++b) {
                                float val = 0.f;
                                if (b0 + b < B) {
                                    const long long off = (tb + b) * H + c;
                                    val = acc[m][b] * act_grad_from_out(a.st_u[hd][off], a.act);
                                    g.d_u[hd][off] = val;
                                }
                                gu[idx * BT + b] = val;
                            }
                        }
                    }
                }
                W.release();
            }
        }
        // ---- c. G_h = g_beliefs + carry + sum_heads W1_h^T gu ------------------------------------
        {
            const int k = tid;
            float acc[BT];
            if (k < D) {
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    float v = cgh[k * BT + b];
                    if (g.g_beliefs && b0 + b < B) v += g.g_beliefs[(tb + b) * D + k];
                    acc[b] = v;
                }
            }
            // summation order of rollout_simt.cu: heads outer, hidden units inner.  Tiles carry KCc hidden units of every
            // head, so per-head partial sums are kept and added in head order at the end.
            float part[MRSSM_MAX_HEADS][BT];
#pragma unroll
            for (int hd = 0; hd < MRSSM_MAX_HEADS; ++hd)
#pragma unroll
                for (int b = 0; b < BT; ++b) part[hd][b] = 0.f;
            for (int c0 = 0; c0 < H; c0 += KCc) {
                const float* ws = W.acquire();
                const int kc = min(KCc, H - c0);
                if (k < D) {
#pragma unroll
                    for (int hd = 0; hd < MRSSM_MAX_HEADS; ++hd) {
                        if (hd < NH) {
                            const float* w = ws + hd * KCc * D + k;
                            const float* gg = gu + (hd * H + c0) * BT;
#pragma unroll 4
                            for (int cc = 0; cc < kc; ++cc) {
                                const float wv = w[cc * D];
                                float v[BT];
                                load_bt<BT>(v, gg + cc * BT);
#pragma unroll
                                for (int b = 0; b < BT; ++b) part[hd][b] = fmaf(wv, v[b], part[hd][b]);
                            }
                        }
                    }
                }
                W.release();
            }
            if (k < D) {
#pragma unroll
                for (int hd = 0; hd < MRSSM_MAX_HEADS; ++hd)
                    if (hd < NH) {
#pragma unroll
                        for (int b = 0; b < BT; ++b) acc[b] += part[hd][b];
                    }
#pragma unroll
                for (int b = 0; b < BT; ++b) Gh[k * BT + b] = acc[b];
            }
            __syncthreads();
        }
        // ---- d. GRU gate backward (elementwise) --------------------------------------------------
        for (int idx = tid; idx < D * BT; idx += NT) {
            int b = idx / D, j = idx % D;
            float gpr = 0.f, gpz = 0.f, gpn = 0.f, gpnr = 0.f, direct = 0.f;
            if (b0 + b < B) {
                long long off = (tb + b) * D + j;
                float r = a.st_r[off], z = a.st_z[off], n = a.st_n[off], ghn = a.st_ghn[off];
                float hp = (t > 0) ? a.beliefs[off - (long long)B * D] : a.prev_belief[(long long)(b0 + b) * D + j];
                float G = Gh[j * BT + b];
                float gn = G * (1.f - z), gz = G * (hp - n);
                direct = G * z;
                gpn = gn * (1.f - n * n);
                gpr = gpn * ghn * r * (1.f - r);
                gpz = gz * z * (1.f - z);
                gpnr = gpn * r;
                long long o3 = (tb + b) * 3 * D + j;
                g.d_gi[o3] = gpr; g.d_gi[o3 + D] = gpz; g.d_gi[o3 + 2 * D] = gpn;
                g.d_gh[o3] = gpr; g.d_gh[o3 + D] = gpz; g.d_gh[o3 + 2 * D] = gpnr;
            }
            dgi[j * BT + b] = gpr;
            dgi[(D + j) * BT + b] = gpz;
            dgi[(2 * D + j) * BT + b] = gpn;
            dghn[j * BT + b] = gpnr;
            cgh[j * BT + b] = direct;      // Gh already consumed the old carry
        }
        __syncthreads();
        // ---- e. gx = W_ih^T dgi ; carry_gh += W_hh^T dgh ; dxpre = gx * act'(x) -------------------
        {
            const int k = tid;
            float ax[BT], ah[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) { ax[b] = 0.f; ah[b] = 0.f; }
            for (int j0 = 0; j0 < 3 * D; j0 += KCe) {
                const float* ws = W.acquire();
                const int kc = min(KCe, 3 * D - j0);
                if (k < D) {
                    const float* wi = ws + k;
                    const float* wh = ws + KCe * D + k;
#pragma unroll 2
                    for (int jj = 0; jj < kc; ++jj) {
                        const int j = j0 + jj;
                        const float wiv = wi[jj * D], whv = wh[jj * D];
                        float v[BT], v2[BT];
                        load_bt<BT>(v, dgi + j * BT);
                        if (j < 2 * D) {
#pragma unroll
                            for (int b = 0; b < BT; ++b) v2[b] = v[b];
                        } else {
                            load_bt<BT>(v2, dghn + (j - 2 * D) * BT);
                        }
#pragma unroll
                        for (int b = 0; b < BT; ++b) {
                            ax[b] = fmaf(wiv, v[b], ax[b]);
                            ah[b] = fmaf(whv, v2[b], ah[b]);
                        }
                    }
                    if (j0 + KCe >= 3 * D) {
#pragma unroll
                        for (int b = 0; b < BT; ++b) {
                            float val = 0.f;
                            if (b0 + b < B) {
                                const long long off = (tb + b) * D + k;
                                val = ax[b] * act_grad_from_out(a.st_x[off], a.act);
                                g.d_xpre[off] = val;
                            }
                            dxp[k * BT + b] = val;
                            cgh[k * BT + b] += ah[b];
                        }
                    }
                }
                W.release();
            }
        }
        // ---- f. g_xin = W_sa^T dxpre -> state carry (masked) and action grads; emit xin ----------
        {
            const float* ws = W.acquire();
            const int i = tid;
            if (i < S + A) {
                float acc[BT];
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = 0.f;
                const float* w = ws + i;                // PyTorch layout [D][S+A]
                for (int k = 0; k < D; ++k) {
                    const float wv = w[k * (S + A)];
                    float v[BT];
                    load_bt<BT>(v, dxp + k * BT);
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] = fmaf(wv, v[b], acc[b]);
                }
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    if (b0 + b >= B) { if (i < S) cgs[i * BT + b] = 0.f; continue; }
                    float m = a.nonterminals ? a.nonterminals[tb + b] : 1.f;
                    if (i < S) {
                        cgs[i * BT + b] = acc[b] * m;
                        if (g.xin) {
                            const float* src = (E > 0) ? a.post_states : a.prior_states;
                            float sp = (t > 0) ? src[(tb + b - B) * S + i] : a.prev_state[(long long)(b0 + b) * S + i];
                            g.xin[(tb + b) * (S + A) + i] = sp * m;
                        }
                    } else {
                        float av = a.actions[(tb + b) * A + (i - S)];
                        if (g.g_actions) g.g_actions[(tb + b) * A + (i - S)] = acc[b];
                        if (g.xin) g.xin[(tb + b) * (S + A) + i] = av;
                    }
                }
            }
            W.release();
        }
    }
    for (int i = tid; i < D * BT; i += NT) {
        int k = i / BT, b = i % BT;
        if (b0 + b < B && g.g_prev_belief) g.g_prev_belief[(long long)(b0 + b) * D + k] = cgh[i];
    }
    for (int i = tid; i < S * BT; i += NT) {
        int k = i / BT, b = i % BT;
        if (b0 + b < B && g.g_prev_state) g.g_prev_state[(long long)(b0 + b) * S + k] = cgs[i];
    }
}

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// returns 0 when the staged kernel ran, -1 when the shapes are not eligible (caller falls back), >0 on error
int rollout_fwd_staged(const mrssm_rollout_args* a, cudaStream_t st) {
    const int D = a->D, S = a->S, H = a->H, A = a->A, NH = 1 + a->n_experts;
    constexpr int BT = 8;
    if (D > NT || NH * H > FC1_PER_THREAD * NT || NH * 2 * S > NT || D % 4 || H % 4 || (2 * S) % 4 || ((S + A) * D) % 4) return -1;
    if ((S + A) * D > TILE_FLOATS) return -1;
    const int KC2 = TILE_FLOATS / (6 * D), KC3 = TILE_FLOATS / (NH * H), KC4 = TILE_FLOATS / (NH * 2 * S);
    if (KC2 < 1 || KC3 < 1 || KC4 < 1) return -1;
    const int ntiles = 1 + (D + KC2 - 1) / KC2 + (D + KC3 - 1) / KC3 + (H + KC4 - 1) / KC4;
    if (ntiles > MAX_TILES - 1) return -1;
    bool ok = al16(a->w_sa) && al16(a->w_ih) && al16(a->w_hh);
    for (int h = 0; h < NH; ++h) ok = ok && al16(a->w1[h]) && al16(a->w2[h]);
    if (!ok) return -1;
    const size_t bytes = sizeof(float) * ((size_t)RING_FWD * TILE_FLOATS + (size_t)BT * ((S + A) + 3 * D + NH * H + NH * 2 * S + S));
    if (bytes > 220 * 1024) return -1;
    MRSSM_CUDA(cudaFuncSetAttribute(rollout_fwd_staged_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    rollout_fwd_staged_kernel<BT><<<(a->B + BT - 1) / BT, NT, bytes, st>>>(*a, KC2, KC3, KC4);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

// w1c: contiguous [H][D] copies of the belief columns of every head's fc1 (made by the caller with mrssm_copy2d)
int rollout_bwd_staged(const mrssm_rollout_bwd_args* g, const float* const* w1c, cudaStream_t st) {
    const mrssm_rollout_args* a = &g->f;
    const int D = a->D, S = a->S, H = a->H, A = a->A, NH = 1 + a->n_experts;
    constexpr int BT = 8;
    if (!w1c) return -1;
    if (D > NT || NH * H > FC1_PER_THREAD * NT || S + A > NT || D % 4 || H % 4 || (D * (S + A)) % 4) return -1;
    if (D * (S + A) > TILE_FLOATS) return -1;
    BwdExtra X;
    X.KCb = TILE_FLOATS / (NH * H);
    X.KCc = TILE_FLOATS / (NH * D);
    X.KCe = TILE_FLOATS / (2 * D);
    if (X.KCb < 1 || X.KCc < 1 || X.KCe < 1) return -1;
    const int ntiles = (2 * S + X.KCb - 1) / X.KCb + (H + X.KCc - 1) / X.KCc + (3 * D + X.KCe - 1) / X.KCe + 1;
    if (ntiles > MAX_TILES - 1) return -1;
    bool ok = al16(a->w_sa) && al16(a->w_ih) && al16(a->w_hh);
    for (int h = 0; h < NH; ++h) {
        ok = ok && w1c[h] && al16(w1c[h]) && al16(a->w2[h]);
        X.w1c[h] = w1c[h];
    }
    if (!ok) return -1;
    const size_t bytes = sizeof(float) * ((size_t)RING_BWD * TILE_FLOATS + (size_t)BT * (D + S + NH * 2 * S + NH * H + D + 3 * D + D + D));
    if (bytes > 220 * 1024) return -1;
    MRSSM_CUDA(cudaFuncSetAttribute(rollout_bwd_staged_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    rollout_bwd_staged_kernel<BT><<<(a->B + BT - 1) / BT, NT, bytes, st>>>(*g, X);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
