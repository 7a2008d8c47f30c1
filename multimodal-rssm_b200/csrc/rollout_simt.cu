// Fused RSSM rollout, forward and BPTT backward, CUDA-core fp32 version.
//
// One launch runs all T steps.  The batch is tiled across CTAs (BT sequences per CTA, no
// inter-CTA communication); within a CTA every stage of the per-step chain
//   [s*mask, a] -> fc_embed+act -> GRUCell -> (1+E) Gaussian heads -> PoE/MoPoE fusion -> rsample
// keeps its activations in shared memory in feature-major [feature][BT] form, so a thread that
// owns an output feature streams one weight column (coalesced, L2-resident after the first step)
// against BT batch rows held in registers.  Only the API-visible outputs and the BPTT stash go
// to HBM.  Reference: utils/models/transition_model.py:200-285 (+ :50-114), encoder.py:50-155.
#include "common.cuh"

namespace {

constexpr int NT = 256;

struct FwdSmem {
    float *xin, *x, *h0, *h1, *u, *o, *sprev;
};

template <int BT>
__device__ __forceinline__ void load_bt(float (&v)[BT], const float* p) {
#pragma unroll
    for (int b = 0; b < BT; b += 4) {
        float4 q = *reinterpret_cast<const float4*>(p + b);
        v[b] = q.x; v[b + 1] = q.y; v[b + 2] = q.z; v[b + 3] = q.w;
    }
}

template <int BT>
__global__ void __launch_bounds__(NT) rollout_fwd_kernel(mrssm_rollout_args a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int D = a.D, S = a.S, H = a.H, A = a.A, E = a.n_experts, NH = 1 + E, B = a.B;
    const int b0 = blockIdx.x * BT;
    float* xin = smem;                       // [S+A][BT]
    float* x = xin + (S + A) * BT;           // [D][BT]
    float* hA = x + D * BT;                  // [D][BT]
    float* hB = hA + D * BT;                 // [D][BT]
    float* u = hB + D * BT;                  // [NH*H][BT]
    float* o = u + NH * H * BT;              // [NH*2S][BT]
    float* sprev = o + NH * 2 * S * BT;      // [S][BT]
    float* hprev = hA;
    float* hcur = hB;

    for (int i = tid; i < D * BT; i += NT) {
        int k = i / BT, b = i % BT;
        hprev[i] = (b0 + b < B) ? a.prev_belief[(long long)(b0 + b) * D + k] : 0.f;
    }
    for (int i = tid; i < S * BT; i += NT) {
        int k = i / BT, b = i % BT;
        sprev[i] = (b0 + b < B) ? a.prev_state[(long long)(b0 + b) * S + k] : 0.f;
    }
    __syncthreads();

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
        // ---- 1. x = act(W_sa xin + b) ---------------------------------------------------------------
        for (int j = tid; j < D; j += NT) {
            float acc[BT];
            float bj = a.b_sa[j];
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] = bj;
            for (int k = 0; k < S + A; ++k) {
                float w = __ldg(a.w_sa + (long long)k * D + j);
                float v[BT];
                load_bt<BT>(v, xin + k * BT);
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = fmaf(w, v[b], acc[b]);
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                float xv = act_apply(acc[b], a.act);
                x[j * BT + b] = xv;
                if (a.st_x && b0 + b < B) a.st_x[(tb + b) * D + j] = xv;
            }
        }
        __syncthreads();
        // ---- 2. GRUCell ------------------------------------------------------------------------------
        for (int j = tid; j < D; j += NT) {
            float ar[BT], az[BT], ain[BT], ahn[BT];
            float br = a.b_ih[j] + a.b_hh[j], bz = a.b_ih[D + j] + a.b_hh[D + j];
            float bin = a.b_ih[2 * D + j], bhn = a.b_hh[2 * D + j];
#pragma unroll
            for (int b = 0; b < BT; ++b) { ar[b] = br; az[b] = bz; ain[b] = bin; ahn[b] = bhn; }
            const float* wi = a.w_ih + j;
            const float* wh = a.w_hh + j;
#pragma unroll 2
            for (int k = 0; k < D; ++k) {
                float wir = __ldg(wi), wiz = __ldg(wi + D), win = __ldg(wi + 2 * D);
                float whr = __ldg(wh), whz = __ldg(wh + D), whn = __ldg(wh + 2 * D);
                wi += 3 * D;
                wh += 3 * D;
                float xv[BT], hv[BT];
                load_bt<BT>(xv, x + k * BT);
                load_bt<BT>(hv, hprev + k * BT);
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    ar[b] = fmaf(wir, xv[b], fmaf(whr, hv[b], ar[b]));
                    az[b] = fmaf(wiz, xv[b], fmaf(whz, hv[b], az[b]));
                    ain[b] = fmaf(win, xv[b], ain[b]);
                    ahn[b] = fmaf(whn, hv[b], ahn[b]);
                }
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                float r = sigmoidf_(ar[b]), z = sigmoidf_(az[b]);
                float n = tanhf(ain[b] + r * ahn[b]);
                float hp = hprev[j * BT + b];
                float hn = (1.f - z) * n + z * hp;
                hcur[j * BT + b] = hn;
                if (b0 + b < B) {
                    long long off = (tb + b) * D + j;
                    a.beliefs[off] = hn;
                    if (a.st_r) { a.st_r[off] = r; a.st_z[off] = z; a.st_n[off] = n; a.st_ghn[off] = ahn[b]; }
                }
            }
        }
        __syncthreads();
        // ---- 3. heads fc1 (+ hoisted embedding half) + act -----------------------------------------------
        for (int idx = tid; idx < NH * H; idx += NT) {
            int hd = idx / H, c = idx % H;
            float acc[BT];
            float bj = a.b1[hd] ? a.b1[hd][c] : 0.f;
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] = bj;
            if (a.emb_pre[hd]) {
#pragma unroll
                for (int b = 0; b < BT; ++b)
                    if (b0 + b < B) acc[b] += a.emb_pre[hd][(tb + b) * H + c];
            }
            const float* w = a.w1[hd] + c;
#pragma unroll 4
            for (int k = 0; k < D; ++k) {
                float wv = __ldg(w + (long long)k * H);
                float hv[BT];
                load_bt<BT>(hv, hcur + k * BT);
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = fmaf(wv, hv[b], acc[b]);
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                float uv = act_apply(acc[b], a.act);
                u[idx * BT + b] = uv;
                if (a.st_u[hd] && b0 + b < B) a.st_u[hd][(tb + b) * H + c] = uv;
            }
        }
        __syncthreads();
        // ---- 4. heads fc2 -> (mean, softplus+min_std) ----------------------------------------------------
        for (int idx = tid; idx < NH * 2 * S; idx += NT) {
            int hd = idx / (2 * S), c = idx % (2 * S);
            float acc[BT];
            float bj = a.b2[hd][c];
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] = bj;
            const float* w = a.w2[hd] + c;
            const float* uh = u + hd * H * BT;
#pragma unroll 4
            for (int k = 0; k < H; ++k) {
                float wv = __ldg(w + (long long)k * 2 * S);
                float uv[BT];
                load_bt<BT>(uv, uh + k * BT);
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = fmaf(wv, uv[b], acc[b]);
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) o[idx * BT + b] = (c < S) ? acc[b] : softplusf_(acc[b]) + a.min_std;
        }
        __syncthreads();
        // ---- 5. prior sample, expert outputs, fusion, posterior sample -----------------------------------
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
                if (a.n_subsets == 0) {          // single-modal: posterior = the one expert
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

// ------------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(NT) rollout_bwd_kernel(mrssm_rollout_bwd_args g) {
    extern __shared__ __align__(16) float smem[];
    const mrssm_rollout_args& a = g.f;
    const int tid = threadIdx.x;
    const int D = a.D, S = a.S, H = a.H, A = a.A, E = a.n_experts, NH = 1 + E, B = a.B;
    const int b0 = blockIdx.x * BT;
    float* cgh = smem;                        // [D][BT]   carried grad wrt h_t
    float* cgs = cgh + D * BT;                // [S][BT]   carried grad wrt the fed-back state
    float* go = cgs + S * BT;                 // [NH*2S][BT]
    float* gu = go + NH * 2 * S * BT;         // [NH*H][BT]
    float* Gh = gu + NH * H * BT;             // [D][BT]
    float* dgi = Gh + D * BT;                 // [3D][BT]
    float* dghn = dgi + 3 * D * BT;           // [D][BT]
    float* dxp = dghn + D * BT;               // [D][BT]

    for (int i = tid; i < D * BT; i += NT) cgh[i] = 0.f;
    for (int i = tid; i < S * BT; i += NT) cgs[i] = 0.f;
    __syncthreads();

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
        // ---- b. gu = (W2^T go) * act'(u) ---------------------------------------------------------------
        for (int idx = tid; idx < NH * H; idx += NT) {
            int hd = idx / H, c = idx % H;
            float acc[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] = 0.f;
            const float* w = a.w2[hd] + c;              // PyTorch layout [2S][H]
            const float* gh = go + hd * 2 * S * BT;
            for (int j = 0; j < 2 * S; ++j) {
                float wv = __ldg(w + (long long)j * H);
                float v[BT];
                load_bt<BT>(v, gh + j * BT);
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = fmaf(wv, v[b], acc[b]);
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                float val = 0.f;
                if (b0 + b < B) {
                    long long off = (tb + b) * H + c;
                    val = acc[b] * act_grad_from_out(a.st_u[hd][off], a.act);
                    g.d_u[hd][off] = val;
                }
                gu[idx * BT + b] = val;
            }
        }
        __syncthreads();
        // ---- c. G_h = g_beliefs + carry + sum_heads W1_h^T gu ------------------------------------------
        for (int k = tid; k < D; k += NT) {
            float acc[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                float v = cgh[k * BT + b];
                if (g.g_beliefs && b0 + b < B) v += g.g_beliefs[(tb + b) * D + k];
                acc[b] = v;
            }
            for (int hd = 0; hd < NH; ++hd) {
                const float* w = a.w1[hd] + k;          // PyTorch layout [H][ld1], belief columns first
                const long long ld = a.ld1[hd];
                const float* gg = gu + hd * H * BT;
#pragma unroll 4
                for (int c = 0; c < H; ++c) {
                    float wv = __ldg(w + c * ld);
                    float v[BT];
                    load_bt<BT>(v, gg + c * BT);
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] = fmaf(wv, v[b], acc[b]);
                }
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) Gh[k * BT + b] = acc[b];
        }
        __syncthreads();
        // ---- d. GRU gate backward (elementwise) ----------------------------------------------------------
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
        // ---- e. gx = W_ih^T dgi ; carry_gh += W_hh^T dgh ; dxpre = gx * act'(x) ---------------------------
        for (int k = tid; k < D; k += NT) {
            float ax[BT], ah[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) { ax[b] = 0.f; ah[b] = 0.f; }
            const float* wi = a.w_ih + k;               // PyTorch layout [3D][D]
            const float* wh = a.w_hh + k;
#pragma unroll 2
            for (int j = 0; j < 3 * D; ++j) {
                float wiv = __ldg(wi + (long long)j * D), whv = __ldg(wh + (long long)j * D);
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
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                float val = 0.f;
                if (b0 + b < B) {
                    long long off = (tb + b) * D + k;
                    val = ax[b] * act_grad_from_out(a.st_x[off], a.act);
                    g.d_xpre[off] = val;
                }
                dxp[k * BT + b] = val;
                cgh[k * BT + b] += ah[b];
            }
        }
        __syncthreads();
        // ---- f. g_xin = W_sa^T dxpre -> state carry (masked) and action grads; emit xin ------------------
        for (int i = tid; i < S + A; i += NT) {
            float acc[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] = 0.f;
            const float* w = a.w_sa + i;                // PyTorch layout [D][S+A]
            for (int k = 0; k < D; ++k) {
                float wv = __ldg(w + (long long)k * (S + A));
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
        __syncthreads();
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

int check_common(const mrssm_rollout_args* a) {
    MRSSM_CHECK(a != nullptr, "rollout: null args");
    MRSSM_CHECK(a->T > 0 && a->B > 0 && a->D > 0 && a->S > 0 && a->H > 0 && a->A > 0, "rollout: bad sizes");
    MRSSM_CHECK(a->n_experts >= 0 && a->n_experts < MRSSM_MAX_HEADS, "rollout: n_experts=%d unsupported", a->n_experts);
    MRSSM_CHECK(a->S <= MRSSM_MAX_STATE, "rollout: state_size %d > %d", a->S, MRSSM_MAX_STATE);
    MRSSM_CHECK(a->n_subsets >= 0 && a->n_subsets <= MRSSM_MAX_SUBSETS, "rollout: n_subsets=%d", a->n_subsets);
    MRSSM_CHECK(a->prev_state && a->prev_belief && a->actions, "rollout: null input");
    MRSSM_CHECK(a->det || a->eps_prior, "rollout: eps_prior required unless det");
    MRSSM_CHECK(a->det || a->n_experts == 0 || a->eps_post, "rollout: eps_post required unless det");
    MRSSM_CHECK(a->beliefs && a->prior_states && a->prior_means && a->prior_stds, "rollout: null output");
    for (int h = 0; h <= a->n_experts; ++h) {
        MRSSM_CHECK(a->w1[h] && a->w2[h] && a->b2[h], "rollout: head %d weights missing", h);
        MRSSM_CHECK(a->b1[h] || a->emb_pre[h], "rollout: head %d has neither bias nor emb_pre", h);
        if (h > 0) MRSSM_CHECK(a->exp_means[h] && a->exp_stds[h], "rollout: expert %d outputs missing", h);
    }
    if (a->n_experts > 0) MRSSM_CHECK(a->post_states && a->post_means && a->post_stds, "rollout: null posterior output");
    return 0;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) MRSSM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

}  // namespace

int rollout_check(const mrssm_rollout_args* a) { return check_common(a); }

int rollout_bwd_check(const mrssm_rollout_bwd_args* g) {
    const mrssm_rollout_args* a = &g->f;
    if (int e = check_common(a)) return e;
    MRSSM_CHECK(a->st_x && a->st_r && a->st_z && a->st_n && a->st_ghn, "rollout_bwd: stash missing");
    MRSSM_CHECK(g->d_xpre && g->d_gi && g->d_gh, "rollout_bwd: null grad output");
    for (int h = 0; h <= a->n_experts; ++h)
        MRSSM_CHECK(a->st_u[h] && g->d_u[h] && g->d_o[h] && a->ld1[h] >= a->D, "rollout_bwd: head %d buffers missing", h);
    return 0;
}

int rollout_fwd_simt(const mrssm_rollout_args* a, cudaStream_t st) {
    if (int e = check_common(a)) return e;
    const int NH = 1 + a->n_experts;
    auto bytes = [&](int bt) {
        return sizeof(float) * (size_t)bt * ((a->S + a->A) + 3 * a->D + NH * a->H + NH * 2 * a->S + a->S);
    };
    if (bytes(8) <= 200 * 1024) {
        if (int e = set_smem(rollout_fwd_kernel<8>, bytes(8))) return e;
        rollout_fwd_kernel<8><<<(a->B + 7) / 8, NT, bytes(8), st>>>(*a);
    } else {
        MRSSM_CHECK(bytes(4) <= 227 * 1024, "rollout: sizes D=%d H=%d exceed shared memory", a->D, a->H);
        if (int e = set_smem(rollout_fwd_kernel<4>, bytes(4))) return e;
        rollout_fwd_kernel<4><<<(a->B + 3) / 4, NT, bytes(4), st>>>(*a);
    }
    MRSSM_LAUNCH_CHECK();
    return 0;
}

int rollout_bwd_simt(const mrssm_rollout_bwd_args* g, cudaStream_t st) {
    const mrssm_rollout_args* a = &g->f;
    if (int e = check_common(a)) return e;
    MRSSM_CHECK(a->st_x && a->st_r && a->st_z && a->st_n && a->st_ghn, "rollout_bwd: stash missing");
    MRSSM_CHECK(g->d_xpre && g->d_gi && g->d_gh, "rollout_bwd: null grad output");
    for (int h = 0; h <= a->n_experts; ++h)
        MRSSM_CHECK(a->st_u[h] && g->d_u[h] && g->d_o[h] && a->ld1[h] >= a->D, "rollout_bwd: head %d buffers missing", h);
    const int NH = 1 + a->n_experts;
    auto bytes = [&](int bt) {
        return sizeof(float) * (size_t)bt * (a->D + a->S + NH * 2 * a->S + NH * a->H + a->D + 3 * a->D + a->D + a->D);
    };
    if (bytes(8) <= 200 * 1024) {
        if (int e = set_smem(rollout_bwd_kernel<8>, bytes(8))) return e;
        rollout_bwd_kernel<8><<<(a->B + 7) / 8, NT, bytes(8), st>>>(*g);
    } else {
        MRSSM_CHECK(bytes(4) <= 227 * 1024, "rollout_bwd: sizes D=%d H=%d exceed shared memory", a->D, a->H);
        if (int e = set_smem(rollout_bwd_kernel<4>, bytes(4))) return e;
        rollout_bwd_kernel<4><<<(a->B + 3) / 4, NT, bytes(4), st>>>(*g);
    }
    MRSSM_LAUNCH_CHECK();
    return 0;
}
