// Per-step elementwise kernels of the LARGE-MODEL rollout (BASELINE config 5: deter = hidden = 1024).  When the recurrent
// weights are far too large for one CTA (the tcgen05 rollout of rollout_tc.cu holds D, H <= 208), every dense contraction of
// a time step — fc_embed_state_action, the two GRU projections, every head's fc1 / fc2 (transition_model.py:226-270,
// encoder.py:126-190) and their dgrads — is a batch-stationary tcgen05 GEMM of dense_tc.cu over all B sequences with the
// bf16 weights resident in L2, and the kernels here do what sits between the GEMMs: building the masked [state, action]
// operand, the GRU gate arithmetic, softplus / PoE-MoPoE fusion / rsample, and the matching backward pieces.  They read and
// write the same [T,B,*] API tensors and BPTT stash as the fused kernels (mrssm_rollout_args), one time step per launch.
#include <algorithm>
#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;
constexpr int NT = 256;

// xin_b[b][0..KX) = [s_{t-1} * mask, a_t, 0..]  (s_{t-1}: prev_state at t = 0, else the posterior (prior if no experts) sample)
__global__ void rstep_xin_kernel(mrssm_rollout_args a, int t, int KX, bf16* __restrict__ xin_b) {
    const int S = a.S, A = a.A, B = a.B;
    const long long total = (long long)B * KX;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / KX), c = (int)(i - (long long)b * KX);
        float v = 0.f;
        if (c < S) {
            const float* src = a.n_experts > 0 ? a.post_states : a.prior_states;
            const float sp = t > 0 ? src[((long long)(t - 1) * B + b) * S + c] : a.prev_state[(long long)b * S + c];
            const float m = a.nonterminals ? a.nonterminals[(long long)t * B + b] : 1.f;
            v = sp * m;
        } else if (c < S + A) {
            v = a.actions[((long long)t * B + b) * A + (c - S)];
        }
        xin_b[i] = __float2bfloat16(v);
    }
}

// GRU cell (nn.GRUCell, gate order r, z, n; b_hn inside the r * (.) term): gi = W_ih x + b_ih, gh = W_hh h + b_hh (fp32 [B,3D])
__global__ void rstep_gate_fwd_kernel(mrssm_rollout_args a, int t, const float* __restrict__ gi, const float* __restrict__ gh,
                                      bf16* __restrict__ hb_out) {
    const int D = a.D, B = a.B;
    const long long total = (long long)B * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / D), j = (int)(i - (long long)b * D);
        const long long o3 = (long long)b * 3 * D + j, off = ((long long)t * B + b) * D + j;
        const float hp = t > 0 ? a.beliefs[off - (long long)B * D] : a.prev_belief[(long long)b * D + j];
        const float r = sigmoidf_(gi[o3] + gh[o3]);
        const float z = sigmoidf_(gi[o3 + D] + gh[o3 + D]);
        const float ghn = gh[o3 + 2 * D];
        const float n = tanhf(gi[o3 + 2 * D] + r * ghn);
        const float h = (1.f - z) * n + z * hp;
        a.beliefs[off] = h;
        if (a.st_r) {
            a.st_r[off] = r; a.st_z[off] = z; a.st_n[off] = n; a.st_ghn[off] = ghn;
        }
        hb_out[i] = __float2bfloat16(h);
    }
}

// o[b][hd*hs + c], c < 2S: fc2 outputs of every head (bias included), row stride ldo.  -> prior / expert / posterior statistics, samples (step 5 of
// the fused kernels: softplus + min_std, 1/sigma-weighted PoE over the subset of each state dimension, rsample)
__global__ void rstep_heads_fwd_kernel(mrssm_rollout_args a, int t, const float* __restrict__ o, int ldo, int hs, bf16* __restrict__ xin_next,
                                       int KX) {
    const int S = a.S, B = a.B, E = a.n_experts;
    const long long total = (long long)B * S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / S), s = (int)(i - (long long)b * S);
        const float* ob = o + (long long)b * ldo;
        const long long off = ((long long)t * B + b) * S + s;
        const float pm = ob[s], ps = softplusf_(ob[S + s]) + a.min_std;
        const float pst = a.det ? pm : fmaf(ps, a.eps_prior[off], pm);
        a.prior_means[off] = pm;
        a.prior_stds[off] = ps;
        a.prior_states[off] = pst;
        if (E > 0) {
            float qm, qs;
            float sumT = 0.f, sumMT = 0.f;
            const unsigned mask = a.n_subsets ? a.subset_mask[a.dim_subset[s]] : 1u;
            for (int e = 1; e <= E; ++e) {
                const float em = ob[e * hs + s], es = softplusf_(ob[e * hs + S + s]) + a.min_std;
                a.exp_means[e][off] = em;
                a.exp_stds[e][off] = es;
                if (mask & (1u << (e - 1))) {
                    const float tt = 1.f / es;
                    sumT += tt;
                    sumMT = fmaf(em, tt, sumMT);
                }
            }
            if (a.n_subsets == 0) {          // single-modal: the posterior is the one expert
                qm = a.exp_means[1][off];
                qs = a.exp_stds[1][off];
            } else {
                qm = sumMT / sumT;
                qs = 1.f / sumT;
            }
            a.post_means[off] = qm;
            a.post_stds[off] = qs;
            const float qst = a.det ? qm : fmaf(qs, a.eps_post[off], qm);
            a.post_states[off] = qst;
            if (xin_next) {
                const float m = a.nonterminals ? a.nonterminals[(long long)(t + 1) * B + b] : 1.f;
                xin_next[(long long)b * KX + s] = __float2bfloat16(qst * m);
            }
        } else if (xin_next) {
            const float m = a.nonterminals ? a.nonterminals[(long long)(t + 1) * B + b] : 1.f;
            xin_next[(long long)b * KX + s] = __float2bfloat16(pst * m);
        }
        if (xin_next && s == 0) {            // the row's action columns and zero padding (the [s * mask, a, 0..] operand of step t + 1)
            for (int c = S; c < KX; ++c)
                xin_next[(long long)b * KX + c] =
                    __float2bfloat16(c < S + a.A ? a.actions[((long long)(t + 1) * B + b) * a.A + (c - S)] : 0.f);
        }
    }
}

// backward of rstep_heads_fwd: upstream grads of the statistics / samples at step t plus the state-gradient carry from step
// t + 1 (cgs, [B,S]) -> grads of every head's fc2 output, written as the bf16 GEMM operand d_o[hd][b * ld + 0..S2p)
struct HeadsBwdOut {
    bf16* d_o[MRSSM_MAX_HEADS];
};
// cgs: the state-gradient carry [B,S], or — dxin_next given — taken straight from step t + 1's dxin = dxpre W_sa (fp32 [B, S+A]): its state
// columns times the mask are the carry, its action columns are g_actions[t + 1] (the xin_bwd of step t + 1 folded in)
__global__ void rstep_heads_bwd_kernel(mrssm_rollout_bwd_args g, int t, const float* __restrict__ cgs, const float* __restrict__ dxin_next,
                                       HeadsBwdOut out, int ld, int S2p) {
    const mrssm_rollout_args& a = g.f;
    const int S = a.S, B = a.B, E = a.n_experts;
    const long long total = (long long)B * S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / S), s = (int)(i - (long long)b * S);
        const long long off = ((long long)t * B + b) * S + s, ro = (long long)b * ld;
        float carry;
        if (dxin_next) {
            const float m = a.nonterminals ? a.nonterminals[(long long)(t + 1) * B + b] : 1.f;
            carry = dxin_next[(long long)b * (S + a.A) + s] * m;
            if (s == 0 && g.g_actions)
                for (int c = 0; c < a.A; ++c) g.g_actions[((long long)(t + 1) * B + b) * a.A + c] = dxin_next[(long long)b * (S + a.A) + S + c];
        } else {
            carry = cgs[i];
        }
        float gps = g.g_prior_states ? g.g_prior_states[off] : 0.f;
        if (E == 0) gps += carry;
        const float gpm = gps + (g.g_prior_means ? g.g_prior_means[off] : 0.f);
        float gpsd = g.g_prior_stds ? g.g_prior_stds[off] : 0.f;
        if (!a.det) gpsd = fmaf(gps, a.eps_prior[off], gpsd);
        const float psd = a.prior_stds[off];
        out.d_o[0][ro + s] = __float2bfloat16(gpm);
        out.d_o[0][ro + S + s] = __float2bfloat16(gpsd * (1.f - expf(-(psd - a.min_std))));
        if (E > 0) {
            const float gq = (g.g_post_states ? g.g_post_states[off] : 0.f) + carry;
            const float gqm = gq + (g.g_post_means ? g.g_post_means[off] : 0.f);
            float gqs = g.g_post_stds ? g.g_post_stds[off] : 0.f;
            if (!a.det) gqs = fmaf(gq, a.eps_post[off], gqs);
            const unsigned mask = a.n_subsets ? a.subset_mask[a.dim_subset[s]] : 1u;
            float P = 0.f, qm = 0.f;
            if (a.n_subsets) {
                P = 1.f / a.post_stds[off];
                qm = a.post_means[off];
            }
            for (int e = 1; e <= E; ++e) {
                float gm = g.g_exp_means[e] ? g.g_exp_means[e][off] : 0.f;
                float gs = g.g_exp_stds[e] ? g.g_exp_stds[e][off] : 0.f;
                const float sd = a.exp_stds[e][off];
                if (mask & (1u << (e - 1))) {
                    if (a.n_subsets == 0) {
                        gm += gqm;
                        gs += gqs;
                    } else {
                        const float te = 1.f / sd, mu = a.exp_means[e][off];
                        gm = fmaf(gqm, te / P, gm);
                        const float gT = gqm * (mu - qm) / P - gqs / (P * P);
                        gs = fmaf(-gT, te * te, gs);
                    }
                }
                out.d_o[e][ro + s] = __float2bfloat16(gm);
                out.d_o[e][ro + S + s] = __float2bfloat16(gs * (1.f - expf(-(sd - a.min_std))));
            }
        }
        // zero the padding columns once per row (c >= 2S)
        if (s == 0)
            for (int hd = 0; hd <= E; ++hd)
                for (int c = 2 * S; c < S2p; ++c) out.d_o[hd][ro + c] = __float2bfloat16(0.f);
    }
}

// GRU gate backward at step t: G = g_beliefs[t] + carry_a (G*z of step t+1) + carry_b (dgh W_hh of step t+1) + dh_heads ->
// d_gi, d_gh (bf16 [B,3D], operands of the next dgrad GEMMs and of the deferred weight gradients) and the new carry_a
__global__ void rstep_gate_bwd_kernel(mrssm_rollout_bwd_args g, int t, const float* __restrict__ dh_heads, float* __restrict__ carry_a,
                                      const float* __restrict__ carry_b, bf16* __restrict__ dgi, bf16* __restrict__ dgh) {
    const mrssm_rollout_args& a = g.f;
    const int D = a.D, B = a.B;
    const long long total = (long long)B * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / D), j = (int)(i - (long long)b * D);
        const long long off = ((long long)t * B + b) * D + j, o3 = (long long)b * 3 * D + j;
        const float r = a.st_r[off], z = a.st_z[off], n = a.st_n[off], ghn = a.st_ghn[off];
        const float hp = t > 0 ? a.beliefs[off - (long long)B * D] : a.prev_belief[(long long)b * D + j];
        const float G = (g.g_beliefs ? g.g_beliefs[off] : 0.f) + carry_a[i] + carry_b[i] + dh_heads[i];
        const float gn = G * (1.f - z), gz = G * (hp - n);
        const float gpn = gn * (1.f - n * n);
        const float gpr = gpn * ghn * r * (1.f - r);
        const float gpz = gz * z * (1.f - z);
        const float gpnr = gpn * r;
        dgi[o3] = __float2bfloat16(gpr); dgi[o3 + D] = __float2bfloat16(gpz); dgi[o3 + 2 * D] = __float2bfloat16(gpn);
        dgh[o3] = __float2bfloat16(gpr); dgh[o3 + D] = __float2bfloat16(gpz); dgh[o3 + 2 * D] = __float2bfloat16(gpnr);
        carry_a[i] = G * z;
    }
}

// dxin = dxpre W_sa (fp32 [B, ld]): the state columns, masked, are the carry into step t - 1's sample; the rest the action grads
__global__ void rstep_xin_bwd_kernel(mrssm_rollout_bwd_args g, int t, const float* __restrict__ dxin, int ld, float* __restrict__ cgs) {
    const mrssm_rollout_args& a = g.f;
    const int S = a.S, A = a.A, B = a.B;
    const long long total = (long long)B * (S + A);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / (S + A)), c = (int)(i - (long long)b * (S + A));
        const float v = dxin[(long long)b * ld + c];
        if (c < S) {
            const float m = a.nonterminals ? a.nonterminals[(long long)t * B + b] : 1.f;
            cgs[(long long)b * S + c] = v * m;
        } else if (g.g_actions) {
            g.g_actions[((long long)t * B + b) * A + (c - S)] = v;
        }
    }
}

__global__ void rstep_add2_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = x[i] + y[i];
}

inline int blocks_for(long long total) { return (int)std::max<long long>(1, std::min<long long>(148 * 8, (total + NT - 1) / NT)); }

}  // namespace

extern "C" int mrssm_rstep_xin(const mrssm_rollout_args* a, int32_t t, int32_t KX, void* xin_b, void* stream) {
    MRSSM_CHECK(a && xin_b && t >= 0 && t < a->T && KX >= a->S + a->A, "rstep_xin: bad arguments");
    rstep_xin_kernel<<<blocks_for((long long)a->B * KX), NT, 0, (cudaStream_t)stream>>>(*a, t, KX, (bf16*)xin_b);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_rstep_gate_fwd(const mrssm_rollout_args* a, int32_t t, const float* gi, const float* gh, void* hb_out, void* stream) {
    MRSSM_CHECK(a && gi && gh && hb_out && a->beliefs && t >= 0 && t < a->T, "rstep_gate_fwd: bad arguments");
    rstep_gate_fwd_kernel<<<blocks_for((long long)a->B * a->D), NT, 0, (cudaStream_t)stream>>>(*a, t, gi, gh, (bf16*)hb_out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_rstep_heads_fwd(const mrssm_rollout_args* a, int32_t t, const float* o, int32_t ldo, int32_t head_stride, void* xin_next,
                                     int32_t KX, void* stream) {
    MRSSM_CHECK(a && o && t >= 0 && t < a->T && a->S <= MRSSM_MAX_STATE && head_stride >= 2 * a->S && ldo >= (1 + a->n_experts) * head_stride,
                "rstep_heads_fwd: bad arguments");
    MRSSM_CHECK(a->prior_states && a->prior_means && a->prior_stds && (a->det || a->eps_prior), "rstep_heads_fwd: null prior output / noise");
    MRSSM_CHECK(a->n_experts >= 0 && a->n_experts < MRSSM_MAX_HEADS && a->n_subsets >= 0 && a->n_subsets <= MRSSM_MAX_SUBSETS,
                "rstep_heads_fwd: n_experts=%d n_subsets=%d unsupported", a->n_experts, a->n_subsets);
    for (int h = 1; h <= a->n_experts; ++h) MRSSM_CHECK(a->exp_means[h] && a->exp_stds[h], "rstep_heads_fwd: expert %d outputs missing", h);
    if (a->n_experts > 0)
        MRSSM_CHECK(a->post_states && a->post_means && a->post_stds && (a->det || a->eps_post), "rstep_heads_fwd: null posterior output / noise");
    MRSSM_CHECK(!xin_next || (t + 1 < a->T && KX >= a->S + a->A), "rstep_heads_fwd: xin_next needs a step t + 1 and KX >= S + A");
    rstep_heads_fwd_kernel<<<blocks_for((long long)a->B * a->S), NT, 0, (cudaStream_t)stream>>>(*a, t, o, ldo, head_stride, (bf16*)xin_next, KX);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_rstep_heads_bwd(const mrssm_rollout_bwd_args* g, int32_t t, const float* cgs, const float* dxin_next, void* const* d_o,
                                     int32_t ld, int32_t S2p, void* stream) {
    MRSSM_CHECK(g && (cgs || dxin_next) && d_o && t >= 0 && t < g->f.T && S2p >= 2 * g->f.S && ld >= S2p, "rstep_heads_bwd: bad arguments");
    MRSSM_CHECK(!dxin_next || t + 1 < g->f.T, "rstep_heads_bwd: dxin_next needs a step t + 1");
    HeadsBwdOut out;
    for (int h = 0; h < MRSSM_MAX_HEADS; ++h) out.d_o[h] = h <= g->f.n_experts ? (bf16*)d_o[h] : nullptr;
    for (int h = 0; h <= g->f.n_experts; ++h) MRSSM_CHECK(out.d_o[h], "rstep_heads_bwd: head %d output missing", h);
    rstep_heads_bwd_kernel<<<blocks_for((long long)g->f.B * g->f.S), NT, 0, (cudaStream_t)stream>>>(*g, t, cgs, dxin_next, out, ld, S2p);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_rstep_gate_bwd(const mrssm_rollout_bwd_args* g, int32_t t, const float* dh_heads, float* carry_a, const float* carry_b,
                                    void* dgi, void* dgh, void* stream) {
    MRSSM_CHECK(g && dh_heads && carry_a && carry_b && dgi && dgh && g->f.st_r && t >= 0 && t < g->f.T, "rstep_gate_bwd: bad arguments");
    rstep_gate_bwd_kernel<<<blocks_for((long long)g->f.B * g->f.D), NT, 0, (cudaStream_t)stream>>>(*g, t, dh_heads, carry_a, carry_b, (bf16*)dgi,
                                                                                                   (bf16*)dgh);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_rstep_xin_bwd(const mrssm_rollout_bwd_args* g, int32_t t, const float* dxin, int32_t ld, float* cgs, void* stream) {
    MRSSM_CHECK(g && dxin && cgs && t >= 0 && t < g->f.T && ld >= g->f.S + g->f.A, "rstep_xin_bwd: bad arguments");
    rstep_xin_bwd_kernel<<<blocks_for((long long)g->f.B * (g->f.S + g->f.A)), NT, 0, (cudaStream_t)stream>>>(*g, t, dxin, ld, cgs);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
extern "C" int mrssm_add2(const float* x, const float* y, int64_t n, float* out, void* stream) {
    MRSSM_CHECK(x && y && out && n > 0, "add2: bad arguments");
    rstep_add2_kernel<<<blocks_for(n), NT, 0, (cudaStream_t)stream>>>(x, y, n, out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

// ---- all T steps from C ------------------------------------------------------------------------------------------------------
int dense_tc_launch(const void* A, long long lda, int M, int K, const void* wpacked, int Npad, int Kpad, int n_valid, const float* bias,
                    int bias_mod, int act, const void* mask, long long ldm, int mask_mode, void* out, long long ldc, long long cstride,
                    int out_f32, cudaStream_t st, const float* addend, long long addend_ld, int group_n, int group_k);

namespace {
// C[M][n_valid] = act(A[M][K] W^T + bias (+ addend)) (* act'(mask)); block-diagonal when gn > 0 (see dense_tc.cu)
inline int gemm(cudaStream_t st, int M, const void* A, long long lda, int K, const void* wp, int Npad, int n_valid, const float* bias, int act,
                const void* mask, long long ldm, int mask_mode, void* out, long long ldc, int out_f32, const float* addend = nullptr,
                long long addend_ld = 0, int gn = 0, int gk = 0) {
    const int Kpad = (int)(ceil_div64(gn > 0 ? gk : K, 64) * 64);
    return dense_tc_launch(A, lda, M, K, wp, Npad, Kpad, n_valid, bias, 0, act, mask, ldm, mask_mode, out, ldc, 1, out_f32, st, addend, addend_ld,
                           gn, gk);
}
inline bf16* b16(void* p, long long off) { return (bf16*)p + off; }

// A second stream for the one GEMM of a step that does not sit on the step's dependency chain: W_hh h_{t-1} (forward; needed only by
// the gate kernel) and d_gh W_hh (backward; needed only by the previous step's gate kernel).  One process drives one GPU.
struct Aux {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok = false;
};
Aux& aux_stream() {
    static Aux a = [] {
        Aux x;
        x.ok = cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) == cudaSuccess;
        return x;
    }();
    return a;
}
bool g_step_aux = true;
}  // namespace

extern "C" int mrssm_rstep_set_aux_stream(int32_t on) {
    g_step_aux = on != 0;
    return 0;
}

#define RSTEP_TRY(expr)             \
    do {                            \
        if (int rc_ = (expr)) return rc_; \
    } while (0)

extern "C" int mrssm_rollout_steps_fwd(const mrssm_rollout_args* a, const mrssm_rstep_ws* w, void* stream) {
    MRSSM_CHECK(a && w && w->xin_all && w->x_all && w->u_cat && w->hb_all && w->gi && w->gh && w->o_cat && w->wp_sa && w->wp_ih && w->wp_hh && w->w2f &&
                    w->b2, "rollout_steps_fwd: null workspace / weights");
    MRSSM_CHECK(a->D % 16 == 0 && a->H % 16 == 0 && w->NH == 1 + a->n_experts && w->n_chunks >= 1 && w->n_chunks <= MRSSM_MAX_HEADS &&
                    w->KX >= a->S + a->A && w->S2p >= 2 * a->S, "rollout_steps_fwd: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const int T = a->T, B = a->B, D = a->D, H = a->H, NH = w->NH, KX = w->KX, S2p = w->S2p;
    const long long BD = (long long)B * D, BU = (long long)B * NH * H;
    Aux& ax = aux_stream();
    const bool use_aux = g_step_aux && ax.ok;
    auto gh_gemm = [&](cudaStream_t s_, int t_) {
        return gemm(s_, B, b16(w->hb_all, t_ * BD), D, D, w->wp_hh, 3 * D, 3 * D, a->b_hh, 0, nullptr, 0, 0, w->gh, 3 * D, 1);
    };
    if (use_aux) {          // W_hh h_{-1} for step 0, off the chain
        MRSSM_CUDA(cudaEventRecord(ax.fork, st));
        MRSSM_CUDA(cudaStreamWaitEvent(ax.s, ax.fork, 0));
        RSTEP_TRY(gh_gemm(ax.s, 0));
        MRSSM_CUDA(cudaEventRecord(ax.join, ax.s));
    }
    for (int t = 0; t < T; ++t) {
        const int ts = w->keep_all ? t : 0;
        bf16* xin_t = b16(w->xin_all, (long long)t * B * KX);
        bf16* x_t = b16(w->x_all, ts * BD);
        bf16* u_t = b16(w->u_cat, ts * BU);
        bf16* hb_n = b16(w->hb_all, (t + 1) * BD);
        if (t == 0) RSTEP_TRY(mrssm_rstep_xin(a, t, KX, xin_t, stream));          // later steps: written by the previous step's heads kernel
        RSTEP_TRY(gemm(st, B, xin_t, KX, KX, w->wp_sa, D, D, a->b_sa, a->act, nullptr, 0, 0, x_t, D, 0));
        RSTEP_TRY(gemm(st, B, x_t, D, D, w->wp_ih, 3 * D, 3 * D, a->b_ih, 0, nullptr, 0, 0, w->gi, 3 * D, 1));
        if (use_aux) MRSSM_CUDA(cudaStreamWaitEvent(st, ax.join, 0));          // W_hh h_{t-1} has run beside the kernels above
        else RSTEP_TRY(gh_gemm(st, t));
        RSTEP_TRY(mrssm_rstep_gate_fwd(a, t, w->gi, w->gh, hb_n, stream));
        if (use_aux && t + 1 < T) {          // the next step's W_hh h_t beside this step's heads and the next step's W_sa, W_ih GEMMs
            MRSSM_CUDA(cudaEventRecord(ax.fork, st));
            MRSSM_CUDA(cudaStreamWaitEvent(ax.s, ax.fork, 0));
            RSTEP_TRY(gh_gemm(ax.s, t + 1));
            MRSSM_CUDA(cudaEventRecord(ax.join, ax.s));
        }
        for (int ci = 0; ci < w->n_chunks; ++ci) {
            const int c0 = w->chunk_c0[ci], n = (w->chunk_c1[ci] - c0) * H;
            const float* add = w->pre_cat ? w->pre_cat + ((long long)t * B * NH * H + (long long)c0 * H) : nullptr;
            RSTEP_TRY(gemm(st, B, hb_n, D, D, w->w1f[ci], n, n, w->b1[ci], a->act, nullptr, 0, 0, u_t + (long long)c0 * H, (long long)NH * H, 0, add,
                           (long long)NH * H));
        }
        RSTEP_TRY(gemm(st, B, u_t, (long long)NH * H, NH * H, w->w2f, NH * S2p, NH * S2p, w->b2, 0, nullptr, 0, 0, w->o_cat, (long long)NH * S2p, 1,
                       nullptr, 0, S2p, H));
        RSTEP_TRY(mrssm_rstep_heads_fwd(a, t, w->o_cat, NH * S2p, S2p, t + 1 < T ? (void*)b16(w->xin_all, (long long)(t + 1) * B * KX) : nullptr, KX,
                                        stream));
    }
    return 0;
}

extern "C" int mrssm_rollout_steps_bwd(const mrssm_rollout_bwd_args* g, const mrssm_rstep_ws* w, void* stream) {
    MRSSM_CHECK(g && w && w->x_all && w->u_cat && w->d_o && w->du_all && w->d_gi && w->d_gh && w->d_xpre && w->dh_heads && w->carry_a && w->carry_b &&
                    w->dxin && w->cgs && w->g_prev_belief && w->wp_sa_b && w->wp_ih_b && w->wp_hh_b && w->w1b && w->keep_all,
                "rollout_steps_bwd: null workspace / weights (the forward must have kept its stash)");
    const mrssm_rollout_args* a = &g->f;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = a->T, B = a->B, D = a->D, H = a->H, S = a->S, A = a->A, NH = w->NH, S2p = w->S2p;
    const long long BD = (long long)B * D, BU = (long long)B * NH * H, BO = (long long)B * NH * S2p;
    void* d_o_ptrs[MRSSM_MAX_HEADS];
    Aux& ax = aux_stream();
    const bool use_aux = g_step_aux && ax.ok;
    bool pending = false;                     // a d_gh W_hh GEMM is in flight on the second stream
    for (int t = T - 1; t >= 0; --t) {
        bf16* d_o_t = b16(w->d_o, t * BO);
        bf16* du_t = b16(w->du_all, t * BU);
        bf16* u_t = b16(w->u_cat, t * BU);
        bf16* x_t = b16(w->x_all, t * BD);
        bf16* dgi_t = b16(w->d_gi, 3 * t * BD);
        bf16* dgh_t = b16(w->d_gh, 3 * t * BD);
        bf16* dxp_t = b16(w->d_xpre, t * BD);
        for (int hd = 0; hd < MRSSM_MAX_HEADS; ++hd) d_o_ptrs[hd] = hd < NH ? (void*)(d_o_t + (long long)hd * S2p) : nullptr;
        // (the state-gradient carry and g_actions[t + 1] come straight from step t + 1's dxin; the last step has no successor: cgs = 0)
        RSTEP_TRY(mrssm_rstep_heads_bwd(g, t, w->cgs, t + 1 < T ? w->dxin : nullptr, d_o_ptrs, NH * S2p, S2p, stream));
        for (int ci = 0; ci < w->n_chunks; ++ci) {
            const int c0 = w->chunk_c0[ci], nc = w->chunk_c1[ci] - c0;
            RSTEP_TRY(gemm(st, B, d_o_t + (long long)c0 * S2p, (long long)NH * S2p, nc * S2p, w->w2b[ci], nc * H, nc * H, nullptr, 0,
                           u_t + (long long)c0 * H, (long long)NH * H, a->act, du_t + (long long)c0 * H, (long long)NH * H, 0, nullptr, 0, H, S2p));
        }
        RSTEP_TRY(gemm(st, B, du_t, (long long)NH * H, NH * H, w->w1b, D, D, nullptr, 0, nullptr, 0, 0, w->dh_heads, D, 1));
        if (pending) {                           // carry_b = d_gh W_hh of step t + 1
            MRSSM_CUDA(cudaStreamWaitEvent(st, ax.join, 0));
            pending = false;
        }
        RSTEP_TRY(mrssm_rstep_gate_bwd(g, t, w->dh_heads, w->carry_a, w->carry_b, dgi_t, dgh_t, stream));
        if (use_aux) {                           // d_gh W_hh feeds only the previous step's gate kernel: beside the W_ih / W_sa dgrads and the heads of step t - 1
            MRSSM_CUDA(cudaEventRecord(ax.fork, st));
            MRSSM_CUDA(cudaStreamWaitEvent(ax.s, ax.fork, 0));
            RSTEP_TRY(gemm(ax.s, B, dgh_t, 3 * D, 3 * D, w->wp_hh_b, D, D, nullptr, 0, nullptr, 0, 0, w->carry_b, D, 1));
            MRSSM_CUDA(cudaEventRecord(ax.join, ax.s));
            pending = true;
        } else {
            RSTEP_TRY(gemm(st, B, dgh_t, 3 * D, 3 * D, w->wp_hh_b, D, D, nullptr, 0, nullptr, 0, 0, w->carry_b, D, 1));
        }
        RSTEP_TRY(gemm(st, B, dgi_t, 3 * D, 3 * D, w->wp_ih_b, D, D, nullptr, 0, x_t, D, a->act, dxp_t, D, 0));
        RSTEP_TRY(gemm(st, B, dxp_t, D, D, w->wp_sa_b, w->KXo, S + A, nullptr, 0, nullptr, 0, 0, w->dxin, S + A, 1));
    }
    if (pending) MRSSM_CUDA(cudaStreamWaitEvent(st, ax.join, 0));
    RSTEP_TRY(mrssm_rstep_xin_bwd(g, 0, w->dxin, S + A, w->cgs, stream));      // step 0's dxin: gradient of prev_state, g_actions[0]
    return mrssm_add2(w->carry_a, w->carry_b, BD, w->g_prev_belief, stream);
}
