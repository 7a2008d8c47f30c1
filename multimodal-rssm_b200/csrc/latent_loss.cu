// Latent half of the ELBO: posterior re-fusion + decoder-latent sample, balanced / MoPoE KL with
// free nats, global KL — forward (row values + deterministic two-stage reduce) and backward.
// One warp per (t,b) row, lanes stride over the S state dims, warp-shuffle reductions.
// Reference: base/algo.py:75-94,157-163,186-188; MRSSM_PoE/algo.py:63-68;
// MRSSM_MoPoE/algo.py:62-67,110-137; encoder.py:50-124.
#include "common.cuh"

namespace {

constexpr int MAXE = MRSSM_MAX_HEADS;   // experts are indexed 1..n_experts

struct Experts {
    float mu[MAXE], sd[MAXE];
};

__device__ __forceinline__ void load_experts(const mrssm_latent_args& a, long long off, Experts& x) {
#pragma unroll
    for (int e = 1; e < MAXE; ++e)
        if (e <= a.n_experts) {
            x.mu[e] = a.exp_means[e][off];
            x.sd[e] = a.exp_stds[e][off];
        }
}

__device__ __forceinline__ void poe_fwd(const mrssm_latent_args& a, const Experts& x, unsigned mask, float& mu, float& sd) {
    float sumT = 0.f, sumMT = 0.f;
#pragma unroll
    for (int e = 1; e < MAXE; ++e)
        if (e <= a.n_experts && (mask & (1u << (e - 1)))) {
            float t = 1.f / x.sd[e];
            sumT += t;
            sumMT = fmaf(x.mu[e], t, sumMT);
        }
    mu = sumMT / sumT;
    sd = 1.f / sumT;
}

__device__ __forceinline__ void poe_bwd(const mrssm_latent_args& a, const Experts& x, unsigned mask, float mu, float sd,
                                        float gmu, float gsd, float (&gm)[MAXE], float (&gs)[MAXE]) {
    float P = 1.f / sd;
#pragma unroll
    for (int e = 1; e < MAXE; ++e)
        if (e <= a.n_experts && (mask & (1u << (e - 1)))) {
            float t = 1.f / x.sd[e];
            gm[e] = fmaf(gmu, t / P, gm[e]);
            float gT = gmu * (x.mu[e] - mu) / P - gsd / (P * P);
            gs[e] = fmaf(-gT, t * t, gs[e]);
        }
}

__device__ __forceinline__ float kl_nn(float mq, float sq, float mp, float sp) {
    float vr = (sq / sp) * (sq / sp);
    float t1 = ((mq - mp) / sp) * ((mq - mp) / sp);
    return 0.5f * (vr + t1 - 1.f - logf(vr));
}

__global__ void latent_fwd_kernel(mrssm_latent_args a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.rows) return;
    const int S = a.S;
    float div[MRSSM_MAX_SUBSETS];
#pragma unroll
    for (int j = 0; j < MRSSM_MAX_SUBSETS; ++j) div[j] = 0.f;
    float glob = 0.f;
    for (int s = lane; s < S; s += 32) {
        long long off = (long long)warp * S + s;
        float pm = a.prior_means[off], ps = a.prior_stds[off];
        Experts x;
        if (a.n_experts && (a.refuse || a.kl_mode == 1)) load_experts(a, off, x);
        float qm, qs;
        if (a.refuse) {
            poe_fwd(a, x, a.subset_mask[a.dim_subset[s]], qm, qs);
            a.q_means[off] = qm;
            a.q_stds[off] = qs;
            a.z_dec[off] = fmaf(qs, a.eps_dec[off], qm);
        } else {
            qm = a.post_means[off];
            qs = a.post_stds[off];
        }
        glob += 0.5f * (qs * qs + qm * qm - 1.f - logf(qs * qs));
        if (a.kl_mode == 1) {
#pragma unroll
            for (int j = 0; j < MRSSM_MAX_SUBSETS; ++j)
                if (j < a.n_subsets) {
                    float m, sd;
                    poe_fwd(a, x, a.subset_mask[j], m, sd);
                    div[j] += kl_nn(m, sd, pm, ps);
                }
        } else {
            div[0] += kl_nn(qm, qs, pm, ps);
        }
    }
    glob = warp_sum(glob);
    float val = 0.f;
    if (a.kl_mode == 1) {
#pragma unroll
        for (int j = 0; j < MRSSM_MAX_SUBSETS; ++j)
            if (j < a.n_subsets) val += fmaxf(warp_sum(div[j]), a.free_nats);
        val /= (float)a.n_subsets;
    } else {
        val = fmaxf(warp_sum(div[0]), a.free_nats);
    }
    if (lane == 0) {
        a.row_scratch[warp] = val;
        a.row_scratch[a.rows + warp] = glob;
    }
}

// deterministic: one block, fixed order
__global__ void reduce2_kernel(const float* scratch, int rows, float scale, float* out) {
    __shared__ double sh[2][32];
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        s0 += scratch[i];
        s1 += scratch[rows + i];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = s0; sh[1][w] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { t0 += sh[0][i]; t1 += sh[1][i]; }
        out[0] = (float)(t0 * scale);
        out[1] = (float)(t1 * scale);
    }
}

__global__ void latent_bwd_kernel(mrssm_latent_args a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.rows) return;
    const int S = a.S;
    const float g_kl = a.g_sums[0] / (float)a.rows, g_glob = a.g_sums[1] / (float)a.rows;
    // pass 1: which clamps are active
    float div[MRSSM_MAX_SUBSETS];
#pragma unroll
    for (int j = 0; j < MRSSM_MAX_SUBSETS; ++j) div[j] = 0.f;
    for (int s = lane; s < S; s += 32) {
        long long off = (long long)warp * S + s;
        float pm = a.prior_means[off], ps = a.prior_stds[off];
        if (a.kl_mode == 1) {
            Experts x;
            load_experts(a, off, x);
#pragma unroll
            for (int j = 0; j < MRSSM_MAX_SUBSETS; ++j)
                if (j < a.n_subsets) {
                    float m, sd;
                    poe_fwd(a, x, a.subset_mask[j], m, sd);
                    div[j] += kl_nn(m, sd, pm, ps);
                }
        } else {
            float qm, qs;
            if (a.refuse) {
                Experts x;
                load_experts(a, off, x);
                poe_fwd(a, x, a.subset_mask[a.dim_subset[s]], qm, qs);
            } else {
                qm = a.post_means[off];
                qs = a.post_stds[off];
            }
            div[0] += kl_nn(qm, qs, pm, ps);
        }
    }
    float w[MRSSM_MAX_SUBSETS];
#pragma unroll
    for (int j = 0; j < MRSSM_MAX_SUBSETS; ++j) {
        int nj = a.kl_mode == 1 ? a.n_subsets : 1;
        w[j] = 0.f;
        if (j < nj) w[j] = (warp_sum(div[j]) > a.free_nats) ? g_kl / (float)nj : 0.f;
    }
    const float wp = a.alpha >= 0.f ? a.alpha : 1.f, wq = a.alpha >= 0.f ? 1.f - a.alpha : 1.f;
    // pass 2: gradients
    for (int s = lane; s < S; s += 32) {
        long long off = (long long)warp * S + s;
        float pm = a.prior_means[off], ps = a.prior_stds[off];
        Experts x;
        float gm[MAXE], gs[MAXE];
#pragma unroll
        for (int e = 0; e < MAXE; ++e) { gm[e] = 0.f; gs[e] = 0.f; }
        if (a.n_experts && (a.refuse || a.kl_mode == 1)) load_experts(a, off, x);
        float qm, qs;
        unsigned qmask = 0;
        if (a.refuse) {
            qmask = a.subset_mask[a.dim_subset[s]];
            poe_fwd(a, x, qmask, qm, qs);
        } else {
            qm = a.post_means[off];
            qs = a.post_stds[off];
        }
        float gqm = g_glob * qm, gqs = g_glob * (qs - 1.f / qs);
        if (a.g_z) {
            float gz = a.g_z[off];
            gqm += gz;
            gqs = fmaf(gz, a.eps_dec[off], gqs);
        }
        float gpm = 0.f, gps = 0.f;
        if (a.kl_mode == 1) {
#pragma unroll
            for (int j = 0; j < MRSSM_MAX_SUBSETS; ++j)
                if (j < a.n_subsets && w[j] != 0.f) {
                    float m, sd;
                    poe_fwd(a, x, a.subset_mask[j], m, sd);
                    float d = m - pm, ip2 = 1.f / (ps * ps);
                    poe_bwd(a, x, a.subset_mask[j], m, sd, w[j] * d * ip2, w[j] * (sd * ip2 - 1.f / sd), gm, gs);
                    gpm -= w[j] * d * ip2;
                    gps += w[j] * (1.f / ps - (sd * sd + d * d) * ip2 / ps);
                }
        } else if (w[0] != 0.f) {
            float d = qm - pm, ip2 = 1.f / (ps * ps);
            gqm += wq * w[0] * d * ip2;
            gqs += wq * w[0] * (qs * ip2 - 1.f / qs);
            gpm -= wp * w[0] * d * ip2;
            gps += wp * w[0] * (1.f / ps - (qs * qs + d * d) * ip2 / ps);
        }
        a.g_prior_means[off] = gpm;
        a.g_prior_stds[off] = gps;
        if (a.refuse) {
            poe_bwd(a, x, qmask, qm, qs, gqm, gqs, gm, gs);
        } else {
            a.g_post_means[off] = gqm;
            a.g_post_stds[off] = gqs;
        }
        if (a.refuse || a.kl_mode == 1) {
#pragma unroll
            for (int e = 1; e < MAXE; ++e)
                if (e <= a.n_experts) {
                    a.g_exp_means[e][off] = gm[e];
                    a.g_exp_stds[e][off] = gs[e];
                }
        }
    }
}

int check(const mrssm_latent_args* a, bool bwd) {
    MRSSM_CHECK(a && a->rows > 0 && a->S > 0 && a->S <= MRSSM_MAX_STATE, "latent: bad sizes");
    MRSSM_CHECK(a->n_experts >= 0 && a->n_experts < MRSSM_MAX_HEADS, "latent: n_experts");
    MRSSM_CHECK(a->prior_means && a->prior_stds && a->row_scratch && a->out_sums, "latent: null tensor");
    if (a->refuse || a->kl_mode == 1) {
        MRSSM_CHECK(a->n_experts > 0 && a->n_subsets > 0 && a->n_subsets <= MRSSM_MAX_SUBSETS, "latent: fusion table missing");
        for (int e = 1; e <= a->n_experts; ++e) MRSSM_CHECK(a->exp_means[e] && a->exp_stds[e], "latent: expert %d missing", e);
    }
    if (a->refuse) MRSSM_CHECK(a->eps_dec && (bwd || (a->z_dec && a->q_means && a->q_stds)), "latent: refuse outputs missing");
    else MRSSM_CHECK(a->post_means && a->post_stds, "latent: posterior missing");
    if (bwd) {
        MRSSM_CHECK(a->g_sums && a->g_prior_means && a->g_prior_stds, "latent_bwd: null grads");
        MRSSM_CHECK(a->refuse || !a->g_z, "latent_bwd: g_z only valid with refuse");
        if (!a->refuse) MRSSM_CHECK(a->g_post_means && a->g_post_stds, "latent_bwd: posterior grads missing");
        if (a->refuse || a->kl_mode == 1)
            for (int e = 1; e <= a->n_experts; ++e) MRSSM_CHECK(a->g_exp_means[e] && a->g_exp_stds[e], "latent_bwd: expert grad %d", e);
    }
    return 0;
}

}  // namespace

extern "C" int mrssm_latent_fwd(const mrssm_latent_args* a, void* stream) {
    if (int e = check(a, false)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    int wpb = 8;
    latent_fwd_kernel<<<(a->rows + wpb - 1) / wpb, wpb * 32, 0, st>>>(*a);
    MRSSM_LAUNCH_CHECK();
    reduce2_kernel<<<1, 1024, 0, st>>>(a->row_scratch, a->rows, 1.f / (float)a->rows, a->out_sums);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_latent_bwd(const mrssm_latent_args* a, void* stream) {
    if (int e = check(a, true)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    int wpb = 8;
    latent_bwd_kernel<<<(a->rows + wpb - 1) / wpb, wpb * 32, 0, st>>>(*a);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
