// Latent overshooting (reference base/algo.py:111-148, MRSSM_MoPoE/algo.py:69-108).
// The reference pads and concatenates one open-loop run per start step t = 1..T-2 along the batch; here
//   * one gather kernel lays out the padded actions / nonterminals / rewards / sequence mask of all runs,
//   * the masked KL reads its (detached) target posterior straight from the [T-1,B,S] rollout tensors by index — for
//     PoE / MoPoE the product of the experts a subset selects is formed in registers — so no padded copy of q exists,
//   * rows past a run's end contribute exactly `free_nats` (KL * 0 = 0, then the clamp) and no gradient.
// Row r = k*N + n of the imagination rollout [OD, N=(T-2)B, S]:  n = (t-1)*B + b, source time t+k, valid iff t+k < T-1.
// One warp per row, lanes stride over S; deterministic one-block final reduce.
#include "common.cuh"

namespace {

constexpr int MAXE = MRSSM_MAX_HEADS;

struct Row {
    int k, b, tt;
    bool valid;
};

__device__ __forceinline__ Row row_of(const mrssm_overshoot_args& a, int row) {
    const int N = (a.T - 2) * a.B;
    Row r;
    r.k = row / N;
    int n = row - r.k * N;
    int t = n / a.B + 1;
    r.b = n - (t - 1) * a.B;
    r.tt = t + r.k;
    r.valid = r.tt < a.T - 1;
    return r;
}

// target posterior of one element: the rollout's posterior, or the product of the selected experts (encoder.py:50-71)
__device__ __forceinline__ void target(const mrssm_overshoot_args& a, long long qoff, float& qm, float& qs) {
    if (a.n_experts == 0) {
        qm = a.post_means[qoff];
        qs = a.post_stds[qoff];
        return;
    }
    float sumT = 0.f, sumMT = 0.f;
#pragma unroll
    for (int e = 1; e < MAXE; ++e)
        if (e <= a.n_experts && (a.subset_mask & (1u << (e - 1)))) {
            float t = 1.f / a.exp_stds[e][qoff];
            sumT += t;
            sumMT = fmaf(a.exp_means[e][qoff], t, sumMT);
        }
    qm = sumMT / sumT;
    qs = 1.f / sumT;
}

__device__ __forceinline__ float kl_nn(float mq, float sq, float mp, float sp) {
    float vr = (sq / sp) * (sq / sp);
    float t1 = ((mq - mp) / sp) * ((mq - mp) / sp);
    return 0.5f * (vr + t1 - 1.f - logf(vr));
}

__global__ void overshoot_kl_fwd_kernel(mrssm_overshoot_args a, int rows) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const Row r = row_of(a, row);
    float sum = 0.f;
    if (r.valid) {
        const long long poff = (long long)row * a.S, qoff = ((long long)r.tt * a.B + r.b) * a.S;
        for (int s = lane; s < a.S; s += 32) {
            float qm, qs;
            target(a, qoff + s, qm, qs);
            sum += kl_nn(qm, qs, a.prior_means[poff + s], a.prior_stds[poff + s]);
        }
        sum = warp_sum(sum);
    }
    if (lane == 0) a.row_scratch[row] = fmaxf(sum, a.free_nats);
}

__global__ void reduce1_kernel(const float* scratch, int rows, float scale, float* out) {
    __shared__ double sh[32];
    double s0 = 0.0;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) s0 += scratch[i];
    for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t0 += sh[i];
        out[0] = (float)(t0 * scale);
    }
}

__global__ void overshoot_kl_bwd_kernel(mrssm_overshoot_args a, int rows) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const Row r = row_of(a, row);
    const long long poff = (long long)row * a.S, qoff = ((long long)r.tt * a.B + r.b) * a.S;
    // the clamp passes gradient only where the (valid) row's KL exceeded free_nats; the forward kept max(sum, free_nats)
    const bool active = r.valid && a.row_scratch[row] > a.free_nats;
    const float g = active ? a.g_out[0] * a.scale / (float)rows : 0.f;
    for (int s = lane; s < a.S; s += 32) {
        float gpm = 0.f, gps = 0.f;
        if (active) {
            float qm, qs;
            target(a, qoff + s, qm, qs);
            float pm = a.prior_means[poff + s], ps = a.prior_stds[poff + s];
            float d = qm - pm, ip2 = 1.f / (ps * ps);
            gpm = -g * d * ip2;
            gps = g * (1.f / ps - (qs * qs + d * d) * ip2 / ps);
        }
        a.g_prior_means[poff + s] = gpm;
        a.g_prior_stds[poff + s] = gps;
    }
}

__global__ void overshoot_gather_kernel(mrssm_overshoot_args a, int rows) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const Row r = row_of(a, row);
    const long long src = (long long)r.tt * a.B + r.b;
    for (int j = 0; j < a.A; ++j) a.actions_o[(long long)row * a.A + j] = r.valid ? a.actions[src * a.A + j] : 0.f;
    a.nonterminals_o[row] = r.valid ? a.nonterminals[src] : 0.f;
    if (a.rewards_o) a.rewards_o[row] = (r.valid && a.rewards) ? a.rewards[src] : 0.f;
    if (a.mask_o) a.mask_o[row] = r.valid ? 1.f : 0.f;
}

int check(const mrssm_overshoot_args* a) {
    MRSSM_CHECK(a && a->T >= 3 && a->B > 0 && a->OD > 0, "overshoot: needs T >= 3, B > 0, distance > 0");
    MRSSM_CHECK((long long)a->OD * (a->T - 2) * a->B < (1ll << 26), "overshoot: too many rows");     /* 32 threads per row, 32-bit thread index */
    return 0;
}

int check_kl(const mrssm_overshoot_args* a, bool bwd) {
    if (int e = check(a)) return e;
    MRSSM_CHECK(a->S > 0 && a->S <= MRSSM_MAX_STATE, "overshoot_kl: bad state size");
    MRSSM_CHECK(a->n_experts >= 0 && a->n_experts < MRSSM_MAX_HEADS, "overshoot_kl: n_experts");
    MRSSM_CHECK(a->prior_means && a->prior_stds && a->row_scratch, "overshoot_kl: null tensor");
    if (a->n_experts) {
        MRSSM_CHECK(a->subset_mask != 0 && (a->subset_mask >> a->n_experts) == 0, "overshoot_kl: bad subset mask");
        for (int e = 1; e <= a->n_experts; ++e)
            if (a->subset_mask & (1u << (e - 1))) MRSSM_CHECK(a->exp_means[e] && a->exp_stds[e], "overshoot_kl: expert %d missing", e);
    } else {
        MRSSM_CHECK(a->post_means && a->post_stds, "overshoot_kl: posterior missing");
    }
    if (bwd) MRSSM_CHECK(a->g_out && a->g_prior_means && a->g_prior_stds, "overshoot_kl_bwd: null grads");
    else MRSSM_CHECK(a->out, "overshoot_kl: null output");
    return 0;
}

}  // namespace

extern "C" int mrssm_overshoot_gather(const mrssm_overshoot_args* a, void* stream) {
    if (int e = check(a)) return e;
    MRSSM_CHECK(a->A > 0 && a->actions && a->nonterminals && a->actions_o && a->nonterminals_o, "overshoot_gather: null tensor");
    const int rows = a->OD * (a->T - 2) * a->B;
    overshoot_gather_kernel<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a, rows);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_overshoot_kl_fwd(const mrssm_overshoot_args* a, void* stream) {
    if (int e = check_kl(a, false)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = a->OD * (a->T - 2) * a->B, wpb = 8;
    overshoot_kl_fwd_kernel<<<(rows + wpb - 1) / wpb, wpb * 32, 0, st>>>(*a, rows);
    MRSSM_LAUNCH_CHECK();
    reduce1_kernel<<<1, 1024, 0, st>>>(a->row_scratch, rows, a->scale / (float)rows, a->out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_overshoot_kl_bwd(const mrssm_overshoot_args* a, void* stream) {
    if (int e = check_kl(a, true)) return e;
    const int rows = a->OD * (a->T - 2) * a->B, wpb = 8;
    overshoot_kl_bwd_kernel<<<(rows + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(*a, rows);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
