// extern "C" dispatch for the conv family and the rollout (include/mrssm_b200.h).
#include "common.cuh"

int conv_down_simt(const mrssm_conv_args* a, cudaStream_t st);
int conv_up_simt(const mrssm_conv_args* a, cudaStream_t st);
int conv_wgrad_simt(const mrssm_conv_args* a, cudaStream_t st);
int rollout_fwd_simt(const mrssm_rollout_args* a, cudaStream_t st);
int rollout_bwd_simt(const mrssm_rollout_bwd_args* a, cudaStream_t st);
int rollout_fwd_staged(const mrssm_rollout_args* a, cudaStream_t st);                                   // -1: shapes not eligible
int rollout_bwd_staged(const mrssm_rollout_bwd_args* a, const float* const* w1c, cudaStream_t st);      // -1: shapes not eligible
int rollout_check(const mrssm_rollout_args* a);
int rollout_bwd_check(const mrssm_rollout_bwd_args* g);

extern "C" int mrssm_conv_down(const mrssm_conv_args* a, void* stream) { return conv_down_simt(a, (cudaStream_t)stream); }
extern "C" int mrssm_conv_up(const mrssm_conv_args* a, void* stream) { return conv_up_simt(a, (cudaStream_t)stream); }
extern "C" int mrssm_conv_wgrad(const mrssm_conv_args* a, void* stream) { return conv_wgrad_simt(a, (cudaStream_t)stream); }
// The weight-streaming kernels (rollout_staged.cu) run when the sizes allow one output feature per thread; otherwise the
// L2-streaming kernels (rollout_simt.cu).  Same arithmetic either way.
extern "C" int mrssm_rollout_fwd(const mrssm_rollout_args* a, void* stream) {
    if (int e = rollout_check(a)) return e;
    const int rc = rollout_fwd_staged(a, (cudaStream_t)stream);
    return rc == -1 ? rollout_fwd_simt(a, (cudaStream_t)stream) : rc;
}
extern "C" int mrssm_rollout_bwd(const mrssm_rollout_bwd_args* a, void* stream) {
    if (int e = rollout_bwd_check(a)) return e;
    bool have = true;
    for (int h = 0; h <= a->f.n_experts; ++h) have = have && a->w1_belief[h] != nullptr;
    const int rc = have ? rollout_bwd_staged(a, a->w1_belief, (cudaStream_t)stream) : -1;
    return rc == -1 ? rollout_bwd_simt(a, (cudaStream_t)stream) : rc;
}
