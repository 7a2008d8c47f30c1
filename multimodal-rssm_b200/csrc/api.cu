// extern "C" dispatch for the conv family and the rollout (include/mrssm_b200.h).
#include "common.cuh"

int conv_down_simt(const mrssm_conv_args* a, cudaStream_t st);
int conv_up_simt(const mrssm_conv_args* a, cudaStream_t st);
int conv_wgrad_simt(const mrssm_conv_args* a, cudaStream_t st);
int rollout_fwd_simt(const mrssm_rollout_args* a, cudaStream_t st);
int rollout_bwd_simt(const mrssm_rollout_bwd_args* a, cudaStream_t st);

extern "C" int mrssm_conv_down(const mrssm_conv_args* a, void* stream) { return conv_down_simt(a, (cudaStream_t)stream); }
extern "C" int mrssm_conv_up(const mrssm_conv_args* a, void* stream) { return conv_up_simt(a, (cudaStream_t)stream); }
extern "C" int mrssm_conv_wgrad(const mrssm_conv_args* a, void* stream) { return conv_wgrad_simt(a, (cudaStream_t)stream); }
extern "C" int mrssm_rollout_fwd(const mrssm_rollout_args* a, void* stream) { return rollout_fwd_simt(a, (cudaStream_t)stream); }
extern "C" int mrssm_rollout_bwd(const mrssm_rollout_bwd_args* a, void* stream) { return rollout_bwd_simt(a, (cudaStream_t)stream); }
