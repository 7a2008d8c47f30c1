/* mrssm_b200.h — C ABI of the B200-native MRSSM training hot path.
 *
 * The reference (EmergentSystemLabStudent/Multimodal-RSSM) has no FFI: its boundary for this path
 * is a set of duck-typed PyTorch modules.  Each entry point below replaces the
 * library-kernel work those modules reach through PyTorch; the reference call site is cited on
 * every declaration (paths relative to the reference root).  The Python host side in
 * multimodal-rssm_b200/{algos,utils} mirrors the reference modules and binds these symbols with
 * ctypes (multimodal-rssm_b200/mrssm_b200/_lib.py); INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers unless the name ends in _host.  No torch types.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *  - Every function returns 0 on success, non-zero on failure; mrssm_last_error() then returns a
 *    thread-local human readable message.  Nothing falls back to the CPU.
 *  - Tensors are addressed by (pointer, 4 element strides) so NCHW (reference layout) and NHWC
 *    (internal layout) are both first class.
 *  - dtype: MRSSM_F32 tensors are float; MRSSM_BF16 tensors are __nv_bfloat16 (tensor-core mode).
 */
#ifndef MRSSM_B200_H
#define MRSSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRSSM_ABI_VERSION 1

enum { MRSSM_ACT_NONE = 0, MRSSM_ACT_RELU = 1, MRSSM_ACT_ELU = 2 };
enum { MRSSM_F32 = 0, MRSSM_BF16 = 1 };
#define MRSSM_MAX_HEADS 5   /* prior + prior_expert + up to 3 modality experts */
#define MRSSM_MAX_SUBSETS 8 /* 2^3 */
#define MRSSM_MAX_STATE 256

const char* mrssm_last_error(void);
int mrssm_abi_version(void);
/* 1 if the loaded library carries sm_100a code and a CUDA device of capability 10.x is present */
int mrssm_device_ok(void);

/* ---- strided 4-D tensor view: element (i,h,w,c) at ptr[i*sI + h*sH + w*sW + c*sC] ---------- */
typedef struct mrssm_t4 {
    void* ptr;
    int64_t sI, sH, sW, sC;
} mrssm_t4;

/* ---- convolution family --------------------------------------------------------------------
 * One geometry describes nn.Conv2d(k, stride 2) and nn.ConvTranspose2d(k, stride 2) alike:
 * a "large" tensor [n_img,Hl,Wl,Cl] and a "small" tensor [n_img,Hs,Ws,Cs], Hl >= 2*(Hs-1)+ksz
 * (equality for ConvTranspose2d; Conv2d floors, leaving the last row/column of `large` unused),
 * and a weight W[cs][cl][kh][kw] at weight[cs*w_ss + cl*w_sl + kh*ksz + kw] — which is exactly
 * Conv2d.weight [Cout=Cs,Cin=Cl,k,k] and ConvTranspose2d.weight [Cin=Cs,Cout=Cl,k,k].
 * nn.Linear is the case Hl=Wl=Hs=Ws=ksz=1 (weight [out=Cs,in=Cl]).
 *   down : small = act(bias + sum_{kh,kw,cl} large[2hs+kh,2ws+kw,cl] * W)   Conv2d fwd, ConvT dgrad, Linear fwd
 *   up   : large = act(bias + sum_{kh,kw,cs} small[hs,ws,cs] * W)           ConvT fwd, Conv2d dgrad, Linear dgrad
 *   wgrad: dW   += sum_{img,hs,ws} small * large ; dbias += colsum           all weight gradients
 * Replaces: encoder.py:315-322 (ImageEncoder convs), :423-432 (128x128), :287-296 (SymbolicEncoder),
 * observation_model.py:65-74,99-102 (ImageDecoder), :172-181 (128x128), :37-51 (DenseDecoder),
 * encoder.py:136-141,172-176 (the time-parallel half of the expert heads) and their autograd.
 */
typedef struct mrssm_conv_args {
    int32_t n_img, Hl, Wl, Cl, Hs, Ws, Cs, ksz;
    int32_t dtype;        /* MRSSM_F32 | MRSSM_BF16 (activations); weights/bias are always f32 masters */
    int32_t act;          /* epilogue activation (down/up) */
    int32_t mask_mode;    /* epilogue multiply by act'(mask): MRSSM_ACT_*; mask has the output's strides */
    int32_t accumulate;   /* down/up: out = mask * act(out_old + result + bias) */
    mrssm_t4 large, small;
    float* weight;        /* down/up: input; wgrad: output (accumulated) */
    int64_t w_ss, w_sl;
    float* bias;          /* down/up: input or NULL; wgrad: dbias output (accumulated) or NULL */
    void* mask;           /* or NULL */
} mrssm_conv_args;

int mrssm_conv_down(const mrssm_conv_args* a, void* stream);
int mrssm_conv_up(const mrssm_conv_args* a, void* stream);
int mrssm_conv_wgrad(const mrssm_conv_args* a, void* stream);
/* bias gradient of a ConvTranspose2d: a->bias[cl] += sum over (img,h,w) of a->large (observation_model.py:65-74 autograd) */
int mrssm_colsum_t4(const mrssm_conv_args* a, void* stream);


/* ---- tensor-core (tcgen05) conv family: bf16 activations, fp32 accumulation --------------------------
 * Same geometry as mrssm_conv_args.  Activations are bf16 NHWC with the channel count padded to a
 * multiple of 8 (Cl / Cs below are the PADDED counts; padded channels hold zeros).  Weights for
 * down/up are packed once per optimiser step by mrssm_tc_pack_weight (bf16, K-major, zero padded);
 * wgrad accumulates fp32 straight into the PyTorch-layout master gradient.  The output of down/up
 * is bf16 NHWC (padded channels written as zeros) or, with out_f32, fp32 with arbitrary strides. */
typedef struct mrssm_tc_conv_args {
    int32_t n_img, Hl, Wl, Cl, Hs, Ws, Cs, ksz;
    int32_t act, mask_mode, out_f32;
    int32_t n_out_pad;      /* down/up: padded output channels (multiple of 16) = rows of the packed weight */
    int32_t n_out_valid;    /* down/up: real output channels */
    int32_t bias_mod;       /* down/up: bias index = n % bias_mod (0 -> n_out_valid) */
    int32_t cs_valid, cl_valid; /* wgrad: real channel counts of the master weight */
    mrssm_t4 large, small, mask;    /* mask: bf16, indexed like the output pixel, own strides */
    const void* wpacked;    /* down/up */
    const float* bias;      /* down/up, or NULL */
    float* dweight;         /* wgrad output (accumulated) */
    int64_t w_ss, w_sl;     /* wgrad: master weight strides */
    /* dense layers only: fp32 [rows][addend_ld] added to the accumulator before the activation (the hoisted, time-parallel
     * embedding half of an expert's fc1, encoder.py:172-176, when the belief half runs step by step) */
    const float* addend;
    int64_t addend_ld;
    /* dense layers only: block-diagonal ("grouped") GEMM — output columns [g*group_n, (g+1)*group_n) contract input columns
       [g*group_k, (g+1)*group_k) only (all heads' fc2, or their dgrads, in one launch: encoder.py:126-190 runs one such Linear
       per expert).  The packed weight then holds group_k (padded to 64) columns per output row.  0 = plain GEMM. */
    int32_t group_n;
    int32_t group_k;
} mrssm_tc_conv_args;

int mrssm_tc_conv_down(const mrssm_tc_conv_args* a, void* stream);
int mrssm_tc_conv_up(const mrssm_tc_conv_args* a, void* stream);
int mrssm_tc_conv_wgrad(const mrssm_tc_conv_args* a, void* stream);
/* mode 0: down [Npad][Kpad=(tap,cl_pad)]; mode 1: up, 4 parity classes [4][Npad][Kpad=(th,tw,cs_pad)];
 * mode 2: ConvTranspose2d on a 1x1 input as a dense layer [Npad=(tap,cl_pad)][Kpad=cs] */
int mrssm_tc_pack_weight(const float* w, int64_t w_ss, int64_t w_sl, int32_t Cs_valid, int32_t Cl_valid,
                         int32_t Cs_pad, int32_t Cl_pad, int32_t ksz, int32_t mode, int32_t Npad, int32_t Kpad,
                         void* out, void* stream);
/* fp32 strided [n,H,W,C] -> bf16 NHWC with channels padded to Cpad (x scale), and back */
int mrssm_tc_to_bf16(const mrssm_t4* src, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t Cpad, float scale,
                     void* dst, void* stream);
int mrssm_tc_from_bf16(const void* src, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t Cpad,
                       const mrssm_t4* dst, void* stream);
/* out[c] += sum_rows x[row][c] for bf16 x [rows][Cpad] */
int mrssm_tc_colsum(const void* x, int64_t rows, int32_t Cpad, int32_t Cvalid, float* out, void* stream);

/* ---- "plane" tensor-core conv family (ksz >= 2): TMA-staged activation tile + shifted UMMA descriptors ----
 * Replaces the same reference call sites as mrssm_tc_conv_* (encoder.py:315-322, observation_model.py:65-74 and their
 * autograd).  bf16 activations are addressed through views with channels in chunks of 8 (padded channels hold zeros):
 *   linear        : element (i,y,x,c) at ptr[i*sI + y*sH + x*sW + (c/8)*sK + c%8]   (NHWC: sW=C, sK=8; planar: sW=8, sK=H*W*8)
 *   parity-planar : pixel (y,x) lives in plane (y&1)*2+(x&1) at row y>>1, column x>>1:
 *                   ptr[i*sI + ((y&1)*2+(x&1))*sP + (c/8)*sK + (y>>1)*sH + (x>>1)*sW + c%8]
 * x-contiguous views (sW == 8) load with one TMA row per run of pixels; the operand that is gathered with stride 2
 * (`large` of down / wgrad) should be parity-planar, the others planar.  Geometry, weight layout [cs][cl][kh][kw] and the
 * meaning of down / up / wgrad are those of mrssm_conv_args.  `up`: n_out_pad is the per-parity-class channel padding
 * (multiple of 8) of the output; the gathered tensor of `up` needs channels padded to 16.  Weights are packed by
 * mrssm_pl_pack_weight (op 0 = down, 1 = up) into [N_total][K_total] bf16 (shape from mrssm_pl_packed_shape). */
typedef struct mrssm_tv {
    void* ptr;
    int64_t sI, sH, sW, sK, sP;
    int32_t par, reserved;
} mrssm_tv;

typedef struct mrssm_pl_conv_args {
    int32_t n_img, Hl, Wl, Cl, Hs, Ws, Cs, ksz;   /* Cl / Cs: PADDED channel counts of the bf16 views */
    int32_t act, mask_mode, out_f32;
    int32_t n_out_pad, n_out_valid;
    int32_t cs_valid, cl_valid;                    /* wgrad: real channel counts of the master weight */
    int32_t s2d_cq;                                /* > 0 (down / wgrad): `large` is the SPACE-TO-DEPTH form of a cq-channel
                                                      image (cq <= 4): a linear view [n, ceil(Hl/2), ceil(Wl/2), Cl = 16] whose
                                                      channel (py*2+px)*cq + c is channel c of pixel (2y+py, 2x+px).  Hl, Wl stay
                                                      the image's own size.  Made by mrssm_pl_import_s2d; weights packed with op 2. */
    mrssm_tv large, small, mask;                   /* mask: indexed at the output pixel */
    mrssm_t4 out32;                                /* fp32 output of down/up when out_f32 (arbitrary strides, e.g. NCHW) */
    const void* wpacked;
    const float* bias;
    float* dweight;                                /* wgrad output (accumulated), PyTorch layout */
    int64_t w_ss, w_sl;
    /* Optional output scale (device scalar x host factor): down/up outputs and wgrad contributions are multiplied by
     * (*scale_ptr) * scale_mul when scale_ptr != NULL.  Lets a loss gradient that is known only as a device scalar
     * (autograd's grad_output of the scalar loss) be folded into the kernels that consume the un-scaled residual. */
    const float* scale_ptr;
    float scale_mul;
    /* Fused reconstruction loss for `up` with out_f32 (observation_model.py:28-31 + base/algo.py:381-383 on the last
     * ConvTranspose2d): when mse_target != NULL (fp32, addressed with out32's strides) the epilogue computes the residual
     * r = recon - target, adds mse_scale * sum(r^2) to *mse_sum (atomic) and writes r as bf16 in space-to-depth form
     * (n_out_valid <= 4 channels -> 16) into the view `large`; out32.ptr may then be NULL (no reconstruction written). */
    const float* mse_target;
    float* mse_sum;
    float mse_scale;
    /* wgrad: bias gradient of the same layer from the operand tile that is already in shared memory (autograd of the `+ bias`
     * of nn.Conv2d encoder.py:315-322 / nn.ConvTranspose2d observation_model.py:65-74).  dbias[c] += scale * sum over every
     * pixel of channel c of `small` (dbias_from = 1: Conv2d, the gradient is the small tensor, cs_valid channels) or of
     * `large` (dbias_from = 2: ConvTranspose2d, cl_valid channels; with s2d_cq the four parity copies fold onto channel c).
     * Replaces a separate pass over the gradient tensor (mrssm_pl_colsum). */
    float* dbias;
    int32_t dbias_from;
    /* ReLU sign bits, one bit per element, byte (pixel, 8-channel chunk) at [img][y][x][C/8] of the OUTPUT tensor:
     * relu_bits_out (down / up with act = ReLU): written by the epilogue, bit = output > 0;
     * relu_bits_in  (down / up, mask_mode = ReLU, instead of `mask`): the act'-mask of a dgrad, 1/16 of the bytes of the
     * bf16 activation it stands for. */
    uint8_t* relu_bits_out;
    const uint8_t* relu_bits_in;
    /* Fused reconstruction loss, target given as the bf16 SPACE-TO-DEPTH view of the image (the form mrssm_pl_import_s2d makes for
     * the first encoder convolution — in training the decoder's target IS the encoder's input, base/algo.py:241,270-273) instead
     * of mse_target: two 16-byte loads per position in the residual's own layout, 2 bytes per element instead of 4.  The target is
     * then the image rounded to bf16. */
    mrssm_tv mse_target_s2d;
} mrssm_pl_conv_args;

int mrssm_pl_conv_down(const mrssm_pl_conv_args* a, void* stream);
int mrssm_pl_conv_up(const mrssm_pl_conv_args* a, void* stream);
int mrssm_pl_conv_wgrad(const mrssm_pl_conv_args* a, void* stream);
int mrssm_pl_packed_shape(int32_t op, int32_t Cs_pad, int32_t Cl_pad, int32_t ksz, int32_t* N_total, int32_t* K_total);
int mrssm_pl_pack_weight(const float* w, int64_t w_ss, int64_t w_sl, int32_t Cs_valid, int32_t Cl_valid, int32_t Cs_pad,
                         int32_t Cl_pad, int32_t ksz, int32_t op /* 0 down, 1 up, 2 down over a space-to-depth source */,
                         int32_t s2d_cq, void* out, void* stream);
/* fp32 strided [n,H,W,C] -> bf16 view with channels padded to Cpad (x scale); image_processing / autograd glue */
int mrssm_pl_import(const mrssm_t4* src, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t Cpad, float scale,
                    const mrssm_tv* dst, void* stream);
/* out[c] += sum over (img,y,x) of the view: bias gradients (autograd of encoder.py:315-322, observation_model.py:65-74) */
/* bf16 view -> bf16 view copy of the same logical [n,H,W,Cpad] tensor (layout change, e.g. parity-planar -> NHWC) */
int mrssm_pl_copy(const mrssm_tv* src, int32_t n_img, int32_t H, int32_t W, int32_t Cpad, const mrssm_tv* dst, void* stream);
/* fp32 strided [n,H,W,C<=4] -> the space-to-depth view described at mrssm_pl_conv_args.s2d_cq */
int mrssm_pl_import_s2d(const mrssm_t4* src, int32_t n_img, int32_t H, int32_t W, int32_t C, float scale, const mrssm_tv* dst,
                        void* stream);
/* fold > 0: the view is a space-to-depth view of a fold-channel tensor (H, W = the view's own size) */
int mrssm_pl_colsum(const mrssm_tv* x, int32_t n_img, int32_t H, int32_t W, int32_t Cpad, int32_t Cvalid, int32_t fold,
                    const float* scale_ptr, float scale_mul, float* out, void* stream);
/* host only, no GPU: format the tiling plan of a layer (op 0 down, 1 up, 2 wgrad) */
int mrssm_pl_describe(const mrssm_pl_conv_args* a, int32_t op, char* buf, int32_t buflen);
/* bring-up switches (descriptor-field variants); 0 = production setting */
int mrssm_pl_set_debug(int32_t key, int32_t value);
/* tuning aid: force the forward-type planner's tiling (images per tile, rows per band, activation stages, resident weights
 * 0/1, weight-ring slots); 0 / -1 = planner's choice.  Production plans come from the built-in measured table. */
int mrssm_pl_set_plan_override(int32_t BI, int32_t TH, int32_t NA, int32_t bres, int32_t NB);
/* Persistent plane kernels launch one CTA per SM; n_sm (default 148) caps their grids so that they can share the GPU with a kernel
 * that occupies the other SMs (the weight gradients of the decoder beside the 32-CTA rollout BPTT). Host-side, sticky. */
int mrssm_pl_set_sm_budget(int32_t n_sm);
/* tuning aid: device buffer of 148*16*8 int64 receiving per-tile clock64 stamps of the fwd-type kernel (NULL = off) */
int mrssm_pl_set_profile_buffer(void* dev_buf);

/* ---- the RSSM rollout ------------------------------------------------------------------------
 * Replaces MultimodalTransitionModel.forward (utils/models/transition_model.py:200-285), its
 * single-modal twin TransitionModel.forward (:50-114), nn.GRUCell (:160,235), the prior head and
 * the expert heads (encoder.py:126-155,157-190,196-224), poe/get_poe_state/get_mopoe_state
 * (encoder.py:50-124) and the rsample calls, as ONE launch for all T steps.  n_experts = 0 is the
 * open-loop imagination mode (transition_model.py:226-245 with observations=None).
 *
 * Layout: every [T,B,x] tensor is time-major contiguous.  Head 0 is the transition prior
 * (stochastic_state_model); heads 1..n_experts are the experts in dict order (prior_expert first
 * for the multimodal model).  Weights arrive pre-transposed ([in][out]) for the forward and in
 * PyTorch layout ([out][in]) for the backward; w1 covers only the belief columns of fc1 — the
 * embedding columns are hoisted into emb_pre (time-parallel GEMM, bias folded in).
 */
typedef struct mrssm_rollout_args {
    int32_t T, B, D, S, H, A, n_experts;
    int32_t act, det;
    float min_std;
    /* inputs */
    const float *prev_state, *prev_belief, *actions, *nonterminals, *eps_prior, *eps_post;
    const float* emb_pre[MRSSM_MAX_HEADS];      /* [T,B,H] per head or NULL (index = head) */
    /* weights: fwd uses *_t ([in][out]); bwd uses PyTorch layout */
    const float *w_sa, *b_sa;                   /* [S+A][D] (fwd) / [D][S+A] (bwd) */
    const float *w_ih, *b_ih, *w_hh, *b_hh;     /* [D][3D] (fwd) / [3D][D] (bwd) */
    const float* w1[MRSSM_MAX_HEADS];           /* [D][H] (fwd) / [H][ld1] (bwd) */
    int64_t ld1[MRSSM_MAX_HEADS];               /* bwd: row stride of w1 (= D or D+E_m) */
    const float* b1[MRSSM_MAX_HEADS];           /* NULL when folded into emb_pre */
    const float* w2[MRSSM_MAX_HEADS];           /* [H][2S] (fwd) / [2S][H] (bwd) */
    const float* b2[MRSSM_MAX_HEADS];
    /* fusion table (encoder.py:73-124): subset j = bitmask over experts (bit e-1 = head e) */
    int32_t n_subsets;
    uint32_t subset_mask[MRSSM_MAX_SUBSETS];
    uint8_t dim_subset[MRSSM_MAX_STATE];        /* which subset supplies state dim s */
    /* outputs [T,B,*] */
    float *beliefs, *prior_states, *prior_means, *prior_stds;
    float *post_states, *post_means, *post_stds;
    float* exp_means[MRSSM_MAX_HEADS];          /* index = head (1..n_experts) */
    float* exp_stds[MRSSM_MAX_HEADS];
    /* stash for BPTT (all NULL for inference): x,r,z,n,ghn [T,B,D]; u[head] [T,B,H] */
    float *st_x, *st_r, *st_z, *st_n, *st_ghn;
    float* st_u[MRSSM_MAX_HEADS];
} mrssm_rollout_args;

int mrssm_rollout_fwd(const mrssm_rollout_args* a, void* stream);

/* ---- the same rollout (forward) on the tensor cores -------------------------------------------------
 * tcgen05 / TMEM version of mrssm_rollout_fwd for D, H <= 208, S <= 32, S + A <= 48, <= 4 heads (the shipped
 * configs): 64 sequences per CTA, bf16 GEMM operands, fp32 accumulation, gate math and recurrent state.  Same
 * reference code replaced (transition_model.py:200-285, encoder.py:50-190), same argument struct, same outputs
 * and stash.  The weights are NOT read through the struct: they are packed once per optimiser step into the bf16
 * B-operand stream of one time step by mrssm_rollout_tc_pack (which takes them in PyTorch layout, [out][in] fp32,
 * w1 with row stride ld1), following the per-step MMA program built by mrssm_rollout_tc_plan (host only; the
 * caller uploads the buffer once per shape).  Only the biases, emb_pre, inputs and outputs are read from `a`. */
int mrssm_rollout_tc_eligible(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts);   /* 1 / 0 */
int mrssm_rollout_tc_plan_bytes(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts, int64_t* plan_bytes,
                                int64_t* packed_bytes);
int mrssm_rollout_tc_plan(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts, void* host_buf, int64_t buflen);
int mrssm_rollout_tc_pack(const mrssm_rollout_args* a, const void* plan_dev, int32_t n_pack, void* packed_dev, void* stream);
int mrssm_rollout_tc_fwd(const mrssm_rollout_args* a, const void* plan_dev, const void* packed_dev, void* stream);
/* tuning aids: device buffer of 4*512 int64 receiving clock64 stamps of one time step of CTA 0 (NULL = off); sequences per
 * 16-row MMA group (16 -> 64 sequences per CTA, 8 -> 32, 0 = automatic) */
int mrssm_rollout_tc_set_profile_buffer(void* dev_buf);
int mrssm_rollout_tc_set_rows(int32_t rows_per_group);
/* 0: force the table-driven MMA issue even for the sizes with a statically unrolled program (D=H=200, S=30, A=3; 1, 2, 4 heads) */
int mrssm_rollout_tc_set_static(int32_t on);

/* BPTT through the rollout (autograd of transition_model.py:226-270).  Consumes the forward's
 * outputs/stash plus upstream gradients of every output; produces the data gradients and the
 * per-step pre-activation gradients from which all weight gradients follow as time-parallel
 * mrssm_conv_wgrad calls (deferred wgrad). */
typedef struct mrssm_rollout_bwd_args {
    mrssm_rollout_args f;                       /* same shapes/weights (PyTorch layout)/outputs/stash */
    /* upstream grads, [T,B,*], any may be NULL (= zero) */
    const float *g_beliefs, *g_prior_states, *g_prior_means, *g_prior_stds;
    const float *g_post_states, *g_post_means, *g_post_stds;
    const float* g_exp_means[MRSSM_MAX_HEADS];
    const float* g_exp_stds[MRSSM_MAX_HEADS];
    /* outputs */
    float *g_prev_state, *g_prev_belief;        /* [B,S], [B,D] */
    float* g_actions;                           /* [T,B,A] */
    float *d_xpre;                              /* [T,B,D]  grad wrt fc_embed pre-activation */
    float *d_gi, *d_gh;                         /* [T,B,3D] grad wrt W_ih x+b_ih and W_hh h+b_hh */
    float* d_u[MRSSM_MAX_HEADS];                /* [T,B,H]  grad wrt fc1 pre-activation (also = grad of emb_pre) */
    float* d_o[MRSSM_MAX_HEADS];                /* [T,B,2S] grad wrt fc2 output */
    float* xin;                                 /* [T,B,S+A] the masked [state,action] input, recomputed */
    /* optional: CONTIGUOUS [H][D] copies of the belief columns of every head's fc1 (f.w1[h] with row stride D).  When all
     * are given the weight-streaming kernel runs (weights staged through shared memory by cp.async.bulk). */
    const float* w1_belief[MRSSM_MAX_HEADS];
} mrssm_rollout_bwd_args;

int mrssm_rollout_bwd(const mrssm_rollout_bwd_args* a, void* stream);

/* BPTT on the tensor cores (tcgen05 twin of mrssm_rollout_bwd, same eligibility as mrssm_rollout_tc_fwd, same argument struct
 * and outputs).  The weights are read from the bf16 stream packed by mrssm_rollout_tc_pack with the plan of
 * mrssm_rollout_tc_bwd_plan (the transposed, dgrad-type blocks of fc2, fc1[:, :D], W_ih, W_hh, W_sa), not through the struct. */
int mrssm_rollout_tc_bwd_plan_bytes(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts, int64_t* plan_bytes,
                                    int64_t* packed_bytes);
int mrssm_rollout_tc_bwd_plan(int32_t D, int32_t S, int32_t H, int32_t A, int32_t n_experts, void* host_buf, int64_t buflen);
int mrssm_rollout_tc_bwd(const mrssm_rollout_bwd_args* g, const void* packed_dev, void* stream);

/* ---- large-model rollout, one time step per call (BASELINE config 5: D = H = 1024 — beyond mrssm_rollout_tc_eligible) ---------
 * The dense contractions of a step (transition_model.py:226-270: fc_embed_state_action, nn.GRUCell's two projections, every
 * head's fc1 / fc2, and their dgrads) run as mrssm_tc_conv_down / up GEMMs over all B sequences; these entries do the work
 * between them on the same [T,B,*] tensors and stash as mrssm_rollout_fwd / bwd (same structs), for time step t:
 *   xin      : bf16 [B,KX] = [s_{t-1} * nonterminal_t, a_t, 0..]            (transition_model.py:228-232)
 *   gate_fwd : gi, gh fp32 [B,3D] (biases included) -> h_t (beliefs[t], stash r z n W_hn h) and its bf16 copy   (nn.GRUCell)
 *   heads_fwd: fc2 outputs of all heads fp32 (row stride ldo, head hd at column hd * head_stride) -> prior / expert / posterior
 *              statistics and samples at t
 *   heads_bwd / gate_bwd / xin_bwd: the matching backward pieces (cgs: state-gradient carry [B,S]; carry_a / carry_b: the two
 *   parts of the belief-gradient carry [B,D]); d_o[hd]: bf16 rows of S2p columns (row stride ld), gradient of head hd's fc2
 *   output, padding columns zeroed. */
int mrssm_rstep_xin(const mrssm_rollout_args* a, int32_t t, int32_t KX, void* xin_b, void* stream);
int mrssm_rstep_gate_fwd(const mrssm_rollout_args* a, int32_t t, const float* gi, const float* gh, void* hb_out, void* stream);
/* xin_next (bf16 [B,KX], may be NULL): also write step t + 1's [s_t * nonterminal_{t+1}, a_{t+1}, 0..] operand (saves the xin launch) */
int mrssm_rstep_heads_fwd(const mrssm_rollout_args* a, int32_t t, const float* o, int32_t ldo, int32_t head_stride, void* xin_next, int32_t KX,
                          void* stream);
/* dxin_next (fp32 [B,S+A], may be NULL): step t + 1's dxin; replaces cgs as the carry and yields g_actions[t + 1] (xin_bwd folded in) */
int mrssm_rstep_heads_bwd(const mrssm_rollout_bwd_args* g, int32_t t, const float* cgs, const float* dxin_next, void* const* d_o, int32_t ld,
                          int32_t S2p, void* stream);
int mrssm_rstep_gate_bwd(const mrssm_rollout_bwd_args* g, int32_t t, const float* dh_heads, float* carry_a, const float* carry_b,
                         void* dgi, void* dgh, void* stream);
int mrssm_rstep_xin_bwd(const mrssm_rollout_bwd_args* g, int32_t t, const float* dxin, int32_t ld, float* cgs, void* stream);
int mrssm_add2(const float* x, const float* y, int64_t n, float* out, void* stream);

/* All T steps of the large-model rollout in one call (the per-step entries above, issued from C so that a train step does not
 * pay ~1000 Python-level launches; same kernels, same order).  Weights are the bf16 packings of mrssm_tc_pack_weight (forward:
 * mode 0, backward: mode 1); heads are stacked: fc1 over the belief columns in n_chunks chunks of heads [chunk_c0, chunk_c1)
 * (<= 4096 output columns each), fc2 / its dgrad as block-diagonal GEMMs.  Workspace tensors are caller-allocated:
 *   xin_all bf16 [T,B,KX], x_all bf16 [T or 1,B,D], u_cat bf16 [T or 1,B,NH*H] (keep_all selects T), hb_all bf16 [T+1,B,D],
 *   gi / gh fp32 [B,3D], o_cat fp32 [B,NH*S2p]; backward: d_o bf16 [T,B,NH*S2p], du_all bf16 [T,B,NH*H], d_gi / d_gh bf16 [T,B,3D],
 *   d_xpre bf16 [T,B,D], dh_heads / carry_a / carry_b fp32 [B,D] (carries zeroed by the caller), dxin fp32 [B,S+A], cgs fp32 [B,S]
 *   (zeroed by the caller; holds the gradient of prev_state on return), g_prev_belief fp32 [B,D]. */
typedef struct mrssm_rstep_ws {
    int32_t KX, S2p, NH, n_chunks, keep_all, KXo;
    int32_t chunk_c0[MRSSM_MAX_HEADS], chunk_c1[MRSSM_MAX_HEADS];
    const void* wp_sa; const void* wp_ih; const void* wp_hh; const void* w2f;
    const void* w1f[MRSSM_MAX_HEADS];
    const float* b1[MRSSM_MAX_HEADS];
    const float* b2;
    const float* pre_cat;
    const void* wp_sa_b; const void* wp_ih_b; const void* wp_hh_b; const void* w1b;
    const void* w2b[MRSSM_MAX_HEADS];
    void* xin_all; void* x_all; void* u_cat; void* hb_all;
    float* gi; float* gh; float* o_cat;
    void* d_o; void* du_all; void* d_gi; void* d_gh; void* d_xpre;
    float* dh_heads; float* carry_a; float* carry_b; float* dxin; float* cgs; float* g_prev_belief;
} mrssm_rstep_ws;
/* (the W_hh GEMM of a step is off the step's dependency chain and runs on a second, internal stream; 0 switches that off) */
int mrssm_rstep_set_aux_stream(int32_t on);
int mrssm_rollout_steps_fwd(const mrssm_rollout_args* a, const mrssm_rstep_ws* w, void* stream);
int mrssm_rollout_steps_bwd(const mrssm_rollout_bwd_args* g, const mrssm_rstep_ws* w, void* stream);

/* ---- the shipped YAML's remaining layers: exact fp32 kernels on NCHW tensors --------------------------------
 * General 2-d convolution without bias — rectangular kernel, any stride / zero padding: nn.Conv2d of SoundEncoder_v2 /
 * SoundDecoder_v2.out (encoder.py:661-721, observation_model.py:420-472), nn.Conv1d k1 (H = L, W = 1), and the bias-free
 * Conv2d / ConvTranspose2d of the BatchNorm image stacks (encoder.py:324-337, observation_model.py:75-86).
 *   x [N,Cin,H,W], w [Cout,Cin,KH,KW], y [N,Cout,Ho,Wo], Ho = (H + 2 PH - KH) / SH + 1.
 *   fwd:   y  = conv(x, w)                     (reads x, w; writes y)
 *   dgrad: dx = conv_transpose(y, w)           (reads y as the output gradient, w; writes dx)
 *   wgrad: dw += correlation(x, y)             (reads x and y as the output gradient; ACCUMULATES into dw)
 * nn.ConvTranspose2d(weight [Cin_T, Cout_T, KH, KW]) is the same geometry with the roles swapped: forward = dgrad with
 * y := its input, dx := its output; its input gradient = fwd; its weight gradient = wgrad with x := the output gradient. */
typedef struct mrssm_gconv_args {
    int32_t N, Cin, H, W, Cout, KH, KW, SH, SW, PH, PW, Ho, Wo;
    const float* x;
    const float* w;
    float* y;
    float* dx;
    float* dw;
} mrssm_gconv_args;
int mrssm_gconv_fwd(const mrssm_gconv_args* a, void* stream);
int mrssm_gconv_dgrad(const mrssm_gconv_args* a, void* stream);
int mrssm_gconv_wgrad(const mrssm_gconv_args* a, void* stream);

/* Staging kernels of the bf16 tensor-core route of the same convolutions (bf16 mode, channels in multiples of 8; ops.GConvTCFn):
 * NCHW fp32 -> NHWC bf16, explicit im2col rows col[(n,ho,wo)][(kh,kw,ci)] (and the matching col2im gather, fp32 NHWC out), NHWC fp32 ->
 * NCHW fp32, and the weight / weight-gradient permutation between [Cout][Cin][KH*KW] and the GEMM's tap-major [Cout][(tap, ci)].
 * The GEMMs themselves are mrssm_tc_conv_down / up / wgrad on the rows. */
int mrssm_nchw_to_nhwc_bf16(const float* x, int64_t N, int32_t C, int32_t HW, void* out, void* stream);
int mrssm_nhwc_to_nchw_f32(const float* x, int64_t N, int32_t C, int32_t HW, float* out, void* stream);
int mrssm_im2col_nhwc(const mrssm_gconv_args* a, const void* x_nhwc, void* col, void* stream);
int mrssm_col2im_nhwc(const mrssm_gconv_args* a, const void* dcol, float* dx_nhwc, void* stream);
int mrssm_gconv_weight_perm(const float* w, int64_t Cout, int32_t Cin, int32_t KHW, float* w2, void* stream);
int mrssm_gconv_weight_perm_add(const float* dw2, int64_t Cout, int32_t Cin, int32_t KHW, float* grad, void* stream);

/* nn.BatchNorm2d (instance = 0: one mean / biased variance per channel over (n, h, w)) and nn.InstanceNorm2d / 1d (instance = 1: per
 * (n, c) plane) with affine parameters, x / y [N,C,HW] fp32, optional fused ReLU (the Conv - BatchNorm - ReLU triples).
 *   batch_stats = 1 (train mode; InstanceNorm without tracked statistics always): statistics of this batch are written to
 *     mean / var ([C] or [N*C]) and, when running_mean / running_var are given, those move by `momentum` towards the batch mean /
 *     UNBIASED variance (InstanceNorm: averaged over the N planes) — torch.nn.functional.batch_norm / instance_norm semantics.
 *   batch_stats = 0 (eval mode): normalise with running_mean / running_var.
 * bwd: g = gradient of y -> dx; sum_g / sum_gx: scratch of one float per group; dgamma / dbeta ([C], may be NULL) are ACCUMULATED. */
typedef struct mrssm_norm_args {
    int32_t N, C, HW, instance, batch_stats, relu;
    float eps, momentum;
    const float* x;
    float* y;
    const float* gamma;
    const float* beta;
    float* mean;
    float* var;
    float* running_mean;
    float* running_var;
} mrssm_norm_args;
int mrssm_norm_fwd(const mrssm_norm_args* a, void* stream);
int mrssm_norm_bwd(const mrssm_norm_args* a, const float* g, float* sum_g, float* sum_gx, float* dgamma, float* dbeta, float* dx, void* stream);

/* nn.GLU(dim=1): x [N,2C,L] -> y [N,C,L] = x[:, :C] * sigmoid(x[:, C:]) and its backward (encoder.py:672-700). */
int mrssm_glu_fwd(const float* x, int64_t N, int32_t C, int32_t L, float* y, void* stream);
int mrssm_glu_bwd(const float* x, const float* g, int64_t N, int32_t C, int32_t L, float* dx, void* stream);
/* per-channel bias of an NCHW tensor with an optional ReLU (the biased Conv2d / ConvTranspose2d + ReLU pairs of the 84x84 and
 * 256x256 image stacks, encoder.py:362-413,511-615; the last ConvTranspose2d of the BatchNorm decoders, observation_model.py:85) and
 * the bias gradient, ACCUMULATED into dbias [C]. */
int mrssm_chan_bias_fwd(const float* x, int64_t N, int32_t C, int32_t HW, const float* bias, int32_t relu, float* y, void* stream);
int mrssm_chan_bias_bwd(const float* g, int64_t N, int32_t C, int32_t HW, float* dbias, void* stream);

/* ---- latent part of the ELBO -------------------------------------------------------------------
 * Replaces _get_posterior_states (MRSSM_PoE/algo.py:63-68, MRSSM_MoPoE/algo.py:62-67, base/algo.py:
 * 157-163), _calc_kl (base/algo.py:75-94), _calc_mopoe_kl (MRSSM_MoPoE/algo.py:110-137) and the
 * global KL (base/algo.py:186-188).  rows = (T-1)*B.
 *   kl_mode 0: balanced KL on (post_means,post_stds)   [RSSM, NN, PoE]
 *   kl_mode 1: MoPoE subset-averaged KL on the experts [MoPoE]
 *   refuse   : recompute the fused posterior from the experts and draw z = mu + sigma*eps_dec
 *              (PoE/MoPoE); otherwise z/q are the rollout's own posterior tensors.
 * out_sums[0] = kl_loss, out_sums[1] = global KL term (un-weighted), both already averaged. */
typedef struct mrssm_latent_args {
    int32_t rows, S, n_experts, kl_mode, refuse;
    int32_t n_subsets;
    uint32_t subset_mask[MRSSM_MAX_SUBSETS];
    uint8_t dim_subset[MRSSM_MAX_STATE];
    float free_nats, alpha;                     /* alpha < 0: no balancing */
    const float *prior_means, *prior_stds, *post_means, *post_stds, *eps_dec;
    const float* exp_means[MRSSM_MAX_HEADS];    /* index 1..n_experts */
    const float* exp_stds[MRSSM_MAX_HEADS];
    float *z_dec, *q_means, *q_stds;            /* [rows,S] outputs when refuse */
    float* row_scratch;                         /* [2*rows] */
    float* out_sums;                            /* [2] */
    /* backward only */
    const float *g_sums;                        /* [2] upstream grads of out_sums (device) */
    const float *g_z;                           /* [rows,S] or NULL */
    float *g_prior_means, *g_prior_stds, *g_post_means, *g_post_stds;
    float* g_exp_means[MRSSM_MAX_HEADS];
    float* g_exp_stds[MRSSM_MAX_HEADS];
} mrssm_latent_args;

int mrssm_latent_fwd(const mrssm_latent_args* a, void* stream);
int mrssm_latent_bwd(const mrssm_latent_args* a, void* stream);

/* ---- latent overshooting ---------------------------------------------------------------------------
 * Replaces _latent_overshooting (base/algo.py:111-148; MoPoE override MRSSM_MoPoE/algo.py:69-108).  One open-loop run of
 * up to OD steps per start step t = 1..T-2, all runs side by side: N = (T-2)*B columns, column n = (t-1)*B + b, row
 * (k, n) of every [OD, N, .] tensor belongs to source time t+k and is padding when t+k >= T-1.
 *   gather : actions_o / nonterminals_o / rewards_o / mask_o  <- the zero-padded slices the reference builds with
 *            F.pad + torch.cat (base/algo.py:124-130); rewards_o and mask_o are optional.
 *   kl_fwd : out[0] = scale * mean_rows max(sum_S KL(q || prior) * mask, free_nats)   (base/algo.py:138-141), q = the
 *            DETACHED posterior of step t+k: post_means/post_stds [T-1,B,S] when n_experts == 0, else the product of the
 *            experts `subset_mask` selects (bit e-1 = expert e; encoder.py:50-71).  row_scratch keeps the clamped row values.
 *   kl_bwd : g_prior_means / g_prior_stds [OD,N,S] from g_out[0] and the forward's row_scratch. */
typedef struct mrssm_overshoot_args {
    int32_t T, B, S, A, OD;                     /* T = chunk length (T-1 model steps) */
    int32_t n_experts;
    uint32_t subset_mask;
    float free_nats, scale;
    const float *prior_means, *prior_stds;      /* [OD, N, S] */
    const float *post_means, *post_stds;        /* [T-1, B, S] */
    const float* exp_means[MRSSM_MAX_HEADS];    /* [T-1, B, S], index 1..n_experts */
    const float* exp_stds[MRSSM_MAX_HEADS];
    float* row_scratch;                         /* [OD*N] */
    float* out;                                 /* [1] */
    const float* g_out;                         /* [1] (device) */
    float *g_prior_means, *g_prior_stds;        /* [OD, N, S] */
    const float *actions, *nonterminals, *rewards;              /* [T,B,A], [T,B], [T,B] (rewards may be NULL) */
    float *actions_o, *nonterminals_o, *rewards_o, *mask_o;     /* [OD,N,A], [OD,N] x3 */
} mrssm_overshoot_args;

int mrssm_overshoot_gather(const mrssm_overshoot_args* a, void* stream);
int mrssm_overshoot_kl_fwd(const mrssm_overshoot_args* a, void* stream);
int mrssm_overshoot_kl_bwd(const mrssm_overshoot_args* a, void* stream);

/* ---- reconstruction loss: sum_features mean_{t,b} (y-o)^2  (observation_model.py:28-31,
 * base/algo.py:381-383).  n = element count, rows = (T-1)*B.  fwd writes *out; bwd writes
 * dy = (*g) * 2 (y-o) / rows. */
int mrssm_mse_fwd(const float* y, const float* o, int64_t n, int64_t rows, float* partial, float* out, void* stream);
int mrssm_mse_bwd(const float* y, const float* o, int64_t n, int64_t rows, const float* g, float* dy, void* stream);
/* elementwise (y-o)^2 for the get_mse API (observation_model.py:28-31) */
int mrssm_sqdiff(const float* y, const float* o, int64_t n, float* out, void* stream);

/* ---- optimiser: clip_grad_norm_(max_norm, L2) + Adam, no host sync (base/algo.py:41-42,258-259).
 * One flat parameter/gradient/moment buffer.  norm_out[0] receives the pre-clip total norm. */
int mrssm_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr,
                    float beta1, float beta2, float eps, float max_norm, float grad_scale,
                    float* partial, float* norm_out, void* stream);

/* ---- replay-buffer sampling on the device -------------------------------------------------------------
 * Replaces ExperienceReplay_Multimodal._retrieve_batch for image observations (utils/replay_buffer/memory.py:191-208):
 * gather by index -> crop (data_augment.py:158-174) -> + PCA colour shift + Gaussian noise, clip to [0,255]
 * (data_augment.py:178-210) -> quantise / dequantise (utils/processing/image_processing.py:5-11), one pass:
 *   x   = frames[idx[row], c, y+dh, x+dw]
 *   x   = clip(x + delta[c] + gauss * gauss_scale * 255, 0, 255)        (only when delta / gauss / gauss_scale > 0 is given)
 *   out = floor(x / 2^(8-bits)) / 2^bits - 0.5 + u / 2^bits
 * gauss ~ N(0,1) and u ~ U[0,1) come from the caller's tensors when given (parity tests), else from a counter-based hash of
 * (seed, element index).  idx holds rows = L*n frame slots in the reference's order (row = l*n + b).  bit_depth 0 skips
 * the last line (binary mask images are only gathered and cropped, memory.py:199-201). */
typedef struct mrssm_replay_gather_args {
    const uint8_t* frames;                      /* [size, C, Hs, Ws] device */
    const int64_t* idx;                         /* [rows] device */
    int64_t rows;
    int32_t C, Hs, Ws, H, W, dh, dw, bit_depth; /* output [rows, C, H, W], crop origin (dh, dw) */
    const float* delta;                         /* [C] device or NULL */
    const float* gauss;                         /* [rows, C, H, W] or NULL */
    float gauss_scale;
    const float* uniform;                       /* [rows, C, H, W] or NULL */
    uint64_t seed;
    float* out;
} mrssm_replay_gather_args;

int mrssm_replay_gather_u8(const mrssm_replay_gather_args* a, void* stream);
/* fp32 rows (vector observations, actions, rewards, nonterminals; memory.py:193-196,210-212): out[r,:] = src[idx[r],:] */
int mrssm_gather_rows(const float* src, const int64_t* idx, int64_t rows, int32_t K, float* out, void* stream);

/* ---- input pipeline: replay-buffer frames are uint8 on the host (utils/replay_buffer/memory.py:160-168); sample()
 * moves them to the device and normalises there (memory.py:197-208 -> utils/processing/image_processing.py:5-11):
 *   dst = floor(u8 / 2^(8-bits)) / 2^bits - 0.5 + u / 2^bits,  u ~ U[0,1).
 * `noise` supplies u (parity tests); NULL draws it from a counter-based hash of (seed, element index). */
int mrssm_normalize_image_u8(const uint8_t* src, int64_t n, int32_t bit_depth, const float* noise, uint64_t seed,
                             float* dst, void* stream);

/* ---- small helpers ---- */
int mrssm_transpose(const float* src, int64_t rows, int64_t cols, int64_t src_ld, float* dst, void* stream);
/* dst[r][c] = src[r*ld + c]: contiguous copy of a column block */
int mrssm_copy2d(const float* src, int64_t rows, int64_t cols, int64_t ld, float* dst, void* stream);
int mrssm_concat2(const float* a, int64_t ca, const float* b, int64_t cb, int64_t rows, float* out, void* stream);
int mrssm_colsum_acc(const float* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream);
int mrssm_fill(float* p, int64_t n, float v, void* stream);
/* out = g * act'(y) with the derivative expressed through the activation OUTPUT y */
int mrssm_act_bwd(const float* g, const float* y, int64_t n, int32_t act, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MRSSM_B200_H */
