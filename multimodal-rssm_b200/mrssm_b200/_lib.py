"""ctypes binding of lib/libmrssm_b200.so (include/mrssm_b200.h).

There is no CPU or eager-PyTorch fallback: if the shared library is missing, or the device is not
an sm_100 part, every op raises.  The library is built in-tree by ``__graft_entry__.build()``
(``make -C multimodal-rssm_b200/csrc``).
"""
import ctypes as C
import os

import torch

MAX_HEADS = 5
MAX_SUBSETS = 8
MAX_STATE = 256
ACT = {None: 0, "none": 0, "relu": 1, "elu": 2}
F32, BF16 = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libmrssm_b200.so")

_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p


class T4(C.Structure):
    _fields_ = [("ptr", _vp), ("sI", C.c_int64), ("sH", C.c_int64), ("sW", C.c_int64), ("sC", C.c_int64)]


class ConvArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_img", "Hl", "Wl", "Cl", "Hs", "Ws", "Cs", "ksz", "dtype", "act",
                                         "mask_mode", "accumulate")] + [
        ("large", T4), ("small", T4), ("weight", _vp), ("w_ss", C.c_int64), ("w_sl", C.c_int64),
        ("bias", _vp), ("mask", _vp)]


class TcConvArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_img", "Hl", "Wl", "Cl", "Hs", "Ws", "Cs", "ksz", "act", "mask_mode",
                                         "out_f32", "n_out_pad", "n_out_valid", "bias_mod", "cs_valid", "cl_valid")] + [
        ("large", T4), ("small", T4), ("mask", T4), ("wpacked", _vp), ("bias", _vp), ("dweight", _vp),
        ("w_ss", C.c_int64), ("w_sl", C.c_int64), ("addend", _vp), ("addend_ld", C.c_int64), ("group_n", C.c_int32), ("group_k", C.c_int32)]


class TV(C.Structure):
    """bf16 activation view with channels in chunks of 8 (mrssm_tv)."""
    _fields_ = [("ptr", _vp), ("sI", C.c_int64), ("sH", C.c_int64), ("sW", C.c_int64), ("sK", C.c_int64), ("sP", C.c_int64),
                ("par", C.c_int32), ("reserved", C.c_int32)]


class PlConvArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_img", "Hl", "Wl", "Cl", "Hs", "Ws", "Cs", "ksz", "act", "mask_mode", "out_f32",
                                         "n_out_pad", "n_out_valid", "cs_valid", "cl_valid", "s2d_cq")] + [
        ("large", TV), ("small", TV), ("mask", TV), ("out32", T4), ("wpacked", _vp), ("bias", _vp), ("dweight", _vp),
        ("w_ss", C.c_int64), ("w_sl", C.c_int64), ("scale_ptr", _vp), ("scale_mul", C.c_float),
        ("mse_target", _vp), ("mse_sum", _vp), ("mse_scale", C.c_float),
        ("dbias", _vp), ("dbias_from", C.c_int32), ("relu_bits_out", _vp), ("relu_bits_in", _vp), ("mse_target_s2d", TV)]


class RolloutArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("T", "B", "D", "S", "H", "A", "n_experts", "act", "det")] + [
        ("min_std", C.c_float)] + [(n, _vp) for n in ("prev_state", "prev_belief", "actions", "nonterminals",
                                                      "eps_prior", "eps_post")] + [
        ("emb_pre", _vp * MAX_HEADS),
        ("w_sa", _vp), ("b_sa", _vp), ("w_ih", _vp), ("b_ih", _vp), ("w_hh", _vp), ("b_hh", _vp),
        ("w1", _vp * MAX_HEADS), ("ld1", C.c_int64 * MAX_HEADS), ("b1", _vp * MAX_HEADS),
        ("w2", _vp * MAX_HEADS), ("b2", _vp * MAX_HEADS),
        ("n_subsets", C.c_int32), ("subset_mask", C.c_uint32 * MAX_SUBSETS), ("dim_subset", C.c_uint8 * MAX_STATE),
        ("beliefs", _vp), ("prior_states", _vp), ("prior_means", _vp), ("prior_stds", _vp),
        ("post_states", _vp), ("post_means", _vp), ("post_stds", _vp),
        ("exp_means", _vp * MAX_HEADS), ("exp_stds", _vp * MAX_HEADS),
        ("st_x", _vp), ("st_r", _vp), ("st_z", _vp), ("st_n", _vp), ("st_ghn", _vp),
        ("st_u", _vp * MAX_HEADS)]


class RstepWs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("KX", "S2p", "NH", "n_chunks", "keep_all", "KXo")] + [
        ("chunk_c0", C.c_int32 * MAX_HEADS), ("chunk_c1", C.c_int32 * MAX_HEADS),
        ("wp_sa", _vp), ("wp_ih", _vp), ("wp_hh", _vp), ("w2f", _vp), ("w1f", _vp * MAX_HEADS), ("b1", _vp * MAX_HEADS), ("b2", _vp),
        ("pre_cat", _vp), ("wp_sa_b", _vp), ("wp_ih_b", _vp), ("wp_hh_b", _vp), ("w1b", _vp), ("w2b", _vp * MAX_HEADS)] + [
        (n, _vp) for n in ("xin_all", "x_all", "u_cat", "hb_all", "gi", "gh", "o_cat", "d_o", "du_all", "d_gi", "d_gh", "d_xpre",
                           "dh_heads", "carry_a", "carry_b", "dxin", "cgs", "g_prev_belief")]


class GConvArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("N", "Cin", "H", "W", "Cout", "KH", "KW", "SH", "SW", "PH", "PW", "Ho", "Wo")] + [
        ("x", _vp), ("w", _vp), ("y", _vp), ("dx", _vp), ("dw", _vp)]


class NormArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("N", "C", "HW", "instance", "batch_stats", "relu")] + [("eps", C.c_float), ("momentum", C.c_float)] + [
        (n, _vp) for n in ("x", "y", "gamma", "beta", "mean", "var", "running_mean", "running_var")]


class RolloutBwdArgs(C.Structure):
    _fields_ = [("f", RolloutArgs)] + [(n, _vp) for n in (
        "g_beliefs", "g_prior_states", "g_prior_means", "g_prior_stds", "g_post_states", "g_post_means",
        "g_post_stds")] + [
        ("g_exp_means", _vp * MAX_HEADS), ("g_exp_stds", _vp * MAX_HEADS),
        ("g_prev_state", _vp), ("g_prev_belief", _vp), ("g_actions", _vp),
        ("d_xpre", _vp), ("d_gi", _vp), ("d_gh", _vp),
        ("d_u", _vp * MAX_HEADS), ("d_o", _vp * MAX_HEADS), ("xin", _vp), ("w1_belief", _vp * MAX_HEADS)]


class LatentArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("rows", "S", "n_experts", "kl_mode", "refuse", "n_subsets")] + [
        ("subset_mask", C.c_uint32 * MAX_SUBSETS), ("dim_subset", C.c_uint8 * MAX_STATE),
        ("free_nats", C.c_float), ("alpha", C.c_float)] + [(n, _vp) for n in (
            "prior_means", "prior_stds", "post_means", "post_stds", "eps_dec")] + [
        ("exp_means", _vp * MAX_HEADS), ("exp_stds", _vp * MAX_HEADS),
        ("z_dec", _vp), ("q_means", _vp), ("q_stds", _vp), ("row_scratch", _vp), ("out_sums", _vp),
        ("g_sums", _vp), ("g_z", _vp),
        ("g_prior_means", _vp), ("g_prior_stds", _vp), ("g_post_means", _vp), ("g_post_stds", _vp),
        ("g_exp_means", _vp * MAX_HEADS), ("g_exp_stds", _vp * MAX_HEADS)]


class OvershootArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("T", "B", "S", "A", "OD", "n_experts")] + [
        ("subset_mask", C.c_uint32), ("free_nats", C.c_float), ("scale", C.c_float)] + [(n, _vp) for n in (
            "prior_means", "prior_stds", "post_means", "post_stds")] + [
        ("exp_means", _vp * MAX_HEADS), ("exp_stds", _vp * MAX_HEADS)] + [(n, _vp) for n in (
            "row_scratch", "out", "g_out", "g_prior_means", "g_prior_stds", "actions", "nonterminals", "rewards",
            "actions_o", "nonterminals_o", "rewards_o", "mask_o")]


class ReplayGatherArgs(C.Structure):
    _fields_ = [("frames", _vp), ("idx", _vp), ("rows", C.c_int64)] + [(n, C.c_int32) for n in (
        "C", "Hs", "Ws", "H", "W", "dh", "dw", "bit_depth")] + [
        ("delta", _vp), ("gauss", _vp), ("gauss_scale", C.c_float), ("uniform", _vp), ("seed", C.c_uint64), ("out", _vp)]


# every symbol include/mrssm_b200.h declares: name -> argtypes (restype is int unless noted)
_i64, _i32, _f = C.c_int64, C.c_int32, C.c_float
SYMBOLS = {
    "mrssm_last_error": None,
    "mrssm_abi_version": [],
    "mrssm_device_ok": [],
    "mrssm_conv_down": [C.POINTER(ConvArgs), _vp],
    "mrssm_conv_up": [C.POINTER(ConvArgs), _vp],
    "mrssm_conv_wgrad": [C.POINTER(ConvArgs), _vp],
    "mrssm_colsum_t4": [C.POINTER(ConvArgs), _vp],
    "mrssm_tc_conv_down": [C.POINTER(TcConvArgs), _vp],
    "mrssm_tc_conv_up": [C.POINTER(TcConvArgs), _vp],
    "mrssm_tc_conv_wgrad": [C.POINTER(TcConvArgs), _vp],
    "mrssm_tc_pack_weight": [_vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp],
    "mrssm_tc_to_bf16": [C.POINTER(T4), _i32, _i32, _i32, _i32, _i32, _f, _vp, _vp],
    "mrssm_tc_from_bf16": [_vp, _i32, _i32, _i32, _i32, _i32, C.POINTER(T4), _vp],
    "mrssm_tc_colsum": [_vp, _i64, _i32, _i32, _vp, _vp],
    "mrssm_pl_conv_down": [C.POINTER(PlConvArgs), _vp],
    "mrssm_pl_conv_up": [C.POINTER(PlConvArgs), _vp],
    "mrssm_pl_conv_wgrad": [C.POINTER(PlConvArgs), _vp],
    "mrssm_pl_import": [C.POINTER(T4), _i32, _i32, _i32, _i32, _i32, _f, C.POINTER(TV), _vp],
    "mrssm_pl_copy": [C.POINTER(TV), _i32, _i32, _i32, _i32, C.POINTER(TV), _vp],
    "mrssm_pl_colsum": [C.POINTER(TV), _i32, _i32, _i32, _i32, _i32, _i32, _vp, _f, _vp, _vp],
    "mrssm_pl_packed_shape": [_i32, _i32, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32)],
    "mrssm_pl_pack_weight": [_vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp],
    "mrssm_pl_import_s2d": [C.POINTER(T4), _i32, _i32, _i32, _i32, _f, C.POINTER(TV), _vp],
    "mrssm_pl_describe": [C.POINTER(PlConvArgs), _i32, C.c_char_p, _i32],
    "mrssm_pl_set_plan_override": [_i32, _i32, _i32, _i32, _i32],
    "mrssm_pl_set_sm_budget": [_i32],
    "mrssm_pl_set_debug": [_i32, _i32],
    "mrssm_pl_set_profile_buffer": [_vp],
    "mrssm_rollout_fwd": [C.POINTER(RolloutArgs), _vp],
    "mrssm_rollout_tc_eligible": [_i32, _i32, _i32, _i32, _i32],
    "mrssm_rollout_tc_plan_bytes": [_i32, _i32, _i32, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64)],
    "mrssm_rollout_tc_plan": [_i32, _i32, _i32, _i32, _i32, _vp, _i64],
    "mrssm_rollout_tc_pack": [C.POINTER(RolloutArgs), _vp, _i32, _vp, _vp],
    "mrssm_rollout_tc_fwd": [C.POINTER(RolloutArgs), _vp, _vp, _vp],
    "mrssm_rollout_tc_set_profile_buffer": [_vp],
    "mrssm_rollout_tc_set_rows": [_i32],
    "mrssm_rollout_tc_set_static": [_i32],
    "mrssm_rollout_tc_bwd_plan_bytes": [_i32, _i32, _i32, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64)],
    "mrssm_rollout_tc_bwd_plan": [_i32, _i32, _i32, _i32, _i32, _vp, _i64],
    "mrssm_rollout_tc_bwd": [C.POINTER(RolloutBwdArgs), _vp, _vp],
    "mrssm_rollout_bwd": [C.POINTER(RolloutBwdArgs), _vp],
    "mrssm_rstep_xin": [C.POINTER(RolloutArgs), _i32, _i32, _vp, _vp],
    "mrssm_rstep_gate_fwd": [C.POINTER(RolloutArgs), _i32, _vp, _vp, _vp, _vp],
    "mrssm_rstep_heads_fwd": [C.POINTER(RolloutArgs), _i32, _vp, _i32, _i32, _vp, _i32, _vp],
    "mrssm_rstep_heads_bwd": [C.POINTER(RolloutBwdArgs), _i32, _vp, _vp, C.POINTER(_vp), _i32, _i32, _vp],
    "mrssm_rstep_gate_bwd": [C.POINTER(RolloutBwdArgs), _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "mrssm_rstep_xin_bwd": [C.POINTER(RolloutBwdArgs), _i32, _vp, _i32, _vp, _vp],
    "mrssm_add2": [_vp, _vp, _i64, _vp, _vp],
    "mrssm_rstep_set_aux_stream": [_i32],
    "mrssm_rollout_steps_fwd": [C.POINTER(RolloutArgs), C.POINTER(RstepWs), _vp],
    "mrssm_rollout_steps_bwd": [C.POINTER(RolloutBwdArgs), C.POINTER(RstepWs), _vp],
    "mrssm_gconv_fwd": [C.POINTER(GConvArgs), _vp],
    "mrssm_gconv_dgrad": [C.POINTER(GConvArgs), _vp],
    "mrssm_gconv_wgrad": [C.POINTER(GConvArgs), _vp],
    "mrssm_nchw_to_nhwc_bf16": [_vp, _i64, _i32, _i32, _vp, _vp],
    "mrssm_nhwc_to_nchw_f32": [_vp, _i64, _i32, _i32, _vp, _vp],
    "mrssm_im2col_nhwc": [C.POINTER(GConvArgs), _vp, _vp, _vp],
    "mrssm_col2im_nhwc": [C.POINTER(GConvArgs), _vp, _vp, _vp],
    "mrssm_gconv_weight_perm": [_vp, _i64, _i32, _i32, _vp, _vp],
    "mrssm_gconv_weight_perm_add": [_vp, _i64, _i32, _i32, _vp, _vp],
    "mrssm_norm_fwd": [C.POINTER(NormArgs), _vp],
    "mrssm_norm_bwd": [C.POINTER(NormArgs), _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mrssm_glu_fwd": [_vp, _i64, _i32, _i32, _vp, _vp],
    "mrssm_glu_bwd": [_vp, _vp, _i64, _i32, _i32, _vp, _vp],
    "mrssm_chan_bias_fwd": [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp],
    "mrssm_chan_bias_bwd": [_vp, _i64, _i32, _i32, _vp, _vp],
    "mrssm_latent_fwd": [C.POINTER(LatentArgs), _vp],
    "mrssm_latent_bwd": [C.POINTER(LatentArgs), _vp],
    "mrssm_overshoot_gather": [C.POINTER(OvershootArgs), _vp],
    "mrssm_overshoot_kl_fwd": [C.POINTER(OvershootArgs), _vp],
    "mrssm_overshoot_kl_bwd": [C.POINTER(OvershootArgs), _vp],
    "mrssm_mse_fwd": [_vp, _vp, _i64, _i64, _vp, _vp, _vp],
    "mrssm_mse_bwd": [_vp, _vp, _i64, _i64, _vp, _vp, _vp],
    "mrssm_sqdiff": [_vp, _vp, _i64, _vp, _vp],
    "mrssm_clip_adam": [_vp, _vp, _vp, _vp, _i64, _i32, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp],
    "mrssm_replay_gather_u8": [C.POINTER(ReplayGatherArgs), _vp],
    "mrssm_gather_rows": [_vp, _vp, _i64, _i32, _vp, _vp],
    "mrssm_normalize_image_u8": [_vp, _i64, _i32, _vp, C.c_uint64, _vp, _vp],
    "mrssm_transpose": [_vp, _i64, _i64, _i64, _vp, _vp],
    "mrssm_copy2d": [_vp, _i64, _i64, _i64, _vp, _vp],
    "mrssm_concat2": [_vp, _i64, _vp, _i64, _i64, _vp, _vp],
    "mrssm_colsum_acc": [_vp, _i64, _i64, _i64, _vp, _vp],
    "mrssm_fill": [_vp, _i64, _f, _vp],
    "mrssm_act_bwd": [_vp, _vp, _i64, _i32, _vp, _vp],
}

_lib = None
launches = 0          # number of C-ABI compute calls issued (bench.py reports kernel launches from it)
kernel_launches = 0   # number of CUDA kernels those calls launched


def load():
    """dlopen the library (no GPU needed) and declare prototypes."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SYMBOLS.items():
            fn = getattr(lib, name)
            if argtypes is None:
                fn.restype = C.c_char_p
                fn.argtypes = []
            else:
                fn.restype = C.c_int
                fn.argtypes = argtypes
        _lib = lib
    return _lib


_dev_checked = False


def _require_device():
    global _dev_checked
    if not _dev_checked:
        lib = load()
        if not torch.cuda.is_available() or not lib.mrssm_device_ok():
            raise RuntimeError("mrssm_b200 needs an sm_100 (B200) CUDA device; there is no CPU fallback: "
                               + lib.mrssm_last_error().decode())
        _dev_checked = True


# kernels launched per C-ABI call (for the gpu_launches claim)
_KERNELS_PER_CALL = {"mrssm_latent_fwd": 2, "mrssm_overshoot_kl_fwd": 2, "mrssm_mse_fwd": 2, "mrssm_clip_adam": 3}


profile = None        # set to a list to record (name, tag, work, start_event, end_event) per call


def call(name, *args, tag=None, work=None):
    """Invoke a C-ABI entry point on torch's current stream; raise RuntimeError on failure.
    tag/work ({"flops":..,"bytes":..} algorithmic work) only feed the optional per-call profile."""
    global launches, kernel_launches
    _require_device()
    lib = load()
    stream = torch.cuda.current_stream().cuda_stream
    if profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args, stream)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.mrssm_last_error().decode()}")
    if profile is not None:
        e1.record()
        profile.append((name, tag, work, e0, e1))
    launches += 1
    kernel_launches += _KERNELS_PER_CALL.get(name, 1)


def call_host(name, *args):
    """Invoke a host-only C-ABI entry point (no stream argument, no device needed)."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.mrssm_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be CUDA fp32/bf16."""
    if t is None:
        return None
    assert t.is_cuda, "mrssm_b200 ops take CUDA tensors only"
    return t.data_ptr()


def ptr_any(t):
    """Device pointer of a CUDA tensor of any dtype."""
    assert t.is_cuda
    return t.data_ptr()


def t4(t, sI, sH, sW, sC):
    return T4(ptr(t), sI, sH, sW, sC)


def nhwc(t, H, W, Cc, pix_stride=None):
    """View a contiguous [N,H,W,C] (or [N, C] when H=W=1) buffer."""
    ps = Cc if pix_stride is None else pix_stride
    return T4(ptr(t), H * W * ps, W * ps, ps, 1)


def nchw(t, H, W, Cc):
    return T4(ptr(t), Cc * H * W, W, 1, H * W)


# ---- bf16 activation views (mrssm_tv) ------------------------------------------------------------------------------
NHWC, PLANAR, PARITY = "nhwc", "planar", "parity"


def view_numel(layout, n, H, W, Cp):
    if layout == PARITY:
        return n * 4 * (Cp // 8) * ((H + 1) // 2) * ((W + 1) // 2) * 8
    return n * H * W * Cp


def tv(p, layout, H, W, Cp):
    """View of a dense bf16 buffer holding [n,H,W,Cp] in the given layout (p: device pointer or tensor)."""
    if p is not None and not isinstance(p, int):
        p = ptr(p)
    if layout == NHWC:
        return TV(p, H * W * Cp, W * Cp, Cp, 8, 0, 0, 0)
    if layout == PLANAR:
        return TV(p, (Cp // 8) * H * W * 8, W * 8, 8, H * W * 8, 0, 0, 0)
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    sK = H2 * W2 * 8
    sP = (Cp // 8) * sK
    return TV(p, 4 * sP, W2 * 8, 8, sK, sP, 1, 0)


NO_TV = TV(None, 0, 0, 0, 0, 0, 0, 0)
NO_T4 = T4(None, 0, 0, 0, 0)
