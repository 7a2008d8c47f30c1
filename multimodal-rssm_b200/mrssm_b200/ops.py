"""torch.autograd.Function wrappers over the C ABI (include/mrssm_b200.h).

Host-side plumbing only: shapes, workspace allocation (torch caching allocator), stream and the
autograd chaining.  All arithmetic happens in libmrssm_b200.so.

Parameter gradients are written by the kernels *directly* into ``param.grad`` (accumulated, like
autograd would), and the Functions return ``None`` for parameter inputs; ``param.grad`` is
normally a view into the optimiser's flat gradient buffer (optim.FlatParams) so one fused
clip+Adam launch consumes them.  Use ``loss.backward()``, not ``torch.autograd.grad`` w.r.t.
parameters.
"""
import ctypes as C
import os

import torch
from torch.autograd import Function

from . import _lib as L

RELU, ELU = 1, 2


def _f32c(t):
    assert t.dtype == torch.float32 and t.is_cuda, (t.dtype, t.device)
    return t if t.is_contiguous() else t.contiguous()


def grad_buf(p):
    """Gradient accumulator of a parameter (created zeroed on first use)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


# ---- side stream: weight-gradient kernels that run beside the serial rollout BPTT ----------------------------------------------
# The rollout BPTT occupies 32 of the 148 SMs for ~2 ms and everything after it depends on it; the decoder's weight gradients
# depend on nothing that follows them.  The decoder's backward therefore only queues them (side_defer); the rollout's backward
# launches them on a second stream right before its own kernel (side_flush: the first one with its grid capped to the SMs the
# rollout leaves free), and whoever reads the gradient buffer next waits for that stream (side_join: optimiser step, DP exchange).
_SIDE = {"stream": None, "jobs": [], "keep": [], "done": None, "active": False,
         "enabled": os.environ.get("MRSSM_SIDE_WGRAD", "1") != "0"}        # (A/B switch for measurements)


def set_side_wgrad(on):
    _SIDE["enabled"] = bool(on)


class side_wgrad_scope:
    """`with side_wgrad_scope(): loss.backward()` — inside, weight gradients may be queued for the side stream; on exit the
    current stream has seen all of them.  Outside such a scope every Function launches its weight gradients in place."""

    def __enter__(self):
        _SIDE["active"] = True
        return self

    def __exit__(self, *exc):
        _SIDE["active"] = False
        side_join()
        return False


def side_enabled():
    return _SIDE["active"] and _SIDE["enabled"] and _STATE["bf16"] and L.profile is None


def side_defer(job, tensors):
    """Queue `job` (a closure launching kernels on torch's current stream) and keep `tensors` alive until it has run."""
    _SIDE["jobs"].append(job)
    _SIDE["keep"] += [t for t in tensors if t is not None]


def side_pending():
    return bool(_SIDE["jobs"]) or _SIDE["done"] is not None


def side_flush(first_sm_budget=0, all_jobs=False):
    """Launch the queued jobs on the side stream, after everything already on the current stream.  first_sm_budget caps the grid of
    the first job's persistent kernels (of every job when all_jobs)."""
    jobs, keep = _SIDE["jobs"], _SIDE["keep"]
    if not jobs:
        return
    if _SIDE["stream"] is None:
        _SIDE["stream"] = torch.cuda.Stream()
    st = _SIDE["stream"]
    ready = torch.cuda.Event()
    ready.record()
    st.wait_event(ready)
    with torch.cuda.stream(st):
        for i, job in enumerate(jobs):
            capped = bool(first_sm_budget) and (i == 0 or all_jobs)
            if capped:
                L.call_host("mrssm_pl_set_sm_budget", first_sm_budget)
            try:
                job()
            finally:
                if capped:
                    L.call_host("mrssm_pl_set_sm_budget", 148)
        done = torch.cuda.Event()
        done.record()
    for t in keep:
        t.record_stream(st)
    _SIDE["done"] = done
    _SIDE["jobs"], _SIDE["keep"] = [], []


def side_join():
    """Make the current stream see every queued / side-stream weight gradient (jobs nobody flushed run here, in order)."""
    jobs = _SIDE["jobs"]
    if jobs:
        for job in jobs:
            job()
        _SIDE["jobs"], _SIDE["keep"] = [], []
    if _SIDE["done"] is not None:
        torch.cuda.current_stream().wait_event(_SIDE["done"])
        _SIDE["done"] = None


def _conv(fn, geom, large, small, weight_ptr, w_ss, w_sl, bias_ptr=None, act=0, mask_ptr=None, mask_mode=0,
          accumulate=0, dtype=L.F32):
    a = L.ConvArgs(*geom, dtype, act, mask_mode, accumulate, large, small, weight_ptr, w_ss, w_sl, bias_ptr, mask_ptr)
    tag = work = None
    if L.profile is not None:
        n, Hl, Wl, Cl, Hs, Ws, Cs, k = geom
        esz = 4 if dtype == L.F32 else 2
        tag = "%s[%dx%dx%d<->%dx%dx%d k%d]" % (fn[11:], Hl, Wl, Cl, Hs, Ws, Cs, k)
        work = dict(flops=2.0 * n * Hs * Ws * Cs * Cl * k * k,
                    bytes=float(n) * (Hl * Wl * Cl + Hs * Ws * Cs) * esz + 4.0 * Cs * Cl * k * k)
    L.call(fn, C.byref(a), tag=tag, work=work)
    if fn == "mrssm_conv_wgrad" and bias_ptr:
        L.kernel_launches += 1


def _row_t4(p, ld):
    """[rows, cols] matrix with row stride ld viewed as n_img=rows, 1x1 spatial."""
    return L.T4(p, ld, 0, 0, 1)


def _off(t, elems):
    return t.data_ptr() + 4 * elems


# ---- dense helpers (nn.Linear semantics, weight [N,K] with row stride ldw) ------------------------
def dense_fwd(x_ptr, ldx, M, K, w_ptr, ldw, N, b_ptr, act, y_ptr, ldy, accumulate=0):
    _conv("mrssm_conv_down", (M, 1, 1, K, 1, 1, N, 1), _row_t4(x_ptr, ldx), _row_t4(y_ptr, ldy), w_ptr, ldw, 1,
          b_ptr, act, None, 0, accumulate)


def dense_dgrad(dy_ptr, ldy, M, N, w_ptr, ldw, K, dx_ptr, ldx, mask_ptr=None, mask_mode=0, accumulate=0):
    _conv("mrssm_conv_up", (M, 1, 1, K, 1, 1, N, 1), _row_t4(dx_ptr, ldx), _row_t4(dy_ptr, ldy), w_ptr, ldw, 1,
          None, 0, mask_ptr, mask_mode, accumulate)


def dense_wgrad(dy_ptr, ldy, x_ptr, ldx, M, N, K, dw_ptr, ldw, db_ptr=None):
    _conv("mrssm_conv_wgrad", (M, 1, 1, K, 1, 1, N, 1), _row_t4(x_ptr, ldx), _row_t4(dy_ptr, ldy), dw_ptr, ldw, 1,
          db_ptr)


def act_bwd(g, y, act):
    out = torch.empty_like(y)
    L.call("mrssm_act_bwd", L.ptr(_f32c(g)), L.ptr(y), y.numel(), act, L.ptr(out))
    return out


def transpose(src_ptr, rows, cols, ld, device):
    dst = torch.empty(cols, rows, device=device, dtype=torch.float32)
    L.call("mrssm_transpose", src_ptr, rows, cols, ld, L.ptr(dst))
    return dst


# ---- MLP (SymbolicEncoder / DenseDecoder / RewardModel) ----------------------------------------------
class MlpFn(Function):
    """y = L_n(...act(L_1([x_0, x_1, ...]))) with `act` after every layer but the last unless
    final_act.  Inputs may be several tensors that the reference concatenates (torch.cat([h, s]))."""

    @staticmethod
    def forward(ctx, act, final_act, n_parts, *args):
        parts = [_f32c(t) for t in args[:n_parts]]
        params = args[n_parts:]
        M = parts[0].shape[0]
        dev = parts[0].device
        n_layers = len(params) // 2
        acts = []
        cur = None
        for i in range(n_layers):
            W, b = params[2 * i], params[2 * i + 1]
            N, K = W.shape
            y = torch.empty(M, N, device=dev, dtype=torch.float32)
            a = act if (i < n_layers - 1 or final_act) else 0
            if i == 0:
                col = 0
                for j, p in enumerate(parts):
                    kp = p.shape[1]
                    lastp = j == len(parts) - 1
                    dense_fwd(L.ptr(p), kp, M, kp, _off(W, col), K, N, L.ptr(b) if lastp else None,
                              a if lastp else 0, L.ptr(y), N, accumulate=int(j > 0))
                    col += kp
                assert col == K
            else:
                dense_fwd(L.ptr(cur), K, M, K, L.ptr(W), K, N, L.ptr(b), a, L.ptr(y), N)
            acts.append(y)
            cur = y
        ctx.act, ctx.final_act, ctx.n_parts = act, final_act, n_parts
        ctx.params = params
        ctx.save_for_backward(*parts, *acts)        # outputs must go through save_for_backward (no ref cycle)
        return cur

    @staticmethod
    def backward(ctx, g):
        act, params = ctx.act, ctx.params
        parts, acts = list(ctx.saved_tensors[:ctx.n_parts]), list(ctx.saved_tensors[ctx.n_parts:])
        n_layers = len(params) // 2
        M = parts[0].shape[0]
        g = _f32c(g)
        if ctx.final_act and act:
            g = act_bwd(g, acts[-1], act)
        gparts = [None] * len(parts)
        for i in reversed(range(n_layers)):
            W, b = params[2 * i], params[2 * i + 1]
            N, K = W.shape
            gW, gb = grad_buf(W), grad_buf(b)
            if i > 0:
                x = acts[i - 1]
                dense_wgrad(L.ptr(g), N, L.ptr(x), K, M, N, K, L.ptr(gW), K, L.ptr(gb))
                gx = torch.empty_like(x)
                dense_dgrad(L.ptr(g), N, M, N, L.ptr(W), K, K, L.ptr(gx), K, L.ptr(x) if act else None, act)
                g = gx
            else:
                col = 0
                for j, p in enumerate(parts):
                    kp = p.shape[1]
                    dense_wgrad(L.ptr(g), N, L.ptr(p), kp, M, N, kp, _off(gW, col), K, L.ptr(gb) if j == 0 else None)
                    if ctx.needs_input_grad[3 + j]:
                        gp = torch.empty_like(p)
                        dense_dgrad(L.ptr(g), N, M, N, _off(W, col), K, kp, L.ptr(gp), kp)
                        gparts[j] = gp
                    col += kp
        return (None, None, None, *gparts, *([None] * len(params)))


# ---- image encoder: stack of Conv2d(k4,s2)+ReLU ------------------------------------------------------
class ConvEncoderFn(Function):
    """[N,C,H,W] fp32 NCHW -> [N, Cout*h*w] flattened in (C,H,W) order (encoder.py:344-346).
    Intermediates are NHWC."""

    @staticmethod
    def forward(ctx, x, *params):
        x = _f32c(x)
        N, Cc, H, W = x.shape
        dev = x.device
        n_layers = len(params) // 2
        tensors = [(x, L.nchw(x, H, W, Cc))]
        geoms = []
        Hl, Wl, Cl = H, W, Cc
        for i in range(n_layers):
            Wt, b = params[2 * i], params[2 * i + 1]
            Cs, _, k, _ = Wt.shape
            Hs, Ws = (Hl - k) // 2 + 1, (Wl - k) // 2 + 1
            if i == n_layers - 1:
                y = torch.empty(N, Cs * Hs * Ws, device=dev, dtype=torch.float32)
                yt = L.nchw(y, Hs, Ws, Cs)
            else:
                y = torch.empty(N, Hs, Ws, Cs, device=dev, dtype=torch.float32)
                yt = L.nhwc(y, Hs, Ws, Cs)
            geom = (N, Hl, Wl, Cl, Hs, Ws, Cs, k)
            _conv("mrssm_conv_down", geom, tensors[-1][1], yt, L.ptr(Wt), Cl * k * k, k * k, L.ptr(b), RELU)
            tensors.append((y, yt))
            geoms.append(geom)
            Hl, Wl, Cl = Hs, Ws, Cs
        ctx.views, ctx.geoms, ctx.params = [_strides(t4) for _, t4 in tensors], geoms, params
        ctx.save_for_backward(*[t for t, _ in tensors])
        return tensors[-1][0]

    @staticmethod
    def backward(ctx, g):
        geoms, params = ctx.geoms, ctx.params
        tensors = [(t, L.T4(L.ptr(t), *v)) for t, v in zip(ctx.saved_tensors, ctx.views)]
        n_layers = len(geoms)
        g = act_bwd(g, tensors[-1][0], RELU)
        gt = L.T4(L.ptr(g), *_strides(tensors[-1][1]))
        for i in reversed(range(n_layers)):
            Wt, b = params[2 * i], params[2 * i + 1]
            geom = geoms[i]
            Cl, k = geom[3], geom[7]
            xin, xt = tensors[i]
            _conv("mrssm_conv_wgrad", geom, xt, gt, L.ptr(grad_buf(Wt)), Cl * k * k, k * k, L.ptr(grad_buf(b)))
            if i > 0:
                gx = torch.empty_like(xin)
                gxt = L.T4(L.ptr(gx), *_strides(xt))
                _conv("mrssm_conv_up", geom, gxt, gt, L.ptr(Wt), Cl * k * k, k * k, None, 0, L.ptr(xin), RELU)
                g, gt = gx, gxt
        return (None, *([None] * len(params)))


def _strides(t4):
    return (t4.sI, t4.sH, t4.sW, t4.sC)


# ---- image decoder: fc([h,s]) -> ConvTranspose2d stack ------------------------------------------------
class ConvDecoderFn(Function):
    """h [R,D], s [R,S] -> recon [R,C,H,W] NCHW fp32 (observation_model.py:91-105)."""

    @staticmethod
    def forward(ctx, h, s, *params):
        h, s = _f32c(h), _f32c(s)
        R, D = h.shape
        S = s.shape[1]
        dev = h.device
        fcw, fcb = params[0], params[1]
        Em = fcw.shape[0]
        y0 = torch.empty(R, Em, device=dev, dtype=torch.float32)
        dense_fwd(L.ptr(h), D, R, D, L.ptr(fcw), D + S, Em, None, 0, L.ptr(y0), Em)
        dense_fwd(L.ptr(s), S, R, S, _off(fcw, D), D + S, Em, L.ptr(fcb), 0, L.ptr(y0), Em, accumulate=1)
        convs = params[2:]
        n_layers = len(convs) // 2
        tensors = [(y0, L.nhwc(y0, 1, 1, Em))]
        geoms = []
        Hs, Ws, Cs = 1, 1, Em
        for i in range(n_layers):
            Wt, b = convs[2 * i], convs[2 * i + 1]
            _, Cl, k, _ = Wt.shape
            Hl, Wl = 2 * (Hs - 1) + k, 2 * (Ws - 1) + k
            last = i == n_layers - 1
            if last:
                y = torch.empty(R, Cl, Hl, Wl, device=dev, dtype=torch.float32)
                yt = L.nchw(y, Hl, Wl, Cl)
            else:
                y = torch.empty(R, Hl, Wl, Cl, device=dev, dtype=torch.float32)
                yt = L.nhwc(y, Hl, Wl, Cl)
            geom = (R, Hl, Wl, Cl, Hs, Ws, Cs, k)
            _conv("mrssm_conv_up", geom, yt, tensors[-1][1], L.ptr(Wt), Cl * k * k, k * k, L.ptr(b), 0 if last else RELU)
            tensors.append((y, yt))
            geoms.append(geom)
            Hs, Ws, Cs = Hl, Wl, Cl
        ctx.views, ctx.geoms, ctx.params = [_strides(t4) for _, t4 in tensors], geoms, params
        ctx.save_for_backward(h, s, *[t for t, _ in tensors])
        return tensors[-1][0]

    @staticmethod
    def backward(ctx, g):
        geoms, params = ctx.geoms, ctx.params
        h, s = ctx.saved_tensors[:2]
        tensors = [(t, L.T4(L.ptr(t), *v)) for t, v in zip(ctx.saved_tensors[2:], ctx.views)]
        convs = params[2:]
        n_layers = len(geoms)
        g = _f32c(g)
        gt = L.T4(L.ptr(g), *_strides(tensors[-1][1]))
        for i in reversed(range(n_layers)):
            Wt, b = convs[2 * i], convs[2 * i + 1]
            geom = geoms[i]
            R, Hl, Wl, Cl, Hs, Ws, Cs, k = geom
            xin, xt = tensors[i]
            _conv("mrssm_conv_wgrad", geom, gt, xt, L.ptr(grad_buf(Wt)), Cl * k * k, k * k, None)
            _colsum_t4(gt, R, Hl, Wl, Cl, grad_buf(b))
            gx = torch.empty_like(xin)
            gxt = L.T4(L.ptr(gx), *_strides(xt))
            _conv("mrssm_conv_down", geom, gt, gxt, L.ptr(Wt), Cl * k * k, k * k, None, 0,
                  L.ptr(xin) if i > 0 else None, RELU if i > 0 else 0)
            g, gt = gx, gxt
        fcw, fcb = params[0], params[1]
        R, D = h.shape
        S = s.shape[1]
        Em = fcw.shape[0]
        gW = grad_buf(fcw)
        dense_wgrad(L.ptr(g), Em, L.ptr(h), D, R, Em, D, L.ptr(gW), D + S, L.ptr(grad_buf(fcb)))
        dense_wgrad(L.ptr(g), Em, L.ptr(s), S, R, Em, S, _off(gW, D), D + S, None)
        gh = gs = None
        if ctx.needs_input_grad[0]:
            gh = torch.empty_like(h)
            dense_dgrad(L.ptr(g), Em, R, Em, L.ptr(fcw), D + S, D, L.ptr(gh), D)
        if ctx.needs_input_grad[1]:
            gs = torch.empty_like(s)
            dense_dgrad(L.ptr(g), Em, R, Em, _off(fcw, D), D + S, S, L.ptr(gs), S)
        return (gh, gs, *([None] * len(params)))


def _colsum_t4(t4, n_img, H, W, Cc, out):
    a = L.ConvArgs(n_img, H, W, Cc, H, W, Cc, 1, L.F32, 0, 0, 0, t4, t4, None, 0, 0, L.ptr(out), None)
    L.call("mrssm_colsum_t4", C.byref(a))


# ---- reconstruction loss ---------------------------------------------------------------------------------
class MseLossFn(Function):
    """sum_features mean_{t,b} (y-o)^2 (observation_model.py:28-31 + base/algo.py:381-383)."""

    @staticmethod
    def forward(ctx, y, o, rows):
        y, o = _f32c(y), _f32c(o)
        assert y.numel() == o.numel()
        partial = torch.empty(2048, device=y.device, dtype=torch.float32)
        out = torch.empty(1, device=y.device, dtype=torch.float32)
        L.call("mrssm_mse_fwd", L.ptr(y), L.ptr(o), y.numel(), rows, L.ptr(partial), L.ptr(out))
        ctx.rows = rows
        ctx.save_for_backward(y, o)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        y, o = ctx.saved_tensors
        dy = torch.empty_like(y)
        L.call("mrssm_mse_bwd", L.ptr(y), L.ptr(o), y.numel(), ctx.rows, L.ptr(_f32c(g)), L.ptr(dy))
        return dy, None, None


def sqdiff(y, o):
    """Elementwise (y-o)^2, forward only (get_mse API)."""
    y, o = _f32c(y), _f32c(o)
    out = torch.empty_like(y)
    L.call("mrssm_sqdiff", L.ptr(y), L.ptr(o), y.numel(), L.ptr(out))
    return out


# ---- the rollout -------------------------------------------------------------------------------------------
class FusionTable:
    """Host-side constant tables of encoder.py:73-124 (Q8).  n_subsets == 0: no fusion (single-modal)."""

    def __init__(self, n_experts, state_size, fusion):
        import itertools
        self.n_experts = n_experts
        if fusion == "single":
            self.masks, self.dim_subset = [], [0] * state_size
        elif fusion == "MoPoE":
            K = n_experts - 1                       # expert 1 = prior_expert, 2.. = modalities
            subs = []
            for n in range(K + 1):
                subs += list(itertools.combinations(range(2, K + 2), n))
            self.masks = [1 | sum(1 << (e - 1) for e in sub) for sub in subs]
            n_sub = len(subs)
            step = int(torch.floor(state_size * torch.tensor(1.0 / float(n_sub), dtype=torch.float32)))
            self.dim_subset = [min(s // step if step > 0 else n_sub - 1, n_sub - 1) for s in range(state_size)]
        else:                                       # PoE / NN: one subset with every expert
            self.masks, self.dim_subset = [(1 << n_experts) - 1], [0] * state_size
        assert len(self.masks) <= L.MAX_SUBSETS and state_size <= L.MAX_STATE

    def fill(self, a):
        a.n_subsets = len(self.masks)
        for i, m in enumerate(self.masks):
            a.subset_mask[i] = m
        for i, d in enumerate(self.dim_subset):
            a.dim_subset[i] = d


class RolloutSpec:
    """Static description handed to RolloutFn: sizes, activation, fusion table, which experts carry
    an embedding, and the parameter order."""

    def __init__(self, D, S, H, A, act, min_std, table, expert_has_emb):
        self.D, self.S, self.H, self.A = D, S, H, A
        self.act, self.min_std, self.table = act, float(min_std), table
        self.expert_has_emb = list(expert_has_emb)      # per expert 1..E
        self.E = len(self.expert_has_emb)


N_FIXED_PARAMS = 6   # w_sa, b_sa, w_ih, w_hh, b_ih, b_hh  then per head (fc1.w, fc1.b, fc2.w, fc2.b)


def rollout_step_enabled():
    return _STATE["bf16"] and _STATE.get("rollout_step", True)


def set_rollout_step(on):
    """bf16 mode only: large models (beyond the fused tcgen05 rollout) run the rollout as per-step tcgen05 GEMMs."""
    _STATE["rollout_step"] = bool(on)


def _step_head_weights(spec, heads, has_pre, dev):
    """All heads' fc1 / fc2 weights stacked for merged launches (cached per weight version): fc1 over the belief columns as one
    GEMM per chunk of heads (<= 4096 output columns), fc2 and its dgrad as block-diagonal GEMMs (one block per head)."""
    D, S, H = spec.D, spec.S, spec.H
    NH, S2p = len(heads), pad16(2 * S)
    key = ("step_heads", heads[0][0].data_ptr(), NH)
    ver = (_STATE["wversion"],) + tuple(p._version for h in heads for p in h) + tuple(has_pre) + tuple(_tok(p) for h in heads for p in h)
    hit = _wcache.get(key)
    if hit is None or hit[0] != ver:
        per = max(1, 4096 // H)
        chunks = [(c0, min(NH, c0 + per)) for c0 in range(0, NH, per)]
        z = lambda n: torch.zeros(n, device=dev, dtype=torch.float32)
        w1f = [torch.cat([packed_cols(heads[h][0], 0, D, 0)[0] for h in range(c0, c1)]) for c0, c1 in chunks]
        b1 = [torch.cat([z(H) if has_pre[h] else heads[h][1].detach() for h in range(c0, c1)]) for c0, c1 in chunks]
        w2f = torch.cat([packed(heads[h][2], 0, S2p, H)[0] for h in range(NH)])
        b2 = torch.cat([torch.cat([heads[h][3].detach(), z(S2p - 2 * S)]) for h in range(NH)])
        w2b = [torch.cat([packed(heads[h][2], 1, S2p, H)[0] for h in range(c0, c1)]) for c0, c1 in chunks]
        w1cat = torch.empty(NH * H, D, device=dev, dtype=torch.float32)
        for hd in range(NH):
            w1 = heads[hd][0]
            L.call("mrssm_copy2d", L.ptr(w1), H, D, w1.shape[1], w1cat.data_ptr() + 4 * hd * H * D)
        w1b = tc_pack_weight(w1cat.reshape(NH * H, D, 1, 1), 1, NH * H, D)           # dh = [du_0 | du_1 | ..] [W1_0[:, :D]; W1_1[:, :D]; ..]
        hit = (ver, dict(chunks=chunks, w1f=w1f, b1=b1, w2f=w2f, b2=b2, w2b=w2b, w1b=w1b))
        _wcache[key] = hit
    return hit[1]


def _step_ws(spec, E, hw, KX):
    w = L.RstepWs()
    NH = 1 + E
    w.KX, w.S2p, w.NH, w.n_chunks = KX, pad16(2 * spec.S), NH, len(hw["chunks"])
    for ci, (c0, c1) in enumerate(hw["chunks"]):
        w.chunk_c0[ci], w.chunk_c1[ci] = c0, c1
        w.w1f[ci], w.b1[ci], w.w2b[ci] = L.ptr(hw["w1f"][ci]), L.ptr(hw["b1"][ci]), L.ptr(hw["w2b"][ci])
    w.w2f, w.b2, w.w1b = L.ptr(hw["w2f"]), L.ptr(hw["b2"]), L.ptr(hw["w1b"])
    return w


def _rollout_steps_fwd(a, spec, E, T, B, w_sa, b_sa, w_ih, w_hh, b_ih, b_hh, heads, has_pre, pre_cat, stash, dev):
    """The T steps of transition_model.py:226-270 as per-step tcgen05 GEMMs + the kernels of csrc/rollout_step.cu, issued by one C
    call (mrssm_rollout_steps_fwd).  Fills the output tensors and the stash named in `a`; returns the extra stash tensors
    [hb_all, xin_all]."""
    D, S, H, A = spec.D, spec.S, spec.H, spec.A
    NH, S2p = 1 + E, pad16(2 * S)
    KX = pad8(S + A)
    nb16 = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.bfloat16)
    xin_all = nb16(T, B, KX)
    x_all = stash["x"] if stash is not None else nb16(1, B, D)
    u_cat = stash["u"][0] if stash is not None else nb16(1, B, NH * H)
    hb_all = nb16(T + 1, B, D)
    L.call("mrssm_tc_to_bf16", C.byref(_row_t4(a.prev_belief, D)), B, 1, 1, D, D, 1.0, hb_all.data_ptr())
    gi = torch.empty(B, 3 * D, device=dev, dtype=torch.float32)
    gh = torch.empty(B, 3 * D, device=dev, dtype=torch.float32)
    o_cat = torch.empty(B, NH * S2p, device=dev, dtype=torch.float32)
    w = _step_ws(spec, E, _step_head_weights(spec, heads, has_pre, dev), KX)
    w.keep_all = int(stash is not None)
    w.wp_sa, w.wp_ih, w.wp_hh = L.ptr(packed(w_sa, 0, pad16(D), KX)), L.ptr(packed(w_ih, 0, pad16(3 * D), D)), L.ptr(packed(w_hh, 0, pad16(3 * D), D))
    w.pre_cat = L.ptr(pre_cat)
    w.xin_all, w.x_all, w.u_cat, w.hb_all = xin_all.data_ptr(), x_all.data_ptr(), u_cat.data_ptr(), hb_all.data_ptr()
    w.gi, w.gh, w.o_cat = gi.data_ptr(), gh.data_ptr(), o_cat.data_ptr()
    work = None
    if L.profile is not None:
        macs = (S + A) * D + 6 * D * D + NH * (D * H + H * 2 * S)
        work = dict(flops=2.0 * macs * T * B, bytes=0.0)
    L.call("mrssm_rollout_steps_fwd", C.byref(a), C.byref(w), tag="rollout_steps_fwd", work=work)
    L.kernel_launches += T * (6 + w.n_chunks)            # (+ the one L.call counts: the t = 0 xin kernel)
    return [hb_all, xin_all]


def _rollout_steps_bwd(ctx, ins, embs, outs, st, gouts):
    """BPTT of _rollout_steps_fwd: per step the dgrad GEMMs and the backward kernels of csrc/rollout_step.cu, then the deferred,
    time-parallel weight-gradient GEMMs over the bf16 per-step gradients."""
    spec, observe, E = ctx.spec, ctx.observe, ctx.E
    prev_state, actions, prev_belief, nonterminals, eps_prior, eps_post = ins
    params = ctx.params
    D, S, H, A = spec.D, spec.S, spec.H, spec.A
    T, B = actions.shape[0], actions.shape[1]
    R, NH = T * B, 1 + E
    dev = actions.device
    w_sa, b_sa, w_ih, w_hh, b_ih, b_hh = params[:N_FIXED_PARAMS]
    heads = [params[N_FIXED_PARAMS + 4 * i: N_FIXED_PARAMS + 4 * i + 4] for i in range(NH)]
    x_all, r_, z_, n_, ghn_, u_cat, hb_all, xin_all = st
    gouts = [None if g_ is None else _f32c(g_) for g_ in gouts]
    KX, S2p = xin_all.shape[-1], pad16(2 * S)
    if side_pending():
        # the decoder's queued weight gradients share the GPU with the (launch-latency-bound) per-step BPTT below
        side_flush(int(os.environ.get("MRSSM_STEP_SIDE_SMS", "100")), all_jobs=True)

    g = L.RolloutBwdArgs()
    a = g.f
    a.T, a.B, a.D, a.S, a.H, a.A, a.n_experts = T, B, D, S, H, A, E
    a.act, a.det, a.min_std = spec.act, int(bool(ctx.det)), spec.min_std
    a.prev_state, a.prev_belief, a.actions = L.ptr(prev_state), L.ptr(prev_belief), L.ptr(actions)
    a.nonterminals, a.eps_prior, a.eps_post = L.ptr(nonterminals), L.ptr(eps_prior), L.ptr(eps_post)
    if observe:
        spec.table.fill(a)
    a.beliefs, a.prior_states, a.prior_means, a.prior_stds = [L.ptr(t_) for t_ in outs[:4]]
    g.g_beliefs, g.g_prior_states, g.g_prior_means, g.g_prior_stds = [L.ptr(t_) for t_ in gouts[:4]]
    if observe:
        a.post_states, a.post_means, a.post_stds = [L.ptr(t_) for t_ in outs[4:7]]
        g.g_post_states, g.g_post_means, g.g_post_stds = [L.ptr(t_) for t_ in gouts[4:7]]
        for e in range(E):
            a.exp_means[e + 1], a.exp_stds[e + 1] = L.ptr(outs[7 + e]), L.ptr(outs[7 + E + e])
            g.g_exp_means[e + 1], g.g_exp_stds[e + 1] = L.ptr(gouts[7 + e]), L.ptr(gouts[7 + E + e])
    a.st_r, a.st_z, a.st_n, a.st_ghn = L.ptr(r_), L.ptr(z_), L.ptr(n_), L.ptr(ghn_)
    f32 = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.float32)
    nb16 = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.bfloat16)
    g_actions = f32(T, B, A)
    g.g_actions = L.ptr(g_actions)
    d_o = nb16(T, B, NH * S2p)
    du_all = nb16(T, B, NH * H)
    d_gi, d_gh, d_xpre = nb16(T, B, 3 * D), nb16(T, B, 3 * D), nb16(T, B, D)
    dh_heads, carry_b, dxin = f32(B, D), torch.zeros(B, D, device=dev), f32(B, S + A)
    carry_a, cgs = torch.zeros(B, D, device=dev), torch.zeros(B, S, device=dev)
    w = _step_ws(spec, E, _step_head_weights(spec, heads, [hd > 0 and spec.expert_has_emb[hd - 1] for hd in range(NH)], dev), KX)
    w.keep_all = 1
    wp_sa = packed(w_sa, 1, pad16(D), KX)
    w.wp_sa_b, w.wp_ih_b, w.wp_hh_b = L.ptr(wp_sa), L.ptr(packed(w_ih, 1, pad16(3 * D), D)), L.ptr(packed(w_hh, 1, pad16(3 * D), D))
    w.KXo = wp_sa.shape[1]
    g_prev_belief = f32(B, D)
    w.x_all, w.u_cat = x_all.data_ptr(), u_cat.data_ptr()
    w.d_o, w.du_all, w.d_gi, w.d_gh, w.d_xpre = d_o.data_ptr(), du_all.data_ptr(), d_gi.data_ptr(), d_gh.data_ptr(), d_xpre.data_ptr()
    w.dh_heads, w.carry_a, w.carry_b, w.dxin, w.cgs = dh_heads.data_ptr(), carry_a.data_ptr(), carry_b.data_ptr(), dxin.data_ptr(), cgs.data_ptr()
    w.g_prev_belief = g_prev_belief.data_ptr()
    work = None
    if L.profile is not None:
        macs = (S + A) * D + 6 * D * D + NH * (D * H + H * 2 * S)
        work = dict(flops=2.0 * macs * T * B, bytes=0.0)
    L.call("mrssm_rollout_steps_bwd", C.byref(g), C.byref(w), tag="rollout_steps_bwd", work=work)
    L.kernel_launches += T * (6 + w.n_chunks) + 1        # (+ xin_bwd of step 0; the one L.call counts: add2)
    g_prev_state = cgs

    # deferred, time-parallel weight gradients: dW += dY^T X over all (t, b) rows, bf16 operands as they are
    def wgrad(dy_ptr, ldy, N, x_ptr, ldx, K, rows, gw_ptr, ld, gb):
        tc_conv_wgrad((rows, 1, 1, pad8(K), 1, 1, pad8(N), 1), L.T4(x_ptr, ldx, 0, 0, 1), L.T4(dy_ptr, ldy, 0, 0, 1), gw_ptr, ld, 1, N, K)
        if gb is not None:
            L.call("mrssm_tc_colsum", dy_ptr, rows, ldy, N, L.ptr(gb))

    wgrad(d_xpre.data_ptr(), D, D, xin_all.data_ptr(), KX, S + A, R, L.ptr(grad_buf(w_sa)), S + A, grad_buf(b_sa))
    wgrad(d_gi.data_ptr(), 3 * D, 3 * D, x_all.data_ptr(), D, D, R, L.ptr(grad_buf(w_ih)), D, grad_buf(b_ih))
    wgrad(d_gh.data_ptr(), 3 * D, 3 * D, hb_all.data_ptr(), D, D, R, L.ptr(grad_buf(w_hh)), D, grad_buf(b_hh))
    g_embs, ei = [], 0
    for hd in range(NH):
        w1, b1, w2, b2 = heads[hd]
        ld = w1.shape[1]
        wgrad(d_o.data_ptr() + 2 * hd * S2p, NH * S2p, 2 * S, u_cat.data_ptr() + 2 * hd * H, NH * H, H, R, L.ptr(grad_buf(w2)), H, grad_buf(b2))
        du_ptr = du_all.data_ptr() + 2 * hd * H
        gw1 = grad_buf(w1)
        wgrad(du_ptr, NH * H, H, hb_all[1].data_ptr(), D, D, R, L.ptr(gw1), ld, grad_buf(b1))
        if hd > 0 and spec.expert_has_emb[hd - 1]:
            emb = embs[ei]
            Em = emb.shape[-1]
            Emp = pad8(Em)
            eb = pl_import(L.nhwc(emb, 1, 1, Em), R, 1, 1, Em, Emp, L.NHWC, dev)[0]
            wgrad(du_ptr, NH * H, H, eb.data_ptr(), Emp, Em, R, _off(gw1, D), ld, None)
            del eb
            ge = None
            if ctx.needs_input_grad[9 + ei]:
                ge = torch.empty_like(emb)
                tc_conv_up((R, 1, 1, Emp, 1, 1, pad8(H), 1), L.nhwc(ge, 1, 1, Em), L.T4(du_ptr, NH * H, 0, 0, 1),
                           packed_cols(w1, D, Em, 1), None, Em, out_f32=1, valid=(H, Em))
            g_embs.append(ge)
            ei += 1
    return (None, None, None, g_prev_state, g_actions, g_prev_belief, None, None, None, *g_embs, *([None] * len(params)))


class RolloutFn(Function):
    """All T steps of transition_model.py:226-270 in one launch (+ hoisted expert-embedding GEMMs).

    apply(spec, observe, det, prev_state, actions, prev_belief, nonterminals|None, eps_prior|None,
          eps_post|None, *embs (one per expert with an embedding, [T,B,E_m]), *params)
    params: fc_embed.weight, fc_embed.bias, rnn.weight_ih, rnn.weight_hh, rnn.bias_ih, rnn.bias_hh,
            then for head 0 (prior) and each expert: fc1.weight, fc1.bias, fc2.weight, fc2.bias.
    returns observe: (beliefs, prior_states, prior_means, prior_stds, post_states, post_means,
            post_stds, *exp_means, *exp_stds); imagine: first four.
    """

    @staticmethod
    def forward(ctx, spec, observe, det, prev_state, actions, prev_belief, nonterminals, eps_prior, eps_post, *rest):
        E = spec.E if observe else 0
        n_emb = sum(spec.expert_has_emb) if observe else 0
        embs = [_f32c(e) for e in rest[:n_emb]]
        params = rest[n_emb:]
        D, S, H, A = spec.D, spec.S, spec.H, spec.A
        T, B = actions.shape[0], actions.shape[1]
        dev = actions.device
        prev_state, actions, prev_belief = _f32c(prev_state), _f32c(actions), _f32c(prev_belief)
        nonterminals = None if nonterminals is None else _f32c(nonterminals)
        eps_prior = None if eps_prior is None else _f32c(eps_prior)
        eps_post = None if eps_post is None else _f32c(eps_post)
        need_grad = any(ctx.needs_input_grad)
        w_sa, b_sa, w_ih, w_hh, b_ih, b_hh = params[:N_FIXED_PARAMS]
        heads = [params[N_FIXED_PARAMS + 4 * i: N_FIXED_PARAMS + 4 * i + 4] for i in range(1 + E)]

        a = L.RolloutArgs()
        a.T, a.B, a.D, a.S, a.H, a.A, a.n_experts = T, B, D, S, H, A, E
        a.act, a.det, a.min_std = spec.act, int(bool(det)), spec.min_std
        a.prev_state, a.prev_belief, a.actions = L.ptr(prev_state), L.ptr(prev_belief), L.ptr(actions)
        a.nonterminals, a.eps_prior, a.eps_post = L.ptr(nonterminals), L.ptr(eps_prior), L.ptr(eps_post)
        keep = []
        use_tc = rollout_tc_enabled() and bool(L.load().mrssm_rollout_tc_eligible(D, S, H, A, E))
        # models too large for the fused tcgen05 rollout (BASELINE config 5: D = H = 1024): one tcgen05 GEMM per contraction per
        # time step over all B sequences, small kernels in between (csrc/rollout_step.cu)
        use_step = (not use_tc) and rollout_step_enabled() and D % 16 == 0 and H % 16 == 0 and D >= 256
        if use_tc:
            # tcgen05 rollout: the kernel reads the packed bf16 weight stream, only the biases come through `a`
            a.b_sa, a.b_ih, a.b_hh = L.ptr(b_sa), L.ptr(b_ih), L.ptr(b_hh)
            tc_plan, tc_packed = rollout_tc_weights(spec, E, w_sa, w_ih, w_hh, heads, dev)
        elif use_step:
            a.b_sa, a.b_ih, a.b_hh = L.ptr(b_sa), L.ptr(b_ih), L.ptr(b_hh)
        else:
            wsaT = transpose(L.ptr(w_sa), D, S + A, S + A, dev)
            wihT = transpose(L.ptr(w_ih), 3 * D, D, D, dev)
            whhT = transpose(L.ptr(w_hh), 3 * D, D, D, dev)
            keep += [wsaT, wihT, whhT]
            a.w_sa, a.b_sa, a.w_ih, a.b_ih, a.w_hh, a.b_hh = (L.ptr(wsaT), L.ptr(b_sa), L.ptr(wihT), L.ptr(b_ih),
                                                              L.ptr(whhT), L.ptr(b_hh))
        emb_pre = [None] * (1 + E)
        has_pre = [hd > 0 and bool(spec.expert_has_emb[hd - 1]) for hd in range(1 + E)]
        # step path: the hoisted embedding halves of all heads side by side (zero columns for heads without an embedding), the
        # addend of the merged fc1 GEMM
        pre_cat = torch.zeros(T, B, (1 + E) * H, device=dev, dtype=torch.float32) if (use_step and any(has_pre)) else None
        ei = 0
        for hd in range(1 + E):
            w1, b1, w2, b2 = heads[hd]
            ld = w1.shape[1]
            if not use_tc and not use_step:
                w1T = transpose(L.ptr(w1), H, D, ld, dev)
                w2T = transpose(L.ptr(w2), 2 * S, H, H, dev)
                keep += [w1T, w2T]
                a.w1[hd], a.w2[hd] = L.ptr(w1T), L.ptr(w2T)
            a.b2[hd], a.ld1[hd] = L.ptr(b2), ld
            if hd > 0 and spec.expert_has_emb[hd - 1]:
                emb = embs[ei]
                ei += 1
                Em = emb.shape[-1]
                assert ld == D + Em, (ld, D, Em)
                if pre_cat is not None:
                    pre, pre_ptr, pre_ld = pre_cat, pre_cat.data_ptr() + 4 * hd * H, (1 + E) * H
                else:
                    pre = torch.empty(T, B, H, device=dev, dtype=torch.float32)
                    pre_ptr, pre_ld = pre.data_ptr(), H
                if bf16_mode() and Em >= 64:
                    # hoisted embedding half of fc1 as a tcgen05 GEMM: bf16 operands, fp32 accumulate and output
                    Emp, Hp = pad8(Em), pad16(H)
                    eb = pl_import(L.nhwc(emb, 1, 1, Em), T * B, 1, 1, Em, Emp, L.NHWC, dev)[0]
                    tc_conv_down((T * B, 1, 1, Emp, 1, 1, Hp, 1), L.T4(eb.data_ptr(), Emp, 0, 0, 1), L.T4(pre_ptr, pre_ld, 0, 0, 1),
                                 packed_cols(w1, D, Em, 0), b1, H, out_f32=1, valid=(H, Em))
                    del eb
                else:
                    dense_fwd(L.ptr(emb), Em, T * B, Em, _off(w1, D), ld, H, L.ptr(b1), 0, pre_ptr, pre_ld)
                emb_pre[hd] = pre
                a.emb_pre[hd] = pre_ptr
                a.b1[hd] = None
            else:
                assert ld == D
                a.b1[hd] = L.ptr(b1)
        if observe:
            spec.table.fill(a)
        new = lambda n: torch.empty(T, B, n, device=dev, dtype=torch.float32)
        outs = [new(D), new(S), new(S), new(S)]
        a.beliefs, a.prior_states, a.prior_means, a.prior_stds = [L.ptr(t) for t in outs]
        if observe:
            post = [new(S), new(S), new(S)]
            a.post_states, a.post_means, a.post_stds = [L.ptr(t) for t in post]
            em = [new(S) for _ in range(E)]
            es = [new(S) for _ in range(E)]
            for e in range(E):
                a.exp_means[e + 1], a.exp_stds[e + 1] = L.ptr(em[e]), L.ptr(es[e])
            outs += post + em + es
        stash = None
        if need_grad and use_step:
            # bf16 stash of the step path: x and u are the GEMM operands themselves; + the bf16 beliefs and [state, action] rows
            nb16 = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.bfloat16)
            stash = dict(x=nb16(T, B, D), r=new(D), z=new(D), n=new(D), ghn=new(D), u=[nb16(T, B, (1 + E) * H)])
            a.st_r, a.st_z, a.st_n, a.st_ghn = [L.ptr(stash[k]) for k in ("r", "z", "n", "ghn")]
        elif need_grad:
            stash = dict(x=new(D), r=new(D), z=new(D), n=new(D), ghn=new(D), u=[new(H) for _ in range(1 + E)])
            a.st_x, a.st_r, a.st_z, a.st_n, a.st_ghn = [L.ptr(stash[k]) for k in ("x", "r", "z", "n", "ghn")]
            for hd in range(1 + E):
                a.st_u[hd] = L.ptr(stash["u"][hd])
        work = None
        if L.profile is not None:
            # compulsory HBM bytes per (b,t) = 4*(sum E_m + A + 1 + 2S) read + 4*(D + 3S + 3S + 2 E S) write
            n_w = sum(p.numel() for p in params)
            if observe:
                per = 4.0 * (sum(e.shape[-1] for e in embs) + A + 1 + 2 * S) + 4.0 * (D + 6 * S + 2 * E * S)
            else:
                per = 4.0 * (A + S) + 4.0 * (D + 3 * S)
            macs = (S + A) * D + 6 * D * D + (1 + E) * (D * H + H * 2 * S) + sum(e.shape[-1] * H for e in embs)
            work = dict(bytes=per * T * B + 4.0 * n_w, flops=2.0 * macs * T * B)
        extra = []
        if use_tc:
            L.call("mrssm_rollout_tc_fwd", C.byref(a), L.ptr_any(tc_plan), L.ptr_any(tc_packed),
                   tag="observe" if observe else "imagine", work=work)
        elif use_step:
            extra = _rollout_steps_fwd(a, spec, E, T, B, w_sa, b_sa, w_ih, w_hh, b_ih, b_hh, heads, has_pre, pre_cat, stash, dev)
        else:
            L.call("mrssm_rollout_fwd", C.byref(a), tag="observe" if observe else "imagine", work=work)
        del keep
        ctx.spec, ctx.observe, ctx.det, ctx.E = spec, observe, det, E
        ctx.params = params
        ctx.step_path = bool(use_step)
        ins = [prev_state, actions, prev_belief, nonterminals, eps_prior, eps_post]
        ctx.in_mask = [t is not None for t in ins]
        st = [] if stash is None else [stash[k] for k in ("x", "r", "z", "n", "ghn")] + stash["u"] + extra
        ctx.counts = (len(embs), len(outs), len(st))
        ctx.save_for_backward(*[t for t in ins if t is not None], *embs, *outs, *st)
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        spec, observe, E = ctx.spec, ctx.observe, ctx.E
        saved = list(ctx.saved_tensors)
        ins = [saved.pop(0) if m else None for m in ctx.in_mask]
        prev_state, actions, prev_belief, nonterminals, eps_prior, eps_post = ins
        n_emb, n_out, n_st = ctx.counts
        embs, outs, st = saved[:n_emb], saved[n_emb:n_emb + n_out], saved[n_emb + n_out:]
        assert n_st, "rollout forward ran without grad"
        if ctx.step_path:
            return _rollout_steps_bwd(ctx, ins, embs, outs, st, gouts)
        stash = dict(zip(("x", "r", "z", "n", "ghn"), st[:5]), u=st[5:])
        params = ctx.params
        D, S, H, A = spec.D, spec.S, spec.H, spec.A
        T, B = actions.shape[0], actions.shape[1]
        R = T * B
        dev = actions.device
        w_sa, b_sa, w_ih, w_hh, b_ih, b_hh = params[:N_FIXED_PARAMS]
        heads = [params[N_FIXED_PARAMS + 4 * i: N_FIXED_PARAMS + 4 * i + 4] for i in range(1 + E)]
        gouts = [None if g is None else _f32c(g) for g in gouts]

        g = L.RolloutBwdArgs()
        a = g.f
        a.T, a.B, a.D, a.S, a.H, a.A, a.n_experts = T, B, D, S, H, A, E
        a.act, a.det, a.min_std = spec.act, int(bool(ctx.det)), spec.min_std
        a.prev_state, a.prev_belief, a.actions = L.ptr(prev_state), L.ptr(prev_belief), L.ptr(actions)
        a.nonterminals, a.eps_prior, a.eps_post = L.ptr(nonterminals), L.ptr(eps_prior), L.ptr(eps_post)
        a.w_sa, a.b_sa, a.w_ih, a.b_ih, a.w_hh, a.b_hh = (L.ptr(w_sa), L.ptr(b_sa), L.ptr(w_ih), L.ptr(b_ih),
                                                          L.ptr(w_hh), L.ptr(b_hh))
        keep = []
        use_tc = rollout_tc_enabled() and bool(L.load().mrssm_rollout_tc_eligible(D, S, H, A, E))
        if use_tc:
            tc_packed = rollout_tc_weights(spec, E, w_sa, w_ih, w_hh, heads, dev, bwd=True)[1]
        for hd in range(1 + E):
            w1, b1, w2, b2 = heads[hd]
            a.w1[hd], a.ld1[hd], a.b1[hd], a.w2[hd], a.b2[hd] = L.ptr(w1), w1.shape[1], L.ptr(b1), L.ptr(w2), L.ptr(b2)
            a.st_u[hd] = L.ptr(stash["u"][hd])
            if use_tc:
                continue
            if w1.shape[1] == D:
                g.w1_belief[hd] = L.ptr(w1)
            else:                                       # contiguous copy of the belief columns for the weight streamer
                wc = torch.empty(H, D, device=dev, dtype=torch.float32)
                L.call("mrssm_copy2d", L.ptr(w1), H, D, w1.shape[1], L.ptr(wc))
                keep.append(wc)
                g.w1_belief[hd] = L.ptr(wc)
        if observe:
            spec.table.fill(a)
        a.beliefs, a.prior_states, a.prior_means, a.prior_stds = [L.ptr(t) for t in outs[:4]]
        g.g_beliefs, g.g_prior_states, g.g_prior_means, g.g_prior_stds = [L.ptr(t) for t in gouts[:4]]
        if observe:
            a.post_states, a.post_means, a.post_stds = [L.ptr(t) for t in outs[4:7]]
            g.g_post_states, g.g_post_means, g.g_post_stds = [L.ptr(t) for t in gouts[4:7]]
            for e in range(E):
                a.exp_means[e + 1], a.exp_stds[e + 1] = L.ptr(outs[7 + e]), L.ptr(outs[7 + E + e])
                g.g_exp_means[e + 1], g.g_exp_stds[e + 1] = L.ptr(gouts[7 + e]), L.ptr(gouts[7 + E + e])
        a.st_x, a.st_r, a.st_z, a.st_n, a.st_ghn = [L.ptr(stash[k]) for k in ("x", "r", "z", "n", "ghn")]
        new = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        g_prev_state, g_prev_belief, g_actions = new(B, S), new(B, D), new(T, B, A)
        d_xpre, d_gi, d_gh, xin = new(T, B, D), new(T, B, 3 * D), new(T, B, 3 * D), new(T, B, S + A)
        d_u = [new(T, B, H) for _ in range(1 + E)]
        d_o = [new(T, B, 2 * S) for _ in range(1 + E)]
        g.g_prev_state, g.g_prev_belief, g.g_actions = L.ptr(g_prev_state), L.ptr(g_prev_belief), L.ptr(g_actions)
        g.d_xpre, g.d_gi, g.d_gh, g.xin = L.ptr(d_xpre), L.ptr(d_gi), L.ptr(d_gh), L.ptr(xin)
        for hd in range(1 + E):
            g.d_u[hd], g.d_o[hd] = L.ptr(d_u[hd]), L.ptr(d_o[hd])
        work = None
        if L.profile is not None:
            # HBM bytes per (b,t) of the BPTT kernel: forward outputs and stash read back, upstream gradients read,
            # pre-activation gradients written (the deferred weight-gradient GEMMs consume them), fp32.
            NHh = 1 + E
            rd = (1 + 2 + 2 * E) * S + 5 * D + NHh * H + D + (D + 6 * S + 2 * E * S) + 2 * S + A + 1
            wr = 7 * D + NHh * H + NHh * 2 * S + (S + A) + A
            n_w = sum(p.numel() for p in params)
            macs = (S + A) * D + 6 * D * D + 2 * NHh * (D * H + H * 2 * S) - NHh * D * H      # dgrad GEMVs of one step
            work = dict(bytes=4.0 * (rd + wr) * T * B + 4.0 * n_w, flops=2.0 * macs * T * B)
        if use_tc:
            if side_pending():
                # the decoder's queued weight gradients go to the side stream now: the first of them on the SMs this kernel leaves free
                n_cta = -(-B // 32) if -(-B // 64) < 74 else -(-B // 64)
                side_flush(148 - n_cta if n_cta <= 74 else 0)
            L.call("mrssm_rollout_tc_bwd", C.byref(g), L.ptr_any(tc_packed), tag="observe" if observe else "imagine", work=work)
        else:
            L.call("mrssm_rollout_bwd", C.byref(g), tag="observe" if observe else "imagine", work=work)
        del keep

        # deferred, time-parallel weight gradients: dW += dY^T X over all (t,b) rows.  fp32 mode: exact CUDA-core GEMMs;
        # bf16 mode: operands rounded to bf16 once, tcgen05 GEMMs with fp32 accumulation into the fp32 master gradient.
        beliefs = outs[0]
        tc = bf16_mode()
        b16 = {}

        def as_bf16(t, cols):
            key = t.data_ptr()
            if key not in b16:
                rows = t.numel() // cols
                b16[key] = pl_import(L.nhwc(t, 1, 1, cols), rows, 1, 1, cols, pad8(cols), L.NHWC, dev)[0]
            return b16[key]

        def wgrad(dy, N, x, K, rows, gw_ptr, ld, gb, dy_row0=0, x_row0=0):
            if not tc:
                dense_wgrad(_off(dy, dy_row0 * N), N, _off(x, x_row0 * K), K, rows, N, K, gw_ptr, ld,
                            None if gb is None else L.ptr(gb))
                return
            Np, Kp = pad8(N), pad8(K)
            dyb, xb = as_bf16(dy, N), as_bf16(x, K)
            tc_conv_wgrad((rows, 1, 1, Kp, 1, 1, Np, 1), L.T4(xb.data_ptr() + 2 * x_row0 * Kp, Kp, 0, 0, 1),
                          L.T4(dyb.data_ptr() + 2 * dy_row0 * Np, Np, 0, 0, 1), gw_ptr, ld, 1, N, K)
            if gb is not None:
                _colsum(_off(dy, dy_row0 * N), rows, N, gb)

        wgrad(d_xpre, D, xin, S + A, R, L.ptr(grad_buf(w_sa)), S + A, grad_buf(b_sa))
        wgrad(d_gi, 3 * D, stash["x"], D, R, L.ptr(grad_buf(w_ih)), D, grad_buf(b_ih))
        gwhh = grad_buf(w_hh)
        wgrad(d_gh, 3 * D, prev_belief, D, B, L.ptr(gwhh), D, grad_buf(b_hh))
        if T > 1:
            wgrad(d_gh, 3 * D, beliefs, D, R - B, L.ptr(gwhh), D, None, dy_row0=B)
            _colsum(_off(d_gh, B * 3 * D), R - B, 3 * D, grad_buf(b_hh))
        g_embs = []
        ei = 0
        for hd in range(1 + E):
            w1, b1, w2, b2 = heads[hd]
            ld = w1.shape[1]
            wgrad(d_o[hd], 2 * S, stash["u"][hd], H, R, L.ptr(grad_buf(w2)), H, grad_buf(b2))
            gw1 = grad_buf(w1)
            wgrad(d_u[hd], H, beliefs, D, R, L.ptr(gw1), ld, grad_buf(b1))
            if hd > 0 and spec.expert_has_emb[hd - 1]:
                emb = embs[ei]
                Em = emb.shape[-1]
                wgrad(d_u[hd], H, emb, Em, R, _off(gw1, D), ld, None)
                ge = None
                if ctx.needs_input_grad[9 + ei]:
                    ge = torch.empty_like(emb)
                    if tc and Em >= 64:
                        Emp, Hp = pad8(Em), pad8(H)
                        tc_conv_up((R, 1, 1, Emp, 1, 1, Hp, 1), L.nhwc(ge, 1, 1, Em), L.T4(as_bf16(d_u[hd], H).data_ptr(), Hp, 0, 0, 1),
                                   packed_cols(w1, D, Em, 1), None, Em, out_f32=1, valid=(H, Em))
                    else:
                        dense_dgrad(L.ptr(d_u[hd]), H, R, H, _off(w1, D), ld, Em, L.ptr(ge), Em)
                g_embs.append(ge)
                ei += 1
        del b16
        return (None, None, None, g_prev_state, g_actions, g_prev_belief, None, None, None, *g_embs,
                *([None] * len(params)))


def _colsum(x_ptr, rows, cols, out):
    L.call("mrssm_colsum_acc", x_ptr, rows, cols, cols, L.ptr(out))


# ---- latent half of the ELBO -----------------------------------------------------------------------------------
class LatentSpec:
    def __init__(self, S, table, kl_mode, refuse, free_nats, alpha):
        self.S, self.table, self.kl_mode, self.refuse = S, table, kl_mode, refuse
        self.free_nats = float(free_nats)
        self.alpha = -1.0 if alpha is None else float(alpha)
        self.E = table.n_experts


class LatentFn(Function):
    """apply(spec, prior_means, prior_stds, post_means, post_stds, eps_dec|None, *exp_means, *exp_stds)
    -> refuse: (z_dec, q_means, q_stds, sums[2]); else (sums[2],)   sums = (kl_loss, global KL)."""

    @staticmethod
    def forward(ctx, spec, prior_means, prior_stds, post_means, post_stds, eps_dec, *experts):
        E = spec.E if (spec.refuse or spec.kl_mode == 1) else 0
        tens = [_f32c(t) for t in (prior_means, prior_stds, post_means, post_stds)]
        eps_dec = None if eps_dec is None else _f32c(eps_dec)
        ex = [_f32c(t) for t in experts]
        shape = tens[0].shape
        rows = tens[0].numel() // spec.S
        dev = tens[0].device
        a = LatentFn._args(spec, rows, tens, eps_dec, ex, E)
        scratch = torch.empty(2 * rows, device=dev, dtype=torch.float32)
        sums = torch.empty(2, device=dev, dtype=torch.float32)
        a.row_scratch, a.out_sums = L.ptr(scratch), L.ptr(sums)
        outs = []
        if spec.refuse:
            z, qm, qs = (torch.empty(shape, device=dev, dtype=torch.float32) for _ in range(3))
            a.z_dec, a.q_means, a.q_stds = L.ptr(z), L.ptr(qm), L.ptr(qs)
            outs = [z, qm, qs]
        L.call("mrssm_latent_fwd", C.byref(a))
        ctx.spec, ctx.rows, ctx.E, ctx.has_eps, ctx.n_ex = spec, rows, E, eps_dec is not None, len(ex)
        ctx.save_for_backward(*tens, *([eps_dec] if eps_dec is not None else []), *ex)
        ctx.set_materialize_grads(False)
        if spec.refuse:
            ctx.mark_non_differentiable(qm, qs)
        return (*outs, sums)

    @staticmethod
    def _args(spec, rows, tens, eps_dec, ex, E):
        a = L.LatentArgs()
        a.rows, a.S, a.n_experts, a.kl_mode, a.refuse = rows, spec.S, E, spec.kl_mode, int(spec.refuse)
        if E:
            spec.table.fill(a)
        a.free_nats, a.alpha = spec.free_nats, spec.alpha
        a.prior_means, a.prior_stds, a.post_means, a.post_stds = [L.ptr(t) for t in tens]
        a.eps_dec = L.ptr(eps_dec)
        for e in range(E):
            a.exp_means[e + 1], a.exp_stds[e + 1] = L.ptr(ex[e]), L.ptr(ex[E + e])
        return a

    @staticmethod
    def backward(ctx, *gouts):
        spec, rows, E = ctx.spec, ctx.rows, ctx.E
        saved = list(ctx.saved_tensors)
        tens, saved = saved[:4], saved[4:]
        eps_dec = saved.pop(0) if ctx.has_eps else None
        ex = saved
        dev = tens[0].device
        g_sums = gouts[-1]
        g_z = gouts[0] if spec.refuse else None
        if g_sums is None:
            g_sums = torch.zeros(2, device=dev, dtype=torch.float32)
        a = LatentFn._args(spec, rows, tens, eps_dec, ex, E)
        a.g_sums = L.ptr(_f32c(g_sums))
        a.g_z = None if g_z is None else L.ptr(_f32c(g_z))
        new = lambda: torch.empty_like(tens[0])
        gpm, gps = new(), new()
        a.g_prior_means, a.g_prior_stds = L.ptr(gpm), L.ptr(gps)
        gqm = gqs = None
        if not spec.refuse:
            gqm, gqs = new(), new()
            a.g_post_means, a.g_post_stds = L.ptr(gqm), L.ptr(gqs)
        gex = [new() for _ in range(2 * E)]
        for e in range(E):
            a.g_exp_means[e + 1], a.g_exp_stds[e + 1] = L.ptr(gex[e]), L.ptr(gex[E + e])
        scratch = torch.empty(1, device=dev, dtype=torch.float32)
        a.row_scratch, a.out_sums = L.ptr(scratch), L.ptr(scratch)
        L.call("mrssm_latent_bwd", C.byref(a))
        n_ex_in = len(ex)
        gex_full = gex if E else [None] * n_ex_in
        return (None, gpm, gps, gqm, gqs, None, *gex_full)


# ---- latent overshooting -----------------------------------------------------------------------------------------
class OvershootSpec:
    """T = chunk length (T-1 model steps), OD = overshooting distance; `mask` selects the experts whose product is the
    target posterior (0: the rollout's own posterior tensors); out = scale * mean_rows max(masked KL, free_nats)."""

    def __init__(self, T, B, S, A, OD, free_nats, scale=1.0, n_experts=0, mask=0):
        self.T, self.B, self.S, self.A, self.OD = T, B, S, A, OD
        self.free_nats, self.scale, self.n_experts, self.mask = float(free_nats), float(scale), n_experts, mask
        self.N = (T - 2) * B

    def args(self):
        a = L.OvershootArgs()
        a.T, a.B, a.S, a.A, a.OD = self.T, self.B, self.S, self.A, self.OD
        a.n_experts, a.subset_mask, a.free_nats, a.scale = self.n_experts, self.mask, self.free_nats, self.scale
        return a


def overshoot_gather(spec, actions, nonterminals, rewards=None, want_mask=False):
    """The padded, batch-concatenated open-loop inputs of base/algo.py:124-130 in one launch:
    -> actions [OD,N,A], nonterminals [OD,N,1], rewards [OD,N]|None, seq mask [OD,N]|None."""
    dev = actions.device
    actions, nonterminals = _f32c(actions), _f32c(nonterminals)
    assert actions.shape == (spec.T, spec.B, spec.A) and nonterminals.numel() == spec.T * spec.B
    new = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
    a = spec.args()
    act_o, nt_o = new(spec.OD, spec.N, spec.A), new(spec.OD, spec.N, 1)
    rw_o = new(spec.OD, spec.N) if rewards is not None else None
    mk_o = new(spec.OD, spec.N) if want_mask else None
    rewards = None if rewards is None else _f32c(rewards)
    a.actions, a.nonterminals, a.rewards = L.ptr(actions), L.ptr(nonterminals), L.ptr(rewards)
    a.actions_o, a.nonterminals_o, a.rewards_o, a.mask_o = L.ptr(act_o), L.ptr(nt_o), L.ptr(rw_o), L.ptr(mk_o)
    L.call("mrssm_overshoot_gather", C.byref(a))
    return act_o, nt_o, rw_o, mk_o


class OvershootKlFn(Function):
    """apply(spec, prior_means, prior_stds [OD,N,S], *targets) -> scalar.  targets: (post_means, post_stds) when
    spec.n_experts == 0, else (*exp_means, *exp_stds), each [T-1,B,S]; the targets are constants (detached in the
    reference, base/algo.py:129)."""

    @staticmethod
    def forward(ctx, spec, prior_means, prior_stds, *targets):
        pm, ps = _f32c(prior_means), _f32c(prior_stds)
        tg = [_f32c(t.detach()) for t in targets]
        assert pm.shape == (spec.OD, spec.N, spec.S) and all(t.shape == (spec.T - 1, spec.B, spec.S) for t in tg)
        dev = pm.device
        scratch = torch.empty(spec.OD * spec.N, device=dev, dtype=torch.float32)
        out = torch.empty((), device=dev, dtype=torch.float32)
        a = OvershootKlFn._args(spec, pm, ps, tg, scratch)
        a.out = L.ptr(out)
        L.call("mrssm_overshoot_kl_fwd", C.byref(a))
        ctx.spec = spec
        ctx.save_for_backward(pm, ps, scratch, *tg)
        return out

    @staticmethod
    def _args(spec, pm, ps, tg, scratch):
        a = spec.args()
        a.prior_means, a.prior_stds, a.row_scratch = L.ptr(pm), L.ptr(ps), L.ptr(scratch)
        if spec.n_experts:
            E = spec.n_experts
            assert len(tg) == 2 * E
            for e in range(E):
                a.exp_means[e + 1], a.exp_stds[e + 1] = L.ptr(tg[e]), L.ptr(tg[E + e])
        else:
            a.post_means, a.post_stds = L.ptr(tg[0]), L.ptr(tg[1])
        return a

    @staticmethod
    def backward(ctx, g):
        spec = ctx.spec
        pm, ps, scratch, *tg = ctx.saved_tensors
        a = OvershootKlFn._args(spec, pm, ps, tg, scratch)
        g = _f32c(g)
        gpm, gps = torch.empty_like(pm), torch.empty_like(ps)
        a.g_out, a.g_prior_means, a.g_prior_stds = L.ptr(g), L.ptr(gpm), L.ptr(gps)
        L.call("mrssm_overshoot_kl_bwd", C.byref(a))
        return (None, gpm, gps, *([None] * len(tg)))


# ---- tensor-core (bf16) primitives --------------------------------------------------------------------------
def pad8(c):
    return (c + 7) // 8 * 8


def pad16(c):
    return (c + 15) // 16 * 16


def pad64(c):
    return (c + 63) // 64 * 64


def tc_to_bf16(src_t4, n_img, H, W, Cc, device, scale=1.0, Cpad=None):
    """fp32 strided [n,H,W,C] -> bf16 NHWC [n,H,W,pad8(C)]."""
    Cpad = pad8(Cc) if Cpad is None else Cpad
    dst = torch.empty(n_img, H, W, Cpad, device=device, dtype=torch.bfloat16)
    L.call("mrssm_tc_to_bf16", C.byref(src_t4), n_img, H, W, Cc, Cpad, float(scale), L.ptr(dst))
    return dst


def tc_from_bf16(src, Cc, dst_t4):
    n_img, H, W, Cpad = src.shape
    L.call("mrssm_tc_from_bf16", L.ptr(src), n_img, H, W, Cc, Cpad, C.byref(dst_t4))


def tc_pack_weight(w, mode, Cs_pad, Cl_pad):
    """w: fp32 master [Cs,Cl,k,k] (or [out,in] Linear = k 1).  mode 0 down, 1 up (4 parity classes), 2 up on 1x1 input."""
    Cs, Cl = w.shape[0], w.shape[1]
    k = w.shape[2] if w.dim() == 4 else 1
    nt = (k + 1) // 2
    if mode == 0:
        Npad, Kpad, classes = pad16(Cs), pad64(k * k * Cl_pad), 1
    elif mode == 1:
        Npad, Kpad, classes = pad16(Cl), pad64(nt * nt * Cs_pad), 4
    else:
        Npad, Kpad, classes = pad16(k * k * Cl_pad), pad64(Cs_pad), 1
    out = torch.empty(classes, Npad, Kpad, device=w.device, dtype=torch.bfloat16)
    L.call("mrssm_tc_pack_weight", L.ptr(w), Cl * k * k, k * k, Cs, Cl, Cs_pad, Cl_pad, k, mode, Npad, Kpad, L.ptr(out))
    return out


def _tc_args(geom, large, small, act=0, mask=None, mask_mode=0, out_f32=0, n_out_pad=0, n_out_valid=0, bias_mod=0,
             cs_valid=0, cl_valid=0, wpacked=None, bias=None, dweight=None, w_ss=0, w_sl=0):
    return L.TcConvArgs(*geom, act, mask_mode, out_f32, n_out_pad, n_out_valid, bias_mod, cs_valid, cl_valid,
                        large, small, mask if mask is not None else L.T4(None, 0, 0, 0, 0), wpacked, bias, dweight, w_ss, w_sl)


def _tc_work(fn, geom, n_valid_pairs):
    if L.profile is None:
        return None, None
    n, Hl, Wl, Cl, Hs, Ws, Cs, k = geom
    cs, cl = n_valid_pairs
    tag = "%s[%dx%dx%d<->%dx%dx%d k%d]" % (fn[9:], Hl, Wl, cl, Hs, Ws, cs, k)
    return tag, dict(flops=2.0 * n * Hs * Ws * cs * cl * k * k, bytes=2.0 * n * (Hl * Wl * Cl + Hs * Ws * Cs) + 2.0 * cs * cl * k * k)


def tc_conv_down(geom, large, small, wpacked, bias, n_out_valid, act=0, mask=None, mask_mode=0, out_f32=0, bias_mod=0,
                 valid=None):
    a = _tc_args(geom, large, small, act, mask, mask_mode, out_f32, wpacked.shape[1], n_out_valid, bias_mod,
                 wpacked=L.ptr(wpacked), bias=L.ptr(bias))
    tag, work = _tc_work("mrssm_tc_conv_down", geom, valid or (geom[6], geom[3]))
    L.call("mrssm_tc_conv_down", C.byref(a), tag=tag, work=work)


def tc_conv_up(geom, large, small, wpacked, bias, n_out_valid, act=0, mask=None, mask_mode=0, out_f32=0, valid=None):
    a = _tc_args(geom, large, small, act, mask, mask_mode, out_f32, wpacked.shape[1], n_out_valid, 0,
                 wpacked=L.ptr(wpacked), bias=L.ptr(bias))
    tag, work = _tc_work("mrssm_tc_conv_up", geom, valid or (geom[6], geom[3]))
    L.call("mrssm_tc_conv_up", C.byref(a), tag=tag, work=work)


def tc_conv_wgrad(geom, large, small, dweight_ptr, w_ss, w_sl, cs_valid, cl_valid):
    a = _tc_args(geom, large, small, cs_valid=cs_valid, cl_valid=cl_valid, dweight=dweight_ptr, w_ss=w_ss, w_sl=w_sl)
    tag, work = _tc_work("mrssm_tc_conv_wgrad", geom, (cs_valid, cl_valid))
    L.call("mrssm_tc_conv_wgrad", C.byref(a), tag=tag, work=work)


def tc_colsum(x, Cvalid, out):
    rows = x.numel() // x.shape[-1]
    L.call("mrssm_tc_colsum", L.ptr(x), rows, x.shape[-1], Cvalid, L.ptr(out))


# ---- plane (TMA + shifted-descriptor) tensor-core convs, ksz >= 2 ------------------------------------------
DOWN, UP = 0, 1


DOWN_S2D = 2


def pl_pack_weight(w, op, Cs_pad, Cl_pad, s2d_cq=0):
    """fp32 master [Cs,Cl,k,k] -> bf16 [N_total, K_total] in the K-step order of csrc/conv_plane.cu.
    op DOWN_S2D: `down` over the space-to-depth form (Cl_pad = 16 channels) of an s2d_cq-channel image."""
    Cs, Cl, k, _ = w.shape
    n, kk = C.c_int32(), C.c_int32()
    L.call_host("mrssm_pl_packed_shape", op, Cs_pad, Cl_pad, k, C.byref(n), C.byref(kk))
    out = torch.empty(n.value, kk.value, device=w.device, dtype=torch.bfloat16)
    L.call("mrssm_pl_pack_weight", L.ptr(w), Cl * k * k, k * k, Cs, Cl, Cs_pad, Cl_pad, k, op, s2d_cq, L.ptr(out))
    return out


def _pl_args(geom, large=None, small=None, mask=None, act=0, mask_mode=0, out32=None, n_out_pad=0, n_out_valid=0,
             cs_valid=0, cl_valid=0, wpacked=None, bias=None, dweight=None, w_ss=0, w_sl=0, s2d_cq=0, scale=None, mse=None):
    """scale: (device float tensor, host factor) folded into the outputs; mse: (target fp32 tensor, sum buffer, factor)."""
    sp, sm = (L.ptr(scale[0]), float(scale[1])) if scale is not None else (None, 1.0)
    mt, ms, mf = (L.ptr(mse[0]), L.ptr(mse[1]), float(mse[2])) if mse is not None else (None, None, 0.0)
    f32 = int(out32 is not None)
    return L.PlConvArgs(*geom, act, mask_mode, f32, n_out_pad, n_out_valid, cs_valid, cl_valid, s2d_cq,
                        large or L.NO_TV, small or L.NO_TV, mask or L.NO_TV, out32 or L.NO_T4, wpacked, bias, dweight, w_ss, w_sl,
                        sp, sm, mt, ms, mf)


def new_relu_bits(n, H, W, Cp, device):
    """ReLU sign bits of an [n,H,W,Cp] activation: one byte per (pixel, 8-channel chunk), [n][H][W][Cp/8] (+ 16 bytes of slack
    for the 16-byte L2 prefetch granules).  Written by a forward epilogue, read by the matching dgrad epilogue."""
    return torch.empty(n * H * W * (Cp // 8) + 16, device=device, dtype=torch.uint8)


def _set_bits(a, bits_out, bits_in):
    if bits_out is not None:
        a.relu_bits_out = L.ptr_any(bits_out)
    if bits_in is not None:
        a.relu_bits_in, a.mask_mode = L.ptr_any(bits_in), RELU


def pl_conv_down(geom, large, out, wpacked, bias, n_out_valid, n_out_pad, act=0, mask=None, mask_mode=0, valid=None, s2d_cq=0,
                 scale=None, bits_out=None, bits_in=None):
    """large: L.TV view of the gathered tensor (parity-planar preferred; the space-to-depth view when s2d_cq);
    out: L.TV (bf16) or L.T4 (fp32, any strides).  bits_out: buffer of new_relu_bits for the sign bits of the (ReLU) output;
    bits_in: the sign bits standing in for the ReLU act'-mask (instead of `mask`)."""
    f32 = isinstance(out, L.T4)
    a = _pl_args(geom, large=large, small=None if f32 else out, mask=mask, act=act, mask_mode=mask_mode,
                 out32=out if f32 else None, n_out_pad=n_out_pad, n_out_valid=n_out_valid, wpacked=L.ptr(wpacked), bias=L.ptr(bias),
                 s2d_cq=s2d_cq, scale=scale)
    _set_bits(a, bits_out, bits_in)
    tag, work = _tc_work("mrssm_pl_conv_down", geom, valid or (geom[6], geom[3]))
    L.call("mrssm_pl_conv_down", C.byref(a), tag=tag, work=work)


def pl_conv_up(geom, out, small, wpacked, bias, n_out_valid, n_out_pad, act=0, mask=None, mask_mode=0, valid=None, bits_out=None,
               bits_in=None):
    """small: L.TV view of the gathered tensor (planar preferred, channels padded to 16); out: L.TV or L.T4 (fp32)."""
    f32 = isinstance(out, L.T4)
    a = _pl_args(geom, large=None if f32 else out, small=small, mask=mask, act=act, mask_mode=mask_mode,
                 out32=out if f32 else None, n_out_pad=n_out_pad, n_out_valid=n_out_valid, wpacked=L.ptr(wpacked), bias=L.ptr(bias))
    _set_bits(a, bits_out, bits_in)
    tag, work = _tc_work("mrssm_pl_conv_up", geom, valid or (geom[6], geom[3]))
    L.call("mrssm_pl_conv_up", C.byref(a), tag=tag, work=work)


def pl_conv_up_mse(geom, resid, small, wpacked, bias, n_out_valid, n_out_pad, target, target_t4, sum_buf, factor, recon_t4=None,
                   valid=None, target_s2d=None):
    """Last ConvTranspose2d + reconstruction loss in one kernel: adds factor * sum((recon - target)^2) to sum_buf[0] and writes
    the residual as bf16 in space-to-depth form into the view `resid`; recon_t4 (fp32, target's strides) is optional.
    target_s2d: L.TV view of the target in bf16 space-to-depth form (pl_import_s2d) — read instead of the fp32 target."""
    out32 = L.T4(recon_t4.ptr if recon_t4 is not None else None, target_t4.sI, target_t4.sH, target_t4.sW, target_t4.sC)
    a = _pl_args(geom, large=resid, small=small, out32=out32, n_out_pad=n_out_pad, n_out_valid=n_out_valid, wpacked=L.ptr(wpacked),
                 bias=L.ptr(bias), mse=(target, sum_buf, factor))
    if target_s2d is not None:
        a.mse_target_s2d = target_s2d
    tag, work = _tc_work("mrssm_pl_conv_up", geom, valid or (geom[6], geom[3]))
    L.call("mrssm_pl_conv_up", C.byref(a), tag=tag, work=work)


def pl_conv_wgrad(geom, large, small, dweight_ptr, w_ss, w_sl, cs_valid, cl_valid, s2d_cq=0, scale=None, dbias=None, dbias_from=0):
    """dbias (fp32 [channels], accumulated) + dbias_from (1: per-channel sums of `small`, the gradient of a Conv2d; 2: of `large`,
    the gradient of a ConvTranspose2d): the layer's bias gradient from the tile the kernel already holds in shared memory."""
    a = _pl_args(geom, large=large, small=small, cs_valid=cs_valid, cl_valid=cl_valid, dweight=dweight_ptr, w_ss=w_ss, w_sl=w_sl,
                 s2d_cq=s2d_cq, scale=scale)
    if dbias is not None:
        a.dbias, a.dbias_from = L.ptr(dbias), dbias_from
    tag, work = _tc_work("mrssm_pl_conv_wgrad", geom, (cs_valid, cl_valid))
    L.call("mrssm_pl_conv_wgrad", C.byref(a), tag=tag, work=work)


def new_act(n, H, W, Cp, layout, device):
    """Uninitialised bf16 activation buffer [n,H,W,Cp] in `layout` -> (flat tensor, L.TV view)."""
    t = torch.empty(L.view_numel(layout, n, H, W, Cp), device=device, dtype=torch.bfloat16)
    return t, L.tv(t, layout, H, W, Cp)


def pl_import(src_t4, n, H, W, Cc, Cp, layout, device, scale=1.0):
    """fp32 strided [n,H,W,C] -> bf16 view in `layout` with channels zero-padded to Cp."""
    t, v = new_act(n, H, W, Cp, layout, device)
    L.call("mrssm_pl_import", C.byref(src_t4), n, H, W, Cc, Cp, float(scale), C.byref(v))
    return t, v


def pl_copy(src_view, n, H, W, Cp, layout, device):
    """bf16 view -> new bf16 buffer of the same logical tensor in `layout`."""
    t, v = new_act(n, H, W, Cp, layout, device)
    L.call("mrssm_pl_copy", C.byref(src_view), n, H, W, Cp, C.byref(v))
    return t, v


def pl_colsum(view, n, H, W, Cp, Cvalid, out, fold=0, scale=None):
    """out[c] += sum over pixels.  fold: the view is the space-to-depth view (H, W its own size) of a fold-channel tensor."""
    sp, sm = (L.ptr(scale[0]), float(scale[1])) if scale is not None else (None, 1.0)
    L.call("mrssm_pl_colsum", C.byref(view), n, H, W, Cp, Cvalid, fold, sp, sm, L.ptr(out))


def pl_import_s2d(src_t4, n, H, W, Cc, device, scale=1.0):
    """fp32 strided [n,H,W,C<=4] -> planar bf16 space-to-depth view [n,ceil(H/2),ceil(W/2),16] (channel = parity*C + c)."""
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    t, v = new_act(n, H2, W2, 16, L.PLANAR, device)
    L.call("mrssm_pl_import_s2d", C.byref(src_t4), n, H, W, Cc, float(scale), C.byref(v))
    return t, v


# ---- bf16 tensor-core mode ------------------------------------------------------------------------------------
_STATE = {"bf16": False, "wversion": 0}
_wcache = {}
_tok_counter = [0]


def _tok(w):
    """Identity token of a weight tensor OBJECT: cached packings are keyed by address + versions, and a new tensor that lands on a
    freed tensor's address (a second model, a test's next parameter set) starts at the same versions — its token differs."""
    t = getattr(w, "_mrssm_tok", None)
    if t is None:
        _tok_counter[0] += 1
        t = w._mrssm_tok = _tok_counter[0]
    return t


def set_bf16_mode(on):
    """True: conv stacks / big GEMMs run on the tcgen05 kernels (bf16 operands, fp32 accumulate, fp32 master
    weights); False: exact fp32 CUDA-core kernels.  Selected from cfg.train.use_amp by the algorithm layer."""
    _STATE["bf16"] = bool(on)


def bf16_mode():
    return _STATE["bf16"]


def set_rollout_tc(on):
    """bf16 mode only: run the rollout forward on the tcgen05 kernel (csrc/rollout_tc.cu) when the sizes are eligible."""
    _STATE["rollout_tc"] = bool(on)


def rollout_tc_enabled():
    return _STATE["bf16"] and _STATE.get("rollout_tc", True)


_tc_plans = {}


def rollout_tc_plan(D, S, H, A, E, dev, bwd=False):
    """The per-step MMA program of the tcgen05 rollout (host-built once per shape and direction, kept on the device)."""
    key = (D, S, H, A, E, str(dev), bwd)
    hit = _tc_plans.get(key)
    if hit is None:
        pb, kb = C.c_int64(), C.c_int64()
        stem = "mrssm_rollout_tc_bwd_plan" if bwd else "mrssm_rollout_tc_plan"
        L.call_host(stem + "_bytes", D, S, H, A, E, C.byref(pb), C.byref(kb))
        host = torch.zeros(pb.value, dtype=torch.uint8)
        L.call_host(stem, D, S, H, A, E, host.data_ptr(), pb.value)
        n_pack = int(host[:64].view(torch.int32)[8])
        hit = (host.to(dev), n_pack, int(kb.value))
        _tc_plans[key] = hit
    return hit


def rollout_tc_weights(spec, E, w_sa, w_ih, w_hh, heads, dev, bwd=False):
    """(plan, packed bf16 weight stream) of the tcgen05 rollout (forward or BPTT); the stream is re-packed when the masters
    change."""
    D, S, H, A = spec.D, spec.S, spec.H, spec.A
    plan, n_pack, packed_bytes = rollout_tc_plan(D, S, H, A, E, dev, bwd)
    ws = [w_sa, w_ih, w_hh] + [w for hd in heads for w in (hd[0], hd[2])]
    key = ("rollout_tc", E, bwd) + tuple(w.data_ptr() for w in ws)
    ver = (_STATE["wversion"],) + tuple(w._version for w in ws) + tuple(_tok(w) for w in ws)
    hit = _wcache.get(key)
    if hit is None or hit[0] != ver:
        pa = L.RolloutArgs()
        pa.D, pa.S, pa.H, pa.A, pa.n_experts = D, S, H, A, E
        pa.w_sa, pa.w_ih, pa.w_hh = L.ptr(w_sa), L.ptr(w_ih), L.ptr(w_hh)
        for i, hd in enumerate(heads):
            pa.w1[i], pa.ld1[i], pa.w2[i] = L.ptr(hd[0]), hd[0].shape[1], L.ptr(hd[2])
        out = hit[1] if hit is not None else torch.empty(packed_bytes, device=dev, dtype=torch.uint8)
        L.call("mrssm_rollout_tc_pack", C.byref(pa), L.ptr_any(plan), n_pack, L.ptr_any(out))
        hit = (ver, out)
        _wcache[key] = hit
    return plan, hit[1]


def bump_weight_version():
    """Called after every optimiser step / checkpoint load: packed bf16 weight copies are stale."""
    _STATE["wversion"] += 1


def packed(w, mode, Cs_pad, Cl_pad):
    key = (w.data_ptr(), mode, Cs_pad, Cl_pad)
    ver = (_STATE["wversion"], w._version, _tok(w))
    hit = _wcache.get(key)
    if hit is None or hit[0] != ver:
        w4 = w if w.dim() == 4 else w.reshape(w.shape[0], w.shape[1], 1, 1)
        hit = (ver, tc_pack_weight(w4.detach(), mode, Cs_pad, Cl_pad))
        _wcache[key] = hit
    return hit[1]


def packed_cols(w, col0, ncols, mode):
    """Cached tcgen05 packing of the column block w[:, col0:col0+ncols] of a Linear weight [out, in] (mode 0: y = x W^T,
    mode 1: dx = dy W)."""
    key = (w.data_ptr(), "cols", col0, ncols, mode)
    ver = (_STATE["wversion"], w._version, _tok(w))
    hit = _wcache.get(key)
    if hit is None or hit[0] != ver:
        Cs, ld = w.shape
        Cl = ncols
        if mode == 0:
            Npad, Kpad, classes = pad16(Cs), pad64(pad8(Cl)), 1
        else:
            Npad, Kpad, classes = pad16(Cl), pad64(pad8(Cs)), 4
        out = torch.empty(classes, Npad, Kpad, device=w.device, dtype=torch.bfloat16)
        L.call("mrssm_tc_pack_weight", w.data_ptr() + 4 * col0, ld, 1, Cs, Cl, pad8(Cs), pad8(Cl), 1, mode, Npad, Kpad, L.ptr(out))
        hit = (ver, out)
        _wcache[key] = hit
    return hit[1]


def _bf16(*shape, device):
    return torch.empty(*shape, device=device, dtype=torch.bfloat16)


def packed_pl(w, op, Cs_pad, Cl_pad, s2d_cq=0):
    """Cached plane-kernel packing of a conv weight (re-packed when the master changes)."""
    key = (w.data_ptr(), "pl", op, Cs_pad, Cl_pad, s2d_cq)
    ver = (_STATE["wversion"], w._version, _tok(w))
    hit = _wcache.get(key)
    if hit is None or hit[0] != ver:
        hit = (ver, pl_pack_weight(w.detach(), op, Cs_pad, Cl_pad, s2d_cq))
        _wcache[key] = hit
    return hit[1]


class MlpTCFn(Function):
    """MlpFn in bf16 tensor-core mode: every Linear is a tcgen05 GEMM (bf16 operands, fp32 accumulate; activation in the epilogue,
    act' in the dgrad epilogue, bf16 intermediates), the result is fp32.  apply(act, final_act, n_parts, *parts, *params) like
    MlpFn.  Replaces eleven per-layer fp32 CUDA-core launches of the SymbolicEncoder / DenseDecoder / RewardModel stacks
    (reference encoder.py:282-305, observation_model.py:33-54, reward_model.py:20-35)."""

    @staticmethod
    def forward(ctx, act, final_act, n_parts, *args):
        parts = [_f32c(t) for t in args[:n_parts]]
        params = args[n_parts:]
        M = parts[0].shape[0]
        dev = parts[0].device
        n_layers = len(params) // 2
        if n_parts == 1:
            x0, K0 = parts[0], parts[0].shape[1]
        else:
            assert n_parts == 2, "MlpTCFn concatenates at most two inputs"
            K0 = parts[0].shape[1] + parts[1].shape[1]
            x0 = torch.empty(M, K0, device=dev, dtype=torch.float32)
            L.call("mrssm_concat2", L.ptr(parts[0]), parts[0].shape[1], L.ptr(parts[1]), parts[1].shape[1], M, L.ptr(x0))
        cur = tc_to_bf16(L.nhwc(x0, 1, 1, K0), M, 1, 1, K0, dev)            # [M,1,1,pad8(K0)]
        Kp = cur.shape[-1]
        acts = [cur]
        out = None
        for i in range(n_layers):
            W, b = params[2 * i], params[2 * i + 1]
            N, K = W.shape
            Np = pad16(N)
            last = i == n_layers - 1
            a = act if (not last or final_act) else 0
            geom = (M, 1, 1, Kp, 1, 1, Np, 1)
            if last:
                out = torch.empty(M, N, device=dev, dtype=torch.float32)
                tc_conv_down(geom, L.nhwc(cur, 1, 1, Kp), L.nhwc(out, 1, 1, N), packed(W, 0, Np, Kp), b, N, act=a, out_f32=1, valid=(N, K))
            else:
                y = _bf16(M, 1, 1, Np, device=dev)
                tc_conv_down(geom, L.nhwc(cur, 1, 1, Kp), L.nhwc(y, 1, 1, Np), packed(W, 0, Np, Kp), b, N, act=a, valid=(N, K))
                acts.append(y)
                cur, Kp = y, Np
        ctx.act, ctx.final_act, ctx.n_parts, ctx.params = act, final_act, n_parts, params
        ctx.part_cols = [p.shape[1] for p in parts]
        ctx.save_for_backward(out, *acts)
        return out

    @staticmethod
    def backward(ctx, g):
        act, params = ctx.act, ctx.params
        out, *acts = ctx.saved_tensors
        n_layers = len(params) // 2
        M = out.shape[0]
        dev = out.device
        g = _f32c(g)
        if ctx.final_act and act:
            g = act_bwd(g, out, act)
        N_last = out.shape[1]
        dy = tc_to_bf16(L.nhwc(g, 1, 1, N_last), M, 1, 1, N_last, dev, Cpad=pad16(N_last))
        gparts = [None] * ctx.n_parts
        for i in reversed(range(n_layers)):
            W, b = params[2 * i], params[2 * i + 1]
            N, K = W.shape
            x = acts[i]
            Kp, Np = x.shape[-1], dy.shape[-1]
            geom = (M, 1, 1, Kp, 1, 1, Np, 1)
            tc_conv_wgrad(geom, L.nhwc(x, 1, 1, Kp), L.nhwc(dy, 1, 1, Np), L.ptr(grad_buf(W)), K, 1, N, K)
            tc_colsum(dy.view(M, Np), N, grad_buf(b))
            if i > 0:
                dx = _bf16(M, 1, 1, Kp, device=dev)
                tc_conv_up(geom, L.nhwc(dx, 1, 1, Kp), L.nhwc(dy, 1, 1, Np), packed(W, 1, Np, Kp), None, K,
                           mask=L.nhwc(x, 1, 1, Kp) if act else None, mask_mode=act, valid=(N, K))
                dy = dx
            elif any(ctx.needs_input_grad[3:3 + ctx.n_parts]):
                gx = torch.empty(M, K, device=dev, dtype=torch.float32)
                tc_conv_up(geom, L.nhwc(gx, 1, 1, K), L.nhwc(dy, 1, 1, Np), packed(W, 1, Np, Kp), None, K, out_f32=1, valid=(N, K))
                col = 0
                for j, kp in enumerate(ctx.part_cols):
                    if ctx.needs_input_grad[3 + j]:
                        gparts[j] = gx[:, col:col + kp].contiguous() if ctx.n_parts > 1 else gx
                    col += kp
        return (None, None, None, *gparts, *([None] * len(params)))


def mlp(act, final_act, n_parts, *args):
    """MLP stack on the tensor cores in bf16 mode (large row counts), on the exact fp32 CUDA-core kernels otherwise."""
    rows = args[0].shape[0]
    fn = MlpTCFn if (bf16_mode() and rows >= 1024 and n_parts <= 2) else MlpFn
    return fn.apply(act, final_act, n_parts, *args)


# The bf16 space-to-depth form of the image the encoder just imported, keyed by the fp32 tensor it came from.  In a training step
# the decoder's reconstruction target is the encoder's input (base/algo.py:241,270-273), so the fused loss kernel can read this
# copy — half the bytes of the fp32 target, in the residual's own layout — instead.  One entry; a changed tensor never matches.
_S2D = {}


def remember_s2d(x, s2d):
    """s2d: the flat bf16 tensor of pl_import_s2d(x).  Entries live until the next optimize / validation (forget_s2d)."""
    _S2D[(x.data_ptr(), tuple(x.shape), x._version)] = s2d


def recall_s2d(target):
    return _S2D.get((target.data_ptr(), tuple(target.shape), target._version))


def forget_s2d():
    """Called at the start of every optimize / validation: an entry never outlives the iteration whose encoder made it (a later
    tensor may land on the same address)."""
    _S2D.clear()


class ConvEncoderTCFn(Function):
    """ConvEncoderFn on tensor cores: NCHW fp32 image -> bf16 parity-planar (8 ch) -> plane conv stack (parity-planar
    bf16 intermediates, each consumed with stride 2 by the next layer) -> fp32 [N, C*h*w] embedding in (C,H,W) order.
    Backward: gradients flow in planar bf16 (they are the stride-1 operand of dgrad / wgrad)."""

    @staticmethod
    def forward(ctx, x, *params):
        x = _f32c(x)
        N, Cc, H, W = x.shape
        dev = x.device
        n_layers = len(params) // 2
        # a <= 4-channel image goes in as its space-to-depth form (16 channels = 4 parities x C): no channel padding
        # to 8 per parity, half the K steps in the first conv and its weight gradient
        cq0 = Cc if Cc <= 4 else 0
        if cq0:
            pre = recall_s2d(x)          # made ahead of time by the input pipeline (mrssm_b200.data.PinnedChunkSource), off the critical path
            if pre is not None:
                acts = [(pre, L.tv(pre, L.PLANAR, (H + 1) // 2, (W + 1) // 2, 16))]
            else:
                acts = [pl_import_s2d(L.nchw(x, H, W, Cc), N, H, W, Cc, dev)]
                remember_s2d(x, acts[0][0])
            Clp = 16
        else:
            acts = [pl_import(L.nchw(x, H, W, Cc), N, H, W, Cc, pad8(Cc), L.PARITY, dev)]
            Clp = pad8(Cc)
        geoms = []
        bits = []
        Hl, Wl = H, W
        y = None
        for i in range(n_layers):
            Wt, b = params[2 * i], params[2 * i + 1]
            Cs, Cl, k, _ = Wt.shape
            Csp = pad16(Cs)
            Hs, Ws = (Hl - k) // 2 + 1, (Wl - k) // 2 + 1
            geom = (N, Hl, Wl, Clp, Hs, Ws, Csp, k)
            cq = cq0 if i == 0 else 0
            wp = packed_pl(Wt, DOWN_S2D if cq else DOWN, Csp, Clp, cq)
            if i == n_layers - 1:
                y = torch.empty(N, Cs * Hs * Ws, device=dev, dtype=torch.float32)
                pl_conv_down(geom, acts[-1][1], L.nchw(y, Hs, Ws, Cs), wp, b, Cs, Csp, act=RELU, valid=(Cs, Cl), s2d_cq=cq)
            else:
                o = new_act(N, Hs, Ws, Csp, L.PARITY, dev)
                bits.append(new_relu_bits(N, Hs, Ws, Csp, dev))       # 1 bit per element: the dgrad's act'-mask
                pl_conv_down(geom, acts[-1][1], o[1], wp, b, Cs, Csp, act=RELU, valid=(Cs, Cl), s2d_cq=cq, bits_out=bits[-1])
                acts.append(o)
            geoms.append((geom, Cs, Cl, cq))
            Hl, Wl, Clp = Hs, Ws, Csp
        ctx.geoms, ctx.params = geoms, params
        ctx.save_for_backward(y, *[t for t, _ in acts], *bits)
        return y

    @staticmethod
    def backward(ctx, g):
        geoms, params = ctx.geoms, ctx.params
        n_layers = len(geoms)
        y, *rest = ctx.saved_tensors
        acts, bits = rest[:n_layers], rest[n_layers:]        # bits[i - 1]: sign bits of acts[i] (the output of layer i - 1)
        dev = y.device
        (N, _, _, _, Hs, Ws, Csp, _), Cs, _, _ = geoms[-1]
        gm = act_bwd(g, y, RELU)
        gb = pl_import(L.nchw(gm, Hs, Ws, Cs), N, Hs, Ws, Cs, Csp, L.PLANAR, dev)
        for i in reversed(range(n_layers)):
            Wt, b = params[2 * i], params[2 * i + 1]
            geom, Cs, Cl, cq = geoms[i]
            N, Hl, Wl, Clp, Hs, Ws, Csp, k = geom
            xv = L.tv(acts[i], L.PLANAR, (Hl + 1) // 2, (Wl + 1) // 2, Clp) if cq else L.tv(acts[i], L.PARITY, Hl, Wl, Clp)
            # weight and bias gradient in one kernel (the bias sums read the gradient tile from shared memory); with grouped TMA
            # boxes and the passes split by row parity the 2x2-map layer (E4) runs here as well (0.9 ms; round 1 sent it to the NHWC
            # implicit-GEMM kernel: 1.2 ms + two layout copies)
            pl_conv_wgrad(geom, xv, gb[1], L.ptr(grad_buf(Wt)), Cl * k * k, k * k, Cs, Cl, s2d_cq=cq, dbias=grad_buf(b), dbias_from=1)
            if i > 0:
                gx = new_act(N, Hl, Wl, Clp, L.PLANAR, dev)
                pl_conv_up(geom, gx[1], gb[1], packed_pl(Wt, UP, Csp, Clp), None, Cl, Clp, bits_in=bits[i - 1], valid=(Cs, Cl))
                gb = gx
        return (None, *([None] * len(params)))


class ConvDecoderTCFn(Function):
    """ConvDecoderFn on tensor cores.  fc([h,s]) and the ConvTranspose on the 1x1 map are dense tcgen05 GEMMs (NHWC), the
    remaining layers are plane `up` kernels (all four output parities in one pass) with planar bf16 activations; output
    is fp32 NCHW.  Backward: gradients flow in parity-planar bf16 (stride-2 operand of dgrad / wgrad)."""

    @staticmethod
    def forward(ctx, h, s, *params):
        return ConvDecoderTCFn._forward(ctx, h, s, None, params)

    @staticmethod
    def _forward(ctx, h, s, target, params):
        """target None: returns the fp32 NCHW reconstruction.  Otherwise (fused loss): returns sum_features mean_rows
        (recon - target)^2 as a 0-dim tensor and keeps the bf16 residual for the backward pass."""
        h, s = _f32c(h), _f32c(s)
        R, D = h.shape
        S = s.shape[1]
        dev = h.device
        fcw, fcb = params[0], params[1]
        Em = fcw.shape[0]
        hs = torch.empty(R, D + S, device=dev, dtype=torch.float32)
        L.call("mrssm_concat2", L.ptr(h), D, L.ptr(s), S, R, L.ptr(hs))
        Kp = pad8(D + S)
        hsb = tc_to_bf16(L.nhwc(hs, 1, 1, D + S), R, 1, 1, D + S, dev)
        y0 = _bf16(R, 1, 1, pad16(Em), device=dev)
        g0 = (R, 1, 1, Kp, 1, 1, pad16(Em), 1)
        tc_conv_down(g0, L.nhwc(hsb, 1, 1, Kp), L.nhwc(y0, 1, 1, pad16(Em)), packed(fcw, 0, pad16(Em), Kp), fcb, Em,
                     valid=(Em, D + S))
        convs = params[2:]
        n_layers = len(convs) // 2
        acts = [(y0, L.NHWC)]
        bits = {}                                    # index into acts -> ReLU sign bits of that activation (plane layers)
        geoms = []
        Hs, Ws, Csp = 1, 1, pad16(Em)
        out = None
        for i in range(n_layers):
            Wt, b = convs[2 * i], convs[2 * i + 1]
            Cs, Cl, k, _ = Wt.shape
            Hl, Wl = 2 * (Hs - 1) + k, 2 * (Ws - 1) + k
            last = i == n_layers - 1
            Clp = pad8(Cl) if last else pad16(Cl)
            geom = (R, Hl, Wl, Clp, Hs, Ws, Csp, k)
            xt, xl = acts[-1]
            cq = Cl if (last and Cl <= 4) else 0          # backward: the image gradient travels in space-to-depth form
            if i == 0 and Hs == 1 and not last:
                assert Clp == Cl, "dense lowering of the first ConvTranspose needs Cout % 16 == 0"
                o = _bf16(R, Hl, Wl, Clp, device=dev)
                wp = packed(Wt, 2, Csp, Clp)
                tc_conv_down((R, 1, 1, Csp, 1, 1, k * k * Clp, 1), L.nhwc(xt, 1, 1, Csp), L.nhwc(o, 1, 1, k * k * Clp), wp, b,
                             k * k * Clp, act=RELU, bias_mod=Clp, valid=(k * k * Cl, Cs))
                # the plane kernels that gather this tensor (next ConvTranspose2d forward, its weight gradient) load x-contiguous
                # chunk planes with one grouped TMA box per tile; an NHWC source costs one box per plane (D2 wgrad: 4.4 vs 1.8 ms)
                op_ = pl_copy(L.tv(o, L.NHWC, Hl, Wl, Clp), R, Hl, Wl, Clp, L.PLANAR, dev)
                del o
                acts.append((op_[0], L.PLANAR))
            elif last and target is not None:
                assert cq, "the fused reconstruction loss needs an image of <= 4 channels"
                target = _f32c(target)
                resid = new_act(R, (Hl + 1) // 2, (Wl + 1) // 2, 16, L.PLANAR, dev)
                out = torch.zeros(1, device=dev, dtype=torch.float32)
                ts = recall_s2d(target)          # the encoder's (or the input pipeline's) bf16 s2d copy of these very frames
                tsv = L.tv(ts, L.PLANAR, (Hl + 1) // 2, (Wl + 1) // 2, 16) if ts is not None else None
                pl_conv_up_mse(geom, resid[1], L.tv(xt, xl, Hs, Ws, Csp), packed_pl(Wt, UP, Csp, Clp), b, Cl, Clp, target,
                               L.nchw(target, Hl, Wl, Cl), out, 1.0 / R, valid=(Cs, Cl), target_s2d=tsv)
                acts.append((resid[0], "s2d"))
            elif last:
                out = torch.empty(R, Cl, Hl, Wl, device=dev, dtype=torch.float32)
                pl_conv_up(geom, L.nchw(out, Hl, Wl, Cl), L.tv(xt, xl, Hs, Ws, Csp), packed_pl(Wt, UP, Csp, Clp), b, Cl, Clp,
                           valid=(Cs, Cl))
            else:
                o = new_act(R, Hl, Wl, Clp, L.PLANAR, dev)
                bits[len(acts)] = new_relu_bits(R, Hl, Wl, Clp, dev)
                pl_conv_up(geom, o[1], L.tv(xt, xl, Hs, Ws, Csp), packed_pl(Wt, UP, Csp, Clp), b, Cl, Clp, act=RELU, valid=(Cs, Cl),
                           bits_out=bits[len(acts)])
                acts.append((o[0], L.PLANAR))
            geoms.append((geom, Cs, Cl, cq))
            Hs, Ws, Csp = Hl, Wl, Clp
        ctx.geoms, ctx.params, ctx.dims = geoms, params, (R, D, S, Em, Kp)
        ctx.layouts = [l for _, l in acts]
        ctx.fused = target is not None
        ctx.bit_slots = sorted(bits)
        ctx.save_for_backward(hsb, *[t for t, _ in acts], *[bits[i] for i in ctx.bit_slots])
        return out.reshape(()) if target is not None else out

    @staticmethod
    def backward(ctx, g):
        return ConvDecoderTCFn._backward(ctx, g, 2)

    @staticmethod
    def _backward(ctx, g, n_lead):
        geoms, params, layouts = ctx.geoms, ctx.params, ctx.layouts
        R, D, S, Em, Kp = ctx.dims
        hsb, *acts = ctx.saved_tensors
        nb = len(ctx.bit_slots)
        bits = dict(zip(ctx.bit_slots, acts[len(acts) - nb:])) if nb else {}
        acts = acts[:len(acts) - nb]
        dev = hsb.device
        convs = params[2:]
        n_layers = len(geoms)
        g = _f32c(g)
        (_, Hl, Wl, Clp, _, _, _, _), _, Cl, cq_last = geoms[-1]
        scale = None
        if ctx.fused:
            # the saved residual r = recon - target; dLoss/drecon = g * 2 r / R, folded into the kernels that consume r
            gb, gl = acts.pop(), "s2d"
            layouts = layouts[:-1]
            scale = (g.reshape(1), 2.0 / R)
        elif cq_last:
            gb, gl = pl_import_s2d(L.nchw(g, Hl, Wl, Cl), R, Hl, Wl, Cl, dev)[0], "s2d"
        else:
            gb, gl = pl_import(L.nchw(g, Hl, Wl, Cl), R, Hl, Wl, Cl, Clp, L.PARITY, dev)[0], L.PARITY
        for i in reversed(range(n_layers)):
            Wt, b = convs[2 * i], convs[2 * i + 1]
            geom, Cs, Cl, cq = geoms[i]
            _, Hl, Wl, Clg, Hs, Ws, Csp, k = geom
            xi, xl = acts[i], layouts[i]
            dense = Hs == 1 and Ws == 1              # ConvTranspose on the 1x1 map: the dense kernels (NHWC)
            if dense:
                assert gl == L.NHWC
                gx = _bf16(R, Hs, Ws, Csp, device=dev)
                tc_conv_wgrad(geom, L.nhwc(gb, Hl, Wl, Clg), L.nhwc(xi, Hs, Ws, Csp), L.ptr(grad_buf(Wt)), Cl * k * k, k * k, Cs, Cl)
                tc_conv_down(geom, L.nhwc(gb, Hl, Wl, Clg), L.nhwc(gx, Hs, Ws, Csp), packed(Wt, 0, Csp, Clg), None, Cs,
                             mask=L.nhwc(xi, Hs, Ws, Csp) if i > 0 else None, mask_mode=RELU if i > 0 else 0, valid=(Cs, Cl))
                tc_colsum(gb.view(-1, Clg), Cl, grad_buf(b))
                nl = L.NHWC
            else:
                s2d = gl == "s2d"
                H2, W2 = (Hl + 1) // 2, (Wl + 1) // 2
                gg = (R, Hl, Wl, 16, Hs, Ws, Csp, k) if s2d else geom
                gv = L.tv(gb, L.PLANAR, H2, W2, 16) if s2d else L.tv(gb, gl, Hl, Wl, Clg)
                xv = L.tv(xi, xl, Hs, Ws, Csp)
                # the gradient w.r.t. this layer's input feeds a dense layer (NHWC) when the layer below sits on the 1x1 map
                nl = L.NHWC if (i > 0 and geoms[i - 1][0][4] == 1) or i == 0 else L.PARITY
                gxt, gxv = new_act(R, Hs, Ws, Csp, nl, dev)
                sc = scale if i == n_layers - 1 else None          # only the first consumer of the raw residual applies the scale
                def wgrad_job(gg=gg, gv=gv, xv=xv, Wt=Wt, b=b, Cl=Cl, Cs=Cs, k=k, cq=cq, s2d=s2d, sc=sc):
                    pl_conv_wgrad(gg, gv, xv, L.ptr(grad_buf(Wt)), Cl * k * k, k * k, Cs, Cl, s2d_cq=cq if s2d else 0, scale=sc,
                                  dbias=grad_buf(b), dbias_from=2)
                if side_enabled():
                    grad_buf(Wt), grad_buf(b)                       # (allocated on this stream if this is the first step)
                    side_defer(wgrad_job, [gb, xi, sc[0] if sc else None])
                else:
                    wgrad_job()
                xbits = bits.get(i)                    # sign bits of this layer's input (written by the plane layer below)
                pl_conv_down(gg, gv, gxv, packed_pl(Wt, DOWN_S2D if s2d else DOWN, Csp, gg[3], cq if s2d else 0), None, Cs, Csp,
                             mask=xv if (i > 0 and xbits is None) else None, mask_mode=RELU if i > 0 else 0, valid=(Cs, Cl),
                             s2d_cq=cq if s2d else 0, scale=sc, bits_in=xbits)
                gx = gxt
            gb, gl = gx, nl
        fcw, fcb = params[0], params[1]
        Emp = pad16(Em)
        g0 = (R, 1, 1, Kp, 1, 1, Emp, 1)
        tc_conv_wgrad(g0, L.nhwc(hsb, 1, 1, Kp), L.nhwc(gb, 1, 1, Emp), L.ptr(grad_buf(fcw)), D + S, 1, Em, D + S)
        tc_colsum(gb.view(R, Emp), Em, grad_buf(fcb))
        ghs = torch.empty(R, D + S, device=dev, dtype=torch.float32)
        tc_conv_up(g0, L.nhwc(ghs, 1, 1, D + S), L.nhwc(gb, 1, 1, Emp), packed(fcw, 1, Emp, Kp), None, D + S, out_f32=1,
                   valid=(Em, D + S))
        gh = ghs[:, :D].contiguous() if ctx.needs_input_grad[0] else None
        gs = ghs[:, D:].contiguous() if ctx.needs_input_grad[1] else None
        return (gh, gs, *([None] * (n_lead - 2)), *([None] * len(params)))


class ConvDecoderMseTCFn(Function):
    """Image decoder + reconstruction loss (observation_model.py:28-31 with base/algo.py:381-383) as one autograd node:
    apply(h, s, target [R,C,H,W] fp32, *params) -> sum_features mean_rows (recon - target)^2.  The reconstruction itself is
    never written: the last ConvTranspose2d's epilogue forms the residual, reduces the loss and stores the residual in bf16,
    which the backward pass consumes as the image gradient (scaled by the incoming grad on the fly)."""

    @staticmethod
    def forward(ctx, h, s, target, *params):
        return ConvDecoderTCFn._forward(ctx, h, s, target, params)

    @staticmethod
    def backward(ctx, g):
        return ConvDecoderTCFn._backward(ctx, g, 3)


# ---- general NCHW fp32 layers: the sound modality and the BatchNorm image stacks of the shipped YAML (csrc/generic_nchw.cu) ----
def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _gconv_call(fn, N, Cin, H, W, Cout, KH, KW, stride, padding, Ho, Wo, x=None, w=None, y=None, dx=None, dw=None):
    a = L.GConvArgs(N, Cin, H, W, Cout, KH, KW, stride[0], stride[1], padding[0], padding[1], Ho, Wo, x, w, y, dx, dw)
    tag = work = None
    if L.profile is not None:
        tag = "%s[%dx%dx%d->%dx%dx%d k%dx%d]" % (fn[6:], Cin, H, W, Cout, Ho, Wo, KH, KW)
        work = dict(flops=2.0 * N * Ho * Wo * Cout * Cin * KH * KW, bytes=4.0 * N * (Cin * H * W + Cout * Ho * Wo) + 4.0 * Cout * Cin * KH * KW)
    L.call(fn, C.byref(a), tag=tag, work=work)


class GConvFn(Function):
    """nn.Conv2d / nn.ConvTranspose2d without bias on fp32 NCHW tensors, any kernel / stride / zero padding (exact CUDA-core
    kernels).  apply(x, weight, stride, padding, transposed); the weight gradient is accumulated into weight.grad."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, transposed):
        x = _f32c(x)
        stride, padding = _pair(stride), _pair(padding)
        KH, KW = w.shape[2], (w.shape[3] if w.dim() == 4 else 1)           # Conv1d weights [out, in, k] are k x 1 kernels
        N = x.shape[0]
        if not transposed:
            Cin, H, W = x.shape[1:]
            Cout = w.shape[0]
            assert w.shape[1] == Cin, (w.shape, x.shape)
            Ho, Wo = (H + 2 * padding[0] - KH) // stride[0] + 1, (W + 2 * padding[1] - KW) // stride[1] + 1
            out = torch.empty(N, Cout, Ho, Wo, device=x.device, dtype=torch.float32)
            geom = (N, Cin, H, W, Cout, KH, KW, stride, padding, Ho, Wo)
            _gconv_call("mrssm_gconv_fwd", *geom, x=L.ptr(x), w=L.ptr(w), y=L.ptr(out))
        else:
            Cout, Ho, Wo = x.shape[1:]                     # conv geometry: the ConvT input is the conv's output side
            Cin = w.shape[1]
            assert w.shape[0] == Cout, (w.shape, x.shape)
            H, W = (Ho - 1) * stride[0] - 2 * padding[0] + KH, (Wo - 1) * stride[1] - 2 * padding[1] + KW
            out = torch.empty(N, Cin, H, W, device=x.device, dtype=torch.float32)
            geom = (N, Cin, H, W, Cout, KH, KW, stride, padding, Ho, Wo)
            _gconv_call("mrssm_gconv_dgrad", *geom, w=L.ptr(w), y=L.ptr(x), dx=L.ptr(out))
        ctx.geom, ctx.transposed, ctx.w = geom, transposed, w
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _f32c(g)
        w, geom = ctx.w, ctx.geom
        gx = None
        if not ctx.transposed:
            if w.requires_grad:
                _gconv_call("mrssm_gconv_wgrad", *geom, x=L.ptr(x), y=L.ptr(g), dw=L.ptr(grad_buf(w)))
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                _gconv_call("mrssm_gconv_dgrad", *geom, w=L.ptr(w), y=L.ptr(g), dx=L.ptr(gx))
        else:
            if w.requires_grad:
                _gconv_call("mrssm_gconv_wgrad", *geom, x=L.ptr(g), y=L.ptr(x), dw=L.ptr(grad_buf(w)))
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                _gconv_call("mrssm_gconv_fwd", *geom, x=L.ptr(g), w=L.ptr(w), y=L.ptr(gx))
        return gx, None, None, None, None


BN_EPS, BN_MOMENTUM = 1e-5, 0.1          # nn.BatchNorm2d / nn.InstanceNorm2d defaults, as the reference constructs them


class NormFn(Function):
    """nn.BatchNorm2d (instance=False) / nn.InstanceNorm2d / 1d (instance=True), affine, with an optional fused ReLU.
    apply(x [N,C,*], gamma, beta, running_mean | None, running_var | None, instance, train, relu).  train (or no running
    buffers): statistics of this batch, running buffers moved by momentum 0.1; else the running statistics.  The gradients of
    gamma / beta are accumulated into their .grad."""

    @staticmethod
    def forward(ctx, x, gamma, beta, rmean, rvar, instance, train, relu):
        x = _f32c(x)
        N, Cn = x.shape[0], x.shape[1]
        HW = x.numel() // (N * Cn)
        batch_stats = bool(train or rmean is None)
        groups = N * Cn if instance else Cn
        y = torch.empty_like(x)
        mean = torch.empty(groups, device=x.device, dtype=torch.float32) if batch_stats else None
        var = torch.empty(groups, device=x.device, dtype=torch.float32) if batch_stats else None
        a = L.NormArgs(N, Cn, HW, int(instance), int(batch_stats), int(relu), BN_EPS, BN_MOMENTUM, L.ptr(x), L.ptr(y), L.ptr(gamma), L.ptr(beta),
                       L.ptr(mean), L.ptr(var), L.ptr(rmean) if (rmean is not None and (train or not batch_stats)) else None,
                       L.ptr(rvar) if (rvar is not None and (train or not batch_stats)) else None)
        if not batch_stats:
            a.mean, a.var = L.ptr(rmean), L.ptr(rvar)       # (unused by the kernel in this mode; keeps the checks uniform)
        L.call("mrssm_norm_fwd", C.byref(a), tag="norm_fwd", work=dict(flops=0.0, bytes=8.0 * x.numel()) if L.profile is not None else None)
        L.kernel_launches += 2 if batch_stats else 0
        ctx.cfg = (N, Cn, HW, int(instance), int(batch_stats), int(relu))
        ctx.params = (gamma, beta, rmean, rvar)
        ctx.save_for_backward(x, y, mean, var)
        return y

    @staticmethod
    def backward(ctx, g):
        x, y, mean, var = ctx.saved_tensors
        gamma, beta, rmean, rvar = ctx.params
        N, Cn, HW, instance, batch_stats, relu = ctx.cfg
        g = _f32c(g)
        groups = N * Cn if instance else Cn
        scratch = torch.empty(2, groups, device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x)
        a = L.NormArgs(N, Cn, HW, instance, batch_stats, relu, BN_EPS, BN_MOMENTUM, L.ptr(x), L.ptr(y), L.ptr(gamma), L.ptr(beta),
                       L.ptr(mean), L.ptr(var), L.ptr(rmean), L.ptr(rvar))
        L.call("mrssm_norm_bwd", C.byref(a), L.ptr(g), scratch[0].data_ptr(), scratch[1].data_ptr(),
               L.ptr(grad_buf(gamma)) if gamma.requires_grad else None, L.ptr(grad_buf(beta)) if beta.requires_grad else None, L.ptr(dx),
               tag="norm_bwd", work=dict(flops=0.0, bytes=16.0 * x.numel()) if L.profile is not None else None)
        L.kernel_launches += 1
        return dx, None, None, None, None, None, None, None


class GluFn(Function):
    """nn.GLU(dim=1) on [N, 2C, *] fp32."""

    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        N, C2 = x.shape[0], x.shape[1]
        assert C2 % 2 == 0
        Lr = x.numel() // (N * C2)
        y = torch.empty(N, C2 // 2, *x.shape[2:], device=x.device, dtype=torch.float32)
        L.call("mrssm_glu_fwd", L.ptr(x), N, C2 // 2, Lr, L.ptr(y))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        N, C2 = x.shape[0], x.shape[1]
        dx = torch.empty_like(x)
        L.call("mrssm_glu_bwd", L.ptr(x), L.ptr(_f32c(g)), N, C2 // 2, x.numel() // (N * C2), L.ptr(dx))
        return dx


def conv2d_nobias(x, w, stride=1, padding=0):
    s_, p_ = _pair(stride), _pair(padding)
    KH, KW = w.shape[2], (w.shape[3] if w.dim() == 4 else 1)
    rows = x.shape[0] * ((x.shape[2] + 2 * p_[0] - KH) // s_[0] + 1) * ((x.shape[3] + 2 * p_[1] - KW) // s_[1] + 1)
    fn = GConvTCFn if gconv_tc_eligible(w.shape[1], w.shape[0], KH, KW, rows) else GConvFn
    return fn.apply(x, w, stride, padding, False)


def conv_transpose2d_nobias(x, w, stride=1, padding=0):
    KH, KW = w.shape[2], w.shape[3]
    rows = x.shape[0] * x.shape[2] * x.shape[3]
    fn = GConvTCFn if gconv_tc_eligible(w.shape[1], w.shape[0], KH, KW, rows) else GConvFn     # conv view: Cin = w.shape[1], Cout = w.shape[0]
    return fn.apply(x, w, stride, padding, True)


def batch_norm(x, bn, relu=False):
    """bn: the nn.BatchNorm2d holding the parameters / running buffers (its .training flag selects the statistics)."""
    y = NormFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, False, bn.training, relu)
    if bn.training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return y


def instance_norm(x, m):
    """m: nn.InstanceNorm2d / 1d (affine).  Tracked layers normalise with the running statistics in eval mode; torch leaves their
    num_batches_tracked untouched."""
    tracked = m.running_mean is not None
    return NormFn.apply(x, m.weight, m.bias, m.running_mean if tracked else None, m.running_var if tracked else None, True,
                        m.training or not tracked, False)


def conv1d_k1(x, w):
    """nn.Conv1d(kernel_size=1, bias=False) on [N, C, L]: the general kernel with H = L, W = 1."""
    assert w.dim() == 3 and w.shape[2] == 1
    return GConvFn.apply(x.unsqueeze(-1), w, 1, 0, False).squeeze(-1)


class ChannelBiasFn(Function):
    """y[n,c,...] = [relu](x[n,c,...] + bias[c]); the bias gradient is accumulated into bias.grad."""

    @staticmethod
    def forward(ctx, x, bias, relu):
        x = _f32c(x)
        N, Cn = x.shape[0], x.shape[1]
        y = torch.empty_like(x)
        L.call("mrssm_chan_bias_fwd", L.ptr(x), N, Cn, x.numel() // (N * Cn), L.ptr(bias), int(relu), L.ptr(y))
        ctx.bias, ctx.relu = bias, bool(relu)
        if relu:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        N, Cn = g.shape[0], g.shape[1]
        if ctx.relu:
            (y,) = ctx.saved_tensors
            g = act_bwd(g, y, RELU)
        if ctx.bias.requires_grad:
            L.call("mrssm_chan_bias_bwd", L.ptr(g), N, Cn, g.numel() // (N * Cn), L.ptr(grad_buf(ctx.bias)))
        return g, None, None


def add_channel_bias(x, bias, relu=False):
    return ChannelBiasFn.apply(x, bias, relu)


# ---- the same convolutions on the tensor cores (bf16 mode): NHWC bf16 staging, explicit im2col rows, dense tcgen05 GEMMs ------------
_GC_COL_BYTES = 2 << 30          # im2col rows materialised per chunk of images


def set_gconv_tc(on):
    _STATE["gconv_tc"] = bool(on)


def gconv_tc_eligible(Cin, Cout, KH, KW, rows):
    """bf16 mode, channels in tensor-core granules, enough work to fill the GEMM tiles (the 128 .. 512-channel sound layers, the
    BatchNorm image stacks' inner layers); everything else stays on the exact fp32 kernels."""
    return (_STATE["bf16"] and _STATE.get("gconv_tc", True) and Cin % 8 == 0 and Cout % 16 == 0 and Cout <= 4096 and Cin * KH * KW >= 256
            and (Cin * KH * KW) % 16 == 0
            and Cout >= 32 and rows >= 1024)


class _GcTc:
    """One convolution geometry (conv view: x [N,Cin,H,W], w [Cout,Cin,KH,KW], y [N,Cout,Ho,Wo]) in chunks of images."""

    def __init__(self, geom, w, dev):
        (self.N, self.Cin, self.H, self.W, self.Cout, self.KH, self.KW, self.stride, self.padding, self.Ho, self.Wo) = geom
        self.K = self.Cin * self.KH * self.KW
        self.dev = dev
        self.nb = max(1, min(self.N, 65535, _GC_COL_BYTES // (self.Ho * self.Wo * self.K * 2)))
        self.w = w
        self._w2 = None

    def w2(self):
        """fp32 [Cout, (tap, ci)] copy of the weight: the K order of the NHWC im2col rows."""
        if self._w2 is None:
            self._w2 = torch.empty(self.Cout, self.K, device=self.dev, dtype=torch.float32)
            L.call("mrssm_gconv_weight_perm", L.ptr(self.w), self.Cout, self.Cin, self.KH * self.KW, L.ptr(self._w2))
        return self._w2

    def args(self, n):
        return L.GConvArgs(n, self.Cin, self.H, self.W, self.Cout, self.KH, self.KW, self.stride[0], self.stride[1], self.padding[0],
                           self.padding[1], self.Ho, self.Wo, None, None, None, None, None)

    def chunks(self):
        for n0 in range(0, self.N, self.nb):
            yield n0, min(self.nb, self.N - n0)

    def nhwc_bf16(self, t, n0, n, Cc, HW):
        """images n0 .. n0+n of an fp32 NCHW tensor -> bf16 [n*HW, Cc] rows"""
        out = torch.empty(n * HW, Cc, device=self.dev, dtype=torch.bfloat16)
        L.call("mrssm_nchw_to_nhwc_bf16", t.data_ptr() + 4 * n0 * Cc * HW, n, Cc, HW, L.ptr(out))
        return out

    def to_nchw(self, rows, t, n0, n, Cc, HW):
        """fp32 [n*HW, Cc] rows -> images n0 .. n0+n of the fp32 NCHW tensor t"""
        L.call("mrssm_nhwc_to_nchw_f32", L.ptr(rows), n, Cc, HW, t.data_ptr() + 4 * n0 * Cc * HW)

    def im2col(self, xh, n):
        col = torch.empty(n * self.Ho * self.Wo, self.K, device=self.dev, dtype=torch.bfloat16)
        L.call("mrssm_im2col_nhwc", C.byref(self.args(n)), L.ptr(xh), L.ptr(col))
        return col

    def rows_times_wT(self, col, M):
        """[M, K] bf16 rows x w2^T -> fp32 [M, Cout]"""
        out = torch.empty(M, self.Cout, device=self.dev, dtype=torch.float32)
        wp = tc_pack_weight(self.w2().reshape(self.Cout, self.K, 1, 1), 0, self.Cout, self.K) if not hasattr(self, "_wp0") else self._wp0
        self._wp0 = wp
        tc_conv_down((M, 1, 1, self.K, 1, 1, self.Cout, 1), _row_t4(col.data_ptr(), self.K), _row_t4(out.data_ptr(), self.Cout), wp, None,
                     self.Cout, out_f32=1, valid=(self.Cout, self.K))
        return out

    def rows_times_w(self, yh, M):
        """[M, Cout] bf16 rows x w2 -> bf16 [M, K] (the gradient of the im2col rows), <= 4096 columns per GEMM"""
        dcol = torch.empty(M, self.K, device=self.dev, dtype=torch.bfloat16)
        if not hasattr(self, "_wp1"):
            self._wp1 = []
            for c0 in range(0, self.K, 4096):
                nc = min(4096, self.K - c0)
                Npad, Kpad = pad16(nc), pad64(self.Cout)
                wp = torch.empty(4, Npad, Kpad, device=self.dev, dtype=torch.bfloat16)
                L.call("mrssm_tc_pack_weight", self.w2().data_ptr() + 4 * c0, self.K, 1, self.Cout, nc, self.Cout, pad8(nc), 1, 1, Npad, Kpad,
                       L.ptr(wp))
                self._wp1.append((c0, nc, wp))
        for c0, nc, wp in self._wp1:
            tc_conv_up((M, 1, 1, pad16(nc), 1, 1, self.Cout, 1), _row_t4(dcol.data_ptr() + 2 * c0, self.K), _row_t4(yh.data_ptr(), self.Cout),
                       wp, None, nc, valid=(self.Cout, nc))
        return dcol

    def col2im(self, dcol, n):
        out = torch.empty(n * self.H * self.W, self.Cin, device=self.dev, dtype=torch.float32)
        L.call("mrssm_col2im_nhwc", C.byref(self.args(n)), L.ptr(dcol), L.ptr(out))
        return out

    def wgrad(self, col, yh, M, dw2):
        """dw2[Cout, K] += yh^T col"""
        tc_conv_wgrad((M, 1, 1, self.K, 1, 1, self.Cout, 1), _row_t4(col.data_ptr(), self.K), _row_t4(yh.data_ptr(), self.Cout), L.ptr(dw2),
                      self.K, 1, self.Cout, self.K)

    def add_wgrad(self, dw2):
        L.call("mrssm_gconv_weight_perm_add", L.ptr(dw2), self.Cout, self.Cin, self.KH * self.KW, L.ptr(grad_buf(self.w)))


class GConvTCFn(Function):
    """GConvFn on the tensor cores: same arguments and results (fp32 NCHW in / out, weight gradient accumulated into weight.grad),
    bf16 operands, fp32 accumulation.  The im2col rows of a chunk of images are materialised in bf16 (recomputed in backward
    rather than kept)."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, transposed):
        x = _f32c(x)
        stride, padding = _pair(stride), _pair(padding)
        KH, KW = w.shape[2], (w.shape[3] if w.dim() == 4 else 1)
        N = x.shape[0]
        if not transposed:
            Cin, H, W = x.shape[1:]
            Cout = w.shape[0]
            Ho, Wo = (H + 2 * padding[0] - KH) // stride[0] + 1, (W + 2 * padding[1] - KW) // stride[1] + 1
        else:
            Cout, Ho, Wo = x.shape[1:]
            Cin = w.shape[1]
            H, W = (Ho - 1) * stride[0] - 2 * padding[0] + KH, (Wo - 1) * stride[1] - 2 * padding[1] + KW
        geom = (N, Cin, H, W, Cout, KH, KW, stride, padding, Ho, Wo)
        g = _GcTc(geom, w, x.device)
        if not transposed:
            out = torch.empty(N, Cout, Ho, Wo, device=x.device, dtype=torch.float32)
            for n0, n in g.chunks():
                col = g.im2col(g.nhwc_bf16(x, n0, n, Cin, H * W), n)
                g.to_nchw(g.rows_times_wT(col, n * Ho * Wo), out, n0, n, Cout, Ho * Wo)
                del col
        else:
            out = torch.empty(N, Cin, H, W, device=x.device, dtype=torch.float32)
            for n0, n in g.chunks():
                dcol = g.rows_times_w(g.nhwc_bf16(x, n0, n, Cout, Ho * Wo), n * Ho * Wo)
                g.to_nchw(g.col2im(dcol, n), out, n0, n, Cin, H * W)
                del dcol
        ctx.geom, ctx.transposed, ctx.w = geom, transposed, w
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, go):
        (x,) = ctx.saved_tensors
        go = _f32c(go)
        w, geom = ctx.w, ctx.geom
        N, Cin, H, W, Cout, KH, KW, stride, padding, Ho, Wo = geom
        g = _GcTc(geom, w, x.device)
        need_x, need_w = ctx.needs_input_grad[0], w.requires_grad
        gx = torch.empty_like(x) if need_x else None
        dw2 = torch.zeros(Cout, g.K, device=x.device, dtype=torch.float32) if need_w else None
        for n0, n in g.chunks():
            M = n * Ho * Wo
            if not ctx.transposed:
                yh = g.nhwc_bf16(go, n0, n, Cout, Ho * Wo)                 # output gradient rows [M, Cout]
                if need_w:
                    col = g.im2col(g.nhwc_bf16(x, n0, n, Cin, H * W), n)
                    g.wgrad(col, yh, M, dw2)
                    del col
                if need_x:
                    dcol = g.rows_times_w(yh, M)
                    g.to_nchw(g.col2im(dcol, n), gx, n0, n, Cin, H * W)
                    del dcol
            else:
                col = g.im2col(g.nhwc_bf16(go, n0, n, Cin, H * W), n)      # im2col rows of the (large) output gradient
                if need_w:
                    g.wgrad(col, g.nhwc_bf16(x, n0, n, Cout, Ho * Wo), M, dw2)
                if need_x:
                    g.to_nchw(g.rows_times_wT(col, M), gx, n0, n, Cout, Ho * Wo)
                del col
        if need_w:
            g.add_wgrad(dw2)
        return gx, None, None, None, None
