"""Flat parameter storage + fused global-norm clip + Adam (one C-ABI call, no host sync).

Replaces, for the hot path, ``nn.utils.clip_grad_norm_(param_list, grad_clip_norm, norm_type=2)``
followed by ``optim.Adam(param_list, lr, eps=adam_epsilon).step()`` (reference
algos/MRSSM/base/algo.py:41-42,258-259).  Parameters stay ordinary ``nn.Parameter`` objects in
PyTorch layout (checkpoint compatible); their ``.data`` and ``.grad`` are re-pointed to views of
three flat fp32 buffers so that the optimiser (and the DP all-reduce) touch one contiguous range.
"""
import torch

from . import _lib as L


class FusedClipAdam(torch.optim.Optimizer):
    """Drop-in for the reference's ``model_optimizer`` (torch.optim.Adam): same param_groups keys and
    an Adam-format ``state_dict`` (step / exp_avg / exp_avg_sq per parameter)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-7, max_grad_norm=100.0):
        params = list(params)
        defaults = dict(lr=lr, betas=betas, eps=eps, max_grad_norm=max_grad_norm, weight_decay=0, amsgrad=False,
                        maximize=False, foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        assert len(self.param_groups) == 1
        dev = params[0].device
        sizes = [p.numel() for p in params]
        pad = [(-n) % 4 for n in sizes]                      # keep every view 16-byte aligned
        total = sum(n + q for n, q in zip(sizes, pad))
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(total, device=dev, dtype=torch.float32)
        self._partial = torch.empty(2048, device=dev, dtype=torch.float32)
        self.grad_norm = torch.zeros(1, device=dev, dtype=torch.float32)   # pre-clip total norm of the last step
        self.step_count = 0
        self.grad_scale = 1.0                                # 1/world_size under DP (sum all-reduce)
        off = 0
        self._views = []
        with torch.no_grad():
            for p, n, q in zip(params, sizes, pad):
                assert p.dtype == torch.float32 and p.device == dev
                self.flat_p[off:off + n].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + n].view(p.shape)
                p.grad = self.flat_g[off:off + n].view(p.shape)
                self.state[p] = dict(step=torch.tensor(0.0), exp_avg=self.flat_m[off:off + n].view(p.shape),
                                     exp_avg_sq=self.flat_v[off:off + n].view(p.shape))
                self._views.append((off, n))
                off += n + q

    def zero_grad(self, set_to_none=False):
        L.call("mrssm_fill", L.ptr(self.flat_g), self.flat_g.numel(), 0.0)

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        base, end = self.flat_g.data_ptr(), self.flat_g.data_ptr() + 4 * self.flat_g.numel()
        for p in g["params"]:       # the kernels accumulate into p.grad; it must still be the view of the flat buffer
            if p.grad is None or not (base <= p.grad.data_ptr() < end):
                raise RuntimeError("FusedClipAdam: a parameter's .grad no longer aliases the flat gradient buffer "
                                   "(zero_grad(set_to_none=True) or an external optimiser replaced it)")
        self.step_count += 1
        L.call("mrssm_clip_adam", L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.flat_m), L.ptr(self.flat_v),
               self.flat_p.numel(), self.step_count, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
               float(g["eps"]), float(g["max_grad_norm"]), float(self.grad_scale), L.ptr(self._partial),
               L.ptr(self.grad_norm))
        from . import ops
        ops.bump_weight_version()                            # packed bf16 weight copies are now stale
        for st in self.state.values():
            st["step"] = torch.tensor(float(self.step_count))

    def load_state_dict(self, state_dict):
        """Accept a torch.optim.Adam state_dict (reference checkpoints) — moments are copied into the
        flat buffers.  Parameters the reference never stepped (reward model, a24) have no entry."""
        params = self.param_groups[0]["params"]
        st = state_dict.get("state", {})
        step = 0
        with torch.no_grad():
            for i, p in enumerate(params):
                if i in st:
                    self.state[p]["exp_avg"].copy_(st[i]["exp_avg"])
                    self.state[p]["exp_avg_sq"].copy_(st[i]["exp_avg_sq"])
                    step = max(step, int(st[i]["step"]))
        self.step_count = step
        if state_dict.get("param_groups"):
            for k in ("lr", "betas", "eps"):
                if k in state_dict["param_groups"][0]:
                    self.param_groups[0][k] = state_dict["param_groups"][0][k]
