"""Mirror of ``pmgt/optimizers.py``: ``DenseSparseAdamW`` (dense branch, fused).

PMGT's only embeddings are frozen, so only the dense branch of the reference
optimizer (optimizers.py:256-270) is ever exercised by the pre-training path:

    p *= 1 - lr * weight_decay
    m  = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2
    p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)

When every parameter (and its gradient) is a view into one flat buffer -- which
is how ``pmgt_b200.PMGT`` stores them -- a step is ONE kernel launch over the
flat buffers; otherwise it is one launch per tensor.
"""
import torch
from torch.optim import Optimizer

from . import ops


class DenseSparseAdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if not 0.0 <= weight_decay:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._flat = None  # (p_flat, m_flat, v_flat, decay_mask, step, span)

    # -- flat fast path ---------------------------------------------------------------
    def _try_flat(self):
        """All params contiguous in one allocation, same lr/betas/eps, decay in {wd, 0}."""
        ps = [(p, g) for g in self.param_groups for p in g["params"]]
        if not ps or any(p.dtype != torch.float32 or not p.is_cuda for p, _ in ps):
            return None
        g0 = self.param_groups[0]
        for g in self.param_groups:
            if g["lr"] != g0["lr"] or g["betas"] != g0["betas"] or g["eps"] != g0["eps"]:
                return None
        wds = sorted({g["weight_decay"] for g in self.param_groups if g["weight_decay"] != 0.0})
        if len(wds) > 1:
            return None
        ps.sort(key=lambda t: t[0].data_ptr())
        base = ps[0][0].data_ptr()
        end = max(p.data_ptr() + 4 * p.numel() for p, _ in ps)
        span = (end - base) // 4
        storage_ok = all(p.untyped_storage().data_ptr() == ps[0][0].untyped_storage().data_ptr() for p, _ in ps)
        if not storage_ok or span > 2 * sum(p.numel() for p, _ in ps) + 1024:
            return None
        dev = ps[0][0].device
        p0 = ps[0][0]
        p_flat = torch.as_strided(p0.data, (span,), (1,), storage_offset=p0.storage_offset())
        mask = torch.zeros(span, dtype=torch.uint8)
        for p, g in ps:
            if g["weight_decay"] != 0.0:
                o = (p.data_ptr() - base) // 4
                mask[o: o + p.numel()] = 1
        m = torch.zeros(span, dtype=torch.float32, device=dev)
        v = torch.zeros(span, dtype=torch.float32, device=dev)
        for p, _ in ps:  # expose the moments through the usual per-parameter state
            o = (p.data_ptr() - base) // 4
            st = self.state[p]
            if "exp_avg" in st:  # resume from per-tensor state
                m[o: o + p.numel()].view_as(p).copy_(st["exp_avg"])
                v[o: o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
            st.setdefault("step", 0)
            st["exp_avg"] = m[o: o + p.numel()].view_as(p)
            st["exp_avg_sq"] = v[o: o + p.numel()].view_as(p)
        return dict(base=base, span=span, p=p_flat, m=m, v=v, mask=mask.to(dev), wd=(wds[0] if wds else 0.0),
                    params=[p for p, _ in ps])

    def _flat_grad(self, fl):
        """The gradients as one flat vector aligned with ``fl['p']`` (zero-copy when they
        are views of one arena laid out like the parameters)."""
        ps = fl["params"]
        g0 = ps[0].grad
        if g0 is None:
            return None
        gbase = g0.data_ptr()
        ok = True
        for p in ps:
            if p.grad is None or p.grad.dtype != torch.float32 or \
                    p.grad.data_ptr() - gbase != p.data_ptr() - fl["base"] or not p.grad.is_contiguous():
                ok = False
                break
        if ok and g0.untyped_storage().nbytes() - 4 * g0.storage_offset() >= 4 * fl["span"]:
            return torch.as_strided(g0, (fl["span"],), (1,), storage_offset=g0.storage_offset())
        flat = torch.zeros(fl["span"], dtype=torch.float32, device=g0.device)
        for p in ps:
            if p.grad is not None:
                o = (p.data_ptr() - fl["base"]) // 4
                flat[o: o + p.numel()].view_as(p).copy_(p.grad)
        return flat

    def flat_views(self):
        """(params, grads) as flat fp32 vectors, or None when the fast path does not apply.
        Used by the trainer for a single gradient allreduce."""
        if self._flat is None or self._flat["params"][0].data_ptr() != self._flat["base"]:
            self._flat = self._try_flat()
        if self._flat is None:
            return None
        return self._flat["p"], self._flat_grad(self._flat)

    def load_state_dict(self, state_dict):
        """``torch.optim.Optimizer.load_state_dict`` + rebuild of the flat moment buffers: the flat fast path caches its
        own ``m`` / ``v`` vectors, so after a resume they are re-derived from the loaded per-parameter state."""
        super().load_state_dict(state_dict)
        self._flat = None

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, grad_scale_dev=None, flat_grad=None):
        """``flat_grad``: the flat gradient vector to step with (as returned by ``flat_views()[1]``, e.g. after an
        allreduce).  Passing it makes the step use exactly the reduced buffer even when ``flat_views`` had to COPY the
        per-parameter gradients into a temporary (gradients that are not zero-copy views of one arena)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None and p.grad.is_sparse:
                    raise NotImplementedError("sparse gradients: PMGT's embeddings are frozen, the sparse branch of "
                                              "the reference optimizer is outside the pre-training path")
        if flat_grad is not None:
            if self._flat is None:
                self.flat_views()
            fv = (self._flat["p"], flat_grad) if self._flat is not None else None
        else:
            fv = self.flat_views()
        if fv is not None and fv[1] is not None:
            fl = self._flat
            if fv[1].numel() != fl["span"]:
                raise ValueError("flat_grad does not match the flat parameter buffer")
            g0 = self.param_groups[0]
            step = self.state[fl["params"][0]]["step"] + 1
            for p in fl["params"]:
                self.state[p]["step"] = step
            ops.adamw_step(fl["p"], fv[1], fl["m"], fl["v"], fl["mask"], g0["lr"], g0["betas"][0], g0["betas"][1],
                           g0["eps"], fl["wd"], step, grad_scale, grad_scale_dev)
            return loss
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("pmgt_b200.DenseSparseAdamW runs on CUDA parameters only (no CPU fallback)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                ops.adamw_step(p.data, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"], None, group["lr"], b1, b2,
                               group["eps"], group["weight_decay"], st["step"], grad_scale, grad_scale_dev)
        return loss
