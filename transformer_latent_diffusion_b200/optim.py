"""``FusedAdamEMA`` — ``optimizer.step()`` + ``update_ema`` of the reference's training loop (tld/train.py:170,172-173,55-58)
as ONE CUDA kernel over flat fp32 arenas (``tld_adam_ema_step``, include/tld_b200.h).

The library's backward leaves every parameter gradient in one device arena (``tld_train_grad_layout``).  This optimiser lays
the model's parameters, both Adam moments and (on rank 0) the EMA weights out in arenas with exactly that layout, re-points
the ``nn.Parameter``s (and the EMA model's) at views of them, and then one elementwise pass updates everything: 28-36 bytes
of HBM traffic per parameter instead of torch's multi-tensor Adam plus two foreach EMA passes.

Arithmetic = ``torch.optim.Adam`` (lr, betas, eps, L2 weight decay; no amsgrad) followed by
``ema = alpha * ema + (1 - alpha) * param``.  ``state_dict()`` / ``load_state_dict()`` use ``torch.optim.Adam``'s layout, so
a checkpoint's ``"opt_state"`` (tld/train.py:147,99) moves between this optimiser and the reference's in both directions.
There is no CPU path: the model must live on a CUDA device.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib


class FusedAdamEMA:
    def __init__(self, model: nn.Module, lr: float = 3e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 ema_model: Optional[nn.Module] = None, alpha: float = 0.999):
        self.model, self.ema_model, self.alpha = model, ema_model, float(alpha)
        self._names = [k for k, _ in model.named_parameters()]
        self.param_groups = [{"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay, "amsgrad": False,
                              "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                              "fused": None, "decoupled_weight_decay": False, "params": list(range(len(self._names)))}]
        self.step_count = 0
        self._arenas = None      # dict(p, m, v, ema, offsets, total)
        self._pending_state = None

    # ------------------------------------------------------------------ arenas
    def _params(self):
        return [p for _, p in self.model.named_parameters()]

    def _materialize(self):
        from .train import _grad_arena

        params = self._params()
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.TldError("FusedAdamEMA needs the model on a CUDA device (no CPU fallback)")
        if any(p.dtype != torch.float32 for p in params):
            raise _lib.TldError("FusedAdamEMA keeps fp32 master weights: the model must be float32")
        lib = _lib.load()
        h = self.model._ensure_handle(dev)
        _lib.check(lib.tld_train_prepare(h), "tld_train_prepare")
        view, offsets = _grad_arena(self.model, h, dev, self._names)
        total = view.numel()
        ar = self._arenas
        if ar is not None and ar["total"] == total and ar["p"].device != dev:   # the model moved: the moments follow it
            for k in ("p", "m", "v", "ema"):
                ar[k] = ar[k].to(dev) if ar[k] is not None else None
        if ar is None or ar["total"] != total:
            ar = {"total": total, "offsets": offsets, "p": torch.zeros(total, device=dev), "m": torch.zeros(total, device=dev),
                  "v": torch.zeros(total, device=dev), "ema": None}
            self._arenas = ar
        with torch.no_grad():
            for p, (off, n) in zip(params, offsets):       # adopt the current values, then alias the arena
                if p.data_ptr() != ar["p"].data_ptr() + 4 * off:
                    ar["p"][off:off + n].copy_(p.detach().reshape(-1))
                    p.data = ar["p"][off:off + n].view(p.shape)
            if self.ema_model is not None:
                if ar["ema"] is None:
                    ar["ema"] = torch.zeros(total, device=dev)
                for (k, p), (off, n) in zip(self.ema_model.named_parameters(), offsets):
                    if p.data_ptr() != ar["ema"].data_ptr() + 4 * off:
                        ar["ema"][off:off + n].copy_(p.detach().to(device=dev, dtype=torch.float32).reshape(-1))
                        p.data = ar["ema"][off:off + n].view(p.shape)
            if self._pending_state is not None:
                self._apply_state(self._pending_state)
                self._pending_state = None
        return ar

    def _aliased(self) -> bool:
        ar = self._arenas
        if ar is None:
            return False
        base = ar["p"].data_ptr()
        ok = all(p.data_ptr() == base + 4 * off for p, (off, _) in zip(self._params(), ar["offsets"]))
        if ok and self.ema_model is not None:
            eb = ar["ema"].data_ptr() if ar["ema"] is not None else None
            ok = eb is not None and all(p.data_ptr() == eb + 4 * off
                                        for p, (off, _) in zip(self.ema_model.parameters(), ar["offsets"]))
        return ok

    # ------------------------------------------------------------------ torch.optim surface
    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self._params():
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_().zero_()

    @torch.no_grad()
    def step(self) -> None:
        ar = self._arenas if self._aliased() else self._materialize()   # .to() / load on another device un-aliases
        params, offsets = self._params(), ar["offsets"]
        grads = [p.grad for p in params]
        if any(g is None for g in grads):
            raise _lib.TldError("FusedAdamEMA.step: every parameter needs a gradient (run backward first)")
        # the backward hands out views of ONE flat buffer laid out like the arena (the snapshot, or the arena itself with
        # module.grad_views): use it in place; gradients that were replaced / re-allocated are gathered first
        base = grads[0].data_ptr() - 4 * offsets[0][0]
        if all(g.dtype == torch.float32 and g.is_contiguous() and g.data_ptr() == base + 4 * off
               for g, (off, _) in zip(grads, offsets)):
            gptr = base
            keep = grads[0]
        else:
            flat = torch.zeros(ar["total"], device=ar["p"].device)
            for g, (off, n) in zip(grads, offsets):
                flat[off:off + n].copy_(g.reshape(-1))
            gptr, keep = flat.data_ptr(), flat
        g = self.param_groups[0]
        self.step_count += 1
        dev = ar["p"].device
        with torch.cuda.device(dev):
            _lib.check(_lib.load().tld_adam_ema_step(
                ar["p"].data_ptr(), gptr, ar["m"].data_ptr(), ar["v"].data_ptr(),
                ar["ema"].data_ptr() if ar["ema"] is not None else None, ar["total"], float(g["lr"]), float(g["betas"][0]),
                float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self.step_count, self.alpha, 1.0,
                _lib.current_stream_ptr(dev)), "tld_adam_ema_step")
        del keep

    # ------------------------------------------------------------------ checkpoint compatibility with torch.optim.Adam
    def state_dict(self) -> dict:
        state = {}
        if self._arenas is not None and self.step_count > 0:
            ar = self._arenas
            for i, (off, n) in enumerate(ar["offsets"]):
                shape = self._params()[i].shape
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": ar["m"][off:off + n].view(shape).clone(),
                            "exp_avg_sq": ar["v"][off:off + n].view(shape).clone()}
        return {"state": state, "param_groups": [dict(self.param_groups[0])]}

    def _apply_state(self, sd: dict) -> None:
        ar = self._arenas
        steps = set()
        for i, st in sd["state"].items():
            off, n = ar["offsets"][int(i)]
            ar["m"][off:off + n].copy_(st["exp_avg"].to(ar["m"].device, torch.float32).reshape(-1))
            ar["v"][off:off + n].copy_(st["exp_avg_sq"].to(ar["v"].device, torch.float32).reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise _lib.TldError("FusedAdamEMA.load_state_dict: per-parameter step counts differ")
        self.step_count = steps.pop() if steps else 0

    def load_state_dict(self, sd: dict) -> None:
        grp = sd["param_groups"][0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in grp:
                self.param_groups[0][k] = tuple(grp[k]) if k == "betas" else grp[k]
        if grp.get("amsgrad"):
            raise _lib.TldError("FusedAdamEMA: amsgrad is not supported")
        if self._arenas is None:
            # arenas are created at the first step (they need the device + library handle); snapshot now: torch's state_dict()
            # hands out the live step / moment tensors, which the source optimiser keeps updating in place
            self._pending_state = {"state": {i: {k: (t.clone() if torch.is_tensor(t) else t) for k, t in st.items()}
                                             for i, st in sd["state"].items()}, "param_groups": sd["param_groups"]}
            steps = {int(float(st["step"])) for st in sd["state"].values()}
            self.step_count = steps.pop() if len(steps) == 1 else 0
        else:
            with torch.no_grad():
                self._apply_state(sd)
