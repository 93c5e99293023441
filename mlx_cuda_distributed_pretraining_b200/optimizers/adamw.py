"""AdamW with mlx.optimizers.AdamW's semantics (the optimizer core/training.py:821 instantiates
for `optimizer: adamw`; formula mirrored in optimizers/enhanced_optimizers.py:157-184):

    p *= 1 - lr*wd ;  m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  p -= lr * m / (sqrt(v) + eps)

mlx 0.25.0 has bias_correction=False by default (third-party behaviour, stated in DESIGN.md);
`bias_correction=True` gives lr/(1-b1^t) and sqrt(v)/sqrt(1-b2^t).  One fused multi-tensor launch
covers the whole flat parameter range (fp32 master + bf16 shadow refreshed in the same pass).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple, Union

import torch

from .. import ops
from ..flat import ParamStore, get_store


class AdamW:
    def __init__(self, learning_rate: Union[float, Callable] = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.01, bias_correction: bool = False):
        self._learning_rate = learning_rate
        self.betas = tuple(betas)
        self.eps = eps
        self.weight_decay = weight_decay
        self.bias_correction = bias_correction
        self.state: Dict[str, Dict[str, torch.Tensor]] = {}
        self.count = 0
        self.grad_scale = 1.0
        self.use_accumulated = False
        self._store: Optional[ParamStore] = None
        self._range = None

    @property
    def learning_rate(self) -> float:
        return self._lr(self.count)

    def _lr(self, count: int) -> float:
        lr = self._learning_rate(count) if callable(self._learning_rate) else self._learning_rate
        return float(lr)

    def init_range(self, store: ParamStore, lo: int, hi: int) -> None:
        if self._store is store and self._range == (lo, hi):
            return
        self._store, self._range = store, (lo, hi)
        self._m = torch.zeros(store.total, dtype=torch.float32, device=store.device)
        self._v = torch.zeros(store.total, dtype=torch.float32, device=store.device)
        self.state = {}
        for name, (o, _shape) in store.index.items():
            if lo <= o < hi:
                self.state[name] = {"m": store.view(self._m, name), "v": store.view(self._v, name)}

    def init(self, model) -> None:
        store = get_store(model)
        self.init_range(store, 0, store.vec_end)

    @torch.no_grad()
    def update_range(self, store: ParamStore, lo: int, hi: int, gsrc: torch.Tensor, grad_scale: float) -> None:
        lr = self._lr(self.count)
        t = self.count + 1
        b1, b2 = self.betas
        bc1 = 1.0 - b1 ** t if self.bias_correction else 1.0
        bc2 = 1.0 - b2 ** t if self.bias_correction else 1.0
        p16 = store.shadow[lo:hi] if store.mixed else None
        ops.adamw(store.master[lo:hi], p16, gsrc[lo:hi], self._m[lo:hi], self._v[lo:hi], lr, b1, b2,
                  self.eps, self.weight_decay, bc1, bc2, grad_scale)
        self.count += 1

    @torch.no_grad()
    def update(self, model, gradients=None) -> None:
        self.init(model)
        store = self._store
        if gradients is not None:
            store.load_gradients(gradients)
        gsrc = store.acc if self.use_accumulated else store.grad
        lo, hi = self._range
        self.update_range(store, lo, hi, gsrc, self.grad_scale)

    def apply_gradients(self, gradients, model):
        self.update(model, gradients)
        return model

    def step(self, model) -> None:
        self.update(model, None)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {}
        for n, s in self.state.items():
            out[f"{n}.m"] = s["m"]
            out[f"{n}.v"] = s["v"]
        return out
