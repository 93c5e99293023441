"""Shampoo with the reference's surface (optimizers/shampoo.py:20-45,143-378) on the sm_100a GEMM
engine.

    opt = Shampoo(learning_rate=float|callable, params=ShampooParams(...))
    opt.update(model, gradients)     # core/training.py:1690,1700

Per step (shampoo.py:314-378, with the wiring defects D1/D2/D9 of SURVEY 2.3 fixed):
  count += 1; lr = sched(count)
  grafting direction d = Adam step (mlx Adam, no bias correction) computed WITHOUT applying it
  L = b2 L + (1-b2) G G^T ; R = b2 R + (1-b2) G^T G       on G[:1024,:1024]   (:229-255)
  every `update_period` steps after `start_preconditioning_step`:
        P = matrix_inverse_pth_root(stat)  -- the reference's literal formula (:88-126):
        Z0 = (M+eps I)/tr ; 6x Z <- Z + Z Z / p ; P = Z * tr^(-1/p^2)        (not a true inverse root)
  m = b1 m + (1-b1) g ; mhat = m/(1-b1^t) ; pre = mhat with pre[:1024,:1024] = PL mhat PR   (:257-295)
  upd = -lr*pre rescaled to ||d||_F (grafting, :297-312) ; decoupled wd: upd -= lr*wd*p     (:372-373)

All matrix products are batched over the same-shape parameters of a flat.ParamStore group and run
on tcgen05.  fp32 matrices enter the tensor cores as bf16 hi+lo pairs (hi*hi + hi*lo + lo*hi, three
accumulating GEMMs, ~16 mantissa bits) so the statistics/roots stay close to the reference's fp32
matmuls; gradients (already bf16 under mixed precision) are used as they are.
Parameters whose preconditioned block side is not a multiple of 8 (only the byte-level embedding
[259, hidden]) are not preconditioned (momentum + grafting only) -- see DESIGN.md.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Union

import torch

from .. import ops
from ..flat import MatGroup, ParamStore, get_store


@dataclass
class ShampooParams:
    beta1: float = 0.9
    beta2: float = 0.99
    epsilon: float = 1e-8
    weight_decay: float = 0.0
    update_period: int = 1
    start_preconditioning_step: int = 10
    preconditioner_epsilon: float = 1e-6
    max_preconditioner_dim: int = 1024
    exponent_override: float = 0.75
    use_bias_correction: bool = True
    grafting_optimizer: str = "adam"
    use_decoupled_weight_decay: bool = True

    def __post_init__(self):
        assert 0.0 <= self.beta1 < 1.0, "beta1 must be in [0, 1)"
        assert 0.0 <= self.beta2 < 1.0, "beta2 must be in [0, 1)"
        assert self.epsilon > 0.0, "epsilon must be positive"
        assert self.update_period > 0, "update_period must be positive"
        assert self.start_preconditioning_step >= 0, "start_preconditioning_step must be non-negative"
        assert self.max_preconditioner_dim > 0, "max_preconditioner_dim must be positive"
        assert 0.0 < self.exponent_override <= 1.0, "exponent_override must be in (0, 1]"
        assert self.grafting_optimizer in ["sgd", "adam", "momentum"], \
            "grafting_optimizer must be one of 'sgd', 'adam', 'momentum'"


class _GroupState:
    def __init__(self, g: MatGroup, cap: int, dev):
        self.k1, self.k2 = min(g.rows, cap), min(g.cols, cap)
        self.enabled = self.k1 % 8 == 0 and self.k2 % 8 == 0
        if not self.enabled:
            return
        b = g.batch
        self.L = torch.zeros(b, self.k1, self.k1, dtype=torch.float32, device=dev)
        self.R = torch.zeros(b, self.k2, self.k2, dtype=torch.float32, device=dev)
        self.PLh = torch.zeros(b, self.k1, self.k1, dtype=torch.bfloat16, device=dev)
        self.PLl = torch.zeros_like(self.PLh)
        self.PRh = torch.zeros(b, self.k2, self.k2, dtype=torch.bfloat16, device=dev)
        self.PRl = torch.zeros_like(self.PRh)
        self.has_precond = False


class Shampoo:
    def __init__(self, learning_rate: Union[float, Callable] = 0.01, params: Optional[ShampooParams] = None,
                 use_distributed: bool = False):
        self.params = params or ShampooParams()
        self._learning_rate = learning_rate
        self.use_distributed = use_distributed
        self.state: Dict[str, Dict[str, torch.Tensor]] = {}
        self.count = 0
        self.grad_scale = 1.0
        self.use_accumulated = False
        self._store: Optional[ParamStore] = None
        if self.params.weight_decay > 0.0 and not self.params.use_decoupled_weight_decay:
            raise ValueError("coupled weight decay is not supported on the fused path (reference factory "
                             "always sets use_decoupled_weight_decay=True, core/training.py:855)")

    def _lr(self, count: int) -> float:
        lr = self._learning_rate(count) if callable(self._learning_rate) else self._learning_rate
        return float(lr)

    @property
    def learning_rate(self) -> float:
        return self._lr(self.count)

    # ------------------------------------------------------------------------------------------
    def init(self, model) -> None:
        store = get_store(model)
        if self._store is store:
            return
        self._store = store
        dev, n = store.device, store.total
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        self._mom = torch.zeros(n, **f32)
        self._gm = torch.zeros(n, **f32)
        self._gv = torch.zeros(n, **f32)
        self._d = torch.zeros(n, **f32)
        self._pre = torch.zeros(n, **f32)
        self._mh = torch.zeros(n, **bf)
        self._ml = torch.zeros(n, **bf)
        self._g16 = None
        cap = self.params.max_preconditioner_dim
        self._gs: List[_GroupState] = [_GroupState(g, cap, dev) for g in store.mat_groups]
        en = [(g, s) for g, s in zip(store.mat_groups, self._gs) if s.enabled]
        max_t = max((g.batch * s.k1 * s.k2 for g, s in en), default=8)
        max_z = max((g.batch * max(s.k1, s.k2) ** 2 for g, s in en), default=8)
        max_b = max([g.batch for g in store.mat_groups] + [1])
        self._T = torch.empty(max_t, **f32)
        self._Th = torch.empty(max_t, **bf)
        self._Tl = torch.empty(max_t, **bf)
        self._Z = [torch.empty(max_z, **f32), torch.empty(max_z, **f32)]
        self._Zh = torch.empty(max_z, **bf)
        self._Zl = torch.empty(max_z, **bf)
        self._n1 = torch.empty(max_b, **f32)
        self._n2 = torch.empty(max_b, **f32)
        self.state = {}
        for name in store.index:
            self.state[name] = {"momentum": store.view(self._mom, name), "graft_m": store.view(self._gm, name),
                                "graft_v": store.view(self._gv, name)}
        for g, s in zip(store.mat_groups, self._gs):
            if s.enabled:
                for i, name in enumerate(g.names):
                    self.state[name]["statistics.0"] = s.L[i]
                    self.state[name]["statistics.1"] = s.R[i]

    # ------------------------------------------------------------------------------------------
    def _gemm3(self, a_mn, b_mn, M, N, K, batch, Ah, Al, lda, sA, Bh, Bl, ldb, sB, C, ldc, sC, D, ldd, sD,
               alpha, beta, av=None, bv=None):
        """D = alpha*(Ah Bh + Ah Bl + Al Bh) + beta*C  (fp32 out): bf16x3 product of two split matrices."""
        ops.gemm_raw(a_mn, b_mn, M, N, K, batch, Ah, lda, sA, Bh, ldb, sB, C, ldc, sC, D, ldd, sD,
                     alpha, beta, av, bv)
        ops.gemm_raw(a_mn, b_mn, M, N, K, batch, Ah, lda, sA, Bl, ldb, sB, D, ldd, sD, D, ldd, sD,
                     alpha, 1.0, av, None)
        ops.gemm_raw(a_mn, b_mn, M, N, K, batch, Al, lda, sA, Bh, ldb, sB, D, ldd, sD, D, ldd, sD,
                     alpha, 1.0, av, None)

    def _inverse_pth_root(self, stat: torch.Tensor, Ph: torch.Tensor, Pl: torch.Tensor) -> None:
        """MatrixSqrt.matrix_inverse_pth_root (shampoo.py:88-126), batched over stat [b,k,k]."""
        hp = self.params
        b, k, _ = stat.shape
        p = hp.exponent_override
        kk = k * k
        # tiny per-matrix scalars (trace, final scale) stay in torch; all k^3 work is tcgen05
        diag = stat.diagonal(dim1=1, dim2=2)
        tr = diag.sum(-1) + k * hp.preconditioner_epsilon
        inv_tr = (1.0 / tr).contiguous()
        scale_fin = tr.pow(-1.0 / (p * p)).contiguous()
        Z, Z2 = self._Z[0][:b * kk], self._Z[1][:b * kk]
        Zh, Zl = self._Zh[:b * kk], self._Zl[:b * kk]
        # Z0 = (M + eps I)/tr  (elementwise set-up, once per update_period), then its bf16 hi/lo split
        Z.copy_(((stat + hp.preconditioner_epsilon * torch.eye(k, device=stat.device)) *
                 inv_tr[:, None, None]).reshape(-1))
        ops.check(ops.lib().b200_split_bf16(Z.data_ptr(), k, Zh.data_ptr(), Zl.data_ptr(), k, b * k, k, 1.0, 0.0,
                                            ops._stream()), "b200_split_bf16")
        iters = 6
        for it in range(iters):
            last = it == iters - 1
            av = scale_fin if last else None
            self._gemm3(False, False, k, k, k, b, Zh, Zl, k, kk, Zh, Zl, k, kk, Z, k, kk, Z2, k, kk,
                        1.0 / p, 1.0, av, av)
            Z, Z2 = Z2, Z
            if not last:
                ops.check(ops.lib().b200_split_bf16(Z.data_ptr(), k, Zh.data_ptr(), Zl.data_ptr(), k, b * k, k,
                                                    1.0, 0.0, ops._stream()), "b200_split_bf16")
        ops.check(ops.lib().b200_split_bf16(Z.data_ptr(), k, Ph.data_ptr(), Pl.data_ptr(), k, b * k, k, 1.0, 0.0,
                                            ops._stream()), "b200_split_bf16")

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def update(self, model, gradients=None) -> None:
        self.init(model)
        store, hp = self._store, self.params
        if gradients is not None:
            store.load_gradients(gradients)
        gsrc = store.acc if self.use_accumulated else store.grad
        self.count += 1
        t = self.count
        lr = self._lr(t)
        gs = self.grad_scale
        end = store.vec_end
        # grafting direction (shampoo.py:162-178,326): computed, not applied
        if hp.grafting_optimizer == "adam":
            ops.adam_direction(self._d[:end], gsrc[:end], self._gm[:end], self._gv[:end], lr, hp.beta1, hp.beta2,
                               hp.epsilon, 1.0, 1.0, gs)
        elif hp.grafting_optimizer == "momentum":   # mlx SGD momentum: v = mu v + g ; step = -lr v
            self._gm[:end].mul_(hp.beta1).add_(gsrc[:end].float(), alpha=gs)
            torch.mul(self._gm[:end], -lr, out=self._d[:end])
        else:
            torch.mul(gsrc[:end].float(), -lr * gs, out=self._d[:end])
        inv_bc = 1.0 / (1.0 - hp.beta1 ** t) if hp.use_bias_correction else 1.0
        # pre = -lr * mhat (the reference's `update = -lr * preconditioned_grad`, shampoo.py:365); the
        # preconditioned blocks are overwritten below with -lr * PL mhat PR
        ops.ema_split(gsrc[:end], self._mom[:end], self._pre[:end], self._mh[:end], self._ml[:end], hp.beta1, gs,
                      inv_bc, -lr)
        if gsrc.dtype != torch.bfloat16:
            if self._g16 is None:
                self._g16 = torch.empty(store.total, dtype=torch.bfloat16, device=store.device)
            ops.check(ops.lib().b200_split_bf16(gsrc.data_ptr(), 8, self._g16.data_ptr(), None, 8,
                                                store.mat_end // 8, 8, 1.0, 0.0, ops._stream()), "b200_split_bf16")
            g16 = self._g16
        else:
            g16 = gsrc
        decay = 1.0 - lr * hp.weight_decay if hp.weight_decay > 0.0 else 1.0
        do_root = t >= hp.start_preconditioning_step and t % hp.update_period == 0
        for g, s in zip(store.mat_groups, self._gs):
            lo, n, rc = g.offset, g.numel, g.rows * g.cols
            b, c = g.batch, g.cols
            if s.enabled:
                k1, k2 = s.k1, s.k2
                G = g16[lo:lo + n]
                a = (1.0 - hp.beta2) * gs * gs
                ops.gemm_raw(False, False, k1, k1, k2, b, G, c, rc, G, c, rc, s.L, k1, k1 * k1, s.L, k1, k1 * k1,
                             a, hp.beta2)
                ops.gemm_raw(True, True, k2, k2, k1, b, G, c, rc, G, c, rc, s.R, k2, k2 * k2, s.R, k2, k2 * k2,
                             a, hp.beta2)
                if do_root:
                    self._inverse_pth_root(s.L, s.PLh, s.PLl)
                    self._inverse_pth_root(s.R, s.PRh, s.PRl)
                    s.has_precond = True
                if t >= hp.start_preconditioning_step and s.has_precond:
                    mh, ml = self._mh[lo:lo + n], self._ml[lo:lo + n]
                    T, Th, Tl = self._T[:b * k1 * k2], self._Th[:b * k1 * k2], self._Tl[:b * k1 * k2]
                    # T = PL @ mhat[:k1,:k2]   (mhat block read in place as an MN-major operand)
                    self._gemm3(False, True, k1, k2, k1, b, s.PLh, s.PLl, k1, k1 * k1, mh, ml, c, rc,
                                None, k2, k1 * k2, T, k2, k1 * k2, 1.0, 0.0)
                    ops.check(ops.lib().b200_split_bf16(T.data_ptr(), k2, Th.data_ptr(), Tl.data_ptr(), k2, b * k1,
                                                        k2, 1.0, 0.0, ops._stream()), "b200_split_bf16")
                    # pre[:k1,:k2] = T @ PR  written straight into the update buffer (ldd = cols)
                    pre = self._pre[lo:lo + n]
                    self._gemm3(False, True, k1, k2, k2, b, Th, Tl, k2, k1 * k2, s.PRh, s.PRl, k2, k2 * k2,
                                None, c, rc, pre, c, rc, -lr, 0.0)
            self._graft_and_apply(store, lo, rc, b, lr, decay)
        for e in store.vec_entries:
            self._graft_and_apply(store, e.offset, e.numel, 1, lr, decay)

    def _graft_and_apply(self, store: ParamStore, lo: int, numel: int, batch: int, lr: float, decay: float) -> None:
        n = numel * batch
        pre, d = self._pre[lo:lo + n], self._d[lo:lo + n]
        n1, n2 = self._n1[:batch], self._n2[:batch]
        ops.sumsq_raw(pre, n1, numel, batch)
        ops.sumsq_raw(d, n2, numel, batch)
        sn, gn = n1.sqrt(), n2.sqrt()      # ||upd|| with upd = -lr*pre (held in self._pre), ||graft step||
        # _apply_grafting (shampoo.py:297-312): sn==0 -> graft step ; gn==0 -> upd ; else upd*(gn/sn).
        # fp32 sqrt(sum(x^2)) overflows exactly like mx.linalg.norm would
        coef = torch.where(sn == 0, torch.zeros_like(sn), torch.where(gn == 0, torch.ones_like(sn), gn / sn))
        coef_d = (sn == 0).float()
        p16 = store.shadow[lo:lo + n] if store.mixed else None
        ops.graft_update(store.master[lo:lo + n], p16, pre, d, numel, batch, coef.contiguous(), coef_d.contiguous(),
                         decay)

    def apply_gradients(self, gradients, model):
        self.update(model, gradients)
        return model

    def step(self, model) -> None:
        self.update(model, None)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {}
        for name, st in self.state.items():
            for k, v in st.items():
                out[f"{name}.{k}"] = v
        return out
