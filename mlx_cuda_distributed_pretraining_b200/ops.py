"""PyTorch-tensor wrappers over the C ABI (include/b200_hotpath.h).

PyTorch is plumbing here: it owns device memory and streams; all arithmetic below runs in the
hand-written sm_100a kernels.  Every wrapper validates dtype/contiguity and raises on failure --
there is no eager fallback.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import torch

from ._lib import B200Error, check, lib, require_device

NS_COEFFS = (3.4445, -4.7750, 2.0315)  # optimizers/muon.py:65
NS_EPS = 1e-7                          # optimizers/muon.py:73


def _stream() -> int:
    # raw handle of torch's current stream on the current device; torch.cuda.current_stream() builds a Stream
    # object (~17 us per call measured, ~100 calls per C2 step: tools/host_profile.py)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class KernelTimer:
    """Optional CUDA-event bracketing of kernel families on the launching stream (bench.py's
    live roofline measurement).  Disabled (None) in normal runs: zero overhead."""

    def __init__(self):
        self.events = {}

    def start(self, name: str):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return (name, e0)

    def stop(self, tok) -> None:
        name, e0 = tok
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.events.setdefault(name, []).append((e0, e1))

    def totals_ms(self):
        torch.cuda.synchronize()
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.events.items()}


TIMER: Optional[KernelTimer] = None


def _t0(name: str):
    return TIMER.start(name) if TIMER is not None else None


def _t1(tok) -> None:
    if tok is not None:
        TIMER.stop(tok)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise B200Error(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")


def _is_bf16(t: torch.Tensor, name: str) -> int:
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise ValueError(f"{name}: unsupported dtype {t.dtype} (bf16 or fp32)")


# ------------------------------------------------------------------------------------------------
# dense contraction engine
# ------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         c: Optional[torch.Tensor] = None, alpha: float = 1.0, beta: float = 0.0,
         alpha_vec: Optional[torch.Tensor] = None, beta_vec: Optional[torch.Tensor] = None,
         out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
         force_bn: int = 0) -> torch.Tensor:
    """D = alpha*op(A) op(B) + beta*C, batched over a leading dim.

    a: [.., M, K] (a_mn=False) or [.., K, M] (a_mn=True); b: [.., N, K] or [.., K, N] (b_mn=True).
    """
    _need(a, torch.bfloat16, "a")
    _need(b, torch.bfloat16, "b")
    a3 = a if a.dim() == 3 else a.unsqueeze(0)
    b3 = b if b.dim() == 3 else b.unsqueeze(0)
    batch = a3.shape[0]
    if b3.shape[0] != batch:
        raise ValueError("gemm: batch mismatch")
    (K, M) = a3.shape[1:] if a_mn else a3.shape[1:][::-1]
    (Kb, N) = b3.shape[1:] if b_mn else b3.shape[1:][::-1]
    if K != Kb:
        raise ValueError(f"gemm: K mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((batch, M, N), device=a.device, dtype=out_dtype)
    out_f32 = _is_bf16(out, "out") == 0
    if tuple(out.shape[-2:]) != (M, N) or not out.is_contiguous():
        raise ValueError("gemm: bad out tensor")
    if c is not None:
        if c.dtype != out.dtype or tuple(c.shape[-2:]) != (M, N) or not c.is_contiguous():
            raise ValueError("gemm: c must match out dtype/shape and be contiguous")
    for v, nm in ((alpha_vec, "alpha_vec"), (beta_vec, "beta_vec")):
        if v is not None:
            _need(v, torch.float32, nm)
            if v.numel() != batch:
                raise ValueError(f"gemm: {nm} must have one entry per batch")
    rc = lib().b200_gemm_bf16(
        int(a_mn), int(b_mn), M, N, K, batch,
        a3.data_ptr(), a3.shape[2], a3.shape[1] * a3.shape[2],
        b3.data_ptr(), b3.shape[2], b3.shape[1] * b3.shape[2],
        _ptr(c), N, M * N,
        out.data_ptr(), N, M * N,
        int(out_f32), float(alpha), float(beta), _ptr(alpha_vec), _ptr(beta_vec), int(force_bn),
        _stream())
    check(rc, "b200_gemm_bf16")
    return out.view(M, N) if (a.dim() == 2 and out.dim() == 3) else out


# ------------------------------------------------------------------------------------------------
# Newton-Schulz / Muon pieces
# ------------------------------------------------------------------------------------------------
def ns_workspace_bytes(batch: int, rows: int, cols: int, steps: int) -> int:
    return int(lib().b200_newton_schulz_workspace_bytes(batch, rows, cols, steps))


_REDUCE_WS = {}


def reduce_workspace(device, batch: int) -> torch.Tensor:
    """Scratch of the deterministic grid reductions (b200_reduce_workspace_bytes), one per (device, stream)."""
    key = (str(device), _stream())
    need = int(lib().b200_reduce_workspace_bytes(int(batch)))
    ws = _REDUCE_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, device=device, dtype=torch.uint8)
        _REDUCE_WS[key] = ws
    return ws


def sumsq(x: torch.Tensor, out: Optional[torch.Tensor] = None, zero_first: bool = True) -> torch.Tensor:
    """Per-matrix sum of squares of x [batch, ...] -> fp32 [batch]."""
    xb = x if x.dim() >= 2 else x.unsqueeze(0)
    batch = xb.shape[0] if x.dim() == 3 else 1
    numel = x.numel() // batch
    if out is None:
        out = torch.empty(batch, device=x.device, dtype=torch.float32)
    ws = reduce_workspace(x.device, batch)
    check(lib().b200_sumsq(x.data_ptr(), _is_bf16(x, "x"), out.data_ptr(), numel, batch,
                           int(zero_first), ws.data_ptr(), ws.numel(), _stream()), "b200_sumsq")
    return out


def ns_scales(ss: torch.Tensor, eps: float = NS_EPS) -> Tuple[torch.Tensor, torch.Tensor]:
    inv = torch.empty_like(ss)
    inv2 = torch.empty_like(ss)
    check(lib().b200_ns_scales(ss.data_ptr(), inv.data_ptr(), inv2.data_ptr(), ss.numel(), eps,
                               _stream()), "b200_ns_scales")
    return inv, inv2


def newton_schulz_raw(x_in: torch.Tensor, x_out: torch.Tensor, inv_norm: torch.Tensor,
                      inv_norm_sq: torch.Tensor, workspace: torch.Tensor, steps: int = 5,
                      coeffs: Sequence[float] = NS_COEFFS) -> torch.Tensor:
    """x_in/x_out: bf16 [batch, rows, cols]; workspace: uint8 buffer from ns_workspace_bytes."""
    _need(x_in, torch.bfloat16, "x_in")
    _need(x_out, torch.bfloat16, "x_out")
    batch, rows, cols = x_in.shape
    a, b, c = coeffs
    check(lib().b200_newton_schulz(x_in.data_ptr(), x_out.data_ptr(), batch, rows, cols, steps,
                                   float(a), float(b), float(c), inv_norm.data_ptr(),
                                   inv_norm_sq.data_ptr(), workspace.data_ptr(),
                                   workspace.numel() * workspace.element_size(), _stream()),
          "b200_newton_schulz")
    return x_out


def zeropower_via_newtonschulz5(g: torch.Tensor, steps: int = 5, eps: float = NS_EPS,
                                coeffs: Sequence[float] = NS_COEFFS) -> torch.Tensor:
    """Drop-in for Muon.zeropower_via_newtonschulz5 (optimizers/muon.py:54-83).

    g: [rows, cols] or [batch, rows, cols], fp32 or bf16 on a B200. Returns bf16 of the same shape.
    """
    require_device()
    squeeze = g.dim() == 2
    g3 = (g.unsqueeze(0) if squeeze else g).contiguous()
    batch, rows, cols = g3.shape
    x_in = g3 if g3.dtype == torch.bfloat16 else g3.to(torch.bfloat16)
    # Frobenius norm of the (fp32) input, folded into iteration 1 by the kernel
    inv, inv2 = ns_scales(sumsq(g3 if g3.dtype in (torch.float32, torch.bfloat16) else x_in), eps)
    ws = torch.empty(ns_workspace_bytes(batch, rows, cols, steps), device=g.device, dtype=torch.uint8)
    x_out = torch.empty_like(x_in)
    newton_schulz_raw(x_in, x_out, inv, inv2, ws, steps, coeffs)
    return x_out[0] if squeeze else x_out


def zeropower_groups(gs: Sequence[torch.Tensor], steps: int = 5, eps: float = NS_EPS,
                     coeffs: Sequence[float] = NS_COEFFS):
    """Newton-Schulz of several [batch, rows, cols] groups (different shapes) through the whole-model chain:
    every stage of the iteration is ONE grouped launch over all groups (b200_newton_schulz_multi).
    Returns a list of bf16 tensors shaped like the inputs."""
    from ._lib import NsGroup
    require_device()
    if not 1 <= len(gs) <= 6:
        raise ValueError("zeropower_groups takes 1..6 groups")
    keep, arr, outs = [], (NsGroup * len(gs))(), []
    for i, g in enumerate(gs):
        g3 = g.contiguous()
        if g3.dim() != 3:
            raise ValueError("each group must be [batch, rows, cols]")
        x_in = g3 if g3.dtype == torch.bfloat16 else g3.to(torch.bfloat16)
        inv, inv2 = ns_scales(sumsq(g3), eps)
        x_out = torch.empty_like(x_in)
        keep += [x_in, inv, inv2]
        outs.append(x_out)
        e = arr[i]
        e.x_in, e.x_out = x_in.data_ptr(), x_out.data_ptr()
        e.batch, e.rows, e.cols = x_in.shape
        e.inv_norm, e.inv_norm_sq = inv.data_ptr(), inv2.data_ptr()
        e.peer_out, e.n_peers = None, 0
    a, b, c = coeffs
    need = int(lib().b200_newton_schulz_multi_workspace_bytes(arr, len(gs), steps))
    ws = torch.empty(max(need, 256), device=gs[0].device, dtype=torch.uint8)
    check(lib().b200_newton_schulz_multi(arr, len(gs), steps, float(a), float(b), float(c), ws.data_ptr(),
                                         ws.numel(), _stream()), "b200_newton_schulz_multi")
    return outs


def muon_momentum(g: torch.Tensor, buf: torch.Tensor, u: torch.Tensor, sumsq_out: torch.Tensor,
                  mu: float, nesterov: bool, gscale: float = 1.0) -> None:
    """g,buf,u: [batch, rows, cols] (g bf16|fp32, buf fp32, u bf16); sumsq_out fp32 [batch]."""
    _need(buf, torch.float32, "buf")
    _need(u, torch.bfloat16, "u")
    batch = g.shape[0] if g.dim() == 3 else 1
    numel = g.numel() // batch
    ws = reduce_workspace(g.device, batch)
    check(lib().b200_muon_momentum(g.data_ptr(), _is_bf16(g, "g"), buf.data_ptr(), u.data_ptr(),
                                   sumsq_out.data_ptr(), numel, batch, float(mu), int(nesterov),
                                   float(gscale), ws.data_ptr(), ws.numel(), _stream()), "b200_muon_momentum")


def axpy_update(p32: torch.Tensor, p16: Optional[torch.Tensor], x: torch.Tensor, s: float) -> None:
    _need(p32, torch.float32, "p32")
    check(lib().b200_axpy_update(p32.data_ptr(), _ptr(p16), x.data_ptr(), _is_bf16(x, "x"),
                                 p32.numel(), float(s), _stream()), "b200_axpy_update")


def sgd_momentum(p32, p16, g, buf, mu: float, nesterov: bool, lr: float, gscale: float = 1.0) -> None:
    check(lib().b200_sgd_momentum(p32.data_ptr(), _ptr(p16), g.data_ptr(), _is_bf16(g, "g"),
                                  buf.data_ptr(), p32.numel(), float(mu), int(nesterov), float(lr),
                                  float(gscale), _stream()), "b200_sgd_momentum")


def adamw(p32, p16, g, m, v, lr: float, b1: float, b2: float, eps: float, wd: float,
          bc1: float = 1.0, bc2: float = 1.0, gscale: float = 1.0) -> None:
    check(lib().b200_adamw(p32.data_ptr(), _ptr(p16), g.data_ptr(), _is_bf16(g, "g"), m.data_ptr(),
                           v.data_ptr(), p32.numel(), float(lr), float(b1), float(b2), float(eps),
                           float(wd), float(bc1), float(bc2), float(gscale), _stream()),
          "b200_adamw")


def adam_direction(d, g, m, v, lr, b1, b2, eps, bc1=1.0, bc2=1.0, gscale=1.0) -> None:
    check(lib().b200_adam_direction(d.data_ptr(), g.data_ptr(), _is_bf16(g, "g"), m.data_ptr(),
                                    v.data_ptr(), d.numel(), float(lr), float(b1), float(b2),
                                    float(eps), float(bc1), float(bc2), float(gscale), _stream()),
          "b200_adam_direction")


def clip_accum(g: torch.Tensor, acc: torch.Tensor, clip: float, scale: float, init: bool) -> None:
    _need(acc, torch.float32, "acc")
    check(lib().b200_clip_accum(g.data_ptr(), _is_bf16(g, "g"), acc.data_ptr(), g.numel(),
                                float(clip), float(scale), int(init), _stream()), "b200_clip_accum")


def split_bf16(src: torch.Tensor, rows: int, cols: int, hi: torch.Tensor,
               lo: Optional[torch.Tensor], scale: float = 1.0, diag_add: float = 0.0) -> None:
    """hi(+lo) = bf16 split of (src[:rows,:cols] + diag_add*I)*scale; src fp32 2-D (strided rows)."""
    _need(hi, torch.bfloat16, "hi")
    check(lib().b200_split_bf16(src.data_ptr(), src.stride(0), hi.data_ptr(), _ptr(lo),
                                hi.stride(0), rows, cols, float(scale), float(diag_add), _stream()),
          "b200_split_bf16")


# ------------------------------------------------------------------------------------------------
# RMSNorm / RoPE / attention as autograd functions
# ------------------------------------------------------------------------------------------------
# Deferred RMSNorm weight gradients: a norm's backward leaves its per-CTA partial sums in a workspace; ONE batched
# launch at the end of the backward pass (an autograd-engine callback) folds all of them into the weights' .grad
# buffers in place.  Used when the weight is a leaf with a preallocated contiguous .grad (the flat ParamStore's
# case); any other weight takes the immediate path and gets its gradient from autograd as usual.
_PENDING_DW: list = []
_DW_DEFER = os.environ.get("B200_DEFER_DW", "1") != "0"


def _dw_target(w: torch.Tensor) -> Optional[torch.Tensor]:
    g = w.grad if (_DW_DEFER and w.is_leaf and w.requires_grad) else None
    if g is None or not g.is_contiguous() or g.dtype not in (torch.bfloat16, torch.float32) or g.shape != w.shape:
        return None
    return g


def flush_deferred_dw() -> None:
    """Runs the pending weight-gradient reductions now (normally called by the autograd engine at the end of the
    backward pass that queued them)."""
    if not _PENDING_DW:
        return
    from ._lib import DwJob
    jobs, _PENDING_DW[:] = list(_PENDING_DW), []
    by_h = {}
    for ws, grad, nparts in jobs:
        by_h.setdefault(grad.numel(), []).append((ws, grad, nparts))
    for H, group in by_h.items():
        arr = (DwJob * len(group))()
        for i, (ws, grad, nparts) in enumerate(group):
            arr[i].partials, arr[i].dw, arr[i].n_partials = ws.data_ptr(), grad.data_ptr(), nparts
            arr[i].dw_is_bf16, arr[i].accumulate = int(grad.dtype == torch.bfloat16), 1
        check(lib().b200_rmsnorm_dw_reduce(arr, len(group), H, _stream()), "b200_rmsnorm_dw_reduce")


def _defer_dw(ws: torch.Tensor, grad: torch.Tensor, nparts: int) -> None:
    if not _PENDING_DW:
        torch.autograd.Variable._execution_engine.queue_callback(flush_deferred_dw)
    _PENDING_DW.append((ws, grad, nparts))


class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        xs = x.contiguous()
        rows, H = xs.numel() // xs.shape[-1], xs.shape[-1]
        y = torch.empty_like(xs)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        wc = w.contiguous().to(xs.dtype)
        check(lib().b200_rmsnorm_fwd(xs.data_ptr(), wc.data_ptr(), y.data_ptr(), rstd.data_ptr(),
                                     rows, H, float(eps), _is_bf16(xs, "x"), _stream()),
              "b200_rmsnorm_fwd")
        ctx.save_for_backward(xs, wc, rstd)
        ctx.w_dtype = w.dtype
        ctx.weight = w
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, wc, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        rows, H = xs.numel() // xs.shape[-1], xs.shape[-1]
        dx = torch.empty_like(xs)
        need_w = ctx.needs_input_grad[1]
        target = _dw_target(ctx.weight) if need_w else None
        dw = None if (target is not None or not need_w) else torch.empty(H, device=xs.device, dtype=torch.float32)
        ws = torch.empty(int(lib().b200_rmsnorm_bwd_workspace_bytes(rows, H)), device=xs.device, dtype=torch.uint8)
        check(lib().b200_rmsnorm_bwd(dy.data_ptr(), xs.data_ptr(), wc.data_ptr(), rstd.data_ptr(),
                                     dx.data_ptr(), _ptr(dw), rows, H, _is_bf16(xs, "x"),
                                     ws.data_ptr(), ws.numel(), _stream()), "b200_rmsnorm_bwd")
        if target is not None:
            _defer_dw(ws, target, int(lib().b200_rmsnorm_bwd_partial_rows(rows, H)))
        return dx, (dw.to(ctx.w_dtype) if dw is not None else None), None


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return _RMSNormFn.apply(x, w, eps)


class _AddRMSNormFn(torch.autograd.Function):
    """(x, delta) -> (s = x + delta, rmsnorm(s) * w): the residual add of arch/llama.py:316-319 fused with
    the norm that reads its result; backward folds the residual-stream gradient into the norm's dx."""

    @staticmethod
    def forward(ctx, x, delta, w, eps):
        xs, ds = x.contiguous(), delta.contiguous()
        rows, H = xs.numel() // xs.shape[-1], xs.shape[-1]
        s = torch.empty_like(xs)
        y = torch.empty_like(xs)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        wc = w.contiguous().to(xs.dtype)
        check(lib().b200_add_rmsnorm_fwd(xs.data_ptr(), ds.data_ptr(), wc.data_ptr(), s.data_ptr(), y.data_ptr(),
                                         rstd.data_ptr(), rows, H, float(eps), _is_bf16(xs, "x"), _stream()),
              "b200_add_rmsnorm_fwd")
        ctx.save_for_backward(s, wc, rstd)
        ctx.w_dtype = w.dtype
        ctx.weight = w
        ctx.set_materialize_grads(False)
        return s, y

    @staticmethod
    def backward(ctx, ds, dy):
        s, wc, rstd = ctx.saved_tensors
        if dy is None:
            return ds, ds, None, None
        dy = dy.contiguous()
        rows, H = s.numel() // s.shape[-1], s.shape[-1]
        dx = torch.empty_like(s)
        need_w = ctx.needs_input_grad[2]
        target = _dw_target(ctx.weight) if need_w else None
        dw = None if (target is not None or not need_w) else torch.empty(H, device=s.device, dtype=torch.float32)
        ws = torch.empty(int(lib().b200_rmsnorm_bwd_workspace_bytes(rows, H)), device=s.device, dtype=torch.uint8)
        dres = ds.contiguous() if ds is not None else None
        check(lib().b200_add_rmsnorm_bwd(dy.data_ptr(), dres.data_ptr() if dres is not None else None,
                                         s.data_ptr(), wc.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                                         _ptr(dw), rows, H, _is_bf16(s, "x"), ws.data_ptr(), ws.numel(),
                                         _stream()), "b200_add_rmsnorm_bwd")
        if target is not None:
            _defer_dw(ws, target, int(lib().b200_rmsnorm_bwd_partial_rows(rows, H)))
        return dx, dx, (dw.to(ctx.w_dtype) if dw is not None else None), None


def add_rmsnorm(x: torch.Tensor, delta: torch.Tensor, w: torch.Tensor, eps: float = 1e-5):
    """Returns (x + delta, rmsnorm(x + delta) * w)."""
    return _AddRMSNormFn.apply(x, delta, w, eps)


def _accumulate_wgrad(w: torch.Tensor, dy2: torch.Tensor, x2: torch.Tensor):
    """dW = dy^T x.  When the weight's .grad is a view of the flat gradient buffer (flat.ParamStore), the
    GEMM accumulates straight into it (beta = 1 epilogue) and autograd gets None: no temporary, no
    separate AccumulateGrad add kernel per parameter."""
    g = w.grad
    if getattr(w, "_b200_flat_grad", False) and g is not None and g.dtype == dy2.dtype:
        g.addmm_(dy2.t(), x2)
        return None
    return dy2.t() @ x2


class _LinearFn(torch.autograd.Function):
    """y = x W^T (bias-free nn.Linear, arch/llama.py projections) through the library GEMM."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return torch.nn.functional.linear(x, w)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = torch.matmul(dy, w) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            dw = _accumulate_wgrad(w, dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1]))
        return dx, dw


class _MultiLinearFn(torch.autograd.Function):
    """Several bias-free projections of the same input (q/k/v, gate/up).  Backward accumulates
    dx = sum_i dy_i W_i inside the GEMM epilogues instead of autograd's extra add passes."""

    @staticmethod
    def forward(ctx, x, *ws):
        ctx.save_for_backward(x, *ws)
        ctx.set_materialize_grads(False)
        return tuple(torch.nn.functional.linear(x, w) for w in ws)

    @staticmethod
    def backward(ctx, *dys):
        x, *ws = ctx.saved_tensors
        x2 = x.reshape(-1, x.shape[-1])
        dx2 = None
        dws = []
        fused = _fuse_adjacent(dys, ws)   # e.g. (dq, Wq), ([dk | dv], [Wk ; Wv]) when memory layouts allow
        if fused is not None:
            for dy2, w, wg in fused:
                if ctx.needs_input_grad[0]:
                    if dx2 is None:
                        dx2 = dy2 @ w
                    else:
                        dx2.addmm_(dy2, w)
                wg.addmm_(dy2.t(), x2)
            return (dx2.view(x.shape) if dx2 is not None else None, *([None] * len(ws)))
        for i, (dy, w) in enumerate(zip(dys, ws)):
            if dy is None:
                dws.append(None)
                continue
            dy2 = dy.reshape(-1, dy.shape[-1])
            if ctx.needs_input_grad[0]:
                if dx2 is None:
                    dx2 = dy2 @ w
                else:
                    dx2.addmm_(dy2, w)
            dws.append(_accumulate_wgrad(w, dy2, x2) if ctx.needs_input_grad[1 + i] else None)
        return (dx2.view(x.shape) if dx2 is not None else None, *dws)


def _fuse_adjacent(dys, ws):
    """Backward of projections sharing an input: where two consecutive weights are adjacent in the flat
    parameter buffer (k_proj / v_proj of a layer), their flat .grad views are adjacent too, and their output
    gradients are the two column halves of one row-major buffer (attention backward writes dk | dv that way),
    [dy_a | dy_b] x [W_a ; W_b] is ONE dgrad GEMM and one wgrad GEMM instead of two each.  Returns a list of
    (dy2 [T, n], W [n, in], W.grad [n, in]) covering all projections, or None when nothing can be merged or
    some weight's gradient is not a flat-buffer view."""
    if any(dy is None for dy in dys) or not all(getattr(w, "_b200_flat_grad", False) and w.grad is not None
                                                 and w.grad.dtype == dy.dtype for w, dy in zip(ws, dys)):
        return None
    out, i, merged = [], 0, False
    while i < len(ws):
        w, dy = ws[i], dys[i]
        n, k_in = w.shape
        es = w.element_size()
        if i + 1 < len(ws):
            w2, dy2 = ws[i + 1], dys[i + 1]
            T = dy.numel() // n
            rows_ok = all(d.dim() >= 2 and d.shape[-1] == n and d.stride(-1) == 1 and d.stride(-2) == 2 * n and
                          all(d.stride(j) == d.stride(j + 1) * d.shape[j + 1] for j in range(d.dim() - 2))
                          for d in (dy, dy2))
            if (w2.shape == w.shape and w.is_contiguous() and w2.is_contiguous() and rows_ok and
                    w2.data_ptr() == w.data_ptr() + n * k_in * es and
                    w2.grad.data_ptr() == w.grad.data_ptr() + n * k_in * es and
                    dy2.data_ptr() == dy.data_ptr() + n * dy.element_size() and dy2.shape == dy.shape):
                out.append((torch.as_strided(dy, (T, 2 * n), (2 * n, 1)),
                            torch.as_strided(w, (2 * n, k_in), (k_in, 1)),
                            torch.as_strided(w.grad, (2 * n, k_in), (k_in, 1))))
                i += 2
                merged = True
                continue
        out.append((dy.reshape(-1, n), w, w.grad))
        i += 1
    return out if merged else None


def linear(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    return _LinearFn.apply(x, w)


def multi_linear(x: torch.Tensor, ws) -> tuple:
    return _MultiLinearFn.apply(x, *ws)


def rope_tables(seq_len: int, head_dim: int, theta: float, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin [S, D/2] fp32: angle = pos * theta^(-2i/D) (arch/llama_standard.py:74-75,83-84)."""
    freqs = torch.pow(torch.tensor(float(theta), dtype=torch.float32),
                      -torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim)
    ang = torch.outer(torch.arange(seq_len, dtype=torch.float32), freqs)
    return torch.cos(ang).to(device).contiguous(), torch.sin(ang).to(device).contiguous()


class _RopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cos_t, sin_t):
        xs = x.contiguous()
        B, S, NH, D = xs.shape
        y = torch.empty_like(xs)
        check(lib().b200_rope(xs.data_ptr(), y.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), B, S,
                              NH, D, 0, _is_bf16(xs, "x"), _stream()), "b200_rope")
        ctx.save_for_backward(cos_t, sin_t)
        return y

    @staticmethod
    def backward(ctx, dy):
        cos_t, sin_t = ctx.saved_tensors
        dy = dy.contiguous()
        B, S, NH, D = dy.shape
        dx = torch.empty_like(dy)
        check(lib().b200_rope(dy.data_ptr(), dx.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), B, S,
                              NH, D, 1, _is_bf16(dy, "dy"), _stream()), "b200_rope(bwd)")
        return dx, None, None


def rope(x: torch.Tensor, cos_t: torch.Tensor, sin_t: torch.Tensor) -> torch.Tensor:
    """x: [B, S, heads, D]; interleaved-pair rotation."""
    return _RopeFn.apply(x, cos_t, sin_t)


def _pad_head_dim(D: int) -> int:
    if D <= 64:
        return 64
    if D <= 128:
        return 128
    raise ValueError(f"attention: head_dim {D} > 128 is not supported")


def attention_fwd_raw(q, k, v, scale: float, causal: bool):
    """q [B,S,H,D], k/v [B,S,Hk,D] bf16 contiguous, D in {64,128}. Returns (o, lse[B,H,S])."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, torch.bfloat16, nm)
    B, S, H, D = q.shape
    Hk = k.shape[2]
    o = torch.empty_like(q)
    lse = torch.empty((B, H, S), device=q.device, dtype=torch.float32)
    tok = _t0("attn_fwd")
    check(lib().b200_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
                              B, S, H, Hk, D, float(scale), int(causal), _stream()), "b200_attn_fwd")
    _t1(tok)
    return o, lse


def attention_bwd_raw(q, k, v, o, do, lse, scale: float, causal: bool):
    B, S, H, D = q.shape
    Hk = k.shape[2]
    dq = torch.empty_like(q)
    # dk and dv share one [B, S, 2*Hk, D] buffer (dk = first Hk heads of every token row, dv = the rest): the
    # k/v projections' backward can then treat [dk | dv] as ONE matrix against the concatenated weight
    dkv = torch.empty((B, S, 2 * Hk, D), device=k.device, dtype=k.dtype)
    dk, dv = dkv[:, :, :Hk], dkv[:, :, Hk:]
    nbytes = int(lib().b200_attn_bwd_workspace_bytes(B, S, H, Hk, D))
    ws = torch.empty(max(nbytes, 16), device=q.device, dtype=torch.uint8)
    tok = _t0("attn_bwd")
    check(lib().b200_attn_bwd_strided(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(),
                                      lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, S, H,
                                      Hk, D, float(scale), int(causal), 2 * Hk, ws.data_ptr(), ws.numel(),
                                      _stream()), "b200_attn_bwd_strided")
    _t1(tok)
    return dq, dk, dv


class _AttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, scale, causal):
        D = q.shape[-1]
        Dp = _pad_head_dim(D)
        if Dp != D:  # zero-padding the head dim leaves QK^T and the first D columns of PV unchanged
            q, k, v = (torch.nn.functional.pad(t, (0, Dp - D)) for t in (q, k, v))
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        o, lse = attention_fwd_raw(q, k, v, scale, causal)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.scale, ctx.causal, ctx.D = scale, causal, D
        return o[..., :D] if Dp != D else o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        Dp = q.shape[-1]
        if Dp != ctx.D:
            do = torch.nn.functional.pad(do, (0, Dp - ctx.D))
        do = do.contiguous()
        dq, dk, dv = attention_bwd_raw(q, k, v, o, do, lse, ctx.scale, ctx.causal)
        if Dp != ctx.D:
            dq, dk, dv = dq[..., :ctx.D], dk[..., :ctx.D], dv[..., :ctx.D]
        return dq, dk, dv, None, None


def _hi_lo(x: torch.Tensor):
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


class _AttentionX3Fn(torch.autograd.Function):
    """fp32 attention on the bf16 tensor-core kernels for small head dims (3*D <= 128), used when the config asks
    for full precision (`mixed_precision: false`, e.g. BASELINE C1 with D = 16).  Each fp32 operand x = hi + lo
    (two bf16) is laid along the head dimension so that ONE bf16 kernel call evaluates the three-term products:

        Q' = [qh | ql | qh]   K' = [kh | kh | kl]   ->  Q'K'^T = qh kh + ql kh + qh kl  ~=  q k^T   (~16 bits)
        V' = [vh | vh | vl]                         ->  O' = [P vh | P vh | P vl],  o = O'[0:D] + O'[2D:3D]
        dO' = [gh | gl | gh]  ->  dP = gh vh + gl vh + gh vl ;  dq = dQ'[0:D] + dQ'[2D:3D],
                                  dk = dK'[0:D] + dK'[D:2D],    dv = dV'[0:D] + dV'[D:2D]

    The logits (where bf16 operand rounding hurts most: an error of |s| * 2^-9 in the exponent) are thus computed
    to fp32-like accuracy; P and the kernel outputs still pass through one bf16 rounding each."""

    @staticmethod
    def forward(ctx, q, k, v, scale, causal):
        D = q.shape[-1]
        Dp = _pad_head_dim(3 * D)
        (qh, ql), (kh, kl), (vh, vl) = _hi_lo(q), _hi_lo(k), _hi_lo(v)

        def cat3(a, b, c):
            t = torch.cat([a, b, c], dim=-1)
            return (torch.nn.functional.pad(t, (0, Dp - 3 * D)) if Dp != 3 * D else t).contiguous()
        q3, k3, v3 = cat3(qh, ql, qh), cat3(kh, kh, kl), cat3(vh, vh, vl)
        o3, lse = attention_fwd_raw(q3, k3, v3, scale, causal)
        ctx.save_for_backward(q3, k3, v3, o3, lse)
        ctx.scale, ctx.causal, ctx.D, ctx.Dp = scale, causal, D, Dp
        return o3[..., :D].float() + o3[..., 2 * D:3 * D].float()

    @staticmethod
    def backward(ctx, do):
        q3, k3, v3, o3, lse = ctx.saved_tensors
        D, Dp = ctx.D, ctx.Dp
        gh, gl = _hi_lo(do.float())
        g3 = torch.cat([gh, gl, gh], dim=-1)
        if Dp != 3 * D:
            g3 = torch.nn.functional.pad(g3, (0, Dp - 3 * D))
        dq3, dk3, dv3 = attention_bwd_raw(q3, k3, v3, o3, g3.contiguous(), lse, ctx.scale, ctx.causal)
        dq = dq3[..., :D].float() + dq3[..., 2 * D:3 * D].float()
        dk = dk3[..., :D].float() + dk3[..., D:2 * D].float()
        dv = dv3[..., :D].float() + dv3[..., D:2 * D].float()
        return dq, dk, dv, None, None


def attention_fp32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: Optional[float] = None,
                   causal: bool = True) -> torch.Tensor:
    """fp32 in / fp32 out attention with bf16x3 logits (see _AttentionX3Fn); head_dim <= 42."""
    if 3 * q.shape[-1] > 128:
        raise ValueError("attention_fp32: head_dim must be <= 42 (three bf16 copies must fit a 128-wide tile)")
    if scale is None:
        scale = q.shape[-1] ** -0.5
    return _AttentionX3Fn.apply(q.float(), k.float(), v.float(), float(scale), bool(causal))


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: Optional[float] = None,
              causal: bool = True) -> torch.Tensor:
    """softmax(q k^T * scale + causal_mask) v with GQA head sharing; [B,S,H,D] layout, bf16."""
    if scale is None:
        scale = q.shape[-1] ** -0.5
    return _AttentionFn.apply(q, k, v, float(scale), bool(causal))


# ------------------------------------------------------------------------------------------------
# raw strided GEMM + Shampoo elementwise helpers
# ------------------------------------------------------------------------------------------------
def gemm_raw(a_mn: bool, b_mn: bool, M: int, N: int, K: int, batch: int,
             A: torch.Tensor, lda: int, strideA: int, B: torch.Tensor, ldb: int, strideB: int,
             C: Optional[torch.Tensor], ldc: int, strideC: int, D: torch.Tensor, ldd: int, strideD: int,
             alpha: float = 1.0, beta: float = 0.0, alpha_vec: Optional[torch.Tensor] = None,
             beta_vec: Optional[torch.Tensor] = None, force_bn: int = 0) -> None:
    """Strided/batched GEMM on views (sub-blocks of larger matrices): explicit leading dims."""
    if A.dtype != torch.bfloat16 or B.dtype != torch.bfloat16:
        raise ValueError("gemm_raw: A and B must be bf16")
    out_f32 = _is_bf16(D, "D") == 0
    if C is not None and C.dtype != D.dtype:
        raise ValueError("gemm_raw: C and D dtypes differ")
    check(lib().b200_gemm_bf16(int(a_mn), int(b_mn), M, N, K, batch, A.data_ptr(), lda, strideA,
                               B.data_ptr(), ldb, strideB, _ptr(C), ldc, strideC, D.data_ptr(), ldd,
                               strideD, int(out_f32), float(alpha), float(beta), _ptr(alpha_vec),
                               _ptr(beta_vec), int(force_bn), _stream()), "b200_gemm_bf16")


def ema_split(g, m, out32, hi, lo, beta: float, gscale: float, inv_bc: float, out_scale: float = 1.0) -> None:
    check(lib().b200_ema_split(g.data_ptr(), _is_bf16(g, "g"), m.data_ptr(), out32.data_ptr(),
                               hi.data_ptr(), _ptr(lo), g.numel(), float(beta), float(gscale),
                               float(inv_bc), float(out_scale), _stream()), "b200_ema_split")


def graft_update(p32, p16, pre, d, numel: int, batch: int, coef, coef_d, decay: float) -> None:
    check(lib().b200_graft_update(p32.data_ptr(), _ptr(p16), pre.data_ptr(), d.data_ptr(), numel, batch,
                                  coef.data_ptr(), coef_d.data_ptr(), float(decay), _stream()),
          "b200_graft_update")


def sumsq_raw(x: torch.Tensor, out: torch.Tensor, numel: int, batch: int, zero_first: bool = True) -> None:
    ws = reduce_workspace(x.device, batch)
    check(lib().b200_sumsq(x.data_ptr(), _is_bf16(x, "x"), out.data_ptr(), numel, batch, int(zero_first),
                           ws.data_ptr(), ws.numel(), _stream()), "b200_sumsq")


# ------------------------------------------------------------------------------------------------
# Shampoo's Kronecker-factor path through the C ABI (b200_shampoo_*); factor matrices are [batch, kp, kp]
# with kp = round_up(k, 8)
# ------------------------------------------------------------------------------------------------
def rup8(k: int) -> int:
    return (int(k) + 7) // 8 * 8


def shampoo_stats(g_hi: torch.Tensor, g_lo: Optional[torch.Tensor], ldg: int, stride_g: int, L: torch.Tensor,
                  R: torch.Tensor, batch: int, k1: int, k2: int, beta2: float, weight: float) -> None:
    _need(L, torch.float32, "L")
    _need(R, torch.float32, "R")
    check(lib().b200_shampoo_stats(g_hi.data_ptr(), _ptr(g_lo), int(ldg), int(stride_g), L.data_ptr(), R.data_ptr(),
                                   int(batch), int(k1), int(k2), float(beta2), float(weight), _stream()),
          "b200_shampoo_stats")


def shampoo_root(M: torch.Tensor, P: torch.Tensor, P_hi: Optional[torch.Tensor], P_lo: Optional[torch.Tensor],
                 k: int, p: float, eps: float, iters: int = 6, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """matrix_inverse_pth_root (optimizers/shampoo.py:88-126) of M [batch, kp, kp] fp32 -> P (same layout)."""
    _need(M, torch.float32, "M")
    _need(P, torch.float32, "P")
    batch = M.shape[0]
    need = int(lib().b200_shampoo_root_workspace_bytes(batch, int(k)))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, device=M.device, dtype=torch.uint8)
    check(lib().b200_shampoo_root(M.data_ptr(), P.data_ptr(), _ptr(P_hi), _ptr(P_lo), batch, int(k), float(p),
                                  float(eps), int(iters), workspace.data_ptr(), workspace.numel(), _stream()),
          "b200_shampoo_root")
    return P


def shampoo_precond(PLh, PLl, PRh, PRl, m_hi, m_lo, ldm: int, stride_m: int, out: torch.Tensor, ldo: int,
                    stride_o: int, batch: int, k1: int, k2: int, alpha: float,
                    workspace: Optional[torch.Tensor] = None) -> None:
    need = int(lib().b200_shampoo_precond_workspace_bytes(int(batch), int(k1), int(k2)))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, device=out.device, dtype=torch.uint8)
    check(lib().b200_shampoo_precond(PLh.data_ptr(), _ptr(PLl), PRh.data_ptr(), _ptr(PRl), m_hi.data_ptr(),
                                     _ptr(m_lo), int(ldm), int(stride_m), out.data_ptr(), int(ldo), int(stride_o),
                                     int(batch), int(k1), int(k2), float(alpha), workspace.data_ptr(),
                                     workspace.numel(), _stream()), "b200_shampoo_precond")


def shampoo_graft(p32: torch.Tensor, p16: Optional[torch.Tensor], upd: torch.Tensor, graft: torch.Tensor,
                  numel: int, batch: int, decay: float, workspace: Optional[torch.Tensor] = None) -> None:
    need = int(lib().b200_shampoo_graft_workspace_bytes(int(batch)))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, device=p32.device, dtype=torch.uint8)
    check(lib().b200_shampoo_graft(p32.data_ptr(), _ptr(p16), upd.data_ptr(), graft.data_ptr(), int(numel),
                                   int(batch), float(decay), workspace.data_ptr(), workspace.numel(), _stream()),
          "b200_shampoo_graft")


# ------------------------------------------------------------------------------------------------
# block-adjacent fused elementwise steps
# ------------------------------------------------------------------------------------------------
class _GluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, up):
        g, u = gate.contiguous(), up.contiguous()
        y = torch.empty_like(g)
        check(lib().b200_glu_fwd(g.data_ptr(), u.data_ptr(), y.data_ptr(), g.numel(), _stream()), "b200_glu_fwd")
        ctx.save_for_backward(g, u)
        return y

    @staticmethod
    def backward(ctx, dy):
        g, u = ctx.saved_tensors
        dy = dy.contiguous()
        dg, du = torch.empty_like(g), torch.empty_like(u)
        check(lib().b200_glu_bwd(dy.data_ptr(), g.data_ptr(), u.data_ptr(), dg.data_ptr(), du.data_ptr(),
                                 g.numel(), _stream()), "b200_glu_bwd")
        return dg, du


_EMB_WS = {}


class _EmbeddingFn(torch.autograd.Function):
    """h = E[tokens] (arch/llama.py:389).  Backward scatters dh into the weight's flat-buffer gradient with fp32
    accumulation per vocabulary row (b200_embedding_bwd) instead of torch's sort + segmented-sum path
    (3 kernels, 0.44 ms per C2 step); autograd sees None for the weight."""

    @staticmethod
    def forward(ctx, tokens, weight):
        ctx.save_for_backward(tokens)
        ctx.weight = weight
        return torch.nn.functional.embedding(tokens, weight)

    @staticmethod
    def backward(ctx, dh):
        (tokens,) = ctx.saved_tensors
        w = ctx.weight
        V, H = w.shape
        dh2 = dh.reshape(-1, H).contiguous()
        tok = tokens.reshape(-1).contiguous()
        need = int(lib().b200_embedding_bwd_workspace_bytes(V, H))
        key = (str(dh.device), _stream())
        ws = _EMB_WS.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, device=dh.device, dtype=torch.uint8)
            _EMB_WS[key] = ws
        check(lib().b200_embedding_bwd(dh2.data_ptr(), tok.data_ptr(), w.grad.data_ptr(), _is_bf16(w.grad, "grad"),
                                       dh2.shape[0], V, H, ws.data_ptr(), ws.numel(), _stream()), "b200_embedding_bwd")
        return None, None


def embedding(tokens: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Embedding lookup whose backward accumulates straight into a flat-buffer gradient (bf16 activations, int64
    tokens, hidden % 8 == 0); other cases use torch's embedding."""
    g = weight.grad
    if (getattr(weight, "_b200_flat_grad", False) and g is not None and weight.is_cuda and weight.dtype == torch.bfloat16
            and tokens.dtype == torch.int64 and weight.shape[1] % 8 == 0 and g.is_contiguous()):
        return _EmbeddingFn.apply(tokens, weight)
    return torch.nn.functional.embedding(tokens, weight)


def _adjacent_rows(w_a: torch.Tensor, w_b: torch.Tensor) -> bool:
    return (w_a.shape == w_b.shape and w_a.is_contiguous() and w_b.is_contiguous() and
            w_b.data_ptr() == w_a.data_ptr() + w_a.numel() * w_a.element_size())


class _MlpFn(torch.autograd.Function):
    """down( gate(x) * sigmoid(up(x)) * 2 )  (arch/llama.py:142-151) with the activation inside the GEMM epilogues:

      forward   [g | u], y = b200_mlp_gateup_glu_fwd(x, [Wg ; Wu])      one tcgen05 kernel (in-tree engine)
                out        = y Wd^T                                      library GEMM
      backward  dWd       += dout^T y                                    library GEMM into the flat gradient
                [dg | du]  = b200_mlp_down_glu_bwd(dout, Wd, [g | u])    one tcgen05 kernel: dout Wd never leaves the chip
                dx         = [dg | du] [Wg ; Wu]                         ONE library GEMM (K = 2I) instead of two
                d[Wg ; Wu] += [dg | du]^T x                              ONE library GEMM into the flat gradient

    Needs gate_proj.weight and up_proj.weight adjacent in memory (flat.ParamStore) with flat-buffer gradients,
    bf16, intermediate size % 128 == 0; `mlp()` falls back to the unfused path otherwise."""

    @staticmethod
    def forward(ctx, x, wg, wu, wd):
        I, K = wg.shape
        x2 = x.reshape(-1, K).contiguous()
        M = x2.shape[0]
        gu = torch.empty((M, 2 * I), device=x.device, dtype=torch.bfloat16)
        y = torch.empty((M, I), device=x.device, dtype=torch.bfloat16)
        tok = _t0("mlp_glu_fwd")
        check(lib().b200_mlp_gateup_glu_fwd(x2.data_ptr(), wg.data_ptr(), gu.data_ptr(), y.data_ptr(), M, K, I,
                                            _stream()), "b200_mlp_gateup_glu_fwd")
        _t1(tok)
        out = y @ wd.t()
        ctx.save_for_backward(x2, gu, y, wg, wu, wd)
        ctx.x_shape = x.shape
        return out.view(*x.shape[:-1], wd.shape[0])

    @staticmethod
    def backward(ctx, dout):
        x2, gu, y, wg, wu, wd = ctx.saved_tensors
        I, K = wg.shape
        H = wd.shape[0]
        M = x2.shape[0]
        d2 = dout.reshape(-1, H).contiguous()
        wd.grad.addmm_(d2.t(), y)                                   # dWd
        dgu = torch.empty((M, 2 * I), device=d2.device, dtype=torch.bfloat16)
        tok = _t0("mlp_glu_bwd")
        check(lib().b200_mlp_down_glu_bwd(d2.data_ptr(), wd.data_ptr(), gu.data_ptr(), dgu.data_ptr(), M, H, I,
                                          _stream()), "b200_mlp_down_glu_bwd")
        _t1(tok)
        w2 = torch.as_strided(wg, (2 * I, K), (K, 1))               # [Wg ; Wu], adjacent in the flat store
        g2 = torch.as_strided(wg.grad, (2 * I, K), (K, 1))
        dx = dgu @ w2 if ctx.needs_input_grad[0] else None
        g2.addmm_(dgu.t(), x2)
        return (dx.view(ctx.x_shape) if dx is not None else None), None, None, None


def mlp_fusable(x: torch.Tensor, wg: torch.Tensor, wu: torch.Tensor, wd: torch.Tensor) -> bool:
    if not (x.is_cuda and x.dtype == torch.bfloat16 and wg.dtype == torch.bfloat16 and wd.dtype == torch.bfloat16):
        return False
    if wg.shape[0] % 128 != 0 or wg.shape[1] % 8 != 0 or wd.shape != (wg.shape[1], wg.shape[0]) or not wd.is_contiguous():
        return False
    flat = all(getattr(w, "_b200_flat_grad", False) and w.grad is not None and w.grad.dtype == torch.bfloat16
               for w in (wg, wu, wd))
    return flat and _adjacent_rows(wg, wu) and _adjacent_rows(wg.grad, wu.grad)


def mlp(x: torch.Tensor, wg: torch.Tensor, wu: torch.Tensor, wd: torch.Tensor) -> torch.Tensor:
    """The reference MLP (arch/llama.py:149-151) on three bias-free projections."""
    if mlp_fusable(x, wg, wu, wd) and _MLP_FUSED:
        return _MlpFn.apply(x, wg, wu, wd)
    g, u = multi_linear(x, (wg, wu))
    return linear(glu(g, u), wd)


# B200_MLP_FUSED=0 restores the library GEMMs + elementwise GLU kernels (A/B switch for measurements; the first
# version of the fused epilogues, with row-per-lane 16-byte stores, LOST to that pair -- see DESIGN.md sections 4.7 and 9)
_MLP_FUSED = os.environ.get("B200_MLP_FUSED", "1") != "0"


def glu(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """gate * sigmoid(up) * 2 (arch/llama.py:151).  bf16 tensors with numel % 8 == 0 take the fused
    kernel; other dtypes (fp32 configs) use the same formula through torch ops."""
    if gate.dtype == torch.bfloat16 and gate.is_cuda and gate.numel() % 8 == 0:
        return _GluFn.apply(gate, up)
    return gate * torch.sigmoid(up) * 2


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits2d, targets, V, pad_token):
        rows, ld = logits2d.shape
        row_loss = torch.empty(rows, device=logits2d.device, dtype=torch.float32)
        row_lse = torch.empty(rows, device=logits2d.device, dtype=torch.float32)
        check(lib().b200_ce_fwd(logits2d.data_ptr(), ld, targets.data_ptr(), rows, V, int(pad_token),
                                row_loss.data_ptr(), row_lse.data_ptr(), _stream()), "b200_ce_fwd")
        ctx.save_for_backward(logits2d, targets, row_lse)
        ctx.V, ctx.pad = V, int(pad_token)
        return row_loss

    @staticmethod
    def backward(ctx, d_row_loss):
        logits2d, targets, row_lse = ctx.saved_tensors
        rows, ld = logits2d.shape
        scale = d_row_loss.contiguous().float()
        # in place: the logits buffer becomes d(loss)/d(logits); its producer (the logits GEMM) does
        # not need its own output for backward
        check(lib().b200_ce_bwd(logits2d.data_ptr(), ld, targets.data_ptr(), rows, ctx.V, ctx.pad,
                                row_lse.data_ptr(), scale.data_ptr(), _stream()), "b200_ce_bwd")
        return logits2d, None, None, None


def cross_entropy_rows(logits2d: torch.Tensor, targets: torch.Tensor, V: int, pad_token: int) -> torch.Tensor:
    """Per-row masked cross entropy of bf16 logits [rows, ld>=V] (contiguous, ld % 8 == 0), fp32 math;
    rows whose target is pad_token contribute 0 (the pad mask of core/training.py:1230-1232)."""
    _need(logits2d, torch.bfloat16, "logits")
    if targets.dtype != torch.int64 or not targets.is_contiguous():
        raise ValueError("targets must be contiguous int64")
    return _CrossEntropyFn.apply(logits2d, targets, int(V), int(pad_token))
