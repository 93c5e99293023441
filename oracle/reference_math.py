"""CPU restatement of the reference's training-step hot path -- TEST INFRASTRUCTURE ONLY.

This module is the parity oracle: a literal PyTorch-CPU (fp32, optionally fp64) restatement of the
arithmetic in arthurcolle/mlx-cuda-distributed-pretraining for the path named in BASELINE.json.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it; the product package never does (tests/test_host_logic.py::test_product_never_imports_oracle).

Pinning status: the reference holds NO golden vectors or numeric asserts for this path (SURVEY.md
section 4) and its runtime (mlx==0.25.0) is not installable here.  The oracle is pinned instead
against the reference's OWN Python source executed over a NumPy shim of the few mlx.core
primitives it calls (oracle/mlx_numpy_shim, fixtures in tests/golden made by
tests/golden/make_golden.py).  mlx.core's primitive kernels themselves (matmul, softmax) are
third-party and unavailable => "parity unpinned" against a real MLX run; see DESIGN.md.

Every function cites the reference file:line it follows (paths relative to the reference root).
Intended-semantics decisions for the reference's wiring defects are listed in DESIGN.md (D1-D14).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

# --------------------------------------------------------------------------------------------
# attention  (arch/flash_attention.py:97-156)
# --------------------------------------------------------------------------------------------
def causal_mask(seq_len: int, dtype=torch.float32) -> torch.Tensor:
    """arch/llama.py:384-387: full(-inf) -> triu(k=1) -> [1,1,S,S]."""
    m = torch.full((seq_len, seq_len), float("-inf"), dtype=dtype)
    return torch.triu(m, diagonal=1)[None, None]


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float,
              mask: Optional[torch.Tensor]) -> torch.Tensor:
    """q [B,S,H,D], k/v [B,S,Hk,D] -> [B,S,H,D].  flash_attention.py:102-153:
    GQA by repeat on a new axis after the kv-head axis (q head h uses kv head h // (H/Hk)),
    q*scale, QK^T, + mask, softmax(-1), PV.  Un-tiled, like the reference."""
    B, S, H, D = q.shape
    Hk = k.shape[2]
    if H > Hk:
        rep = H // Hk
        k = k.reshape(B, S, Hk, 1, D).repeat_interleave(rep, dim=3).reshape(B, S, H, D)
        v = v.reshape(B, S, Hk, 1, D).repeat_interleave(rep, dim=3).reshape(B, S, H, D)
    q = q * scale
    qh, kh, vh = q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)
    scores = torch.matmul(qh, kh.transpose(-1, -2))
    if mask is not None:
        scores = scores + mask
    w = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(w, vh)
    return ctx.permute(0, 2, 1, 3)


def attention_lse(q, k, v, scale, mask) -> torch.Tensor:
    """log-sum-exp of the scaled masked scores, [B,H,S] (what the fused kernel saves)."""
    B, S, H, D = q.shape
    Hk = k.shape[2]
    if H > Hk:
        k = k.reshape(B, S, Hk, 1, D).repeat_interleave(H // Hk, dim=3).reshape(B, S, H, D)
    scores = torch.matmul((q * scale).permute(0, 2, 1, 3), k.permute(0, 2, 3, 1))
    if mask is not None:
        scores = scores + mask
    return torch.logsumexp(scores, dim=-1)


# --------------------------------------------------------------------------------------------
# RMSNorm / RoPE / MLP  (arch/llama.py:50-56, arch/llama_standard.py:74-75,117-127, llama.py:151)
# --------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    xf = x.float()
    rms = torch.sqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + eps)
    return (xf / rms * w.float()).to(dt)


def rope(x: torch.Tensor, theta: float = 10000.0, positions: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [B,S,NH,D]: pairs (x[2i], x[2i+1]) rotated by pos * theta^(-2i/D), re-interleaved
    (intent of llama_standard.py:117-130; the shipped broadcast is broken)."""
    B, S, NH, D = x.shape
    pos = torch.arange(S, dtype=torch.float32) if positions is None else positions.float()
    freqs = torch.pow(torch.tensor(float(theta)), -torch.arange(0, D, 2, dtype=torch.float32) / D)
    ang = torch.outer(pos, freqs)[None, :, None, :]          # [1,S,1,D/2]
    cos, sin = torch.cos(ang).to(x.dtype), torch.sin(ang).to(x.dtype)
    x0, x1 = x[..., 0::2], x[..., 1::2]
    out = torch.stack([x0 * cos - x1 * sin, x0 * sin + x1 * cos], dim=-1)
    return out.reshape(B, S, NH, D)


def mlp(x, w_gate, w_up, w_down):
    """arch/llama.py:151: down(gate(x) * sigmoid(up(x)) * 2)  (sic: not SiLU-GLU)."""
    lin = torch.nn.functional.linear
    return lin(lin(x, w_gate) * torch.sigmoid(lin(x, w_up)) * 2, w_down)


# --------------------------------------------------------------------------------------------
# Llama model (arch/llama.py:298-319, 366-412) on a flat name->tensor parameter dict
# --------------------------------------------------------------------------------------------
@dataclass
class LlamaDims:
    hidden_size: int
    intermediate_size: int
    num_layers: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    vocab_size: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    use_rope: bool = False          # arch/llama.py never applies RoPE
    tie_word_embeddings: bool = True


def param_shapes(d: LlamaDims) -> Dict[str, Tuple[int, ...]]:
    """Flattened parameter names as mlx tree_flatten(model.parameters()) would produce."""
    sh: Dict[str, Tuple[int, ...]] = {"embed_tokens.weight": (d.vocab_size, d.hidden_size)}
    for i in range(d.num_layers):
        p = f"layers.{i}."
        sh[p + "input_layernorm.weight"] = (d.hidden_size,)
        sh[p + "self_attn.attn.q_proj.weight"] = (d.num_heads * d.head_dim, d.hidden_size)
        sh[p + "self_attn.attn.k_proj.weight"] = (d.num_kv_heads * d.head_dim, d.hidden_size)
        sh[p + "self_attn.attn.v_proj.weight"] = (d.num_kv_heads * d.head_dim, d.hidden_size)
        sh[p + "self_attn.attn.o_proj.weight"] = (d.hidden_size, d.num_heads * d.head_dim)
        sh[p + "post_attention_layernorm.weight"] = (d.hidden_size,)
        sh[p + "mlp.gate_proj.weight"] = (d.intermediate_size, d.hidden_size)
        sh[p + "mlp.up_proj.weight"] = (d.intermediate_size, d.hidden_size)
        sh[p + "mlp.down_proj.weight"] = (d.hidden_size, d.intermediate_size)
    sh["norm.weight"] = (d.hidden_size,)
    if not d.tie_word_embeddings:
        sh["lm_head.weight"] = (d.vocab_size, d.hidden_size)
    return sh


def init_params(d: LlamaDims, seed: int = 42, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Harness-owned initialisation shared by oracle and CUDA paths:
    linears U(-1/sqrt(in), 1/sqrt(in)), embedding N(0, 1/hidden), norm gains 1."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(d).items():
        if len(shape) == 1:
            out[name] = torch.ones(shape, dtype=dtype)
        elif name == "embed_tokens.weight":
            out[name] = (torch.randn(shape, generator=g, dtype=torch.float32) * d.hidden_size ** -0.5).to(dtype)
        else:
            bound = 1.0 / math.sqrt(shape[1])
            out[name] = ((torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound).to(dtype)
    return out


def llama_forward(params: Dict[str, torch.Tensor], tokens: torch.Tensor, d: LlamaDims) -> torch.Tensor:
    """tokens [B,S] int64 -> logits [B,S,V].  arch/llama.py:366-412 with FlashAttention.__call__
    (flash_attention.py:158-194) inlined."""
    lin = torch.nn.functional.linear
    B, S = tokens.shape
    h = params["embed_tokens.weight"][tokens]
    mask = causal_mask(S, dtype=h.dtype)
    scale = d.head_dim ** -0.5
    for i in range(d.num_layers):
        p = f"layers.{i}."
        x = rmsnorm(h, params[p + "input_layernorm.weight"], d.rms_norm_eps)
        q = lin(x, params[p + "self_attn.attn.q_proj.weight"]).reshape(B, S, d.num_heads, d.head_dim)
        k = lin(x, params[p + "self_attn.attn.k_proj.weight"]).reshape(B, S, d.num_kv_heads, d.head_dim)
        v = lin(x, params[p + "self_attn.attn.v_proj.weight"]).reshape(B, S, d.num_kv_heads, d.head_dim)
        if d.use_rope:
            q, k = rope(q, d.rope_theta), rope(k, d.rope_theta)
        ctx = attention(q, k, v, scale, mask).reshape(B, S, d.num_heads * d.head_dim)
        h = h + lin(ctx, params[p + "self_attn.attn.o_proj.weight"])
        x = rmsnorm(h, params[p + "post_attention_layernorm.weight"], d.rms_norm_eps)
        h = h + mlp(x, params[p + "mlp.gate_proj.weight"], params[p + "mlp.up_proj.weight"],
                    params[p + "mlp.down_proj.weight"])
    h = rmsnorm(h, params["norm.weight"], d.rms_norm_eps)
    w_out = params["embed_tokens.weight"] if d.tie_word_embeddings else params["lm_head.weight"]
    return lin(h, w_out)


def compute_loss(logits: torch.Tensor, targets: torch.Tensor, pad_token: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """core/training.py:1226-1234: CE in fp32, pad-masked, sum / ntoks."""
    ce = torch.nn.functional.cross_entropy(logits.float().reshape(-1, logits.shape[-1]),
                                           targets.reshape(-1), reduction="none").reshape(targets.shape)
    pad_mask = targets != pad_token
    ntoks = pad_mask.sum()
    return (ce * pad_mask).sum() / ntoks, ntoks


# --------------------------------------------------------------------------------------------
# schedules (mlx_lm_utils.py:5-56 as composed in core/training.py:770-785)
# --------------------------------------------------------------------------------------------
def linear_schedule(start, end, steps):
    def f(step):
        if step >= steps:
            return end
        return start + (end - start) * (step / steps)
    return f


def cosine_decay(start, steps, end=0.0):
    def f(step):
        if step >= steps:
            return end
        return end + (start - end) * 0.5 * (1 + math.cos(math.pi * (step / steps)))
    return f


def join_schedules(schedules, transitions):
    def f(step):
        for i, t in enumerate(transitions):
            if step < t:
                return schedules[i](step)
        return schedules[-1](step - transitions[-1])
    return f


def make_schedule(cfg: dict, lr: float, total_steps: int) -> Callable[[int], float]:
    t = cfg["type"]
    if t == "cosine_with_warmup":
        return join_schedules([linear_schedule(0, lr, cfg["warmup_steps"]),
                               cosine_decay(lr, total_steps, lr * cfg["min_lr_ratio"])],
                              [cfg["warmup_steps"]])
    if t == "cosine":
        return cosine_decay(lr, total_steps, lr * cfg["min_lr_ratio"])
    if t == "linear":
        return linear_schedule(lr, 0, total_steps)
    raise ValueError(f"Unsupported scheduler type: {t}")


# --------------------------------------------------------------------------------------------
# Muon (optimizers/muon.py)
# --------------------------------------------------------------------------------------------
NS_COEFFS = (3.4445, -4.7750, 2.0315)


def newton_schulz5(G: torch.Tensor, steps: int = 5, eps: float = 1e-7,
                   operand_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """muon.py:54-83.  operand_dtype=torch.bfloat16 emulates the kernel's rounding points
    (bf16 GEMM operands / bf16 storage of A, B, X; fp32 accumulation) for tolerance calibration."""
    a, b, c = NS_COEFFS
    transposed = G.shape[-2] > G.shape[-1]
    if transposed:
        G = G.transpose(-1, -2)
    norm = torch.linalg.matrix_norm(G, keepdim=True)           # Frobenius over the last two dims
    X = G / (norm + eps)
    rnd = (lambda t: t.to(operand_dtype).to(G.dtype)) if operand_dtype is not None else (lambda t: t)
    X = rnd(X)
    for _ in range(steps):
        A = rnd(X @ X.transpose(-1, -2))
        B = rnd(b * A + c * (A @ A))
        X = rnd(a * X + B @ X)
    if transposed:
        X = X.transpose(-1, -2)
    return X


class MuonOracle:
    """Intended semantics of Muon.update (muon.py:85-141) with SURVEY D1/D2 fixed: gradients looked
    up by flat name and the returned updates applied in place."""

    def __init__(self, learning_rate, momentum=0.95, nesterov=True, ns_steps=5, alternate=None):
        self.lr, self.momentum, self.nesterov, self.ns_steps = learning_rate, momentum, nesterov, ns_steps
        self.alternate = alternate
        self.state: Dict[str, Dict[str, torch.Tensor]] = {}
        self.count = 0

    def update(self, params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        lr = self.lr(self.count) if callable(self.lr) else self.lr
        updates, nm_p, nm_g = {}, {}, {}
        for name, p in params.items():
            g = grads.get(name)
            if g is None:
                continue
            if p.dim() == 2:
                st = self.state.setdefault(name, {"momentum_buffer": torch.zeros_like(g)})
                buf = (1 - self.momentum) * g + self.momentum * st["momentum_buffer"]
                st["momentum_buffer"] = buf
                u = g + self.momentum * buf if self.nesterov else buf
                X = newton_schulz5(u, self.ns_steps)
                scaling = max(1, p.shape[0] / p.shape[1]) ** 0.5
                updates[name] = -lr * scaling * X
            else:
                nm_p[name], nm_g[name] = p, g
        if self.alternate is not None and nm_p:
            updates.update(self.alternate.directions(nm_p, nm_g))
        else:
            for name, g in nm_g.items():
                st = self.state.setdefault(name, {"momentum_buffer": torch.zeros_like(g)})
                buf = (1 - self.momentum) * g + self.momentum * st["momentum_buffer"]
                st["momentum_buffer"] = buf
                u = g + self.momentum * buf if self.nesterov else buf
                updates[name] = -lr * u
        self.count += 1
        for name, u in updates.items():
            params[name] = params[name] + u
        return updates


class AdamWOracle:
    """mlx.optimizers.AdamW @0.25.0 as stated in SURVEY 8c (third-party; from its published docs):
    p *= (1 - lr*wd); m,v EMAs; p -= lr * m / (sqrt(v) + eps); bias correction off by default."""

    def __init__(self, learning_rate, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, bias_correction=False):
        self.lr, self.betas, self.eps, self.wd, self.bc = learning_rate, betas, eps, weight_decay, bias_correction
        self.state: Dict[str, Dict[str, torch.Tensor]] = {}
        self.count = 0

    def update(self, params, grads):
        lr = self.lr(self.count) if callable(self.lr) else self.lr
        b1, b2 = self.betas
        t = self.count + 1
        for name, p in params.items():
            g = grads.get(name)
            if g is None:
                continue
            st = self.state.setdefault(name, {"m": torch.zeros_like(g), "v": torch.zeros_like(g)})
            st["m"] = b1 * st["m"] + (1 - b1) * g
            st["v"] = b2 * st["v"] + (1 - b2) * g * g
            p = p * (1 - lr * self.wd)
            if self.bc:
                num = lr / (1 - b1 ** t) * st["m"]
                den = torch.sqrt(st["v"]) / math.sqrt(1 - b2 ** t) + self.eps
            else:
                num, den = lr * st["m"], torch.sqrt(st["v"]) + self.eps
            params[name] = p - num / den
        self.count += 1


# --------------------------------------------------------------------------------------------
# Shampoo (optimizers/shampoo.py)
# --------------------------------------------------------------------------------------------
def matrix_inverse_pth_root(M: torch.Tensor, p: float, epsilon: float = 1e-6, num_iters: int = 6) -> torch.Tensor:
    """shampoo.py:88-126, literally (it is NOT a true inverse root)."""
    n = M.shape[0]
    M = M + torch.eye(n, dtype=M.dtype) * epsilon
    alpha = -1.0 / p
    Z = M / torch.trace(M)
    scaling = torch.trace(M) ** (1.0 / p)
    for _ in range(num_iters):
        Z = Z @ (torch.eye(n, dtype=M.dtype) - Z * alpha)
    return Z * (scaling ** alpha)


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


class ShampooOracle:
    """shampoo.py:314-378 with D1/D2/D9 fixed: grafting step = Adam direction*lr computed without
    applying (mlx Adam: no bias correction), updates applied in place."""

    def __init__(self, learning_rate, params: Optional[ShampooParams] = None):
        self.lr, self.hp = learning_rate, params or ShampooParams()
        self.state: Dict[str, dict] = {}
        self.count = 0

    def _graft_direction(self, st, g, lr):
        hp = self.hp
        if hp.grafting_optimizer == "adam":
            st["graft_m"] = hp.beta1 * st["graft_m"] + (1 - hp.beta1) * g
            st["graft_v"] = hp.beta2 * st["graft_v"] + (1 - hp.beta2) * g * g
            return -lr * st["graft_m"] / (torch.sqrt(st["graft_v"]) + hp.epsilon)
        if hp.grafting_optimizer == "momentum":
            st["graft_m"] = hp.beta1 * st["graft_m"] + g
            return -lr * st["graft_m"]
        return -lr * g

    def update(self, params, grads):
        hp = self.hp
        self.count += 1
        lr = self.lr(self.count) if callable(self.lr) else self.lr
        cap = hp.max_preconditioner_dim
        for name, p in params.items():
            g = grads.get(name)
            if g is None:
                continue
            if name not in self.state:
                st = {"momentum": torch.zeros_like(p), "graft_m": torch.zeros_like(p),
                      "graft_v": torch.zeros_like(p), "statistics": None, "preconditioners": None}
                if p.dim() == 2:
                    d1, d2 = min(p.shape[0], cap), min(p.shape[1], cap)
                    st["statistics"] = [torch.zeros(d1, d1), torch.zeros(d2, d2)]
                    st["preconditioners"] = [None, None]
                self.state[name] = st
            st = self.state[name]
            graft = self._graft_direction(st, g, lr)
            if hp.weight_decay > 0 and not hp.use_decoupled_weight_decay:
                g = g + hp.weight_decay * p
            if st["statistics"] is not None:
                m_, n_ = min(g.shape[0], cap), min(g.shape[1], cap)
                lg = g[:m_, :n_]
                st["statistics"][0] = hp.beta2 * st["statistics"][0] + (1 - hp.beta2) * (lg @ lg.T)
                st["statistics"][1] = hp.beta2 * st["statistics"][1] + (1 - hp.beta2) * (lg.T @ lg)
                if self.count >= hp.start_preconditioning_step and self.count % hp.update_period == 0:
                    st["preconditioners"] = [matrix_inverse_pth_root(s, hp.exponent_override,
                                                                     hp.preconditioner_epsilon)
                                             for s in st["statistics"]]
            st["momentum"] = hp.beta1 * st["momentum"] + (1 - hp.beta1) * g
            mom = st["momentum"]
            if hp.use_bias_correction:
                mom = mom / (1.0 - hp.beta1 ** self.count)
            pre = mom.clone()
            if (st["statistics"] is not None and self.count >= hp.start_preconditioning_step
                    and st["preconditioners"][0] is not None and st["preconditioners"][1] is not None):
                m_, n_ = min(g.shape[0], cap), min(g.shape[1], cap)
                pre[:m_, :n_] = st["preconditioners"][0] @ mom[:m_, :n_] @ st["preconditioners"][1]
            upd = -lr * pre
            # mx.linalg.norm (shampoo.py:300-301) is sqrt(sum(x^2)) in fp32: it overflows to inf when the
            # literal "inverse root" blows the preconditioned step up (D10), which zeroes the grafted step
            sn, gn = torch.sqrt((upd * upd).sum()), torch.sqrt((graft * graft).sum())
            if sn == 0:
                upd = graft
            elif gn != 0:
                upd = upd * (gn / sn)
            if hp.weight_decay > 0 and hp.use_decoupled_weight_decay:
                upd = upd - lr * hp.weight_decay * p
            params[name] = p + upd


# --------------------------------------------------------------------------------------------
# gradient post-processing (core/training.py:1664-1696) and DP contract (hybrid_distributed.py)
# --------------------------------------------------------------------------------------------
def clip_elementwise(grads: Dict[str, torch.Tensor], clip: float) -> Dict[str, torch.Tensor]:
    return {k: torch.clamp(g, -clip, clip) for k, g in grads.items()}


def mean_gradients(per_device: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """hybrid_distributed.py:303-354: unweighted mean over device shards."""
    n = len(per_device)
    return {k: sum(d[k] for d in per_device) / n for k in per_device[0]}


def token_weighted_loss(losses: Sequence[float], ntoks: Sequence[int]) -> float:
    """hybrid_distributed.py:504,519-520."""
    return sum(l * n for l, n in zip(losses, ntoks)) / sum(ntoks)


def synthetic_batch(step: int, rank: int, batch: int, seq: int, vocab: int) -> torch.Tensor:
    """SURVEY 8d synthetic tokens: randint(0, normal_vocab, (B, S+1)), seed 42 + 1000*step + rank."""
    g = torch.Generator().manual_seed(42 + 1000 * step + rank)
    return torch.randint(0, vocab, (batch, seq + 1), generator=g, dtype=torch.int64)


def loss_and_grads(params: Dict[str, torch.Tensor], batch: torch.Tensor, d: LlamaDims, pad_token: int):
    """One forward/backward of the reference step via autograd on the oracle graph."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    logits = llama_forward(leaves, batch[:, :-1], d)
    loss, ntoks = compute_loss(logits, batch[:, 1:], pad_token)
    loss.backward()
    return loss.detach(), int(ntoks), {k: v.grad for k, v in leaves.items()}


def synthetic_batch_markov(step: int, rank: int, batch: int, seq: int, vocab: int) -> torch.Tensor:
    """LEARNABLE synthetic tokens for loss-curve parity runs (uniform noise has the constant optimum
    ln(vocab), so its curve cannot tell two optimizers apart): a fixed first-order Markov chain, every
    token having 4 equally likely successors (table drawn once, seed 4242) with 10 % uniform noise;
    per-step draws use the same seed rule as synthetic_batch.  Entropy rate ~ 0.9*ln4 + 0.1*ln(vocab) + H(0.1)."""
    succ = torch.randint(0, vocab, (vocab, 4), generator=torch.Generator().manual_seed(4242), dtype=torch.int64)
    g = torch.Generator().manual_seed(42 + 1000 * step + rank)
    out = torch.empty((batch, seq + 1), dtype=torch.int64)
    out[:, 0] = torch.randint(0, vocab, (batch,), generator=g, dtype=torch.int64)
    choice = torch.randint(0, 4, (batch, seq), generator=g, dtype=torch.int64)
    noisy = torch.rand((batch, seq), generator=g) < 0.1
    rnd = torch.randint(0, vocab, (batch, seq), generator=g, dtype=torch.int64)
    rows = torch.arange(batch)
    for t in range(seq):
        nxt = succ[out[:, t], choice[:, t]]
        out[:, t + 1] = torch.where(noisy[:, t], rnd[:, t], nxt)
    del rows
    return out
