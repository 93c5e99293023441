"""Llama-style decoder with the reference's structure and parameter names (arch/llama.py), hosted
on PyTorch modules whose hot ops are the sm_100a kernels:

  RMSNorm            arch/llama.py:44-56   -> ops.rmsnorm (fused fwd/bwd)
  MLP                arch/llama.py:142-151 -> down(gate(x) * sigmoid(up(x)) * 2)  (sic)
  AttentionModule    arch/llama.py:154-252 -> FlashAttention (fused causal/GQA kernels)
  TransformerBlock   arch/llama.py:255-319 pre-norm residual
  Model              arch/llama.py:322-412 causal mask, tied logits h @ E^T

Flattened parameter names equal mlx's tree_flatten(model.parameters()) names, e.g.
`layers.0.self_attn.attn.q_proj.weight`, so checkpoints and optimizer state keys line up.
RoPE: arch/llama.py constructs a RotaryPositionEncoding and never applies it; the
faithful default is therefore no positional rotation.  `ModelArgs.apply_rope=True` enables the
interleaved-pair rotation of arch/llama_standard.py:213-215 through ops.rope.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
from torch import nn

from .. import ops
from .flash_attention import CausalMask, FlashAttention


@dataclass
class ModelArgs:
    model_type: str
    hidden_size: int
    num_hidden_layers: int
    intermediate_size: int
    num_attention_heads: int
    head_dim: Optional[int] = None
    vocab_size: int = 32000
    num_key_value_heads: Optional[int] = None
    rope_theta: float = 10000.0
    rope_traditional: bool = False
    rope_scaling: Optional[Dict[str, Any]] = None
    rms_norm_eps: float = 1e-5
    max_position_embeddings: int = 4096
    attention_bias: bool = False
    attention_dropout: float = 0.0
    tie_word_embeddings: bool = False
    logit_scale: Optional[float] = None
    mlp_bias: bool = False
    use_flash_attention: bool = True
    use_flex_attention: bool = False
    flash_block_size: int = 128
    num_local_experts: int = 0
    num_experts_per_tok: int = 0
    apply_rope: bool = False  # extension; see module docstring


class RMSNorm(nn.Module):
    def __init__(self, dims: int, eps: float = 1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dims))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.rmsnorm(x, self.weight, self.eps)


class MLP(nn.Module):
    def __init__(self, hidden_size: int, intermediate_size: int, use_bias: bool = False):
        super().__init__()
        self.gate_proj = nn.Linear(hidden_size, intermediate_size, bias=use_bias)
        self.up_proj = nn.Linear(hidden_size, intermediate_size, bias=use_bias)
        self.down_proj = nn.Linear(intermediate_size, hidden_size, bias=use_bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.gate_proj.bias is not None:
            return self.down_proj(ops.glu(self.gate_proj(x), self.up_proj(x)))
        return ops.mlp(x, self.gate_proj.weight, self.up_proj.weight, self.down_proj.weight)


class AttentionModule(nn.Module):
    def __init__(self, hidden_size: int, num_heads: int, num_kv_heads: Optional[int] = None,
                 head_dim: Optional[int] = None, max_positions: int = 4096, rope_theta: float = 10000.0,
                 rope_traditional: bool = False, rope_scaling: Optional[Dict[str, Any]] = None,
                 use_bias: bool = False, use_flash_attention: bool = True, use_flex_attention: bool = False,
                 flash_block_size: int = 128, apply_rope: bool = False):
        super().__init__()
        if use_flex_attention:
            raise NotImplementedError("FlexAttention (arch/flex_attention.py) is out of scope of the hot path")
        self.hidden_size = hidden_size
        self.num_heads = num_heads
        self.num_kv_heads = num_kv_heads or num_heads
        self.head_dim = head_dim or (hidden_size // num_heads)
        # SimpleAttention (use_flash_attention=False) computes the same function; one kernel serves both
        self.attn = FlashAttention(hidden_size, num_heads, self.num_kv_heads, self.head_dim,
                                   use_bias=use_bias, flash_block_size=flash_block_size)
        self.scale = self.head_dim ** -0.5
        self.apply_rope = apply_rope
        self.max_positions = max_positions
        self.rope_theta = rope_theta

    def forward(self, x: torch.Tensor, mask=None, position_ids=None) -> torch.Tensor:
        if self.apply_rope and self.attn.rope_tables is None:
            self.attn.rope_tables = ops.rope_tables(self.max_positions, self.head_dim, self.rope_theta, x.device)
        return self.attn(x, mask=mask)


class TransformerBlock(nn.Module):
    def __init__(self, args: ModelArgs):
        super().__init__()
        inter = args.intermediate_size or 4 * args.hidden_size
        self.input_layernorm = RMSNorm(args.hidden_size, eps=args.rms_norm_eps)
        self.self_attn = AttentionModule(
            hidden_size=args.hidden_size, num_heads=args.num_attention_heads,
            num_kv_heads=args.num_key_value_heads, head_dim=args.head_dim,
            max_positions=args.max_position_embeddings, rope_theta=args.rope_theta,
            rope_traditional=args.rope_traditional, rope_scaling=args.rope_scaling,
            use_bias=args.attention_bias, use_flash_attention=args.use_flash_attention,
            use_flex_attention=args.use_flex_attention, flash_block_size=args.flash_block_size,
            apply_rope=args.apply_rope)
        self.post_attention_layernorm = RMSNorm(args.hidden_size, eps=args.rms_norm_eps)
        self.mlp = MLP(args.hidden_size, inter, use_bias=args.mlp_bias)

    def forward(self, x: torch.Tensor, mask=None, position_ids=None) -> torch.Tensor:
        x = x + self.self_attn(self.input_layernorm(x), mask=mask, position_ids=position_ids)
        return x + self.mlp(self.post_attention_layernorm(x))

    def forward_fused(self, h: torch.Tensor, pending, mask=None, position_ids=None):
        """Same block with every residual add fused into the norm that consumes it.  `pending` is the
        previous block's MLP output not yet added to the stream h (None for the first block); returns
        (h after the attention residual, this block's MLP output still to be added)."""
        n1, n2 = self.input_layernorm, self.post_attention_layernorm
        if pending is None:
            y = n1(h)
        else:
            h, y = ops.add_rmsnorm(h, pending, n1.weight, n1.eps)
        a = self.self_attn(y, mask=mask, position_ids=position_ids)
        h, y = ops.add_rmsnorm(h, a, n2.weight, n2.eps)
        return h, self.mlp(y)


class Model(nn.Module):
    def __init__(self, args: ModelArgs):
        super().__init__()
        self.args = args
        self.vocab_size = args.vocab_size
        self.embed_tokens = nn.Embedding(args.vocab_size, args.hidden_size)
        self.layers = nn.ModuleList([TransformerBlock(args) for _ in range(args.num_hidden_layers)])
        self.norm = RMSNorm(args.hidden_size, eps=args.rms_norm_eps)
        self.logit_scale = args.logit_scale
        self._tied_w = None  # padded alias of the embedding for the logits GEMM (set lazily)
        self.lm_head = None if args.tie_word_embeddings else nn.Linear(args.hidden_size, args.vocab_size, bias=False)

    def hidden_states(self, inputs: torch.Tensor, position_ids=None, attention_mask=None) -> torch.Tensor:
        B, S = inputs.shape
        mask = CausalMask(S) if attention_mask is None else attention_mask
        h = ops.embedding(inputs, self.embed_tokens.weight)
        pending = None
        for layer in self.layers:
            h, pending = layer.forward_fused(h, pending, mask=mask, position_ids=position_ids)
        if pending is None:
            return self.norm(h)
        return ops.add_rmsnorm(h, pending, self.norm.weight, self.norm.eps)[1]

    def _logits_weight(self) -> torch.Tensor:
        """[V or V_pad, hidden] weight of the logits GEMM: lm_head, or the embedding (tied,
        arch/llama.py:401-403) through its zero-row-padded alias when V is not a multiple of 8 (with N = 32003
        cuBLAS falls off its fast path; same values, same gradients, padded classes dropped afterwards)."""
        if self.lm_head is not None:
            return self.lm_head.weight
        store = getattr(self, "_b200_store", None)
        if store is not None and self._tied_w is None:
            w = store.tied_logits_weight()
            self._tied_w = w if w is not None else False
        if self._tied_w is not None and self._tied_w is not False:
            return self._tied_w
        return self.embed_tokens.weight

    def padded_logits(self, inputs: torch.Tensor):
        """(logits [B*S, ld], V) with ld >= V, ld % 8 == 0 when the fused cross-entropy can take them
        (bf16 weights, tied + row-padded embedding alias or an aligned vocabulary); else (None, V).
        Eligibility is decided BEFORE the forward pass, so an ineligible config (fp32, logit_scale,
        unaligned untied vocabulary) never pays for hidden states it then throws away."""
        if self.logit_scale is not None:
            return None, self.vocab_size
        w = self._logits_weight()
        if w.shape[0] % 8 != 0 or w.dtype != torch.bfloat16:
            return None, self.vocab_size
        h = self.hidden_states(inputs)
        logits = torch.nn.functional.linear(h, w)
        return logits.view(-1, logits.shape[-1]), self.vocab_size

    def forward(self, inputs: torch.Tensor, position_ids=None, attention_mask=None) -> torch.Tensor:
        h = self.hidden_states(inputs, position_ids, attention_mask)
        logits = torch.nn.functional.linear(h, self._logits_weight())[..., :self.vocab_size]
        if self.logit_scale is not None:
            logits = logits * self.logit_scale
        return logits

    # mlx nn.Module-style accessors used by the optimizers' update(model, grads) surface
    def parameters_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self.named_parameters())

    @torch.no_grad()
    def load_parameters(self, params: Dict[str, torch.Tensor], strict: bool = True) -> None:
        own = dict(self.named_parameters())
        for name, value in params.items():
            if name not in own:
                if strict:
                    raise KeyError(f"unexpected parameter {name}")
                continue
            own[name].copy_(value.to(own[name].dtype))
        if strict:
            missing = set(own) - set(params)
            if missing:
                raise KeyError(f"missing parameters: {sorted(missing)[:5]}...")
