"""`python -m core.training --config <yaml>` -- the reference's entry point (core/training.py),
re-hosted on the B200 hot path.

Kept from the reference: the YAML schema (dataclasses core/training.py:52-138), `Config.from_yaml`
(:141-167), the byte-level / tokenizers-JSON `TokenizerManager` (:324-440), JSONL chunking
`DataManager` (:442-543), schedules (mlx_lm_utils.py:5-56 as joined in :770-785), the optimizer
factory names (:787-896), `compute_loss` (:1195-1234), the hot loop's order of operations
(:1637-1768: fwd/bwd -> fp32 grads -> elementwise clip -> accumulate -> optimizer.update),
the `Step n: k=v | ...` log line (:257-271) and the three-file checkpoint layout (:1347-1394).

Replaced: mlx arrays/autograd -> PyTorch autograd over the sm_100a kernels; the thread-queue
"devices" (distributed/utils.py) -> one process per GPU with an NCCL all-reduce of the flat
gradient buffer, engaged only when `system.distributed: true` and the job runs under torchrun.

Synthetic data: `data.input_file: synthetic` (or Trainer(..., synthetic=True)) draws
randint(0, normal_vocab_size, (B, S+1)) with seed 42 + 1000*step + rank, which is what
bench.py and the parity tests use; there is no network for real corpora here.
"""
from __future__ import annotations

import argparse
import json
import logging
import math
import os
import random
import sys
import time
from dataclasses import dataclass, field
from datetime import datetime
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import yaml

from .. import ops
from ..arch.llama import Model, ModelArgs
from ..distributed import dp
from ..flat import ParamStore
from ..optimizers import AdamW, HybridOptimizer, Muon, Shampoo, ShampooParams


# ------------------------------------------------------------------------------------------------
# configuration (same sections and keys as the reference)
# ------------------------------------------------------------------------------------------------
@dataclass
class DataConfig:
    input_file: str
    preprocessing: Dict[str, int]
    tokenizer: Dict[str, Any]
    tokenizer_path: Optional[str] = None
    validation_file: Optional[str] = None
    weight_path: Optional[str] = None


@dataclass
class ModelConfig:
    architecture: str
    dimensions: Dict[str, int]
    attention: Dict[str, Any]
    normalization: Dict[str, float]
    rope: Dict[str, Any]
    misc: Dict[str, bool]


@dataclass
class TrainingConfig:
    hyperparameters: Dict[str, Any]
    scheduler: Dict[str, Any]
    optimization: Dict[str, Any]
    epochs: Optional[int] = None
    early_stopping: Dict[str, Any] = field(default_factory=lambda: {
        "enabled": False, "patience": 3, "min_delta": 0.001, "metric": "val_loss", "mode": "min"})
    lr_finder: Dict[str, Any] = field(default_factory=lambda: {
        "enabled": False, "min_lr": 1e-7, "max_lr": 1.0, "num_steps": 100})


@dataclass
class LoggingConfig:
    log_dir: str
    checkpoint_dir: str
    steps: Dict[str, int]
    metrics: Dict[str, bool]
    tensorboard: bool = False
    wandb: bool = False
    wandb_project: Optional[str] = None
    wandb_entity: Optional[str] = None
    log_memory_usage: bool = False
    log_gradient_norm: bool = False
    log_parameter_norm: bool = False
    log_samples: bool = False
    log_samples_count: int = 3


@dataclass
class SystemConfig:
    seed: int
    device: str
    distributed: bool = False
    devices: Optional[List[str]] = None
    cuda_devices: Optional[List[int]] = None
    memory_limit: Optional[int] = None
    mixed_precision: bool = False
    precision: str = "float16"
    gradient_checkpointing: bool = False
    gradient_checkpointing_ratio: float = 0.5
    model_parallel: bool = False
    model_parallel_size: int = 1
    zero_optimization_level: int = 0


@dataclass
class ResumeConfig:
    checkpoint: str
    reset_optimizer: bool = False
    reset_training_state: bool = False


@dataclass
class Config:
    name: str
    data: DataConfig
    model: ModelConfig
    training: TrainingConfig
    logging: LoggingConfig
    system: SystemConfig
    resume: Optional[ResumeConfig] = None
    overwrite: bool = False

    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "Config":
        if "name" not in d:
            raise ValueError("Config must specify a 'name' field at the top level")
        training = dict(d["training"])
        epochs = training.pop("epochs", None)
        resume = ResumeConfig(**d["resume"]) if d.get("resume") else None
        return cls(name=d["name"], overwrite=d.get("overwrite", False), data=DataConfig(**d["data"]),
                   model=ModelConfig(**d["model"]), training=TrainingConfig(**training, epochs=epochs),
                   logging=LoggingConfig(**d["logging"]), system=SystemConfig(**d["system"]), resume=resume)

    @classmethod
    def from_yaml(cls, yaml_path: str) -> "Config":
        with open(yaml_path, "r") as f:
            return cls.from_dict(yaml.safe_load(f))


# ------------------------------------------------------------------------------------------------
# schedules (mlx_lm_utils.py:5-56)
# ------------------------------------------------------------------------------------------------
def linear_schedule(start_value, end_value, steps):
    def schedule(step):
        return end_value if step >= steps else start_value + (end_value - start_value) * (step / steps)
    return schedule


def cosine_decay(start_value, steps, end_value=0.0):
    def schedule(step):
        if step >= steps:
            return end_value
        return end_value + (start_value - end_value) * 0.5 * (1 + math.cos(math.pi * (step / steps)))
    return schedule


def join_schedules(schedules, transition_steps):
    def schedule(step):
        for i, t in enumerate(transition_steps):
            if step < t:
                return schedules[i](step)
        return schedules[-1](step - transition_steps[-1])
    return schedule


# ------------------------------------------------------------------------------------------------
# tokenizer / data
# ------------------------------------------------------------------------------------------------
class TokenizerManager:
    def __init__(self, config: DataConfig, run_dir: Optional[Path] = None, synthetic: bool = False):
        self.config = config
        self.external_tokenizer = None
        if config.tokenizer_path is not None and not synthetic:
            self.use_external_tokenizer(config.tokenizer_path)
        else:
            self.setup_vocabulary()

    def use_external_tokenizer(self, tokenizer_path: str) -> None:
        from tokenizers import Tokenizer
        tokenizer_file = os.path.join(tokenizer_path, "tokenizer.json")
        if not os.path.exists(tokenizer_file):
            raise ValueError(f"Tokenizer file not found at {tokenizer_file}")
        self.external_tokenizer = Tokenizer.from_file(tokenizer_file)
        vocab = self.external_tokenizer.get_vocab()
        sp = self.config.tokenizer["special_tokens"]
        self.PAD_TOKEN, self.BOS_TOKEN, self.EOS_TOKEN = vocab.get(sp["pad"]), vocab.get(sp["bos"]), vocab.get(sp["eos"])
        self.VOCAB_SIZE = len(vocab)
        if None in (self.PAD_TOKEN, self.BOS_TOKEN, self.EOS_TOKEN):
            raise ValueError("One or more special tokens not found in the external tokenizer vocabulary")

    def setup_vocabulary(self) -> None:
        n = self.config.tokenizer["normal_vocab_size"]
        sp = self.config.tokenizer["special_tokens"]
        self.special_token_map = {tok: n + i for i, tok in enumerate(sp.values())}
        self.PAD_TOKEN = self.special_token_map[sp["pad"]]
        self.BOS_TOKEN = self.special_token_map[sp["bos"]]
        self.EOS_TOKEN = self.special_token_map[sp["eos"]]
        self.VOCAB_SIZE = n + len(self.special_token_map)

    def tokenize(self, text: str) -> list:
        if self.external_tokenizer is not None:
            return self.external_tokenizer.encode(text).ids
        return list(text.encode("utf-8"))

    def detokenize(self, tokens) -> str:
        if hasattr(tokens, "tolist"):
            tokens = tokens.tolist()
        if self.external_tokenizer is not None:
            return self.external_tokenizer.decode(tokens)
        return bytes(t for t in tokens if t < 256).decode("utf-8", errors="ignore")

    def tokenize_doc(self, doc: str) -> list:
        max_length = self.config.preprocessing["max_context_size"]
        return [self.BOS_TOKEN] + self.tokenize(doc)[:max_length] + [self.EOS_TOKEN]


class SyntheticData:
    """SURVEY 8d synthetic token stream; same draws for the CUDA path and the CPU oracle.

    `data.input_file: synthetic`         uniform tokens (BASELINE configs), optimum loss = ln(vocab)
    `data.input_file: synthetic:markov`  a fixed first-order Markov chain (4 successors per token, 10 % uniform
                                         noise): LEARNABLE, so a loss curve can tell two optimizers apart; the
                                         oracle's twin is reference_math.synthetic_batch_markov (same draws)."""

    def __init__(self, vocab: int, batch_size: int, seq_len: int, rank: int = 0, num_batches: int = 1 << 30,
                 kind: str = "uniform"):
        self.vocab, self.batch_size, self.seq_len, self.rank = vocab, batch_size, seq_len, rank
        self.train_docs = range(num_batches * batch_size)
        self.has_validation_data = False
        self.num_validation_batches = 0
        self.val_ptr = 0
        self.kind = kind
        if kind not in ("uniform", "markov"):
            raise ValueError(f"unknown synthetic data kind {kind!r} (uniform | markov)")
        self._succ = (torch.randint(0, vocab, (vocab, 4), generator=torch.Generator().manual_seed(4242), dtype=torch.int64)
                      if kind == "markov" else None)

    def generate_batch(self, step: int) -> torch.Tensor:
        g = torch.Generator().manual_seed(42 + 1000 * step + self.rank)
        B, S = self.batch_size, self.seq_len
        if self.kind == "uniform":
            return torch.randint(0, self.vocab, (B, S + 1), generator=g, dtype=torch.int64)
        out = torch.empty((B, S + 1), dtype=torch.int64)
        out[:, 0] = torch.randint(0, self.vocab, (B,), generator=g, dtype=torch.int64)
        choice = torch.randint(0, 4, (B, S), generator=g, dtype=torch.int64)
        noisy = torch.rand((B, S), generator=g) < 0.1
        rnd = torch.randint(0, self.vocab, (B, S), generator=g, dtype=torch.int64)
        for t in range(S):
            out[:, t + 1] = torch.where(noisy[:, t], rnd[:, t], self._succ[out[:, t], choice[:, t]])
        return out


class DataManager:
    def __init__(self, config: DataConfig, tokenizer: TokenizerManager, batch_size: int = 1, rank: int = 0,
                 world: int = 1):
        self.config, self.tokenizer, self.batch_size = config, tokenizer, batch_size
        self.rank, self.world = rank, world
        self.train_docs: List[str] = []
        self.val_docs: List[str] = []
        self.val_ptr = 0
        self._load_file(config.input_file, self.train_docs)
        idx = sorted(range(len(self.train_docs)), key=lambda i: len(self.train_docs[i]))
        random.shuffle(idx)
        self.train_batch_idx = [idx[i:i + batch_size] for i in range(0, len(idx) - batch_size + 1, batch_size)]
        self.train_indices = np.random.permutation(len(self.train_batch_idx))
        self.val_batch_idx: List[List[int]] = []
        if config.validation_file:
            self._load_file(config.validation_file, self.val_docs)
            vidx = sorted(range(len(self.val_docs)), key=lambda i: len(self.val_docs[i]))
            self.val_batch_idx = [vidx[i:min(i + batch_size, len(vidx))] for i in range(0, len(vidx), batch_size)]
            self.val_indices = np.random.permutation(len(self.val_batch_idx))

    def _load_file(self, file_path: str, docs: List[str]) -> None:
        chunk = self.config.preprocessing["max_context_size"]
        stride = chunk - self.config.preprocessing.get("chunk_overlap", 0)
        with open(file_path, "r") as f:
            for line in f:
                text = json.loads(line)["text"]
                for i in range(0, len(text), stride):
                    docs.append(text[i:i + chunk])

    def generate_batch(self, step: int) -> torch.Tensor:
        # data-parallel ranks walk disjoint batches of the shared shuffled order
        j = (step * self.world + self.rank) % len(self.train_indices)
        return self._create_batch([self.train_docs[i] for i in self.train_batch_idx[self.train_indices[j]]])

    def generate_validation_batch(self, batch_idx: int) -> torch.Tensor:
        if not self.val_batch_idx:
            raise ValueError("No validation data available")
        ids = self.val_batch_idx[self.val_indices[self.val_ptr % len(self.val_indices)]]
        self.val_ptr += 1
        return self._create_batch([self.val_docs[i] for i in ids])

    def _create_batch(self, docs: List[str]) -> torch.Tensor:
        batch = [self.tokenizer.tokenize_doc(d) for d in docs]
        max_len = min(max(len(x) for x in batch), self.config.preprocessing.get("max_context_size", 2048))
        rows = [x[:max_len] + [self.tokenizer.PAD_TOKEN] * (max_len - len(x[:max_len])) for x in batch]
        return torch.tensor(rows, dtype=torch.int64)

    @property
    def has_validation_data(self) -> bool:
        return self.config.validation_file is not None and len(self.val_docs) > 0

    @property
    def num_validation_batches(self) -> int:
        return len(self.val_batch_idx)


# ------------------------------------------------------------------------------------------------
# optimizer factory (core/training.py:770-896)
# ------------------------------------------------------------------------------------------------
class OptimizationManager:
    def __init__(self, config: TrainingConfig, num_training_steps: int):
        self.config = config
        self.num_training_steps = num_training_steps

    def create_scheduler(self) -> Callable[[int], float]:
        cfg = self.config.scheduler
        lr = self.config.hyperparameters["learning_rate"]
        if cfg["type"] == "cosine_with_warmup":
            warm = linear_schedule(0, lr, steps=cfg["warmup_steps"])
            cos = cosine_decay(lr, self.num_training_steps, lr * cfg["min_lr_ratio"])
            return join_schedules([warm, cos], [cfg["warmup_steps"]])
        if cfg["type"] == "cosine":
            return cosine_decay(lr, self.num_training_steps, lr * cfg["min_lr_ratio"])
        if cfg["type"] == "linear":
            return linear_schedule(lr, 0, steps=self.num_training_steps)
        raise ValueError(f"Unsupported scheduler type: {cfg['type']}")

    def create_optimizer(self, schedule, cfg: Optional[Dict[str, Any]] = None):
        cfg = self.config.optimization if cfg is None else cfg
        hp = self.config.hyperparameters
        kwargs: Dict[str, Any] = {"learning_rate": schedule}
        if "betas" in cfg:
            kwargs["betas"] = tuple(cfg["betas"])
        if "eps" in cfg:
            kwargs["eps"] = cfg["eps"]
        # the reference only forwards weight_decay when optimization.weight_decay exists;
        # mlx AdamW's default 0.01 otherwise.  Forward the hyperparameter explicitly when present.
        if "weight_decay" in cfg or "weight_decay" in hp:
            kwargs["weight_decay"] = hp.get("weight_decay", cfg.get("weight_decay", 0.01))
        name = cfg["optimizer"]
        if name in ("adamw", "adamw_enhanced"):
            return AdamW(bias_correction=(name == "adamw_enhanced"), **kwargs)
        if name == "adam":
            kwargs.pop("weight_decay", None)
            return AdamW(weight_decay=0.0, **kwargs)
        if name == "muon":
            # `optimizer: muon` binds to the Newton-Schulz Muon of optimizers/muon.py
            return Muon(learning_rate=schedule, momentum=cfg.get("momentum", 0.95),
                        nesterov=cfg.get("nesterov", True), ns_steps=cfg.get("ns_steps", 5),
                        betas=kwargs.get("betas"), eps=kwargs.get("eps"),
                        weight_decay=kwargs.get("weight_decay"))
        if name == "shampoo":
            params = ShampooParams(
                beta1=cfg.get("beta1", 0.9), beta2=cfg.get("beta2", 0.95), epsilon=cfg.get("epsilon", 1e-8),
                weight_decay=kwargs.get("weight_decay", 0.0) if "weight_decay" in cfg else 0.0,
                update_period=cfg.get("update_period", 100),
                start_preconditioning_step=cfg.get("start_preconditioning_step", 1000),
                preconditioner_epsilon=cfg.get("preconditioner_epsilon", 1e-6),
                exponent_override=cfg.get("exponent_override", 0.75), use_bias_correction=True,
                grafting_optimizer=cfg.get("grafting_optimizer", "adam"), use_decoupled_weight_decay=True)
            return Shampoo(learning_rate=schedule, params=params)
        if name == "hybrid":
            m_cfg = {k: v for k, v in cfg.items() if k not in ("optimizer", "non_matrix_optimizer")}
            n_cfg = {k: v for k, v in cfg.items() if k not in ("optimizer", "matrix_optimizer")}
            m_cfg["optimizer"] = cfg.get("matrix_optimizer", "muon")
            n_cfg["optimizer"] = cfg.get("non_matrix_optimizer", "adamw")
            return HybridOptimizer(learning_rate=schedule, matrix_optimizer=self.create_optimizer(schedule, m_cfg),
                                   non_matrix_optimizer=self.create_optimizer(schedule, n_cfg))
        raise ValueError(f"Unsupported optimizer: {name}")


# ------------------------------------------------------------------------------------------------
# trainer
# ------------------------------------------------------------------------------------------------
def _precision_dtype(system: SystemConfig) -> torch.dtype:
    if not system.mixed_precision:
        return torch.float32
    # the tensor-core kernels take bf16 operands; `precision: float16` is served in bf16 (same
    # width, wider exponent) -- documented in DESIGN.md
    return torch.bfloat16


class BatchPrefetcher:
    """Background producer of token batches (the reference builds every batch on the
    training thread, core/training.py:519-543,1651).  A daemon thread runs `make_batch(step)` for the
    coming steps into a rotating pool of PINNED host buffers; `get(step)` hands the training loop the
    buffer for exactly that step, so tokenisation / padding / the pinned staging copy overlap the GPU
    step and the loop's own work is one asynchronous H2D copy.  A buffer is only rewritten after the
    copy that read it has completed (`release(step, event)`)."""

    def __init__(self, make_batch: Callable[[int], torch.Tensor], start: int, stop: int, depth: int = 2,
                 pin: bool = True):
        import queue
        import threading
        self._make, self._stop_step, self._depth = make_batch, stop, max(1, depth)
        self._pin = pin and torch.cuda.is_available()
        self._q: "queue.Queue" = queue.Queue(maxsize=self._depth)
        self._pool: List[Optional[torch.Tensor]] = [None] * (self._depth + 2)
        self._events: List[Any] = [None] * (self._depth + 2)
        self._halt = threading.Event()
        self._next = start
        self._thread = threading.Thread(target=self._run, args=(start,), daemon=True, name="batch-prefetch")
        self._thread.start()

    def _run(self, start: int) -> None:
        try:
            for step in range(start, self._stop_step):
                if self._halt.is_set():
                    return
                batch = self._make(step)
                slot = (step - start) % len(self._pool)
                ev = self._events[slot]
                if ev is not None:
                    ev.synchronize()          # the H2D copy that last read this buffer has finished
                    self._events[slot] = None
                buf = self._pool[slot]
                if buf is None or buf.shape != batch.shape or buf.dtype != batch.dtype:
                    buf = torch.empty(batch.shape, dtype=batch.dtype, pin_memory=self._pin)
                    self._pool[slot] = buf
                buf.copy_(batch)
                while not self._halt.is_set():
                    try:
                        self._q.put((step, slot, buf), timeout=0.1)
                        break
                    except Exception:
                        continue
        except BaseException as e:  # surfaced to the training thread by get()
            self._q.put((None, None, e))

    def get(self, step: int):
        """-> (pinned batch for `step`, slot token to pass to release())."""
        if step != self._next:
            raise RuntimeError(f"prefetcher is sequential: expected step {self._next}, got {step}")
        got_step, slot, buf = self._q.get()
        if got_step is None:
            raise buf
        assert got_step == step
        self._next += 1
        return buf, slot

    def release(self, slot: int, event=None) -> None:
        """Call once the H2D copy of the buffer has been enqueued; `event` (recorded after it) guards reuse."""
        self._events[slot] = event

    def close(self) -> None:
        self._halt.set()
        try:
            while True:
                self._q.get_nowait()
        except Exception:
            pass
        self._thread.join(timeout=5)


class EarlyStoppingMonitor:
    """core/training.py:621-668: stop after `patience` validations without a `min_delta` improvement."""

    def __init__(self, config: Optional[Dict[str, Any]]):
        config = config or {}
        self.enabled = config.get("enabled", False)
        self.patience = config.get("patience", 3)
        self.min_delta = config.get("min_delta", 0.001)
        self.metric = config.get("metric", "val_loss")
        self.mode = config.get("mode", "min")
        self.best_value = float("inf") if self.mode == "min" else float("-inf")
        self.counter = 0

    def update(self, metrics: Dict[str, float]) -> bool:
        if not self.enabled or self.metric not in metrics:
            return False
        cur = metrics[self.metric]
        improved = (self.best_value - cur > self.min_delta) if self.mode == "min" else (cur - self.best_value > self.min_delta)
        if improved:
            self.best_value, self.counter = cur, 0
            return False
        self.counter += 1
        return self.counter >= self.patience


class Trainer:
    def __init__(self, config, for_training: bool = True, synthetic: Optional[bool] = None,
                 run_root: str = "runs", quiet: bool = False, init_params: Optional[Dict[str, torch.Tensor]] = None):
        self.config = config if isinstance(config, Config) else Config.from_yaml(config)
        self.quiet = quiet
        self.synthetic = (self.config.data.input_file.startswith("synthetic")
                          if synthetic is None else synthetic)
        self.rank, self.world, self.local_rank = dp.env_rank_world()
        self.run_dir = Path(run_root) / self.config.name
        self.checkpoint_dir = self.run_dir / "checkpoints"
        self.log_file = self.run_dir / "log.txt"
        if self.rank == 0:
            self.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        self.logger = logging.getLogger(f"trainer.{self.config.name}")
        self.logger.setLevel(logging.WARNING if quiet else logging.INFO)
        if not self.logger.handlers and not quiet:
            self.logger.addHandler(logging.StreamHandler(sys.stdout))
        self.total_tokens = 0
        self.validation_losses: List[Tuple[int, float]] = []
        self.setup_system()
        self.tokenizer = TokenizerManager(self.config.data, self.run_dir, synthetic=self.synthetic)
        self.setup_model(init_params)
        if for_training:
            self.setup_data()
            self.setup_training()

    # -- setup ---------------------------------------------------------------------------------
    def setup_system(self) -> None:
        seed = self.config.system.seed
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        ops.require_device()  # the hot path has no CPU fallback (system.device: cpu is the oracle's job)
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.distributed = bool(self.config.system.distributed) and self.world > 1
        if self.config.system.distributed and self.world == 1:
            self.logger.info("system.distributed is set but WORLD_SIZE=1: running single-GPU "
                             "(launch with torchrun --nproc-per-node N for data parallelism)")
        if self.distributed:
            dp.init_process_group(self.device)

    def setup_model(self, init_params: Optional[Dict[str, torch.Tensor]] = None) -> None:
        m = self.config.model
        if m.architecture not in ("llama", "llama_standard"):
            raise ImportError(f"Model architecture '{m.architecture}' not found")
        args = ModelArgs(
            model_type=m.architecture, hidden_size=m.dimensions["hidden_size"],
            num_hidden_layers=m.dimensions.get("num_layers", 8), intermediate_size=m.dimensions["intermediate_size"],
            num_attention_heads=m.attention["num_heads"], rms_norm_eps=m.normalization["rms_norm_eps"],
            vocab_size=self.tokenizer.VOCAB_SIZE,
            head_dim=m.attention.get("head_dim") or m.dimensions["hidden_size"] // m.attention["num_heads"],
            max_position_embeddings=m.attention["max_position_embeddings"],
            num_key_value_heads=m.attention.get("num_kv_heads") or m.attention["num_heads"],
            attention_bias=m.misc["attention_bias"], mlp_bias=m.misc["mlp_bias"], rope_theta=m.rope["theta"],
            rope_traditional=m.rope["traditional"], rope_scaling=m.rope["scaling"],
            tie_word_embeddings=m.misc["tie_word_embeddings"], logit_scale=m.misc.get("logit_scale"),
            use_flash_attention=m.attention.get("use_flash_attention", True),
            use_flex_attention=m.attention.get("use_flex_attention", False),
            flash_block_size=m.attention.get("flash_block_size", 128),
            apply_rope=bool(m.rope.get("enabled", m.architecture == "llama_standard")))
        self.model_args = args
        model = Model(args)
        init_model_(model, args, self.config.system.seed)
        if init_params is not None:
            model.load_parameters(init_params)
        model.to(self.device)
        self.model = model
        self.store = ParamStore(model, _precision_dtype(self.config.system), self.device)
        if self.distributed:
            dp.broadcast_(self.store.master)
            self.store.refresh_shadow()
        n = sum(p.numel() for p in model.parameters()) / 1e6
        self.logger.info(f"Model has {n:.2f}M parameters (vocab {args.vocab_size})")

    def setup_data(self) -> None:
        hp = self.config.training.hyperparameters
        bs = hp["batch_size"]
        seq = self.config.data.preprocessing["max_context_size"]
        if self.synthetic:
            kind = self.config.data.input_file.split(":", 1)[1] if ":" in self.config.data.input_file else "uniform"
            self.data_manager = SyntheticData(self.config.data.tokenizer["normal_vocab_size"], bs, seq, self.rank,
                                              num_batches=hp.get("iters", 1000), kind=kind)
        else:
            self.data_manager = DataManager(self.config.data, self.tokenizer, bs, self.rank, self.world)

    def setup_training(self) -> None:
        hp = self.config.training.hyperparameters
        steps_per_epoch = max(len(self.data_manager.train_docs) // hp["batch_size"], 1)
        self.steps_per_epoch = steps_per_epoch
        self.total_steps = (steps_per_epoch * self.config.training.epochs if self.config.training.epochs is not None
                            else hp.get("iters", steps_per_epoch))
        om = OptimizationManager(self.config.training, self.total_steps)
        self.lr_schedule = om.create_scheduler()
        self.optimizer = om.create_optimizer(self.lr_schedule)
        self.grad_accum_steps = hp.get("gradient_accumulation_steps", 1)
        self.clip_value = float(hp["gradient_clip"]) if "gradient_clip" in hp else 0.0
        self.use_acc = self.grad_accum_steps > 1 or self.clip_value > 0.0
        if self.use_acc:
            self.store.ensure_acc()
        self.optimizer.use_accumulated = self.use_acc
        self.optimizer.grad_scale = 1.0 / self.world if self.distributed else 1.0
        if hasattr(self.optimizer, "shard_ns"):
            # owner-computes Newton-Schulz across the data-parallel ranks (SURVEY 8e fused mode)
            self.optimizer.shard_ns = bool(self.distributed and self.world > 1 and
                                           os.environ.get("B200_SHARD_NS", "1") != "0")
        self.validation_steps = self.config.logging.steps.get("validation_interval", 0)
        self.early_stopping = EarlyStoppingMonitor(self.config.training.early_stopping)
        self._pin = None

    # -- loss --------------------------------------------------------------------------------------
    def compute_loss(self, model, inputs: torch.Tensor, targets: torch.Tensor):
        max_ctx = self.config.model.attention.get("max_position_embeddings", 2048)
        if inputs.shape[1] > max_ctx:
            inputs, targets = inputs[:, :max_ctx], targets[:, :max_ctx]
        pad_mask = targets != self.tokenizer.PAD_TOKEN
        ntoks = pad_mask.sum()
        logits2d, V = model.padded_logits(inputs) if hasattr(model, "padded_logits") else (None, None)
        if logits2d is not None:
            # fused path: fp32 log-sum-exp over bf16 logits, pad mask applied per row, no fp32 [B,S,V] copy
            rows = ops.cross_entropy_rows(logits2d, targets.reshape(-1).contiguous(), V, self.tokenizer.PAD_TOKEN)
            return rows.sum() / ntoks, ntoks
        logits = model(inputs).float()  # loss always in fp32 (core/training.py:1226)
        ce = torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), targets.reshape(-1),
                                               reduction="none").view(targets.shape)
        return (ce * pad_mask).sum() / ntoks, ntoks

    def _to_device(self, batch: torch.Tensor) -> torch.Tensor:
        """Per-step host->device copy of the token batch from pinned memory.  The CPU can run a whole step
        ahead of the GPU, so a staging buffer is only rewritten after the asynchronous copy that last read it
        has completed (event recorded right after that copy): two rotating buffers, the same guard as
        BatchPrefetcher.release."""
        if self._pin is None:
            self._pin = [[None, None], [None, None]]   # [buffer, event of the last H2D copy that read it]
            self._pin_next = 0
        slot = self._pin[self._pin_next]
        self._pin_next ^= 1
        if slot[1] is not None:
            slot[1].synchronize()
        if slot[0] is None or slot[0].shape != batch.shape or slot[0].dtype != batch.dtype:
            slot[0] = torch.empty(batch.shape, dtype=batch.dtype, pin_memory=True)
        slot[0].copy_(batch)
        dev_batch = slot[0].to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot[1] = ev
        return dev_batch

    # -- one micro-step + (maybe) update; the unit bench.py times ----------------------------------
    def train_step(self, step: int, batch: Optional[torch.Tensor] = None):
        pf = getattr(self, "_prefetcher", None)
        if batch is None and pf is not None:
            pinned, slot = pf.get(step)       # built and staged by the background thread
            dev_batch = pinned.to(self.device, non_blocking=True)
            ev = None
            if self.device.type == "cuda":
                ev = torch.cuda.Event()
                ev.record()
            pf.release(slot, ev)
        else:
            if batch is None:
                batch = self.data_manager.generate_batch(step)
            dev_batch = self._to_device(batch)
        return self.micro_step(step, dev_batch)

    def micro_step(self, step: int, dev_batch: torch.Tensor):
        """Forward + backward of one micro-batch already resident on the device, gradient clamp / 1/k
        accumulation (core/training.py:1664-1696), and -- on the k-th micro-step -- the data-parallel
        exchange and the optimizer update.  Returns (loss, ntoks, did_update)."""
        loss, ntoks = self.compute_loss(self.model, dev_batch[:, :-1], dev_batch[:, 1:])
        loss.backward()
        k = self.grad_accum_steps
        store = self.store
        if self.use_acc:
            first = (self._accum_step == 0)
            ops.clip_accum(store.grad, store.acc, self.clip_value, 1.0 / k, init=first)
            store.zero_grad()
        self._accum_step += 1
        self._accum_tokens = self._accum_tokens + ntoks
        did_update = False
        if self._accum_step == k or step == self.total_steps - 1:
            if self.distributed:
                dp.all_reduce_sum_(store.acc if self.use_acc else store.grad)
            self.optimizer.update(self.model)
            if not self.use_acc:
                store.zero_grad()
            self._accum_step = 0
            did_update = True
        return loss.detach(), ntoks.detach(), did_update

    def validate(self) -> Optional[float]:
        if not self.data_manager.has_validation_data:
            return None
        tot_loss = torch.zeros((), device=self.device)
        tot_tok = torch.zeros((), device=self.device)
        with torch.no_grad():
            for i in range(min(self.data_manager.num_validation_batches, 50)):
                b = self._to_device(self.data_manager.generate_validation_batch(i))
                loss, ntoks = self.compute_loss(self.model, b[:, :-1], b[:, 1:])
                tot_loss += loss * ntoks
                tot_tok += ntoks
        if self.distributed:
            v = torch.stack([tot_loss, tot_tok])
            dp.all_reduce_sum_(v)
            tot_loss, tot_tok = v[0], v[1]
        return float((tot_loss / tot_tok).item())

    # -- main loop (core/training.py:1637-1768) ---------------------------------------------------
    def train(self) -> Dict[str, Any]:
        start_step = 0
        if self.config.resume and self.config.resume.checkpoint:
            start_step = self.load_checkpoint(self.config.resume.checkpoint, self.config.resume.reset_optimizer)
            if self.config.resume.reset_training_state:
                start_step, self.total_tokens, self.validation_losses = 0, 0, []
        self._accum_step, self._accum_tokens = 0, 0
        total_tokens = torch.zeros((), dtype=torch.int64, device=self.device) + int(self.total_tokens)
        log_every = self.config.logging.steps["logging_interval"]
        ckpt_every = self.config.logging.steps.get("checkpoint_interval", 0)
        logf = open(self.log_file, "a" if start_step > 0 else "w") if self.rank == 0 else None
        if logf and start_step == 0:
            logf.write(f"Training started at {datetime.now()}\nTotal steps: {self.total_steps}\n")
            if self.grad_accum_steps > 1:
                logf.write(f"Using gradient accumulation with {self.grad_accum_steps} steps\n")
            logf.write("=" * 50 + "\n\n")
        start_time = time.time()
        val_loss = None
        last: Dict[str, Any] = {}
        if os.environ.get("B200_PREFETCH", "1") != "0":
            self._prefetcher = BatchPrefetcher(self.data_manager.generate_batch, start_step, self.total_steps)
        try:
            last = self._train_loop(start_step, total_tokens, log_every, ckpt_every, logf, start_time, val_loss, last)
        finally:
            if getattr(self, "_prefetcher", None) is not None:
                self._prefetcher.close()
                self._prefetcher = None
        return last

    def _train_loop(self, start_step, total_tokens, log_every, ckpt_every, logf, start_time, val_loss, last):
        pending: Dict[str, Any] = {}   # metrics_to_log of the reference: kept until the next logging step
        for step in range(start_step, self.total_steps):
            loss, ntoks, did_update = self.train_step(step)
            if did_update:
                total_tokens += self._accum_tokens
                self._accum_tokens = 0
            if self.validation_steps > 0 and self.data_manager.has_validation_data and (step + 1) % self.validation_steps == 0:
                val_loss = self.validate()
                self.validation_losses.append((step + 1, val_loss))
                pending.update(val_loss=val_loss, val_ppl=float(np.exp(val_loss)))
                if self.early_stopping.update({"val_loss": val_loss}):   # core/training.py:1726
                    self.logger.info("Early stopping triggered, ending training")
                    if logf:
                        logf.write(f"Step {step}: val_loss={val_loss} | early_stopping=1\n")
                    break
            metrics: Dict[str, Any] = {}
            if step % log_every == 0:
                if self.distributed:  # token-weighted loss across ranks (hybrid_distributed.py:504,519-520)
                    v = torch.stack([loss.float() * ntoks, ntoks.float()])
                    dp.all_reduce_sum_(v)
                    loss_val, ntok_val = float((v[0] / v[1]).item()), int(v[1].item())
                else:
                    loss_val, ntok_val = float(loss.item()), int(ntoks.item())
                tt = int(total_tokens.item()) * (self.world if self.distributed else 1)
                metrics = {**pending, "loss": loss_val, "ppl": float(np.exp(min(loss_val, 80.0))),
                           "lr": float(self.lr_schedule(step)), "tokens": ntok_val, "total_tokens": tt,
                           "tokens_per_sec": float(tt / max(time.time() - start_time, 1e-9))}
                pending = {}
                last = metrics
                if logf:
                    line = f"Step {step}: " + " | ".join(f"{k}={v}" for k, v in metrics.items())
                    logf.write(line + "\n")
                    logf.flush()
                    self.logger.info(line)
            if ckpt_every and (step + 1) % ckpt_every == 0 and self.rank == 0:
                self.total_tokens = int(total_tokens.item())
                self.save_checkpoint(step + 1, val_loss)
        self.total_tokens = int(total_tokens.item())
        if self.rank == 0:
            self.save_checkpoint("final", val_loss)
        if logf:
            logf.write("\n" + "=" * 50 + f"\nTraining completed at {datetime.now()}\n")
            logf.close()
        return last

    # -- checkpoints (core/training.py:1347-1394,1437-1478) ---------------------------------------
    def save_checkpoint(self, step, val_loss: Optional[float] = None) -> None:
        from safetensors.torch import save_file
        torch.cuda.synchronize()
        weights = {n: t.detach().cpu().contiguous() for n, t in self.store.named_master().items()}
        save_file(weights, str(self.checkpoint_dir / f"step_{step}_model.safetensors"))
        opt_state = {n: t.detach().cpu().contiguous().clone() for n, t in self.optimizer.state_dict().items()}
        opt_state["count"] = torch.tensor([self.optimizer.count], dtype=torch.int64)
        save_file(opt_state, str(self.checkpoint_dir / f"step_{step}_optimizer.safetensors"))
        state = {"step": step if isinstance(step, int) else self.total_steps, "val_ptr": self.data_manager.val_ptr,
                 "total_tokens": int(self.total_tokens), "validation_losses": self.validation_losses}
        (self.checkpoint_dir / f"step_{step}_state.json").write_text(json.dumps(state))
        meta_path = self.run_dir / "metadata.json"
        meta = json.loads(meta_path.read_text()) if meta_path.exists() else {"name": self.config.name, "checkpoints": []}
        info = {"step": step, "timestamp": datetime.now().isoformat(),
                "paths": {"model": f"checkpoints/step_{step}_model.safetensors",
                          "optimizer": f"checkpoints/step_{step}_optimizer.safetensors",
                          "state": f"checkpoints/step_{step}_state.json"}}
        if val_loss is not None:
            info["validation_loss"] = val_loss
        meta.setdefault("checkpoints", []).append(info)
        meta_path.write_text(json.dumps(meta, indent=2))

    def load_checkpoint(self, checkpoint_path: str, reset_optimizer: bool = False) -> int:
        from safetensors.torch import load_file
        weights = load_file(f"{checkpoint_path}_model.safetensors")
        with torch.no_grad():
            for n, t in weights.items():
                if n in self.store.index:  # non-strict, like models/llama.py:440-475
                    self.store.view(self.store.master, n).copy_(t.to(self.device))
            self.store.refresh_shadow()
        if not reset_optimizer and os.path.exists(f"{checkpoint_path}_optimizer.safetensors"):
            self.optimizer.init(self.model)   # allocate state so the saved tensors have somewhere to go
            st = load_file(f"{checkpoint_path}_optimizer.safetensors")
            own = self.optimizer.state_dict()
            with torch.no_grad():
                for n, t in st.items():
                    if n == "count":
                        self.optimizer.count = int(t.item())
                    elif n in own and own[n].is_cuda:
                        own[n].copy_(t.to(self.device))
            if hasattr(self.optimizer, "load_extra_state"):   # e.g. the hybrid optimizer's AdamW step counter
                self.optimizer.load_extra_state(st)
            if hasattr(self.optimizer, "after_load"):         # e.g. Shampoo: rebuild bf16 preconditioner operands
                self.optimizer.after_load()
        state = json.loads(Path(f"{checkpoint_path}_state.json").read_text())
        self.total_tokens = state.get("total_tokens", 0)
        self.validation_losses = [tuple(x) for x in state.get("validation_losses", [])]
        if getattr(self, "data_manager", None) is not None:
            self.data_manager.val_ptr = state.get("val_ptr", 0)   # core/training.py:1469
        return int(state["step"])


def init_model_(model: Model, args: ModelArgs, seed: int) -> None:
    """Harness-owned init shared with the oracle (oracle.reference_math.init_params draws the same
    values in the same parameter order): linears U(+-1/sqrt(in)), embedding N(0, 1/hidden), gains 1."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1:
                p.fill_(1.0)
            elif name == "embed_tokens.weight":
                p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float32) * args.hidden_size ** -0.5)
            else:
                bound = 1.0 / math.sqrt(p.shape[1])
                p.copy_((torch.rand(p.shape, generator=g, dtype=torch.float32) * 2 - 1) * bound)


# ------------------------------------------------------------------------------------------------
# CLI (core/training.py:1907-2013) and train() wrapper (:2039-2082)
# ------------------------------------------------------------------------------------------------
def train(config, **trainer_kwargs) -> Dict[str, Any]:
    """Accepts a YAML path or a config dict, like the reference's train(config)."""
    cfg = Config.from_dict(config) if isinstance(config, dict) else Config.from_yaml(config)
    trainer = Trainer(cfg, **trainer_kwargs)
    try:
        return trainer.train()
    finally:
        if trainer.distributed:
            dp.destroy()


def main(argv: Optional[List[str]] = None) -> None:
    ap = argparse.ArgumentParser(description="Train a language model on the B200 hot path")
    ap.add_argument("--config", type=str, required=True, help="Path to YAML config file")
    ap.add_argument("--mixed-precision", action="store_true")
    ap.add_argument("--precision", type=str, default=None, choices=["float16", "bfloat16"])
    ap.add_argument("--gradient-checkpointing", action="store_true")
    ap.add_argument("--find-lr", action="store_true")
    ap.add_argument("--tensorboard", action="store_true")
    ap.add_argument("--wandb", action="store_true")
    ap.add_argument("--log-interval", type=int, default=None)
    ap.add_argument("--run-id", type=str, default=None)
    ap.add_argument("--synthetic", action="store_true", help="use synthetic tokens instead of data.input_file")
    ap.add_argument("--iters", type=int, default=None, help="override training.hyperparameters.iters")
    a = ap.parse_args(argv)
    with open(a.config) as f:
        d = yaml.safe_load(f)
    if a.mixed_precision:
        d["system"]["mixed_precision"] = True
    if a.precision:
        d["system"]["precision"] = a.precision
    if a.log_interval:
        d["logging"]["steps"]["logging_interval"] = a.log_interval
    if a.run_id:
        d["name"] = f"{d['name']}-{a.run_id}"
    if a.iters:
        d["training"]["hyperparameters"]["iters"] = a.iters
    if a.gradient_checkpointing or a.find_lr or a.tensorboard or a.wandb:
        print("note: --gradient-checkpointing/--find-lr/--tensorboard/--wandb are accepted for CLI "
              "compatibility and ignored (outside the hot path)")
    train(d, synthetic=True if a.synthetic else None)


if __name__ == "__main__":
    main()
