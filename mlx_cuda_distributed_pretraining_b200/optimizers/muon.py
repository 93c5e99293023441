"""Muon with the reference's surface (optimizers/muon.py:7-141), running on the sm_100a kernels.

    opt = Muon(learning_rate=float|callable, momentum=0.95, nesterov=True, ns_steps=5,
               alternate_optimizer=None, betas=None, eps=None, weight_decay=None)
    opt.update(model, gradients)      # the one call the trainer makes (core/training.py:1690,1700)
    opt.state[name]["momentum_buffer"], opt.count, opt.zeropower_via_newtonschulz5(G, steps)

Semantics follow the reference's intended math with its wiring defects fixed:
gradients are matched by flattened name and the update is applied to the model in place.
Per 2-D shape group the step is: fused momentum+Nesterov+sum-of-squares pass -> batched
Newton-Schulz on tcgen05 -> fused `p += -lr*max(1,r/c)^0.5 * X` pass (fp32 master + bf16 shadow).
Non-2-D parameters go to `alternate_optimizer` if given, else to the SGD-momentum fallback
(muon.py:119-138).  `betas/eps/weight_decay` are accepted and ignored, as in the reference (:41-52).

Data parallel (`shard_ns=True`, set by the trainer when world > 1): gradients are already
all-reduced, so every rank holds the same momentum; each rank then orthogonalises only its slice
of every shape group (owner computes) and the results are exchanged with one in-place all-gather
per group (a broadcast for groups with fewer matrices than ranks), issued asynchronously so the
gather of group g overlaps the Newton-Schulz GEMMs of group g+1.  When the ranks can map each
other's memory (torch symmetric memory over NVLink; `B200_NS_P2P=0` disables) the gather is not a
separate collective at all: the gather buffers are peer-mapped and the LAST GEMM of each chain
stores its output tiles into every rank's buffer from its epilogue
(`b200_newton_schulz_allgather`), followed by one cross-rank barrier before the apply kernels.  Communication volume equals
the gradient all-reduce's all-gather half; Newton-Schulz work per rank drops by ~1/world
(SURVEY 8e "fused mode", modal/modal_cuda_utils.py:468-490 for the size-balanced ownership idea).
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, Dict, Optional, Tuple, Union

import torch

from .. import ops
from ..distributed import dp
from ..flat import ParamStore, get_store


class Muon:
    def __init__(self, learning_rate: Union[float, Callable] = 0.02, momentum: float = 0.95,
                 nesterov: bool = True, ns_steps: int = 5, alternate_optimizer=None,
                 betas: Optional[Tuple[float, float]] = None, eps: Optional[float] = None,
                 weight_decay: Optional[float] = None):
        self._learning_rate = learning_rate
        self.momentum = momentum
        self.nesterov = nesterov
        self.ns_steps = ns_steps
        self.alternate_optimizer = alternate_optimizer
        self.state: Dict[str, Dict[str, torch.Tensor]] = {}
        self.count = 0
        self.grad_scale = 1.0        # e.g. 1/world_size for the data-parallel mean
        self.use_accumulated = False  # read gradients from store.acc (fp32) instead of store.grad
        self.shard_ns = False         # owner-computes Newton-Schulz across data-parallel ranks
        self._store: Optional[ParamStore] = None
        self._buf = None
        self._exchange = "auto"       # see set_exchange()
        self._xg = self._xg_handles = None
        self._use_multicast = False

    def set_exchange(self, mode: str) -> str:
        """Select the owner-computes exchange at run time (bench.py's cross-check of the implementations):
        "auto" (peer stores; NVSwitch multicast from 4 ranks up), "unicast", "multicast", or "nccl" (equal-chunk
        ownership + all_gather_into_tensor).  Returns the mode actually in effect."""
        if mode not in ("auto", "unicast", "multicast", "nccl"):
            raise ValueError(f"unknown exchange mode {mode}")
        self._exchange = mode
        self._plan_world = None
        if self._xg_handles is not None and mode != "nccl":
            world = dp.world_size()
            mc_ok = all(int(getattr(h, "multicast_ptr", 0) or 0) != 0 for h in self._xg_handles)
            self._use_multicast = mc_ok and (mode == "multicast" or (mode == "auto" and world >= 4))
        return self.exchange_mode

    # -- reference surface ---------------------------------------------------------------------
    @property
    def learning_rate(self) -> float:
        return self._lr(self.count)

    def _lr(self, count: int) -> float:
        lr = self._learning_rate(count) if callable(self._learning_rate) else self._learning_rate
        return float(lr)

    def zeropower_via_newtonschulz5(self, G: torch.Tensor, steps: int) -> torch.Tensor:
        return ops.zeropower_via_newtonschulz5(G, steps)

    # -- state -----------------------------------------------------------------------------------
    def init(self, model) -> None:
        store = get_store(model)
        if self._store is store:
            return
        self._store = store
        dev = store.device
        self._buf = torch.zeros(store.total, dtype=torch.float32, device=dev)  # momentum buffers
        max_numel = max((g.numel for g in store.mat_groups), default=0)
        max_batch = max((g.batch for g in store.mat_groups), default=1)
        ws_bytes = max((ops.ns_workspace_bytes(g.batch, g.rows, g.cols, self.ns_steps)
                        for g in store.mat_groups), default=0)
        self._u = torch.empty(max(max_numel, 8), dtype=torch.bfloat16, device=dev)
        self._x = torch.empty(max(max_numel, 8), dtype=torch.bfloat16, device=dev)
        self._xg = None  # per-group gather buffers of the sharded mode (allocated on first use)
        self._xg_handles = None  # symmetric-memory handles when the buffers are peer-mapped
        self._use_multicast = False
        self._ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        self._ss = torch.empty(max_batch, dtype=torch.float32, device=dev)
        self._inv = torch.empty(max_batch, dtype=torch.float32, device=dev)
        self._inv2 = torch.empty(max_batch, dtype=torch.float32, device=dev)
        self.state = {}
        for g in store.mat_groups:
            for n in g.names:
                self.state[n] = {"momentum_buffer": store.view(self._buf, n)}
        if self.alternate_optimizer is None:
            for e in store.vec_entries:
                self.state[e.name] = {"momentum_buffer": store.view(self._buf, e.name)}
        else:
            self.alternate_optimizer.init_range(store, store.vec_offset, store.vec_end)

    # -- step --------------------------------------------------------------------------------------
    @torch.no_grad()
    def update(self, model, gradients=None) -> None:
        self.init(model)
        store = self._store
        if gradients is not None:
            store.load_gradients(gradients)
        gsrc = store.acc if self.use_accumulated else store.grad
        lr = self._lr(self.count)
        lib = ops.lib()
        stream = ops._stream()
        a, b, c = ops.NS_COEFFS
        world = dp.world_size() if self.shard_ns else 1
        rank = dp.env_rank_world()[0] if world > 1 else 0
        if world > 1 and self._xg is None:
            self._alloc_gather_buffers(store, world)
        p2p = world > 1 and self._xg_handles is not None and self._exchange != "nccl"
        pending = []
        # experimental (B200_NS_STREAMS=2, single process): run the shape groups' chains alternately on two
        # side streams so that one group's tail wave is filled by the other group's tiles
        n_streams = int(os.environ.get("B200_NS_STREAMS", "1")) if world == 1 else 1
        if n_streams > 1:
            self._update_multistream(store, gsrc, lr, lib, a, b, c, n_streams)
            self._update_vectors(store, gsrc, lr)
            self.count += 1
            return
        for gi, g in enumerate(store.mat_groups):
            n = g.numel
            rc = g.rows * g.cols
            gg = gsrc[g.offset:g.offset + n]
            buf = self._buf[g.offset:g.offset + n]
            u = self._u[:n]
            x = self._x[:n] if world == 1 else self._xg[gi]
            rws = ops.reduce_workspace(store.device, g.batch)
            ops.check(lib.b200_muon_momentum(gg.data_ptr(), ops._is_bf16(gg, "grad"), buf.data_ptr(),
                                             u.data_ptr(), self._ss.data_ptr(), rc, g.batch,
                                             float(self.momentum), int(self.nesterov),
                                             float(self.grad_scale), rws.data_ptr(), rws.numel(), stream),
                      "b200_muon_momentum")
            ops.check(lib.b200_ns_scales(self._ss.data_ptr(), self._inv.data_ptr(), self._inv2.data_ptr(),
                                         g.batch, ops.NS_EPS, stream), "b200_ns_scales")
            mine = self.owned_ranges_of(gi, world, rank)   # matrices this rank orthogonalises
            chunk = dp.chunk_ranges(g.batch, world)[0]
            tok = ops._t0("newton_schulz")
            for lo, hi in mine:
                if hi <= lo:
                    continue
                if p2p:   # GEMM -> all-gather in one kernel: the last GEMM stores into every rank's buffer
                    hdl = self._xg_handles[gi]
                    if self._use_multicast:
                        # one store to the NVSwitch multicast address reaches every rank's replica: per-rank
                        # NVLink egress is its own share, not (world - 1) copies of it
                        ptrs = [int(hdl.multicast_ptr) + 2 * lo * rc]
                    else:
                        ptrs = [int(bp) + 2 * lo * rc for r, bp in enumerate(hdl.buffer_ptrs) if r != rank]
                    peers = (ctypes.c_void_p * len(ptrs))(*ptrs)
                    ops.check(lib.b200_newton_schulz_allgather(
                        u.data_ptr() + 2 * lo * rc, x.data_ptr() + 2 * lo * rc, hi - lo, g.rows, g.cols,
                        self.ns_steps, a, b, c, self._inv.data_ptr() + 4 * lo, self._inv2.data_ptr() + 4 * lo,
                        self._ws.data_ptr(), self._ws.numel(), peers, len(ptrs), stream),
                        "b200_newton_schulz_allgather")
                    continue
                ops.check(lib.b200_newton_schulz(u.data_ptr() + 2 * lo * rc, x.data_ptr() + 2 * lo * rc, hi - lo,
                                                 g.rows, g.cols, self.ns_steps, a, b, c,
                                                 self._inv.data_ptr() + 4 * lo, self._inv2.data_ptr() + 4 * lo,
                                                 self._ws.data_ptr(), self._ws.numel(), stream),
                          "b200_newton_schulz")
            ops._t1(tok)
            works = []
            if world > 1 and not p2p:
                if g.batch >= world:
                    works.append(dp.all_gather_chunks_(x, chunk * rc, async_op=True))
                else:
                    works += [dp.broadcast_async_(x[i * rc:(i + 1) * rc], dp.small_group_owner(i, world))
                              for i in range(g.batch)]
            pending.append((g, x, works))
            if world == 1:
                self._apply(store, g, x, lr, lib, stream)   # the single x buffer is reused by the next group
                pending.pop()
        if p2p:
            # every rank's peer stores are stream-ordered before its barrier arrival
            self._xg_handles[0].barrier()
        for g, x, works in pending:
            for w in works:
                if w is not None:
                    w.wait()   # orders the current stream after the collective
            self._apply(store, g, x, lr, lib, stream)
        self._update_vectors(store, gsrc, lr)
        self.count += 1

    def _update_vectors(self, store, gsrc, lr) -> None:
        if store.vec_end > store.vec_offset:
            lo, hi = store.vec_offset, store.vec_end
            if self.alternate_optimizer is not None:
                self.alternate_optimizer.update_range(store, lo, hi, gsrc, self.grad_scale)
            else:
                p16 = store.shadow[lo:hi] if store.mixed else None
                ops.sgd_momentum(store.master[lo:hi], p16, gsrc[lo:hi], self._buf[lo:hi], self.momentum,
                                 self.nesterov, lr, self.grad_scale)

    def _update_multistream(self, store, gsrc, lr, lib, a, b, c, n_streams: int) -> None:
        main = torch.cuda.current_stream()
        if getattr(self, "_side", None) is None or len(self._side) != n_streams:
            self._side = [torch.cuda.Stream(device=store.device) for _ in range(n_streams)]
            self._side_bufs = [tuple(torch.empty_like(t) for t in (self._u, self._x, self._ws, self._ss, self._inv, self._inv2))
                               for _ in range(n_streams)]
        fork = torch.cuda.Event()
        fork.record(main)
        tok = ops._t0("newton_schulz")   # spans the whole forked region on the main stream
        for st in self._side:
            st.wait_event(fork)
        for gi, g in enumerate(store.mat_groups):
            st = self._side[gi % n_streams]
            u_, x_, ws_, ss_, inv_, inv2_ = self._side_bufs[gi % n_streams]
            n, rc = g.numel, g.rows * g.cols
            gg = gsrc[g.offset:g.offset + n]
            buf = self._buf[g.offset:g.offset + n]
            with torch.cuda.stream(st):
                sp = ops._stream()
                rws = ops.reduce_workspace(store.device, g.batch)   # keyed by stream: one per side stream
                ops.check(lib.b200_muon_momentum(gg.data_ptr(), ops._is_bf16(gg, "grad"), buf.data_ptr(), u_.data_ptr(),
                                                 ss_.data_ptr(), rc, g.batch, float(self.momentum), int(self.nesterov),
                                                 float(self.grad_scale), rws.data_ptr(), rws.numel(), sp),
                          "b200_muon_momentum")
                ops.check(lib.b200_ns_scales(ss_.data_ptr(), inv_.data_ptr(), inv2_.data_ptr(), g.batch, ops.NS_EPS, sp),
                          "b200_ns_scales")
                ops.check(lib.b200_newton_schulz(u_.data_ptr(), x_.data_ptr(), g.batch, g.rows, g.cols, self.ns_steps,
                                                 a, b, c, inv_.data_ptr(), inv2_.data_ptr(), ws_.data_ptr(), ws_.numel(),
                                                 sp), "b200_newton_schulz")
                self._apply(store, g, x_, lr, lib, sp)
        for st in self._side:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        ops._t1(tok)

    def _alloc_gather_buffers(self, store, world: int) -> None:
        """One padded buffer per shape group: world * ceil(batch/world) matrices, so exchanges can stay in
        flight.  Peer-mapped (symmetric memory) when possible, plain device memory + NCCL otherwise."""
        sizes = [max(g.batch, world * dp.chunk_ranges(g.batch, world)[0]) * g.rows * g.cols for g in store.mat_groups]
        if os.environ.get("B200_NS_P2P", "1") != "0" and world <= 8:
            try:
                import torch.distributed as dist
                import torch.distributed._symmetric_memory as symm
                bufs = [symm.empty(n, dtype=torch.bfloat16, device=store.device) for n in sizes]
                self._xg_handles = [symm.rendezvous(t, group=dist.group.WORLD) for t in bufs]
                self._xg = bufs
                # multicast pays off once a unicast exchange would send several copies (measured at 2 ranks:
                # one unicast copy 27.5 ms/step vs multicast 28.0); B200_NS_MULTICAST=1/0 forces it on/off
                mc_env = os.environ.get("B200_NS_MULTICAST", "auto")
                mc_ok = all(int(getattr(h, "multicast_ptr", 0) or 0) != 0 for h in self._xg_handles)
                self._use_multicast = mc_ok and (mc_env == "1" or (mc_env == "auto" and world >= 4))
                return
            except Exception as e:  # no peer mapping on this machine: NCCL exchange instead
                self._xg_handles = None
                if os.environ.get("B200_NS_P2P") == "1":
                    raise
                print(f"[muon] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather")
        self._xg = [torch.empty(n, dtype=torch.bfloat16, device=store.device) for n in sizes]

    @property
    def exchange_mode(self) -> str:
        if not self.shard_ns or self._xg is None:
            return "replicated" if not self.shard_ns else "sharded (not started)"
        if self._xg_handles is None or self._exchange == "nccl":
            return "NCCL all-gather"
        return ("fused GEMM+all-gather, NVSwitch multicast stores" if self._use_multicast
                else "fused GEMM+all-gather, unicast peer stores")

    def owned_ranges_of(self, gi: int, world: int, rank: int):
        """[lo, hi) ranges of shape group `gi` this rank orthogonalises.  With the peer-memory exchange any
        ownership works, so the split is flop-balanced over ALL matrices (dp.balanced_ranges: at 8 ranks a
        rank runs one or two long batched chains instead of 3-matrix slivers of every group); the NCCL
        exchange needs the equal-chunk split of `owned_ranges`."""
        g = self._store.mat_groups[gi]
        if world <= 1 or self._xg_handles is None or self._exchange == "nccl":
            return self.owned_ranges(g.batch, world, rank)
        if getattr(self, "_plan_world", None) != world:
            def cost(r, c):
                m, n = min(r, c), max(r, c)
                return 4.0 * m * m * n + 2.0 * m ** 3
            self._plan = dp.balanced_ranges([(q.batch, cost(q.rows, q.cols)) for q in self._store.mat_groups], world)
            self._plan_world = world
        return [(lo, hi) for (gg, lo, hi) in self._plan[rank] if gg == gi]

    @staticmethod
    def owned_ranges(batch: int, world: int, rank: int):
        """[lo, hi) matrix ranges of a shape group that `rank` orthogonalises in the sharded mode."""
        if world <= 1:
            return [(0, batch)]
        if batch >= world:
            return [dp.chunk_ranges(batch, world)[1][rank]]
        return [(i, i + 1) for i in range(batch) if dp.small_group_owner(i, world) == rank]

    @staticmethod
    def _apply(store, g, x, lr, lib, stream) -> None:
        """p += -lr * max(1, rows/cols)^0.5 * X  (muon.py:111-114), fp32 master + bf16 shadow in one pass."""
        n = g.numel
        scaling = max(1.0, g.rows / g.cols) ** 0.5
        p32 = store.master[g.offset:g.offset + n]
        p16 = store.shadow[g.offset:g.offset + n] if store.mixed else None
        ops.check(lib.b200_axpy_update(p32.data_ptr(), ops._ptr(p16), x.data_ptr(), 1, n,
                                       float(-lr * scaling), stream), "b200_axpy_update")

    # MLX-style aliases kept so callers written against mlx.optimizers.Optimizer keep working
    def apply_gradients(self, gradients, model):
        self.update(model, gradients)
        return model

    def step(self, model) -> None:
        """torch-style: consume the gradients autograd left in the flat store."""
        self.update(model, None)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {f"{n}.momentum_buffer": s["momentum_buffer"] for n, s in self.state.items()}
        if self.alternate_optimizer is not None:
            out.update(self.alternate_optimizer.state_dict())
        return out
