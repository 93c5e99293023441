"""Data-parallel plumbing: one process per GPU (torchrun), NCCL over NVLink 5 / NVSwitch.

Replaces the reference's thread-queue "devices" and Python gradient averaging
(distributed/utils.py:8-226, distributed/hybrid_distributed.py:303-354,430-452,495-522).  The DP
contract it defined is kept: every rank processes its own `batch_size` sequences, gradients are the
unweighted mean over ranks (sum all-reduce here, 1/world folded into the optimizer's fused
gradient scale), the logged loss is token-weighted.

Because parameters/gradients live in one shape-grouped flat buffer (flat.ParamStore) the exchange
step is a single all-reduce of one contiguous tensor per optimizer update (not per micro-batch).
CPU tests run the same code over gloo.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(device: torch.device) -> None:
    if dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    backend = "nccl" if device.type == "cuda" else "gloo"
    kwargs = {"device_id": device} if device.type == "cuda" else {}
    dist.init_process_group(backend=backend, **kwargs)


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=src)
    return t


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def shard_batch(batch: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous split of a global batch; the last shard takes the remainder
    (distribute_batch, hybrid_distributed.py:440-450)."""
    n = batch.shape[0]
    per = n // world
    lo = rank * per
    hi = n if rank == world - 1 else lo + per
    return batch[lo:hi]


def token_weighted_loss(loss_sum_and_ntok: torch.Tensor) -> float:
    """[sum(loss_i * ntok_i), sum(ntok_i)] -> mean loss after a sum all-reduce."""
    all_reduce_sum_(loss_sum_and_ntok)
    return float(loss_sum_and_ntok[0] / loss_sum_and_ntok[1])


def partition_by_cost(costs: List[float], world: int) -> List[int]:
    """Static greedy bin-packing (largest first): owner rank per item.  Used by the
    owner-computes Newton-Schulz mode, echoing ModalDistributedOptimizer's size-balanced chunks
    (modal/modal_cuda_utils.py:468-490)."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    load = [0.0] * world
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        owner[i] = r
        load[r] += costs[i]
    return owner


def destroy() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()
