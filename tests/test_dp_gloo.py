"""CPU, world_size 2 over gloo: the data-parallel exchange step (flat-buffer all-reduce with the
1/world mean folded into the consumer, token-weighted loss, parameter broadcast) against the
oracle's DP contract (distributed/hybrid_distributed.py:303-354,430-452,504,519-520)."""
import os
import sys
from pathlib import Path

import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mlx_cuda_distributed_pretraining_b200.arch.llama import Model, ModelArgs
    from mlx_cuda_distributed_pretraining_b200.distributed import dp
    from mlx_cuda_distributed_pretraining_b200.flat import ParamStore
    from oracle import reference_math as R

    dp.init_process_group(torch.device("cpu"))
    assert dp.env_rank_world() == (rank, world, rank)
    torch.manual_seed(100 + rank)  # deliberately different initial weights per rank
    args = ModelArgs(model_type="llama", hidden_size=32, num_hidden_layers=1, intermediate_size=64,
                     num_attention_heads=2, head_dim=16, vocab_size=50, num_key_value_heads=1,
                     tie_word_embeddings=True)
    model = Model(args)
    store = ParamStore(model, torch.float32, torch.device("cpu"))
    dp.broadcast_(store.master)  # rank 0's weights everywhere
    ref = store.master.clone()
    # each rank fabricates its own gradient shard, as if from its own batch
    g = torch.Generator().manual_seed(7 + rank)
    store.grad.copy_(torch.randn(store.total, generator=g))
    local = store.grad.clone()
    dp.all_reduce_sum_(store.grad)
    mean = store.grad / world
    loss_tok = torch.tensor([(1.0 + rank) * (10 + 20 * rank), 10.0 + 20 * rank])
    loss = dp.token_weighted_loss(loss_tok)
    # owner-computes exchange (Muon's sharded Newton-Schulz): 5 "matrices" of 6 elements, chunk = 3 per rank;
    # each rank fills only the slice it owns, the in-place gather completes the buffer everywhere
    chunk, ranges = dp.chunk_ranges(5, world)
    xg = torch.full((world * chunk * 6,), -1.0)
    lo, hi = ranges[rank]
    xg[lo * 6:hi * 6] = torch.arange(lo * 6, hi * 6, dtype=torch.float32)
    w = dp.all_gather_chunks_(xg, chunk * 6, async_op=True)
    w.wait()
    small = torch.full((4,), float(rank))
    bw = dp.broadcast_async_(small, dp.small_group_owner(0, world))
    bw.wait()
    torch.save({"master": ref, "local": local, "mean": mean, "loss": loss, "gathered": xg[:30], "small": small},
               f"{out_dir}/r{rank}.pt")
    dp.barrier()
    dp.destroy()


def test_two_rank_gradient_mean_and_loss(tmp_path):
    from oracle import reference_math as R
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["master"], r1["master"])                       # broadcast
    want = R.mean_gradients([{"g": r0["local"]}, {"g": r1["local"]}])["g"]  # unweighted mean over shards
    assert torch.allclose(r0["mean"], want, rtol=1e-6, atol=1e-7)
    assert torch.equal(r0["mean"], r1["mean"])
    assert abs(r0["loss"] - R.token_weighted_loss([1.0, 2.0], [10, 30])) < 1e-6
    assert r0["loss"] == r1["loss"]
    for r in (r0, r1):
        assert torch.equal(r["gathered"], torch.arange(30, dtype=torch.float32))
        assert torch.equal(r["small"], torch.full((4,), 1.0))           # lone matrix lives on the last rank


def test_owner_partition_covers_every_matrix_once():
    sys.path.insert(0, str(ROOT))
    from mlx_cuda_distributed_pretraining_b200.distributed import dp
    for world in (2, 3, 4, 8):
        for batch in (1, 2, 5, 8, 12, 24, 25):
            if batch >= world:
                c, ranges = dp.chunk_ranges(batch, world)
                assert c * world >= batch and len(ranges) == world
                owned = [i for lo, hi in ranges for i in range(lo, hi)]
                assert all(hi - lo <= c for lo, hi in ranges)
            else:
                owned = sorted(i for r in range(world) for i in range(batch) if dp.small_group_owner(i, world) == r)
                assert dp.small_group_owner(0, world) == world - 1
            assert owned == list(range(batch))
    # flop-balanced contiguous plan (peer-memory exchange): C2's five shape groups over 8 ranks
    def cost(r, c):
        m, n = min(r, c), max(r, c)
        return 4.0 * m * m * n + 2.0 * m ** 3
    groups = [(1, cost(32003, 1024)), (24, cost(1024, 1024)), (24, cost(512, 1024)), (24, cost(2816, 1024)),
              (12, cost(1024, 2816))]
    for world in (2, 4, 8):
        plan = dp.balanced_ranges(groups, world)
        seen = {g: [] for g in range(len(groups))}
        loads = []
        for r in range(world):
            loads.append(sum((hi - lo) * groups[g][1] for g, lo, hi in plan[r]))
            for g, lo, hi in plan[r]:
                seen[g] += list(range(lo, hi))
        assert all(sorted(seen[g]) == list(range(groups[g][0])) for g in seen)       # every matrix exactly once
        biggest = max(c for _, c in groups)
        assert max(loads) <= sum(loads) / world + biggest                            # balanced up to one matrix
        assert max(len(p) for p in plan) <= 4                                        # few chains per rank
