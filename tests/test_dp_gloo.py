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
    torch.save({"master": ref, "local": local, "mean": mean, "loss": loss}, f"{out_dir}/r{rank}.pt")
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
