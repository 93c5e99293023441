"""torchrun worker for the 2-GPU data-parallel test: one Muon update, rank r on synthetic batch
(step 0, rank r); every rank saves its fp32 master weights."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from mlx_cuda_distributed_pretraining_b200.core.training import Config, Trainer  # noqa: E402
from mlx_cuda_distributed_pretraining_b200.distributed import dp  # noqa: E402
from tests.smoke_check import tiny_config  # noqa: E402


def main():
    out = Path(sys.argv[1])
    cfg = tiny_config(optimizer="muon")
    cfg["system"]["distributed"] = True
    tr = Trainer(Config.from_dict(cfg), synthetic=True, quiet=True, run_root=str(out / "runs"))
    tr._accum_step, tr._accum_tokens = 0, 0
    assert tr.distributed and tr.world == 2
    tr.train_step(0)
    torch.cuda.synchronize()
    assert getattr(tr.optimizer, "shard_ns", False), "owner-computes Newton-Schulz should be on under DP"
    torch.save({n: t.detach().cpu() for n, t in tr.store.named_master().items()}, out / f"dp_rank{tr.rank}.pt")
    (out / f"exchange_rank{tr.rank}.txt").write_text(tr.optimizer.exchange_mode)
    dp.barrier()
    dp.destroy()


if __name__ == "__main__":
    main()
