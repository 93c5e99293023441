from mlx_cuda_distributed_pretraining_b200.optimizers.muon import Muon  # noqa: F401
