from mlx_cuda_distributed_pretraining_b200.optimizers.shampoo import Shampoo, ShampooParams  # noqa: F401
