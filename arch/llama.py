from mlx_cuda_distributed_pretraining_b200.arch.llama import *  # noqa: F401,F403
from mlx_cuda_distributed_pretraining_b200.arch.llama import Model, ModelArgs  # noqa: F401
