from mlx_cuda_distributed_pretraining_b200.arch.flash_attention import CausalMask, FlashAttention  # noqa: F401
