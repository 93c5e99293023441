def scatter(*a, **k):
    raise NotImplementedError("mlx_graphs.utils.scatter is only imported by arch/flex_attention.py (out of scope)")
