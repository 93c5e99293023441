"""NumPy stand-in for the few mlx primitives the reference's hot path calls -- TEST INFRASTRUCTURE.

Purpose: execute the reference's OWN Python (optimizers/muon.py, optimizers/shampoo.py,
arch/flash_attention.py, arch/llama.py, mlx_lm_utils.py) unmodified, in this container where
mlx==0.25.0 cannot be installed, to generate golden vectors that pin oracle/reference_math.py
(tests/golden/make_golden.py).  Only primitive ops (matmul, softmax, elementwise, reductions) are
re-implemented, in float32 NumPy; control flow and formulas are the reference's.
Never imported by the product package.
"""
