"""HybridOptimizer (optimizers/hybrid_optimizer.py:78-114): 2-D parameters -> matrix optimizer
(Muon), everything else -> non-matrix optimizer (AdamW).  The reference's factory recursion for
`optimizer: hybrid` crashes; the routing it intended is exactly Muon's
`alternate_optimizer` hook, so this class is a thin composition."""
from __future__ import annotations

from .adamw import AdamW
from .muon import Muon


class HybridOptimizer:
    def __init__(self, learning_rate=None, matrix_optimizer: Muon = None, non_matrix_optimizer: AdamW = None):
        if matrix_optimizer is None or non_matrix_optimizer is None:
            raise ValueError("HybridOptimizer needs a matrix_optimizer and a non_matrix_optimizer")
        if not isinstance(matrix_optimizer, Muon):
            raise ValueError("matrix_optimizer must be Muon (Newton-Schulz); Shampoo routes all params itself")
        self.matrix_optimizer = matrix_optimizer
        self.non_matrix_optimizer = non_matrix_optimizer
        matrix_optimizer.alternate_optimizer = non_matrix_optimizer

    @property
    def count(self):
        return self.matrix_optimizer.count

    @property
    def state(self):
        s = dict(self.matrix_optimizer.state)
        s.update(self.non_matrix_optimizer.state)
        return s

    def __setattr__(self, k, v):
        if k in ("grad_scale", "use_accumulated") and "matrix_optimizer" in self.__dict__:
            setattr(self.matrix_optimizer, k, v)
        object.__setattr__(self, k, v)

    def update(self, model, gradients=None):
        self.matrix_optimizer.update(model, gradients)

    def step(self, model):
        self.update(model, None)

    def state_dict(self):
        return self.matrix_optimizer.state_dict()
