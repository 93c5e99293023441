from .adamw import AdamW
from .hybrid_optimizer import HybridOptimizer
from .muon import Muon
from .shampoo import Shampoo, ShampooParams

__all__ = ["AdamW", "Muon", "Shampoo", "ShampooParams", "HybridOptimizer"]
