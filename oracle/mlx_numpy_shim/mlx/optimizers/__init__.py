class Optimizer:
    def __init__(self, schedulers=None):
        self.state = {}


class _Stub(Optimizer):
    def __init__(self, learning_rate=None, **kw):
        super().__init__()
        self.learning_rate = learning_rate
        self.kw = kw

    def update(self, model, grads):
        raise NotImplementedError("third-party mlx optimizer: not part of the shim")


class Adam(_Stub): pass
class AdamW(_Stub): pass
class SGD(_Stub): pass
