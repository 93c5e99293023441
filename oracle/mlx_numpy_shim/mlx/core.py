import numpy as _np

float32 = _np.float32
float16 = _np.float16
bfloat16 = _np.float32   # numpy has no bf16; golden vectors are generated in fp32
int32 = _np.int32
int64 = _np.int64
pi = _np.pi
inf = _np.inf
cpu = "cpu"
gpu = "gpu"


class _At:
    def __init__(self, arr):
        self._a = arr
        self._idx = None

    def __getitem__(self, idx):
        self._idx = idx
        return self

    def set(self, value):
        out = self._a.copy()
        out[self._idx] = value
        return out

    def add(self, value):
        out = self._a.copy()
        out[self._idx] += value
        return out


class array(_np.ndarray):
    def __new__(cls, data, dtype=None):
        a = _np.asarray(data, dtype=dtype)
        if a.dtype == _np.float64:
            a = a.astype(_np.float32)
        return a.view(cls)

    @property
    def at(self):
        return _At(self)

    def tolist(self):
        return _np.asarray(self).tolist()


def _w(x):
    return x.view(array) if isinstance(x, _np.ndarray) and not isinstance(x, array) else x


def ones(shape, dtype=float32): return _w(_np.ones(shape, dtype=dtype))
def zeros(shape, dtype=float32): return _w(_np.zeros(shape, dtype=dtype))
def zeros_like(a): return _w(_np.zeros_like(a))
def ones_like(a): return _w(_np.ones_like(a))
def full(shape, val, dtype=float32): return _w(_np.full(shape, val, dtype=dtype))
def eye(n, dtype=float32): return _w(_np.eye(n, dtype=dtype))
def arange(*a, dtype=None):
    r = _np.arange(*a)
    return _w(r.astype(dtype) if dtype is not None else r)
def triu(a, k=0): return _w(_np.triu(a, k=k))
def matmul(a, b): return _w(_np.matmul(a, b))
def sqrt(a): return _w(_np.sqrt(a))
def square(a): return _w(_np.square(a))
def mean(a, axis=None, keepdims=False): return _w(_np.mean(a, axis=axis, keepdims=keepdims, dtype=_np.float32))
def sum(a, axis=None, keepdims=False): return _w(_np.sum(a, axis=axis, keepdims=keepdims))
def sigmoid(a): return _w((1.0 / (1.0 + _np.exp(-_np.asarray(a, dtype=_np.float32)))).astype(_np.float32))
def exp(a): return _w(_np.exp(a))
def cos(a): return _w(_np.cos(a)) if isinstance(a, _np.ndarray) else float(_np.cos(a))
def sin(a): return _w(_np.sin(a)) if isinstance(a, _np.ndarray) else float(_np.sin(a))
def power(a, b): return _w(_np.power(a, b).astype(_np.float32))
def outer(a, b): return _w(_np.outer(a, b))
def repeat(a, repeats, axis=None): return _w(_np.repeat(a, repeats, axis=axis))
def transpose(a, axes=None): return _w(_np.transpose(a, axes))
def stack(arrs, axis=0): return _w(_np.stack(arrs, axis=axis))
def concatenate(arrs, axis=0): return _w(_np.concatenate(arrs, axis=axis))
def take(a, idx, axis=None): return _w(_np.take(a, idx, axis=axis))
def clip(a, lo, hi): return _w(_np.clip(a, lo, hi))
def maximum(a, b): return _w(_np.maximum(a, b))
def minimum(a, b): return _w(_np.minimum(a, b))
def abs(a): return _w(_np.abs(a))
def trace(a): return _np.float32(_np.trace(a))
def eval(*a, **k): return None
def set_default_device(d): return None


def softmax(a, axis=-1):
    a = _np.asarray(a, dtype=_np.float32)
    m = _np.max(a, axis=axis, keepdims=True)
    e = _np.exp(a - m)
    return _w((e / _np.sum(e, axis=axis, keepdims=True)).astype(_np.float32))


def norm(a, axis=None, keepdims=False):
    """mx.norm as the reference's muon.py:72 intends it: Frobenius norm over `axis`
    (real mlx only has mx.linalg.norm)."""
    return _w(_np.sqrt(_np.sum(_np.square(_np.asarray(a, dtype=_np.float32)), axis=axis, keepdims=keepdims)))


class linalg:
    @staticmethod
    def norm(a, axis=None, keepdims=False):
        return norm(a, axis=axis, keepdims=keepdims)


class random:
    @staticmethod
    def seed(s):
        _np.random.seed(s)

    @staticmethod
    def normal(shape, dtype=float32):
        return _w(_np.random.standard_normal(shape).astype(dtype))

    @staticmethod
    def uniform(low=0.0, high=1.0, shape=(), dtype=float32):
        return _w(_np.random.uniform(low, high, shape).astype(dtype))
