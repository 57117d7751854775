"""NumPy stand-in for the few TensorFlow symbols the reference's synthesis modules use
(modules/inharm_synth.py, modules/filtered_noise_synth.py:1-42).  Test infrastructure.
Every op computes in the dtype of its input, so float32 inputs follow TF's float32
evaluation (np.cumsum(float32) is a sequential float32 sum like Eigen's CPU scan)."""
import types

import numpy as np

newaxis = None
float32 = np.float32
int32 = np.int32

_noise_queue = []


def push_noise(x):
    """Golden generator hook: next tf.random.uniform() returns this array."""
    _noise_queue.append(np.asarray(x))


def linspace(start, stop, num):
    return np.linspace(start, stop, int(num)).astype(np.float32)


def cumsum(x, axis=0):
    x = np.asarray(x)
    return np.cumsum(x, axis=axis, dtype=x.dtype)


def cos(x):
    return np.cos(x)


def reduce_sum(x, axis=None, keepdims=False):
    x = np.asarray(x)
    return np.sum(x, axis=axis, keepdims=keepdims, dtype=x.dtype)


def greater(a, b):
    a = np.asarray(a)
    return a > a.dtype.type(b)


def zeros_like(x, dtype=None):
    return np.zeros_like(x, dtype=dtype)


def _log(x):
    if isinstance(x, (int, float)):
        return np.log(np.float32(x))
    return np.log(x)


def _maximum(a, b):
    a = np.asarray(a)
    return np.maximum(a, a.dtype.type(b) if np.isscalar(b) else b)


math = types.SimpleNamespace(
    tanh=np.tanh, log=_log, pow=lambda x, y: np.power(x, np.asarray(x).dtype.type(y)),
    sqrt=np.sqrt, maximum=_maximum, exp=np.exp)


def _uniform(shape, minval=0., maxval=1., dtype=np.float32, seed=None):
    if _noise_queue:
        x = _noise_queue.pop(0)
        assert list(x.shape) == [int(s) for s in shape], (x.shape, shape)
        return x
    rng = np.random.default_rng(seed)
    return rng.uniform(minval, maxval, size=[int(s) for s in shape]).astype(dtype)


random = types.SimpleNamespace(uniform=_uniform)


class _Layer:
    def __init__(self, *args, **kwargs):
        self.name = kwargs.get('name')

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


keras = types.SimpleNamespace(layers=types.SimpleNamespace(Layer=_Layer))
