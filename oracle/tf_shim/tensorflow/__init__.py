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


def _pow(x, y):
    x = np.asarray(x)
    y = x.dtype.type(y) if np.isscalar(y) else np.asarray(y, x.dtype)
    return np.power(x, y)


def _minimum(a, b):
    a = np.asarray(a)
    return np.minimum(a, a.dtype.type(b) if np.isscalar(b) else b)


math = types.SimpleNamespace(
    tanh=np.tanh, log=_log, pow=_pow, sqrt=np.sqrt, maximum=_maximum, minimum=_minimum, exp=np.exp,
    abs=np.abs)


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


# ---- symbols used by modules/fdn_reverb.py (executed for tests/golden/fdn_*.npz) ---------------
complex64 = np.complex64


class Tensor(np.ndarray):
    """isinstance(x, tf.Tensor) is False for plain ndarrays, so fdn_reverb.tf_complex64 takes its
    convert_to_tensor branch -- same values."""


def cast(x, dtype=None):
    return np.asarray(x).astype(dtype)


def convert_to_tensor(x, dtype=None):
    return np.asarray(x).astype(dtype) if dtype is not None else np.asarray(x)


def complex(real, imag):  # noqa: A001  (tf.complex)
    real = np.asarray(real, dtype=np.float32)
    return (real + 1j * np.asarray(imag, dtype=np.float32)).astype(np.complex64)


def exp(x):
    return np.exp(x)


def floor(x):
    return np.floor(x)


def pow(x, y):  # noqa: A001
    y = np.asarray(y)
    return np.power(np.asarray(x, dtype=y.dtype) if np.isscalar(x) else x, y)


def stack(values, axis=0):
    return np.stack(values, axis=axis)


def range(n, dtype=np.int32):  # noqa: A001
    return np.arange(n).astype(dtype)


def eye(n, batch_shape=None, dtype=np.float32):
    e = np.eye(n, dtype=dtype)
    if batch_shape:
        e = np.broadcast_to(e, list(batch_shape) + [n, n]).copy()
    return e


def ones(shape, dtype=np.float32):
    return np.ones(shape, dtype=dtype)


def expand_dims(x, axis):
    return np.expand_dims(x, axis)


def squeeze(x, axis=None):
    return np.squeeze(x, axis=axis)


def tile(x, multiples):
    return np.tile(x, multiples)


class _Immutable(np.ndarray):
    """TF tensors are immutable: ``x += y`` rebinds x to a new (broadcast) tensor, it never writes
    in place (surrogate_synth.py:91 relies on it: [B, N, 1] += [B, N, H])."""

    def __iadd__(self, other):
        return np.add(self, other)

    def __imul__(self, other):
        return np.multiply(self, other)


def repeat(x, repeats, axis=None):
    return np.repeat(x, repeats, axis=axis).view(_Immutable)


def transpose(x, perm=None):
    return np.transpose(x, perm)


def pad(x, paddings):
    return np.pad(x, paddings)


def matmul(a, b):
    return np.matmul(a, b)


def reduce_prod(x, axis=None):
    x = np.asarray(x)
    return np.prod(x, axis=axis, dtype=x.dtype)


def _batch_diag(x):
    x = np.asarray(x)
    out = np.zeros(x.shape + (x.shape[-1],), dtype=x.dtype)
    idx = np.arange(x.shape[-1])
    out[..., idx, idx] = x
    return out


linalg = types.SimpleNamespace(diag=_batch_diag, inv=np.linalg.inv)
signal = types.SimpleNamespace(irfft=lambda x: np.fft.irfft(x).astype(np.float32))
keras.activations = types.SimpleNamespace(sigmoid=lambda x: 1.0 / (1.0 + np.exp(-x)))


# ---- symbols used by modules/surrogate_synth.py (executed for tests/golden/surrogate_*.npz) ------
newaxis = None


def where(cond, x, y):
    return np.where(cond, x, y)


def greater_equal(a, b):
    a = np.asarray(a)
    return a >= (a.dtype.type(b) if np.isscalar(b) else b)


def ones_like(x, dtype=None):
    return np.ones_like(x, dtype=dtype)
