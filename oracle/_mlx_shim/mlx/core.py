"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the `mlx.core` calls the
reference's hot path makes.  See package docstring.

Semantics restated from the published MLX 0.30 API:
  * default float dtype is float32 and default int dtype is int32 (no float64
    arithmetic: every float64/int64 result is narrowed, mirroring MLX's type
    promotion where int32 (op) python-scalar / float32 stay 32-bit);
  * `mx.fast.rms_norm(x, w, eps)`  = x * rsqrt(mean(x^2, -1) + eps) * w, fp32 accumulate;
  * `mx.fast.scaled_dot_product_attention(q,k,v,scale,mask)` = softmax(q k^T * scale + mask) v
    over (B,H,T,D) operands, fp32 softmax;
  * `mx.conv2d(x_NHWC, w_OHWI, stride, padding)` = cross-correlation, NHWC in/out.
bfloat16 does not exist in numpy; it is aliased to float32 here (golden vectors
are generated in float32, the reference's default compute dtype, model.py:456).
"""
from __future__ import annotations

import types

import numpy as np

float32 = np.float32
float16 = np.float16
bfloat16 = np.float32  # see module docstring
float64 = np.float64
int32 = np.int32
int64 = np.int64
uint8 = np.uint8
uint32 = np.uint32
bool_ = np.bool_
Dtype = type


def _narrow(x):
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64:
            return x.astype(np.float32)
        if x.dtype == np.int64:
            return x.astype(np.int32)
    return x


class array(np.ndarray):
    """`mx.array` look-alike: an ndarray that never widens to 64-bit."""

    def __new__(cls, data, dtype=None):
        a = np.array(data, dtype=dtype, copy=True)
        if dtype is None:
            a = _narrow(a)
        return a.view(cls)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = tuple(np.asarray(i) if isinstance(i, array) else i for i in inputs)
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, array) else o for o in out)
        res = getattr(ufunc, method)(*ins, **kwargs)
        if isinstance(res, tuple):
            return tuple(_wrap(r) for r in res)
        return _wrap(res)

    # numpy's ndarray.astype keeps the subclass; nothing else to override.


def _wrap(x):
    if isinstance(x, np.ndarray):
        return _narrow(np.asarray(x)).view(array)
    if isinstance(x, np.generic):
        return _narrow(np.asarray(x)).view(array)
    return x


def _np(x):
    return np.asarray(x)


def eval(*args, **kwargs):  # noqa: A001 - mirrors mx.eval (lazy-graph flush); eager here
    return None


def compile(fn=None, **kwargs):  # noqa: A001 - mx.compile is a pure optimisation
    if fn is None:
        return lambda f: f
    return fn


def zeros(shape, dtype=float32):
    return _wrap(np.zeros(shape, dtype=dtype))


def ones(shape, dtype=float32):
    return _wrap(np.ones(shape, dtype=dtype))


def zeros_like(a):
    return _wrap(np.zeros_like(_np(a)))


def ones_like(a):
    return _wrap(np.ones_like(_np(a)))


def arange(*args, dtype=None):
    a = np.arange(*args)
    if dtype is not None:
        a = a.astype(dtype)
    return _wrap(a)


def linspace(start, stop, num=50, dtype=float32):
    return _wrap(np.linspace(start, stop, num).astype(dtype))


def concatenate(arrays, axis=0):
    return _wrap(np.concatenate([_np(a) for a in arrays], axis=axis))


def stack(arrays, axis=0):
    return _wrap(np.stack([_np(a) for a in arrays], axis=axis))


def repeat(a, repeats, axis=None):
    return _wrap(np.repeat(_np(a), repeats, axis=axis))


def tile(a, reps):
    return _wrap(np.tile(_np(a), reps))


def broadcast_to(a, shape):
    return _wrap(np.broadcast_to(_np(a), shape))


def contiguous(a):
    return _wrap(np.ascontiguousarray(_np(a)))


def expand_dims(a, axis):
    return _wrap(np.expand_dims(_np(a), axis))


def transpose(a, axes=None):
    return _wrap(np.transpose(_np(a), axes))


def meshgrid(*xs, indexing="xy"):
    return [_wrap(g) for g in np.meshgrid(*[_np(x) for x in xs], indexing=indexing)]


def where(c, a, b):
    return _wrap(np.where(_np(c), a, b))


def clip(a, lo, hi):
    return _wrap(np.clip(_np(a), lo, hi))


def maximum(a, b):
    return _wrap(np.maximum(a, b))


def minimum(a, b):
    return _wrap(np.minimum(a, b))


def pad(a, pad_width, constant_values=0):
    return _wrap(np.pad(_np(a), pad_width, constant_values=constant_values))


def all(a, axis=None, keepdims=False):  # noqa: A001
    return _wrap(np.all(_np(a), axis=axis, keepdims=keepdims))


def mean(a, axis=None, keepdims=False):
    return _wrap(np.mean(_np(a), axis=axis, keepdims=keepdims, dtype=np.float32))


def sum(a, axis=None, keepdims=False):  # noqa: A001
    return _wrap(np.sum(_np(a), axis=axis, keepdims=keepdims))


def sin(a):
    return _wrap(np.sin(_np(a)))


def cos(a):
    return _wrap(np.cos(_np(a)))


def exp(a):
    return _wrap(np.exp(_np(a)))


def log(a):
    return _wrap(np.log(_np(a)))


def tanh(a):
    return _wrap(np.tanh(_np(a)))


def sqrt(a):
    return _wrap(np.sqrt(_np(a)))


def rsqrt(a):
    return _wrap(1.0 / np.sqrt(_np(a)))


def power(a, b):
    return _wrap(np.power(a, b))


def sigmoid(a):
    a = _np(a)
    return _wrap(1.0 / (1.0 + np.exp(-a)))


def abs(a):  # noqa: A001
    return _wrap(np.abs(_np(a)))


def matmul(a, b):
    return _wrap(np.matmul(_np(a), _np(b)))


def softmax(a, axis=-1):
    a = _np(a).astype(np.float32)
    a = a - a.max(axis=axis, keepdims=True)
    e = np.exp(a)
    return _wrap(e / e.sum(axis=axis, keepdims=True))


def conv2d(x, w, stride=1, padding=0, dilation=1, groups=1):
    """NHWC input, (O, kH, kW, I) weight, cross-correlation -- MLX's documented layout."""
    import torch
    import torch.nn.functional as F

    xt = torch.from_numpy(np.ascontiguousarray(_np(x), dtype=np.float32)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.ascontiguousarray(_np(w), dtype=np.float32)).permute(0, 3, 1, 2)
    y = F.conv2d(xt, wt, None, stride=stride, padding=padding, dilation=dilation, groups=groups)
    return _wrap(y.permute(0, 2, 3, 1).contiguous().numpy())


# --- mx.random -----------------------------------------------------------------
_rng = np.random.default_rng(0)


def _seed(s):
    global _rng
    _rng = np.random.default_rng(int(s))


def _normal(shape=(), dtype=float32, loc=0.0, scale=1.0, key=None):
    return _wrap((_rng.standard_normal(shape) * scale + loc).astype(dtype))


def _key(s):
    return _wrap(np.array([0, int(s)], dtype=np.uint32))


def _uniform(low=0.0, high=1.0, shape=(), dtype=float32, key=None):
    return _wrap((_rng.uniform(low, high, size=shape)).astype(dtype))


random = types.SimpleNamespace(seed=_seed, normal=_normal, key=_key, uniform=_uniform)


# --- mx.fast ---------------------------------------------------------------------
def _rms_norm(x, weight, eps):
    x32 = _np(x).astype(np.float32)
    y = x32 * (1.0 / np.sqrt(np.mean(x32 * x32, axis=-1, keepdims=True, dtype=np.float32) + np.float32(eps)))
    if weight is not None:
        y = y * _np(weight).astype(np.float32)
    return _wrap(y.astype(_np(x).dtype if _np(x).dtype != np.float64 else np.float32))


def _sdpa(q, k, v, scale=None, mask=None):
    q32, k32, v32 = (_np(t).astype(np.float32) for t in (q, k, v))
    if scale is None:
        scale = 1.0 / np.sqrt(q32.shape[-1])
    s = np.matmul(q32, np.swapaxes(k32, -1, -2)) * np.float32(scale)
    if mask is not None:
        s = s + _np(mask).astype(np.float32)
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(axis=-1, keepdims=True)
    return _wrap(np.matmul(p, v32).astype(np.float32))


class _MetalKernelStub:
    """`mx.fast.metal_kernel` is Apple-only; the DiT/VAE hot path never invokes the
    three Metal kernels (SURVEY.md section 2 kernel table), so constructing is allowed and
    calling is an error."""

    def __init__(self, **kw):
        self.name = kw.get("name")

    def __call__(self, *a, **k):
        raise RuntimeError(f"Metal kernel {self.name!r} cannot run in the numpy shim")


fast = types.SimpleNamespace(
    rms_norm=_rms_norm,
    scaled_dot_product_attention=_sdpa,
    metal_kernel=lambda **kw: _MetalKernelStub(**kw),
)

metal = types.SimpleNamespace(clear_cache=lambda: None)
