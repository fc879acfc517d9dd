"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the `mlx.nn` pieces the
reference's hot path uses: Module (attribute container), Linear (y = x W^T + b,
weight stored (out, in) like PyTorch), LayerNorm(affine=False), SiLU, silu,
gelu_approx (tanh form).  See package docstring."""
from __future__ import annotations

import math

import numpy as np

from . import core as mx


class Module:
    def __init__(self):
        pass

    def parameters(self):
        out = {}
        for k, v in vars(self).items():
            if isinstance(v, np.ndarray):
                out[k] = v
            elif isinstance(v, Module):
                out[k] = v.parameters()
            elif isinstance(v, list) and v and isinstance(v[0], Module):
                out[k] = [m.parameters() for m in v]
        return out

    def update(self, tree):
        for k, v in tree.items():
            cur = getattr(self, k, None)
            if isinstance(v, dict) and isinstance(cur, Module):
                cur.update(v)
            elif isinstance(v, (list, dict)) and isinstance(cur, list):
                items = v.items() if isinstance(v, dict) else enumerate(v)
                for i, sub in items:
                    cur[int(i)].update(sub)
            else:
                setattr(self, k, v)
        return self

    def eval(self):
        return self


class Linear(Module):
    def __init__(self, input_dims, output_dims, bias=True):
        super().__init__()
        k = math.sqrt(1.0 / input_dims)
        rng = np.random.default_rng(input_dims * 7919 + output_dims)
        self.weight = mx.array(rng.uniform(-k, k, (output_dims, input_dims)).astype(np.float32))
        if bias:
            self.bias = mx.array(rng.uniform(-k, k, (output_dims,)).astype(np.float32))

    def __call__(self, x):
        y = np.matmul(np.asarray(x), np.asarray(self.weight).T)
        if "bias" in vars(self):
            y = y + np.asarray(self.bias)
        return mx._wrap(y)


class LayerNorm(Module):
    def __init__(self, dims, eps=1e-5, affine=True, bias=True):
        super().__init__()
        self.eps = eps
        self.dims = dims
        if affine:
            self.weight = mx.ones((dims,))
            if bias:
                self.bias = mx.zeros((dims,))

    def __call__(self, x):
        x = np.asarray(x).astype(np.float32)
        mu = x.mean(axis=-1, keepdims=True)
        var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
        y = (x - mu) / np.sqrt(var + np.float32(self.eps))
        if "weight" in vars(self):
            y = y * np.asarray(self.weight)
            if "bias" in vars(self):
                y = y + np.asarray(self.bias)
        return mx._wrap(y)


class GroupNorm(Module):
    """Parameter container only: the reference's upscaler stores weight/bias here and applies the normalisation itself
    (model/upscaler/spatial.py:91-128 group_norm_5d); calling it is not part of any restated path."""

    def __init__(self, num_groups, dims, eps=1e-5, affine=True, pytorch_compatible=False):
        super().__init__()
        self.num_groups = num_groups
        self.dims = dims
        self.eps = eps
        if affine:
            self.weight = mx.ones((dims,))
            self.bias = mx.zeros((dims,))

    def __call__(self, x):
        raise NotImplementedError("mlx shim: nn.GroupNorm.__call__ is not restated (unused by the reference paths)")


def silu(x):
    x = np.asarray(x)
    return mx._wrap(x / (1.0 + np.exp(-x)))


def gelu_approx(x):
    x = np.asarray(x)
    c = np.float32(math.sqrt(2.0 / math.pi))
    return mx._wrap(np.float32(0.5) * x * (1.0 + np.tanh(c * (x + np.float32(0.044715) * x * x * x))))


def gelu(x):
    from math import sqrt

    import torch

    return mx._wrap(torch.nn.functional.gelu(torch.from_numpy(np.asarray(x, dtype=np.float32))).numpy())


class SiLU(Module):
    def __call__(self, x):
        return silu(x)


class GELU(Module):
    def __init__(self, approx="none"):
        super().__init__()
        self.approx = approx

    def __call__(self, x):
        return gelu_approx(x) if self.approx != "none" else gelu(x)
