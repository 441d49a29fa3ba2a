"""Minimal stand-in for the `jax` top-level module (see ../README.md)."""
import torch

from . import numpy  # noqa: F401
from . import scipy  # noqa: F401
from .numpy import _wrap, JArray

lax = None  # imported by the reference (TQ:4), never used


def jit(fn=None, static_argnums=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


def vmap(fn, in_axes=0, out_axes=0):
    assert in_axes == 0 and out_axes == 0

    def mapped(x):
        return _wrap(torch.stack([fn(row) for row in x], dim=0))
    return mapped


def value_and_grad(fn, argnums=0):
    single = isinstance(argnums, int)
    idx = (argnums,) if single else tuple(argnums)

    def wrapped(*args):
        args = list(args)
        leaves = []
        for i in idx:
            leaf = torch.as_tensor(args[i], dtype=torch.float64).detach().clone().requires_grad_(True)
            args[i] = _wrap(leaf)
            leaves.append(leaf)
        val = fn(*args)
        grads = torch.autograd.grad(val, leaves)
        grads = tuple(_wrap(g) for g in grads)
        return _wrap(val.detach()), (grads[0] if single else grads)
    return wrapped
