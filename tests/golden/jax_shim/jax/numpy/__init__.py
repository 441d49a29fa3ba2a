"""Minimal stand-in for `jax.numpy`, backed by torch CPU float64/complex128 (see ../../README.md)."""
import math
import types

import numpy as _np
import torch

float64 = torch.float64
complex128 = torch.complex128
pi = math.pi


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, value):
                out = arr.clone()
                out[idx] = value
                return out
        return _Setter()


class JArray(torch.Tensor):
    """torch.Tensor plus the ndarray methods the reference files call."""

    def astype(self, dtype):
        return self.to(dtype)

    @property
    def at(self):
        return _At(self)

    @property
    def T(self):
        if self.dim() < 2:
            return self
        return self.transpose(-2, -1) if self.dim() == 2 else self.permute(*reversed(range(self.dim())))

    def __lt__(self, other):       # NumPy orders complex numbers lexicographically (real, then imaginary)
        if self.is_complex():
            o = torch.as_tensor(other, dtype=self.dtype)
            return (self.real < o.real) | ((self.real == o.real) & (self.imag < o.imag))
        return torch.Tensor.__lt__(self, other)


ndarray = JArray


def _wrap(t):
    return t.as_subclass(JArray) if isinstance(t, torch.Tensor) else t


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    arr = _np.asarray(x)
    if dtype is None:
        dtype = torch.complex128 if _np.iscomplexobj(arr) else torch.float64
    return torch.as_tensor(arr).to(dtype)


def asarray(x, dtype=None):
    return _wrap(_t(x, dtype))


array = asarray


def zeros(shape, dtype=float64):
    return _wrap(torch.zeros(shape, dtype=dtype))


def eye(n, dtype=float64):
    return _wrap(torch.eye(n, dtype=dtype))


def arange(n, dtype=float64):
    return _wrap(torch.arange(n, dtype=dtype))


def linspace(start, stop, num, dtype=float64):
    return _wrap(torch.as_tensor(_np.linspace(start, stop, num)).to(dtype))    # NumPy's own formula


def conj(x):
    x = _t(x)
    return _wrap(torch.conj(x).resolve_conj() if x.is_complex() else x)


def real(x):
    x = _t(x)
    return _wrap(x.real if x.is_complex() else x)


def imag(x):
    x = _t(x)
    return _wrap(x.imag if x.is_complex() else torch.zeros_like(x))


def _unary(fn):
    return lambda x: _wrap(fn(_t(x)))


exp, tanh, sqrt, square, trace = (_unary(f) for f in (torch.exp, torch.tanh, torch.sqrt, torch.square, torch.trace))


def kron(a, b):
    return _wrap(torch.kron(_t(a), _t(b)))


def outer(a, b):
    return _wrap(torch.outer(_t(a), _t(b)))


def matmul(a, b):
    return _wrap(torch.matmul(_t(a), _t(b)))


def tensordot(a, b, axes):
    return _wrap(torch.tensordot(_t(a), _t(b), dims=(list(axes[0]), list(axes[1]))))


def stack(xs, axis=0):
    return _wrap(torch.stack([_t(x) for x in xs], dim=axis))


def concatenate(xs, axis=0):
    xs = [_t(x) for x in xs]
    dt = torch.complex128 if any(x.is_complex() for x in xs) else torch.float64
    return _wrap(torch.cat([x.to(dt) for x in xs], dim=axis))


def block(rows):
    return _wrap(torch.cat([torch.cat([_t(x) for x in row], dim=1) for row in rows], dim=0))


def pad(x, widths, mode="constant"):
    assert mode == "constant" and len(widths) == 2 and tuple(widths[1]) == (0, 0)
    lo, hi = widths[0]
    x = _t(x)
    z = lambda k: torch.zeros((k,) + tuple(x.shape[1:]), dtype=x.dtype)  # noqa: E731
    return _wrap(torch.cat([z(lo), x, z(hi)], dim=0))


def where(cond, a, b):
    a_t, b_t = isinstance(a, torch.Tensor), isinstance(b, torch.Tensor)
    dt = (a.dtype if a_t else b.dtype) if (a_t or b_t) else torch.float64
    if (a_t and a.is_complex()) or (b_t and b.is_complex()):
        dt = torch.complex128
    return _wrap(torch.where(cond, _t(a, dt), _t(b, dt)))


def squeeze(x):
    return _wrap(torch.squeeze(_t(x)))


fft = types.SimpleNamespace(
    fft=lambda x: _wrap(torch.fft.fft(_t(x))),
    ifft=lambda x: _wrap(torch.fft.ifft(_t(x))),
    fftshift=lambda x: _wrap(torch.fft.fftshift(_t(x))),
    ifftshift=lambda x: _wrap(torch.fft.ifftshift(_t(x))),
    fftfreq=lambda n, d=1.0: _wrap(torch.as_tensor(_np.fft.fftfreq(n, d=d))),
)
