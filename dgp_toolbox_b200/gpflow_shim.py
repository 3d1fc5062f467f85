"""Minimal stand-ins for the GPflow objects the reference's DGP path holds (kernels, likelihood, mean functions,
Parameter). Same attribute names and semantics as GPflow 2.0 so reference call sites keep working
(`kern.K`, `kern.K_diag`, `kern.variance`, `kern.lengthscales`, `likelihood.variance`, `Parameter.numpy/assign`):
  * SquaredExponential / RBF: dgp_dace call sites utils/layers.py:221,230,243,272
  * Gaussian: utils/utils.py:89-93,108-111
  * Zero / Identity / Linear: utils/layer_initializations.py:27,42,52
Values live in float64 CUDA tensors; kernel evaluations run in libdgp_b200 (dgp_kernel_K).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

DEFAULT_JITTER = 1e-6  # gpflow.default_jitter()


def default_jitter():
    return DEFAULT_JITTER


def default_float():
    return torch.float64


class Parameter:
    """A GPflow-style parameter: constrained value on the device + the name of its bijector.

    transform: None (identity) | "positive" (softplus) | "positive_shift" (softplus + 1e-6, GPflow likelihood variance)
               | "triangular" (tfp FillTriangular: a permutation of the lower triangle)
    """

    def __init__(self, value, transform=None, trainable=True, name=None, device=None):
        self.value = _lib.as_device(value, device).clone()
        self.transform = transform
        self.trainable = trainable
        self.name = name

    def numpy(self):
        return self.value.detach().cpu().numpy()

    def assign(self, value):
        v = _lib.as_device(value, self.value.device)
        if v.shape != self.value.shape:
            v = v.reshape(self.value.shape)
        self.value.copy_(v)
        return self

    @property
    def shape(self):
        return tuple(self.value.shape)

    def __mul__(self, other):
        return self.value * other

    __rmul__ = __mul__

    def __array__(self, dtype=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __dlpack__(self, *a, **k):
        return self.value.__dlpack__(*a, **k)

    def __dlpack_device__(self):
        return self.value.__dlpack_device__()

    def __repr__(self):
        return f"Parameter(name={self.name}, shape={self.shape}, transform={self.transform}, trainable={self.trainable})"

    # -- unconstrained-space helpers (what tf.GradientTape differentiates in the reference) --
    def lower(self):
        return 1e-6 if self.transform == "positive_shift" else 0.0

    def unconstrained(self):
        if self.transform in ("positive", "positive_shift"):
            y = self.value - self.lower()
            return y + torch.log(-torch.expm1(-y))          # softplus^-1
        return self.value.clone()

    def set_unconstrained(self, u):
        if self.transform in ("positive", "positive_shift"):
            self.value.copy_(torch.nn.functional.softplus(u) + self.lower())
        elif self.transform == "triangular":
            self.value.copy_(torch.tril(u))
        else:
            self.value.copy_(u)

    def grad_to_unconstrained(self, g):
        """Chain rule d constrained / d unconstrained applied to a constrained-space gradient."""
        if self.transform in ("positive", "positive_shift"):
            return g * (-torch.expm1(-(self.value - self.lower())))   # sigmoid(u) = 1 - exp(-softplus(u))
        if self.transform == "triangular":
            return torch.tril(g)
        return g


def set_trainable(obj, flag: bool):
    if isinstance(obj, Parameter):
        obj.trainable = flag
        return
    for p in getattr(obj, "parameters", []):
        p.trainable = flag


class _Module:
    @property
    def parameters(self):
        out = []
        for v in self.__dict__.values():
            if isinstance(v, Parameter):
                out.append(v)
            elif isinstance(v, _Module):
                out.extend(v.parameters)
            elif isinstance(v, (list, tuple)):
                for e in v:
                    if isinstance(e, _Module):
                        out.extend(e.parameters)
                    elif isinstance(e, Parameter):
                        out.append(e)
        return out

    @property
    def trainable_parameters(self):
        return [p for p in self.parameters if p.trainable]


class SquaredExponential(_Module):
    """GPflow kernels.SquaredExponential: variance * exp(-0.5 * |(x - x') / lengthscales|^2); ARD when lengthscales is a vector."""
    kernel_kind = 0

    def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None, name=None):
        if active_dims is not None:
            raise NotImplementedError("active_dims is out of scope of the accelerated path (SURVEY §8 f2)")
        self.variance = Parameter(np.asarray(variance, dtype=np.float64).reshape(()), transform="positive", name="variance")
        self.lengthscales = Parameter(np.asarray(lengthscales, dtype=np.float64), transform="positive", name="lengthscales")
        self.name = name or "squared_exponential"

    @property
    def ard(self):
        return self.lengthscales.value.dim() > 0 and self.lengthscales.value.numel() > 1

    def lengthscales_vector(self, D):
        ls = self.lengthscales.value.reshape(-1)
        if ls.numel() == 1 and D == 1 and ls.is_contiguous():
            return ls      # a view of the parameter itself: in-library optimiser loops update it in place, no mirror needed
        if ls.numel() == 1:
            # the C ABI takes one lengthscale per input column; the expansion lives in one buffer per kernel object so that
            # its address is stable from call to call (graph replay, dgp_adam_param.mirror)
            buf = getattr(self, "_ls_vec", None)
            if buf is None or buf.numel() != D or buf.device != ls.device:
                buf = self._ls_vec = torch.empty(D, dtype=torch.float64, device=ls.device)
            buf.copy_(ls.expand(D))
            return buf
        if ls.numel() != D:
            raise ValueError(f"lengthscales has {ls.numel()} entries, input has {D} columns")
        return ls.contiguous()

    def K(self, X, X2=None):
        X = _lib.as_device(X)
        X2 = X if X2 is None else _lib.as_device(X2)
        D = X.shape[-1]
        ctx = _lib.get_context(X.device)
        out = torch.empty((X.shape[0], X2.shape[0]), dtype=torch.float64, device=X.device)
        ls = self.lengthscales_vector(D)
        var = self.variance.value.reshape(1)
        ctx.call("dgp_kernel_K", self.kernel_kind, D, _lib.ptr(ls), _lib.ptr(var), _lib.ptr(X), X.shape[0], _lib.ptr(X2), X2.shape[0], _lib.ptr(out))
        return out

    def K_diag(self, X):
        X = _lib.as_device(X)
        return self.variance.value.reshape(()).expand(X.shape[0]).clone()

    __call__ = K


RBF = SquaredExponential


class Matern32(SquaredExponential):
    """GPflow kernels.Matern32: variance (1 + sqrt3 r) exp(-sqrt3 r), r = sqrt(max(r2, 1e-36)) (BO/SO_BO.py:194,241)."""
    kernel_kind = 1

    def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None, name=None):
        SquaredExponential.__init__(self, variance, lengthscales, active_dims, name or "matern32")


class Matern52(SquaredExponential):
    """GPflow kernels.Matern52: variance (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r) (BO/SO_BO.py:196,243)."""
    kernel_kind = 2

    def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None, name=None):
        SquaredExponential.__init__(self, variance, lengthscales, active_dims, name or "matern52")


class Gaussian(_Module):
    """GPflow likelihoods.Gaussian (variance has the softplus + 1e-6 shift transform)."""

    def __init__(self, variance=1.0):
        self.variance = Parameter(np.asarray(variance, dtype=np.float64).reshape(()), transform="positive_shift", name="variance")

    def variational_expectations(self, Fmu, Fvar, Y):
        v = self.variance.value
        return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(v) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / v

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + self.variance.value


class Zero(_Module):
    mean_kind = 0

    def __call__(self, X):
        return torch.zeros(X.shape[:-1] + (1,), dtype=torch.float64, device=X.device)


class Identity(_Module):
    mean_kind = 1

    def __call__(self, X):
        return X


class Linear(_Module):
    mean_kind = 2

    def __init__(self, A=None, b=None):
        A = np.ones((1, 1)) if A is None else np.asarray(A, dtype=np.float64)
        b = np.zeros(A.shape[1]) if b is None else np.asarray(b, dtype=np.float64).reshape(-1)
        if b.size == 1 and A.shape[1] > 1:
            b = np.full(A.shape[1], float(b[0]))
        self.A = Parameter(A, name="A")
        self.b = Parameter(b, name="b")

    def __call__(self, X):
        return X @ self.A.value + self.b.value
