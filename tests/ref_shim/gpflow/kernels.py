"""gpflow.kernels (2.0): stationary kernels with the expanded-square distance, Linear, White, Sum, Product, active_dims."""
import numpy as np
import tensorflow as tf

from .base import Module, Parameter
from .utilities import positive

_c = tf.convert_to_tensor


def square_distance(X, X2):
    """gpflow.utilities.ops.square_distance: ‖x‖² + ‖x'‖² − 2 x·x' (no clamp at zero)."""
    if X2 is None:
        Xs = tf.reduce_sum(tf.square(X), axis=-1, keepdims=True)
        dist = -2 * tf.matmul(X, X, transpose_b=True)
        dist = dist + Xs + tf.linalg.adjoint(Xs)
        return dist
    Xs = tf.reduce_sum(tf.square(X), axis=-1)
    X2s = tf.reduce_sum(tf.square(X2), axis=-1)
    dist = -2 * tf.tensordot(X, X2, [[-1], [-1]])
    dist = dist + Xs[..., :, None] + X2s[..., None, :]
    return dist


class Kernel(Module):
    def __init__(self, active_dims=None, name=None):
        Module.__init__(self, name=name)
        self.active_dims = None if active_dims is None else list(active_dims)

    def slice(self, X, X2=None):
        X = _c(X)
        X2 = None if X2 is None else _c(X2)
        if self.active_dims is not None:
            X = X[..., self.active_dims]
            if X2 is not None:
                X2 = X2[..., self.active_dims]
        return X, X2

    def __call__(self, X, X2=None, *, full_cov=True, presliced=False):
        if not full_cov:
            return self.K_diag(X)
        return self.K(X, X2)

    def __add__(self, other):
        return Sum([self, other])

    def __mul__(self, other):
        return Product([self, other])


class Combination(Kernel):
    """Sum/Product flatten nested combinations of the same type, like GPflow's `Combination._set_kernels`."""

    def __init__(self, kernels, name=None):
        Kernel.__init__(self, name=name)
        flat = []
        for k in kernels:
            flat.extend(k.kernels if isinstance(k, type(self)) else [k])
        self.kernels = flat


class Sum(Combination):
    def K(self, X, X2=None):
        out = self.kernels[0].K(X, X2)
        for k in self.kernels[1:]:
            out = out + k.K(X, X2)
        return out

    def K_diag(self, X):
        out = self.kernels[0].K_diag(X)
        for k in self.kernels[1:]:
            out = out + k.K_diag(X)
        return out


class Product(Combination):
    def K(self, X, X2=None):
        out = self.kernels[0].K(X, X2)
        for k in self.kernels[1:]:
            out = out * k.K(X, X2)
        return out

    def K_diag(self, X):
        out = self.kernels[0].K_diag(X)
        for k in self.kernels[1:]:
            out = out * k.K_diag(X)
        return out


class Stationary(Kernel):
    def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None, name=None, **kwargs):
        Kernel.__init__(self, active_dims=active_dims, name=name)
        lengthscales = kwargs.pop("lengthscale", lengthscales)
        self.variance = Parameter(variance, transform=positive())
        self.lengthscales = Parameter(lengthscales, transform=positive())

    @property
    def ard(self):
        return len(self.lengthscales.shape) > 0

    def scale(self, X):
        return X / self.lengthscales if X is not None else X

    def scaled_squared_euclid_dist(self, X, X2=None):
        return square_distance(self.scale(X), self.scale(X2))

    def K(self, X, X2=None):
        X, X2 = self.slice(X, X2)
        return self.K_r2(self.scaled_squared_euclid_dist(X, X2))

    def K_diag(self, X):
        X = _c(X)
        return tf.fill(tuple(X.shape[:-1]), 1.0) * tf.squeeze(self.variance.value())

    def K_r2(self, r2):
        r = tf.sqrt(tf.maximum(r2, 1e-36))
        return self.K_r(r)


class SquaredExponential(Stationary):
    def K_r2(self, r2):
        return self.variance * tf.exp(-0.5 * r2)


RBF = SquaredExponential


class Matern32(Stationary):
    def K_r(self, r):
        sqrt3 = np.sqrt(3.0)
        return self.variance * (1.0 + sqrt3 * r) * tf.exp(-sqrt3 * r)


class Matern52(Stationary):
    def K_r(self, r):
        sqrt5 = np.sqrt(5.0)
        return self.variance * (1.0 + sqrt5 * r + 5.0 / 3.0 * tf.square(r)) * tf.exp(-sqrt5 * r)


class Linear(Kernel):
    def __init__(self, variance=1.0, active_dims=None, name=None):
        Kernel.__init__(self, active_dims=active_dims, name=name)
        self.variance = Parameter(variance, transform=positive())

    def K(self, X, X2=None):
        X, X2 = self.slice(X, X2)
        if X2 is None:
            return tf.matmul(X * self.variance, X, transpose_b=True)
        return tf.tensordot(X * self.variance, X2, [[-1], [-1]])

    def K_diag(self, X):
        X, _ = self.slice(X, None)
        return tf.reduce_sum(tf.square(X) * self.variance, axis=-1)


class White(Kernel):
    def __init__(self, variance=1.0, active_dims=None, name=None):
        Kernel.__init__(self, active_dims=active_dims, name=name)
        self.variance = Parameter(variance, transform=positive())

    def K(self, X, X2=None):
        X = _c(X)
        if X2 is None:
            return self.variance * tf.eye(X.shape[-2])
        return tf.zeros((X.shape[-2], _c(X2).shape[-2]))

    def K_diag(self, X):
        X = _c(X)
        return tf.fill(tuple(X.shape[:-1]), 1.0) * tf.squeeze(self.variance.value())
