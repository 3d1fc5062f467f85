"""gpflow.mean_functions: Zero, Identity, Linear."""
import numpy as np
import tensorflow as tf

from .base import Module, Parameter


class MeanFunction(Module):
    pass


class Zero(MeanFunction):
    def __init__(self, output_dim=1):
        MeanFunction.__init__(self)
        self.output_dim = output_dim

    def __call__(self, X):
        X = tf.convert_to_tensor(X)
        return tf.zeros(tuple(X.shape[:-1]) + (self.output_dim,), dtype=X.dtype)


class Identity(MeanFunction):
    def __call__(self, X):
        return tf.convert_to_tensor(X)


class Linear(MeanFunction):
    def __init__(self, A=None, b=None):
        MeanFunction.__init__(self)
        A = np.ones((1, 1)) if A is None else A
        b = np.zeros(1) if b is None else b
        self.A = Parameter(np.atleast_2d(A))
        self.b = Parameter(b)

    def __call__(self, X):
        return tf.tensordot(X, self.A, [[-1], [0]]) + self.b
