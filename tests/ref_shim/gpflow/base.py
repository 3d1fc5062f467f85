"""gpflow.base: Parameter (constrained value = bijector(unconstrained tf.Variable)) and Module."""
import numpy as np
import tensorflow as tf
import tensorflow_probability as tfp

_c = tf.convert_to_tensor


class Module(tf.Module):
    @property
    def parameters(self):
        return tuple(p for p in self._walk(set()) if isinstance(p, Parameter))

    @property
    def trainable_parameters(self):
        return tuple(p for p in self.parameters if p.trainable)


class Parameter(tf.Module):
    """gpflow.Parameter(value, transform=None, trainable=True): gradients are taken w.r.t. `unconstrained_variable`
    (= transform.inverse(value)); reading the parameter applies transform.forward."""

    def __init__(self, value, *, transform=None, prior=None, trainable=True, dtype=None, name=None):
        tf.Module.__init__(self, name=name)
        if isinstance(value, Parameter):
            transform = transform or value.transform
            value = value.numpy()
        self.transform = transform or tfp.bijectors.Identity()
        self.prior = prior
        self._unconstrained = tf.Variable(self.transform.inverse(_c(value, tf.float64)), trainable=trainable)

    @property
    def unconstrained_variable(self):
        return self._unconstrained

    @property
    def trainable(self):
        return self._unconstrained.trainable

    @trainable.setter
    def trainable(self, flag):
        self._unconstrained.trainable = bool(flag)

    def _as_tensor(self):
        return self.transform.forward(self._unconstrained._as_tensor())

    def value(self):
        return self._as_tensor()

    read_value = value

    def numpy(self):
        return self._as_tensor().numpy().copy()

    def assign(self, value):
        self._unconstrained.assign(self.transform.inverse(_c(value, tf.float64)))
        return self

    @property
    def shape(self):
        return tuple(self._as_tensor().shape)

    @property
    def dtype(self):
        return tf.float64

    def __array__(self, dtype=None, copy=None):
        return self.numpy()

    def __len__(self):
        return self.shape[0]


def _binop(name):
    def f(self, other):
        return getattr(self._as_tensor(), name)(other)
    return f


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__", "__pow__",
           "__rpow__", "__matmul__", "__getitem__", "__lt__", "__gt__", "__le__", "__ge__"):
    setattr(Parameter, _n, _binop(_n))
Parameter.__neg__ = lambda self: -self._as_tensor()
