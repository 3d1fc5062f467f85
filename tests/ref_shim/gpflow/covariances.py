"""gpflow.covariances: Kuu(inducing, kernel, jitter) = k(Z) + jitter·I ; Kuf(inducing, kernel, Xnew) = k(Z, Xnew)."""
import tensorflow as tf


def Kuu(inducing_variable, kernel, *, jitter=0.0):
    Kzz = kernel(inducing_variable.Z)
    Kzz = Kzz + jitter * tf.eye(len(inducing_variable), dtype=Kzz.dtype)
    return Kzz


def Kuf(inducing_variable, kernel, Xnew):
    return kernel(inducing_variable.Z, Xnew)
