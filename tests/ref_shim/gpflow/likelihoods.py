"""gpflow.likelihoods.Gaussian (2.0): variance Parameter with positive(lower=1e-6)."""
import numpy as np
import tensorflow as tf

from .base import Module, Parameter
from .utilities import positive


class Likelihood(Module):
    pass


class Gaussian(Likelihood):
    DEFAULT_VARIANCE_LOWER_BOUND = 1e-6

    def __init__(self, variance=1.0, variance_lower_bound=DEFAULT_VARIANCE_LOWER_BOUND, **kwargs):
        Module.__init__(self, **kwargs)
        self.variance = Parameter(variance, transform=positive(lower=variance_lower_bound))

    def variational_expectations(self, Fmu, Fvar, Y):
        return -0.5 * np.log(2 * np.pi) - 0.5 * tf.math.log(self.variance) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / self.variance

    def predict_mean_and_var(self, Fmu, Fvar):
        return tf.identity(Fmu), Fvar + self.variance

    def conditional_mean(self, F):
        return tf.identity(F)

    def conditional_variance(self, F):
        return tf.fill(tuple(F.shape), 1.0) * tf.squeeze(self.variance.value())
