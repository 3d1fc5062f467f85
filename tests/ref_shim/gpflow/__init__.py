"""Minimal `gpflow` (2.0 API) stand-in on the tensorflow shim (TEST INFRASTRUCTURE, see ../README.md).
Only the symbols the reference's hot-path modules import (SURVEY.md §2 third-party table)."""
import numpy as np
import tensorflow as tf

from . import base, covariances, inducing_variables, kernels, likelihoods, mean_functions, models, optimizers, utilities
from .base import Module, Parameter
from .utilities import set_trainable

_JITTER = 1e-6


def default_float():
    return tf.float64


def default_jitter():
    return _JITTER


def default_int():
    return tf.int32
