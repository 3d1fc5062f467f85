"""Layer-list construction with the behaviour of dgp_dace/utils/layer_initializations.py:24-68 (pure host set-up, numpy):
which mean function each hidden layer gets and how the inducing inputs follow the width changes."""
from __future__ import annotations

import numpy as np

from .. import gpflow_shim as gpflow
from .layers import SVGP_Layer


def _np(x):
    return np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=np.float64)


def width_change_map(inputs, d_in, d_out):
    """The fixed [d_in, d_out] linear map between two layer widths, or None when they are equal (Identity mean function,
    :41-42). Narrowing projects on the leading principal directions of the layer's inputs (:45-47); widening copies the
    coordinates and pads with zeros (:49-50)."""
    if d_in == d_out:
        return None
    if d_out < d_in:
        principal = np.linalg.svd(inputs, full_matrices=False)[2]     # rows: right singular vectors
        return principal[:d_out].T
    pad = np.zeros((d_in, d_out))
    pad[np.arange(d_in), np.arange(d_in)] = 1.0
    return pad


def init_layers_linear(X, Y, Z, kernels, num_units, num_outputs=None, mean_function=None, Layer=SVGP_Layer, white=False):
    """One SVGP layer per kernel. A hidden layer whose width changes gets the non-trainable Linear mean function of
    width_change_map (:52-55), and the inducing inputs (and the data used for the next projection) are pushed through the same
    map (:59-61); the last layer gets `mean_function` (Zero by default) and `num_outputs` columns (:64-67)."""
    running_X, running_Z = _np(X).copy(), _np(Z).copy()
    widths = [running_X.shape[1], *num_units]
    layers = []
    for (d_in, d_out), kern in zip(zip(widths[:-1], widths[1:]), kernels[:-1]):   # as many hidden layers as both lists allow
        W = width_change_map(running_X, d_in, d_out)
        if W is None:
            mean = gpflow.Identity()
        else:
            mean = gpflow.Linear(W)
            gpflow.set_trainable(mean, False)
        layers.append(Layer(kern, running_Z, d_out, mean, white=white))
        if W is not None:
            running_Z, running_X = running_Z @ W, running_X @ W
    last_mean = gpflow.Zero() if mean_function is None else mean_function
    layers.append(Layer(kernels[-1], running_Z, num_outputs or _np(Y).shape[1], last_mean, white=white))
    return layers
