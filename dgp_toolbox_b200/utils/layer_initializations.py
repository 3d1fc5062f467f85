"""Layer-list construction, mirroring dgp_dace/utils/layer_initializations.py:24-68 (pure host set-up, numpy)."""
from __future__ import annotations

import numpy as np

from .. import gpflow_shim as gpflow
from .layers import SVGP_Layer


def _np(x):
    return np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=np.float64)


def init_layers_linear(X, Y, Z, kernels, num_units, num_outputs=None, mean_function=None, Layer=SVGP_Layer, white=False):
    """Hidden layers get an Identity mean function when widths match (:41-42), a PCA projection when narrowing (:45-47)
    or identity-plus-zero-padding when widening (:49-50), both as non-trainable Linear (:52-55); Z (and the running X)
    are projected with the same W (:59-61). The last layer gets `mean_function` (Zero by default) (:64-67)."""
    X, Z = _np(X), _np(Z)
    num_outputs = num_outputs or _np(Y).shape[1]
    mean_function = gpflow.Zero() if mean_function is None else mean_function
    layers = []
    dims = [X.shape[1]] + list(num_units)
    X_running, Z_running = X.copy(), Z.copy()
    for dim_in, dim_out, kern in zip(dims[:-1], dims[1:], kernels[:-1]):
        if dim_in == dim_out:
            mf = gpflow.Identity()
        else:
            if dim_in > dim_out:
                _, _, V = np.linalg.svd(X_running, full_matrices=False)
                W = V[:dim_out, :].T
            else:
                W = np.concatenate([np.eye(dim_in), np.zeros((dim_in, dim_out - dim_in))], 1)
            mf = gpflow.Linear(W)
            gpflow.set_trainable(mf, False)
        layers.append(Layer(kern, Z_running, dim_out, mf, white=white))
        if dim_in != dim_out:
            Z_running = Z_running.dot(W)
            X_running = X_running.dot(W)
    layers.append(Layer(kernels[-1], Z_running, num_outputs, mean_function, white=white))
    return layers
