"""Host-side mirrors of dgp_dace/utils/utils.py: reparameterize (diag branch) and BroadcastingLikelihood (Gaussian)."""
from __future__ import annotations

import torch

from .. import gpflow_shim as gpflow


def reparameterize(mean, var, z, full_cov=False):
    """dgp_dace/utils/utils.py:22-51, diagonal branch (:40-41): mean + z * (var + jitter) ** 0.5.
    Inside DGP.propagate this arithmetic is fused into the layer kernel; this function serves direct callers."""
    if var is None:
        return mean
    if full_cov:
        raise NotImplementedError("full_cov=True is out of scope of the accelerated path (SURVEY §8 f4)")
    return mean + z * (var + gpflow.default_jitter()) ** 0.5


class BroadcastingLikelihood:
    """dgp_dace/utils/utils.py:54-117 for the Gaussian likelihood: [S,N,D] moments against [N,D] targets need no
    reshape (utils/utils.py:66-74), Y is broadcast on a new leading axis."""

    def __init__(self, likelihood):
        self.likelihood = likelihood
        if not isinstance(likelihood, gpflow.Gaussian):
            raise NotImplementedError("only the Gaussian likelihood is on the accelerated path")

    def variational_expectations(self, Fmu, Fvar, Y):
        return self.likelihood.variational_expectations(Fmu, Fvar, Y[None] if Y.dim() == Fmu.dim() - 1 else Y)

    def predict_mean_and_var(self, Fmu, Fvar):
        return self.likelihood.predict_mean_and_var(Fmu, Fvar)
