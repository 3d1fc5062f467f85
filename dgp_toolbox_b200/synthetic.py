"""Synthetic inputs of the benchmark configurations: X ~ N(0, I), Y = sin(sum_j x_j / sqrt(D0)) + 0.1 eps,
Z_l ~ N(0, I), ARD lengthscales sqrt(D_in), variance 1, q_mu = 0.1 N(0,1), q_sqrt = 0.5 I + 0.05 tril(N(0,1)), sigma_n^2 = 0.1.
All draws from numpy.random.default_rng with fixed seeds, float64. (tests/ check that the oracle's generator, which is
test infrastructure, produces the same arrays.)"""
from __future__ import annotations

import math

import numpy as np

CONFIGS = {
    # name: D0, num_units (hidden widths; L entries -> L+1 SVGP layers), M, S
    "c1": dict(D0=2, num_units=[2], M=50, S=10),
    "c2": dict(D0=8, num_units=[8, 8, 8], M=256, S=32),
    "c3": dict(D0=20, num_units=[20] * 5, M=512, S=64),
}


def synthetic_problem(D0, num_units, M, N, seed_shift=0, lik_var=0.1, ls_scale=1.0):
    rng = np.random.default_rng(0 + seed_shift)
    X = rng.standard_normal((N, D0))
    Y = np.sin(X.sum(1, keepdims=True) / math.sqrt(D0)) + 0.1 * np.random.default_rng(1 + seed_shift).standard_normal((N, 1))
    dims = [D0] + list(num_units) + [1]
    layers = []
    for l, (din, dout) in enumerate(zip(dims[:-1], dims[1:])):
        Z = np.random.default_rng(10 + l + seed_shift).standard_normal((M, din))
        q_mu = 0.1 * np.random.default_rng(20 + l + seed_shift).standard_normal((M, dout))
        R = np.random.default_rng(30 + l + seed_shift).standard_normal((dout, M, M))
        q_sqrt = 0.5 * np.eye(M)[None] + 0.05 * np.tril(R)
        last = l == len(dims) - 2
        if last:
            kind, W = "zero", None
        elif din == dout:
            kind, W = "identity", None
        elif din > dout:
            _, _, V = np.linalg.svd(X if l == 0 else rng.standard_normal((max(N, din), din)), full_matrices=False)
            kind, W = "linear", V[:dout, :].T
        else:
            kind, W = "linear", np.concatenate([np.eye(din), np.zeros((din, dout - din))], 1)
        layers.append(dict(Z=Z, lengthscales=np.full(din, ls_scale * math.sqrt(din)), variance=1.0, q_mu=q_mu, q_sqrt=q_sqrt,
                           mean_kind=kind, mf_W=W, mf_b=None if W is None else np.zeros(dout)))
    return dict(X=X, Y=Y, layers=layers, lik_var=lik_var)


def model_from_problem(prob, num_samples, seed=1234):
    """The problem as a DGP_Base on the current CUDA device."""
    from . import gpflow_shim as G
    from .models.dgp import DGP_Base
    from .utils.layers import SVGP_Layer
    layers = []
    for l in prob["layers"]:
        kern = G.SquaredExponential(variance=l["variance"], lengthscales=l["lengthscales"])
        mf = G.Zero() if l["mean_kind"] == "zero" else G.Identity() if l["mean_kind"] == "identity" else G.Linear(l["mf_W"], l["mf_b"])
        layer = SVGP_Layer(kern, l["Z"], l["q_mu"].shape[1], mf)
        layer.q_mu.assign(l["q_mu"])
        layer.q_sqrt.assign(l["q_sqrt"])
        layers.append(layer)
    return DGP_Base(G.Gaussian(prob["lik_var"]), layers, num_samples=num_samples, seed=seed)


def minibatch(D0, N, index):
    """Minibatch `index` of the synthetic stream (fresh draws per index)."""
    rng = np.random.default_rng(1000 + index)
    X = rng.standard_normal((N, D0))
    Y = np.sin(X.sum(1, keepdims=True) / math.sqrt(D0)) + 0.1 * rng.standard_normal((N, 1))
    return X, Y


def flops_per_layer(D0, num_units, M, vform=False):
    """Algorithmic (triangular-aware, useful) forward FP64 flops of each SVGP layer for ONE evaluation of its conditional
    (SURVEY §8d): (2 + D_out) M^2 + 2 M D_in + 2 M (2 D_out + 1) in the reference's operation order (V, A = Lu^-T V, T_d);
    the V-form (C_d = q_sqrt_d^T Lu^-T folded once per step, no A pass) needs (1 + D_out) M^2 + the same lower-order terms."""
    dims = [D0] + list(num_units) + [1]
    lead = 1 if vform else 2
    return [(lead + dout) * M * M + 2 * M * din + 2 * M * (2 * dout + 1) for din, dout in zip(dims[:-1], dims[1:])]


def param_flops_per_point_sample(D0, num_units, M, S):
    """Flops per point-sample of the contractions over the point-samples that the V-form adjoint needs: the D_out lower-only
    products W_d = V diag(2 Gv_d) V^T (D_out M^2; the reference order and the A-form need (1 + D_out) M^2 — G1 = tril(dV V^T) is
    derived from the W_d once per step), V Gm (2 M D_out) and Gbar [X, 1] (2 M (D_in + 1)); first layer once per point."""
    dims = [D0] + list(num_units) + [1]
    fl = [dout * M * M + 2 * M * dout + 2 * M * (din + 1) for din, dout in zip(dims[:-1], dims[1:])]
    return sum(fl) if len(fl) < 2 else fl[0] / S + sum(fl[1:])


def flops_per_point_sample(D0, num_units, M, S=None):
    """(forward, ELBO+grad = 3x forward) flops per point-sample. S = None: the reference's formulation, every layer evaluated
    for every point-sample in the reference's operation order. S given: what this implementation needs — V-form
    layers, and the first layer (whose input is shared by the S samples of a point) evaluated once per point."""
    fl = flops_per_layer(D0, num_units, M, vform=S is not None)
    f = sum(fl) if S is None or len(fl) < 2 else fl[0] / S + sum(fl[1:])
    return f, 3 * f
