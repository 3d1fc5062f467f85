"""Generates tests/golden/mf_dgp_em.npz by EXECUTING the reference's embedded-mapping multi-fidelity model
(`/root/reference/dgp_dace/models/MF_DGP_EM.py`, `utils/layers_red.py`, unmodified) on the stand-in tensorflow / gpflow of tests/ref_shim:
3 fidelities with input spaces of 2, 3 and 4 dimensions, non-trivial parameters; inputs, parameters, the N(0,1) draws in consumption
order, the ELBO (MF_DGP_EM.py:257-297) and its gradients w.r.t. every trainable parameter (constrained space).
    python tests/golden/make_golden_mf_em.py        (build container only)"""
import contextlib
import importlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refexec as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ns = R.load()
EM = importlib.import_module("dgp_dace.models.MF_DGP_EM")
assert EM.__file__.startswith(R.REFERENCE)
tf = ns.tf

rng = np.random.default_rng(7)
dims, Ns, Ms = [2, 3, 4], [12, 8, 6], [6, 5, 4]
X = [rng.uniform(0, 1, (n, d)) for n, d in zip(Ns, dims)]
f0 = lambda x: np.sin(3 * x[:, :1]) + 0.5 * x[:, 1:2]
Y = [f0(X[0]), 1.2 * f0(X[1]) + 0.3 * X[1][:, 2:3], 1.5 * f0(X[2]) - 0.2 * X[2][:, 3:4] * X[2][:, 2:3]]
X_red = [X[1][:, :2].copy(), X[2][:, :2].copy()]          # nominal mappings of the higher fidelities' inputs into the LOWEST input space (MF_DGP_EM.py:282-287 compares them with the fully projected H)
Z = [rng.uniform(0, 1, (m, d)) for m, d in zip(Ms, dims)]
W = [rng.uniform(0, 1, (5, 4)), rng.uniform(0, 1, (4, 3))]   # inducing inputs of the projection layers 4 -> 3 and 3 -> 2

recorded = []


def recording_source(shape):
    z = rng.standard_normal(shape)
    recorded.append(z)
    return z


tf.random.source = recording_source
with contextlib.redirect_stdout(io.StringIO()):
    model = EM.DGP_Base.make_mf_dgp([x.copy() for x in X], [z.copy() for z in Z], [w.copy() for w in W])
model.num_samples = 3


def kparams(layer, i):
    ks = layer.kern.kernels
    out = {}
    if i == 0:
        out.update(in_var=ks[0].variance, in_ls=ks[0].lengthscales, white_var=ks[1].variance)
    else:
        prod, k_in = ks[0], ks[1]
        k_corr, inner = prod.kernels
        out.update(corr_var=k_corr.variance, corr_ls=k_corr.lengthscales, prev_var=inner.kernels[0].variance,
                   prev_ls=inner.kernels[0].lengthscales, lin_var=inner.kernels[1].variance, in_var=k_in.variance, in_ls=k_in.lengthscales)
        if len(ks) > 2:
            out["white_var"] = ks[2].variance
    return out


params = {}
for i, layer in enumerate(model.layers):
    for name, p in kparams(layer, i).items():
        p.assign(1e-2 * (1 + i) if name == "white_var" else rng.uniform(0.6, 1.4, p.shape))
        params[f"layers.{i}.{name}"] = p
    M = layer.num_inducing
    layer.q_mu.assign(0.3 * rng.standard_normal((M, 1)))
    layer.q_sqrt.assign(0.6 * layer.q_sqrt.numpy() + 0.05 * np.tril(rng.standard_normal((1, M, M))))
    params[f"layers.{i}.q_mu"], params[f"layers.{i}.q_sqrt"] = layer.q_mu, layer.q_sqrt
    params[f"layers.{i}.Z"] = layer.feature.Z if i == 0 else layer.feature.Z_left
for i, layer in enumerate(model.layers_red):
    M, Do = layer.num_inducing, layer.num_outputs
    layer.kern.variance.assign(rng.uniform(0.6, 1.4))
    layer.kern.lengthscales.assign(rng.uniform(0.6, 1.4, layer.kern.lengthscales.shape))
    layer.q_mu.assign(0.3 * rng.standard_normal((M, Do)))
    layer.q_sqrt.assign(0.5 * layer.q_sqrt.numpy() + 0.05 * np.tril(rng.standard_normal((Do, M, M))))
    params[f"red.{i}.in_var"], params[f"red.{i}.in_ls"] = layer.kern.variance, layer.kern.lengthscales
    params[f"red.{i}.q_mu"], params[f"red.{i}.q_sqrt"], params[f"red.{i}.Z"] = layer.q_mu, layer.q_sqrt, layer.feature.Z
model.likelihood.likelihood.variance.assign(0.05)
model.likelihood_projection.likelihood.variance.assign(0.2)
params["lik_var"], params["lik_proj_var"] = model.likelihood.likelihood.variance, model.likelihood_projection.likelihood.variance

out = {"provenance": np.array("reference source (/root/reference/dgp_dace/models/MF_DGP_EM.py, utils/layers_red.py, unmodified) executed under "
                              "tests/ref_shim; InducingPoints(layers=..., layers_red=...) semantics assumed as documented there"),
       "S": np.int64(model.num_samples), "nfid": np.int64(3)}
for i in range(3):
    out[f"X{i}"], out[f"Y{i}"], out[f"Zinit{i}"] = X[i], Y[i], Z[i]
for i in range(2):
    out[f"Xred{i}"], out[f"Winit{i}"] = X_red[i], W[i]
for k, p in params.items():
    out["param_" + k] = p.numpy()

recorded.clear()
tvars = list(model.trainable_variables)
data = ([tf.constant(x) for x in X], [tf.constant(y) for y in Y], [tf.constant(x) for x in X_red])
with tf.GradientTape() as tape:
    elbo = model.ELBO_closure(data)
    grads = tape.gradient(elbo, tvars)
by_var = {id(v): g for v, g in zip(tvars, grads)}
out["elbo"] = np.float64(elbo.numpy())
out["n_draws"] = np.int64(len(recorded))
for j, z in enumerate(recorded):
    out[f"draw{j}"] = z
for k, p in params.items():
    g = by_var.get(id(p.unconstrained_variable))
    if g is None:
        continue
    g = g.as_subclass(torch.Tensor)
    t = type(p.transform).__name__
    if t == "FillTriangular":
        g = p.transform.forward(g).as_subclass(torch.Tensor)
    elif t != "Identity":
        theta = p.value().detach().as_subclass(torch.Tensor)
        if t == "Chain":
            theta = theta - p.transform.bijectors[0].shift
        g = g / (1.0 - torch.exp(-theta))
    out["grad_" + k] = g.numpy().reshape(p.numpy().shape)
np.savez_compressed(os.path.join(HERE, "mf_dgp_em.npz"), **out)
print("mf_dgp_em: ELBO", float(out["elbo"]), "draws", len(recorded), [z.shape for z in recorded], "grads", len([k for k in out if k.startswith("grad_")]))
