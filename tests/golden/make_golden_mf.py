"""Generates tests/golden/mf_dgp.npz by EXECUTING the reference's multi-fidelity model (`/root/reference/dgp_dace/models/MF_DGP.py`,
unmodified) on the stand-in tensorflow / gpflow of tests/ref_shim: a 3-fidelity `DGP_Base.make_mf_dgp` model with non-trivial
parameters; stored are the inputs, every parameter, the N(0,1) draws in the order the reference consumed them, the ELBO
(MF_DGP.py:199-226, Z_right re-sampled through the earlier layers), its gradients w.r.t. every trainable parameter (constrained
space) and a `propagate` with explicit zs. The reference's patched GPflow `InducingPoints(layers=…)` is not in its repository; the
stand-in's assumed semantics are stated in tests/ref_shim/gpflow/inducing_variables.py.
    python tests/golden/make_golden_mf.py        (build container only)"""
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
MF = importlib.import_module("dgp_dace.models.MF_DGP")
assert MF.__file__.startswith(R.REFERENCE)
tf = ns.tf

rng = np.random.default_rng(42)
Din = 2
Ns, Ms = [14, 9, 6], [7, 6, 5]
f_lo = lambda x: np.sin(3 * x[:, :1]) + 0.5 * x[:, 1:2]
X = [rng.uniform(0, 1, (n, Din)) for n in Ns]
Y = [f_lo(X[0]), 1.2 * f_lo(X[1]) + 0.3 * X[1][:, :1] ** 2, 1.5 * f_lo(X[2]) - 0.2 * X[2][:, 1:2] + 0.1]
Z = [rng.uniform(0, 1, (m, Din)) for m in Ms]

recorded = []


def recording_source(shape):
    z = rng.standard_normal(shape)
    recorded.append(z)
    return z


tf.random.source = recording_source
with contextlib.redirect_stdout(io.StringIO()):
    model = MF.DGP_Base.make_mf_dgp([z.copy() for z in Z])
model.num_samples = 4

# ---- non-trivial parameters ----
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


for i, layer in enumerate(model.layers):
    for name, p in kparams(layer, i).items():
        p.assign(1e-2 * (1 + i) if name == "white_var" else rng.uniform(0.6, 1.4, p.shape))
    M = layer.num_inducing
    layer.q_mu.assign(0.3 * rng.standard_normal((M, 1)))
    layer.q_sqrt.assign(0.6 * layer.q_sqrt.numpy() + 0.05 * np.tril(rng.standard_normal((1, M, M))))
model.likelihood.likelihood.variance.assign(0.05)

params = {}
for i, layer in enumerate(model.layers):
    for name, p in kparams(layer, i).items():
        params[f"layers.{i}.{name}"] = p
    params[f"layers.{i}.q_mu"] = layer.q_mu
    params[f"layers.{i}.q_sqrt"] = layer.q_sqrt
    params[f"layers.{i}.Z"] = layer.feature.Z if i == 0 else layer.feature.Z_left
params["lik_var"] = model.likelihood.likelihood.variance
out = {"provenance": np.array("reference source (/root/reference/dgp_dace/models/MF_DGP.py, utils/layers.py, unmodified) executed under "
                              "tests/ref_shim; InducingPoints(layers=...) semantics assumed as documented there"),
       "Din": np.int64(Din), "S": np.int64(model.num_samples), "nfid": np.int64(3)}
for i in range(3):
    out[f"X{i}"], out[f"Y{i}"], out[f"Zinit{i}"] = X[i], Y[i], Z[i]
for k, p in params.items():
    out["param_" + k] = p.numpy()

# ---- ELBO and gradients through the reference's own code (MF_DGP.py:199-226, 381-386) ----
recorded.clear()
tvars = list(model.trainable_variables)
data = ([tf.constant(x) for x in X], [tf.constant(y) for y in Y])
with tf.GradientTape() as tape:
    elbo = model.ELBO_closure(data, tf_sample_Z_right=True)
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
for i in (1, 2):
    out[f"Zright{i}"] = model.layers[i].feature.Z_right.detach().as_subclass(torch.Tensor).numpy()

# ---- propagate / predict_f with explicit zs on test points (uses the Z_right of the ELBO call above) ----
Xt = rng.uniform(0, 1, (5, Din))
zs = [rng.standard_normal((3, 5, 1)) for _ in range(3)]
Fs, Fm, Fv = model.propagate(tf.constant(Xt), S=3, zs=[tf.constant(z) for z in zs])
out["Xt"] = Xt
for i in range(3):
    out[f"zt{i}"] = zs[i]
    out[f"F{i}"], out[f"Fmean{i}"], out[f"Fvar{i}"] = [t.detach().as_subclass(torch.Tensor).numpy() for t in (Fs[i], Fm[i], Fv[i])]
np.savez_compressed(os.path.join(HERE, "mf_dgp.npz"), **out)
print("mf_dgp: ELBO", float(out["elbo"]), "draws", len(recorded), [z.shape for z in recorded][:8], "grads", sorted(k for k in out if k.startswith("grad_")))
