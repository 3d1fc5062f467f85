"""Generates tests/golden/mf_nat_adam.npz by EXECUTING the reference's `MultiFidelityDeepGP.optimize_nat_adam`
(`/root/reference/dgp_dace/models/MF_DGP.py:426-519`, unmodified: Adam on the kernel parameters, then + inducing inputs, then Adam +
GPflow `NaturalGradient` on every layer's (q_mu, q_sqrt), two ELBO evaluations per iteration) on the stand-in tensorflow / gpflow of
tests/ref_shim, for 1 + 1 + 2 iterations of a 3-fidelity model. Stored: inputs, every parameter before and after, the printed ELBO
values, the N(0,1) draws in the order the reference consumed them.
    python tests/golden/make_golden_mf_nat.py        (build container only)"""
import contextlib
import importlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refexec as R  # noqa: E402

HERE = os.environ.get("DGP_GOLDEN_OUT") or os.path.dirname(os.path.abspath(__file__))      # tests/test_reference_exec.py regenerates into a temp dir
ns = R.load()
MF = importlib.import_module("dgp_dace.models.MF_DGP")
assert MF.__file__.startswith(R.REFERENCE)
tf = ns.tf

rng = np.random.default_rng(99)
Din = 2
Ns = [9, 7, 5]
f_lo = lambda x: np.sin(3 * x[:, :1]) + 0.5 * x[:, 1:2]
X = [rng.uniform(0, 1, (n, Din)) for n in Ns]
Y = [f_lo(X[0]), 1.2 * f_lo(X[1]) + 0.3 * X[1][:, :1] ** 2, 1.5 * f_lo(X[2]) - 0.2 * X[2][:, 1:2] + 0.1]

recorded = []


def recording_source(shape):
    z = rng.standard_normal(shape)
    recorded.append(z)
    return z


tf.random.source = recording_source
with contextlib.redirect_stdout(io.StringIO()):
    mf = MF.MultiFidelityDeepGP([x.copy() for x in X], [y.copy() for y in Y])      # Z = X (MF_DGP.py:521-537)
model = mf.model
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


for i, layer in enumerate(model.layers):
    for name, p in kparams(layer, i).items():
        p.assign(1e-2 * (1 + i) if name == "white_var" else rng.uniform(0.6, 1.4, p.shape))
    M = layer.num_inducing
    layer.q_sqrt.assign(np.tril(0.8 * np.eye(M)[None] + 0.05 * rng.standard_normal((1, M, M))))     # explicit: independent of the constructor's draws

params = {}
for i, layer in enumerate(model.layers):
    for name, p in kparams(layer, i).items():
        params[f"layers.{i}.{name}"] = p
    params[f"layers.{i}.q_mu"] = layer.q_mu
    params[f"layers.{i}.q_sqrt"] = layer.q_sqrt
    params[f"layers.{i}.Z"] = layer.feature.Z if i == 0 else layer.feature.Z_left
params["lik_var"] = model.likelihood.likelihood.variance
out = {"provenance": np.array("reference source (/root/reference/dgp_dace/models/MF_DGP.py MultiFidelityDeepGP.optimize_nat_adam, unmodified) "
                              "executed under tests/ref_shim (gpflow.optimizers.NaturalGradient stand-in: XiNat)"),
       "Din": np.int64(Din), "S": np.int64(model.num_samples), "nfid": np.int64(3), "iters": np.array([1, 1, 2]),
       "lr_adam": np.float64(0.01), "lr_gamma": np.float64(0.05)}
for i in range(3):
    out[f"X{i}"], out[f"Y{i}"] = X[i], Y[i]
for k, p in params.items():
    out["param0_" + k] = p.numpy().copy()

recorded.clear()
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    mf.optimize_nat_adam(lr_adam=0.01, lr_gamma=0.05, iterations1=1, iterations2=1, iterations3=2, messages=1)
out["elbo_trace"] = np.array([float(l.split("ELBO:")[1]) for l in buf.getvalue().splitlines() if l.startswith("ELBO:")])
out["n_draws"] = np.int64(len(recorded))
for j, z in enumerate(recorded):
    out[f"draw{j}"] = z
for k, p in params.items():
    out["param1_" + k] = p.numpy().copy()
np.savez_compressed(os.path.join(HERE, "mf_nat_adam.npz"), **out)
print("mf_nat_adam: trace", out["elbo_trace"], "draws", len(recorded),
      "max |dq_mu|", [float(np.abs(out[f"param1_layers.{i}.q_mu"] - out[f"param0_layers.{i}.q_mu"]).max()) for i in range(3)])
