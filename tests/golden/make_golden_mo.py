"""Generates tests/golden/mo_dgp.npz by EXECUTING the reference's multi-objective model class (`DGP_Base` of
`/root/reference/dgp_dace/models/MO_DGP.py`, unmodified) and the `mo_dgp` branch of `/root/reference/dgp_dace/EHVI.py` on the
stand-in tensorflow / gpflow of tests/ref_shim.

What can be executed: `MO_DGP.DGP_Base.propagate / predict_f / E_log_p_Y / ELBO(tf_sample_Z_right=False)` (:88-216 -- the cyclic
objective-0 / objective-1 chain started from an N(0,1) column, the per-objective likelihood terms, the KL sum) and `EHVI.EHVI`
with an object named 'mo_dgp' (:124-130,154-157). What cannot: `make_mf_dgp` takes `Din` from `Z[0]`, whose layer needs `Din + 1`
columns (:257), and `sample_Z_right` feeds layer 0 a `Din`-column input (:29-31), so neither the constructor nor
`ELBO(tf_sample_Z_right=True)` runs in the reference as written. The layers are therefore built here with the reference's own
`SVGP_Layer` (utils/layers.py:180-224) on explicit `Din + 1`-column inducing inputs (both non-augmented), with the kernels
`make_mf_dgp` writes (:262-283).
    python tests/golden/make_golden_mo.py        (build container only)"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refexec as R  # noqa: E402

HERE = os.environ.get("DGP_GOLDEN_OUT") or os.path.dirname(os.path.abspath(__file__))      # tests/test_reference_exec.py regenerates into a temp dir
ns = R.load()
MO = importlib.import_module("dgp_dace.models.MO_DGP")
assert MO.__file__.startswith(R.REFERENCE)
tf, gpflow = ns.tf, ns.gpflow
from gpflow.kernels import RBF, Linear, White  # noqa: E402
from gpflow.likelihoods import Gaussian  # noqa: E402
from gpflow.mean_functions import Zero  # noqa: E402

rng = np.random.default_rng(7)
Din, M, S, LOOP = 2, 6, 3, 2
N = [11, 9]
f0 = lambda x: np.sin(3 * x[:, :1]) + 0.5 * x[:, 1:2]
f1 = lambda x: np.cos(2 * x[:, :1]) * x[:, 1:2] - 0.3
X = [rng.uniform(0, 1, (n, Din)) for n in N]
Y = [f0(X[0]), f1(X[1])]
Z = [np.concatenate([rng.uniform(0, 1, (M, Din)), rng.standard_normal((M, 1))], 1) for _ in range(2)]

recorded = []


def recording_source(shape):
    z = rng.standard_normal(shape)
    recorded.append(z)
    return z


tf.random.source = recording_source

# ---- kernels exactly as MO_DGP.py:262-283 writes them (Din = input dimension of the objectives) ----
kernels = []
for l in range(2):
    D_range = list(range(Din + 1))
    k_corr = RBF(active_dims=D_range[:Din], variance=1.0)
    k_prev = RBF(active_dims=D_range[Din:], variance=1.0)
    k_in = RBF(active_dims=D_range[:Din], variance=1.0)
    kernels.append(k_corr * (k_prev + Linear(active_dims=D_range[Din:], variance=1.0)) + k_in)
kernels[0] += White(variance=1e-6)
layers = [ns.layers.SVGP_Layer(kernels[i], Z[i].copy(), 1, Zero()) for i in range(2)]
model = MO.DGP_Base(Gaussian(), layers, loop=LOOP, num_samples=S)


def kparams(layer):
    ks = layer.kern.kernels
    prod, k_in = ks[0], ks[1]
    k_corr, inner = prod.kernels
    out = dict(corr_var=k_corr.variance, corr_ls=k_corr.lengthscales, prev_var=inner.kernels[0].variance,
               prev_ls=inner.kernels[0].lengthscales, lin_var=inner.kernels[1].variance, in_var=k_in.variance, in_ls=k_in.lengthscales)
    if len(ks) > 2:
        out["white_var"] = ks[2].variance
    return out


for i, layer in enumerate(model.layers):
    for name, p in kparams(layer).items():
        p.assign(2e-2 if name == "white_var" else rng.uniform(0.6, 1.4, p.shape))
    layer.q_mu.assign(0.3 * rng.standard_normal((M, 1)))
    layer.q_sqrt.assign(0.6 * layer.q_sqrt.numpy() + 0.05 * np.tril(rng.standard_normal((1, M, M))))
model.likelihood.likelihood.variance.assign(0.05)

params = {}
for i, layer in enumerate(model.layers):
    for name, p in kparams(layer).items():
        params[f"layers.{i}.{name}"] = p
    params[f"layers.{i}.q_mu"], params[f"layers.{i}.q_sqrt"], params[f"layers.{i}.Z"] = layer.q_mu, layer.q_sqrt, layer.feature.Z
params["lik_var"] = model.likelihood.likelihood.variance
out = {"provenance": np.array("reference source (/root/reference/dgp_dace/models/MO_DGP.py DGP_Base, utils/layers.py SVGP_Layer, EHVI.py, "
                              "unmodified) executed under tests/ref_shim; layers built on explicit (Din+1)-column inducing inputs"),
       "Din": np.int64(Din), "S": np.int64(S), "loop": np.int64(LOOP)}
for i in range(2):
    out[f"X{i}"], out[f"Y{i}"], out[f"Zinit{i}"] = X[i], Y[i], Z[i]
for k, p in params.items():
    out["param_" + k] = p.numpy()

# ---- ELBO (MO_DGP.py:187-216, tf_sample_Z_right=False) and its gradients ----
recorded.clear()
tvars = list(model.trainable_variables)
data = ([tf.constant(x) for x in X], [tf.constant(y) for y in Y])
with tf.GradientTape() as tape:
    elbo = model.ELBO(data, tf_sample_Z_right=False)
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

# ---- propagate with explicit zs (:88-122; the start column is still drawn: one recorded draw), both loop settings ----
Xt = rng.uniform(0, 1, (5, Din))
zs = [rng.standard_normal((S, 5, 1)) for _ in range(2)]
out["Xt"] = Xt
for i in range(2):
    out[f"zt{i}"] = zs[i]
for loop in (LOOP, 0, 1):
    model.loop = loop
    recorded.clear()
    Fs, Fm, Fv = model.propagate(tf.constant(Xt), S=S, zs=[tf.constant(z) for z in zs])
    assert len(recorded) == 1
    out[f"prop{loop}_start"] = recorded[0]
    for i in range(2):
        out[f"prop{loop}_F{i}"], out[f"prop{loop}_Fmean{i}"], out[f"prop{loop}_Fvar{i}"] = \
            [t.detach().as_subclass(torch.Tensor).numpy() for t in (Fs[i], Fm[i], Fv[i])]
model.loop = LOOP

# ---- EHVI.py's mo_dgp branch (:124-130) + exact uncorrelated strip sum (:154-157) ----
Yd = [rng.uniform(0, 1, (8, 1)), rng.uniform(0, 1, (8, 1))]
nd = ns.EHVI.NDC(Yd, np.full((8, 1), -1.0))
YND = ns.EHVI.Y_ND(Yd, nd, [1.5, 1.5], [-1.0, -1.0])
Xc = rng.uniform(0, 1, (7, Din))
recorded.clear()
obj = types.SimpleNamespace(name="mo_dgp", model=model)
Se = 4
ehvi = ns.EHVI.EHVI(obj, tf.constant(Xc), YND, corr=False, approximation='None', S=Se)
out["ehvi_X"], out["ehvi_S"] = Xc, np.int64(Se)
out["ehvi_ynd0"], out["ehvi_ynd1"] = np.asarray(YND[0]), np.asarray(YND[1])
out["ehvi"] = ehvi.detach().as_subclass(torch.Tensor).numpy()
out["ehvi_n_draws"] = np.int64(len(recorded))
for j, z in enumerate(recorded):
    out[f"ehvi_draw{j}"] = z
np.savez_compressed(os.path.join(HERE, "mo_dgp.npz"), **out)
print("mo_dgp: ELBO", float(out["elbo"]), "draws", len(recorded), [z.shape for z in (out[f"draw{j}"] for j in range(int(out["n_draws"])))],
      "grads", sorted(k for k in out if k.startswith("grad_")), "ehvi", out["ehvi"].ravel()[:3], "ehvi draws", int(out["ehvi_n_draws"]))
