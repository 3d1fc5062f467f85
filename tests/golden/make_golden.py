"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE: `/root/reference/dgp_dace` is imported unmodified
on the stand-in tensorflow / gpflow / tfp of tests/ref_shim (tests/refexec.py), and every stored output — forward chain,
ELBO, gradients, predict, EI, EHVI, three `optimize_adam` iterations — is what the reference's code returned for the stored
inputs and draws (`provenance` key). The CPU oracle is run beside it and must agree to 1e-11 or the script stops.
The differential-evolution choices in aux_adam_de.npz are the exception: TFP's random stream cannot be reproduced, those are
the oracle's Philox choices (integer plumbing fixtures).
    python tests/golden/make_golden.py        (build container only: needs /root/reference)"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dgp_oracle as O  # noqa: E402
from tests import refexec as R  # noqa: E402
from tests.helpers import _condition  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"c1_like": (2, [2], 50, 40, 10), "c2_like": (8, [8, 8], 64, 24, 4), "ragged": (5, [3, 6], 20, 17, 3),
         "white_mixed": (3, [3, 2], 40, 45, 4), "matern": (3, [3], 16, 21, 5)}
WHITE = {"white_mixed": [True, False, True]}      # whitened layers (utils/layers.py:246,254-255,296-303), mixed with non-white
KERNELS = {"matern": ["matern32", "matern52"]}    # SO_BO.py:194-197 builds Matern layers
PROVENANCE = "reference source (/root/reference/dgp_dace, unmodified) executed under tests/ref_shim; oracle agrees to 1e-11"

ns = R.load()


def plain(t):
    return t.detach().as_subclass(torch.Tensor).numpy().copy() if isinstance(t, torch.Tensor) else np.asarray(t)


def check(a, b, what, tol=1e-11, scale=0.0):
    a, b = plain(a), plain(b)
    err = float(np.max(np.abs(a - b.reshape(a.shape)))) / max(float(np.max(np.abs(b))), scale, 1e-300)
    assert err < tol, (what, err)


for name, (D0, units, M, N, S) in CASES.items():
    prob = O.synthetic_problem(D0, units, M, N)
    for layer, w in zip(prob["layers"], WHITE.get(name, [])):
        layer["white"] = w
    for layer, k in zip(prob["layers"], KERNELS.get(name, [])):
        layer["kernel"] = k
    prob = _condition(prob)
    om = O.model_from_problem(prob, S)
    rm = R.reference_model(prob, S)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    Xt = ns.tf.constant(prob["X"])
    zs = [torch.as_tensor(O.philox_normal(4321, l, S, N, layer.D_out)) for l, layer in enumerate(om.layers)]
    y_min = float(prob["Y"].min())
    # --- the reference's code ---
    rFs, rFm, rFv = rm.propagate(Xt, S=S, zs=[ns.tf.constant(z.numpy()) for z in zs])                  # models/dgp.py:34-63
    rval, rgrads = R.elbo_and_grads(rm, prob["X"], prob["Y"], zs)                                        # models/dgp.py:89-109,142-145
    with R.fixed_draws(zs):
        rpm, rpv = ns.dgp.DGP.predict(rm, Xt, num_samples=S)                                             # models/dgp.py:362-366
    with R.fixed_draws(zs):
        rei_a = ns.Infill_criteria.EI(y_min, D0).run(rm, Xt, analytic=True, num_samples=S)               # Infill_criteria.py:28-52
    with R.fixed_draws(zs):
        rei_mc = ns.Infill_criteria.EI(y_min, D0).run(rm, Xt, analytic=False, num_samples=S)
    # --- the oracle beside it ---
    Fs, Fm, Fv = O.propagate(om.layers, X, S, zs)
    val, g = O.elbo_and_grads(om, X, Y, zs)
    for l in range(len(om.layers)):
        check(Fs[l], rFs[l], f"{name} F{l}", scale=1.0)
        check(Fm[l], rFm[l], f"{name} Fmean{l}", scale=1.0)
        check(Fv[l], rFv[l], f"{name} Fvar{l}", scale=1.0)
    check(val, rval, name + " elbo")
    for k in g:
        check(g[k], rgrads[k], f"{name} grad {k}", tol=1e-10)
    pm, pv = O.predict(om, X, S, zs)
    check(pm, rpm, name + " predict mean")
    check(pv, rpv, name + " predict var")
    check(O.ei_analytic(Fm[-1], Fv[-1], y_min), rei_a, name + " ei", tol=1e-10)
    check(O.ei_mc(Fs[-1], y_min), rei_mc, name + " ei mc", tol=1e-10)
    out = {"provenance": np.array(PROVENANCE), "X": prob["X"], "Y": prob["Y"], "lik_var": np.float64(prob["lik_var"]), "S": np.int64(S),
           "seed": np.int64(4321), "elbo": np.float64(rval), "ei_analytic": plain(rei_a), "ei_mc": plain(rei_mc),
           "y_min": np.float64(y_min), "predict_mean": plain(rpm), "predict_var": plain(rpv)}
    for l, layer in enumerate(prob["layers"]):
        for k in ("Z", "lengthscales", "q_mu", "q_sqrt"):
            out[f"layer{l}_{k}"] = layer[k]
        out[f"layer{l}_variance"] = np.float64(layer["variance"])
        out[f"layer{l}_mean_kind"] = np.array(layer["mean_kind"])
        out[f"layer{l}_kernel"] = np.array(layer.get("kernel", "rbf"))
        out[f"layer{l}_white"] = np.int64(1 if layer.get("white", False) else 0)
        if layer["mf_W"] is not None:
            out[f"layer{l}_mf_W"], out[f"layer{l}_mf_b"] = layer["mf_W"], layer["mf_b"]
        out[f"F{l}"], out[f"Fmean{l}"], out[f"Fvar{l}"] = plain(rFs[l]), plain(rFm[l]), plain(rFv[l])
    for k, v in rgrads.items():
        out["grad_" + k] = plain(v).reshape(tuple(g[k].shape))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, rval, sum(v.nbytes for v in out.values()) // 1024, "KB")

# ---- EHVI over two DGPs (EHVI.py:107-119,150-157) with the padded front of EHVI.Y_ND (:90-100) ----
pa, pb = _condition(O.synthetic_problem(3, [3], 16, 14)), _condition(O.synthetic_problem(3, [3], 16, 14, seed_shift=5))
S, N = 5, 14
oa, ob = O.model_from_problem(pa, S), O.model_from_problem(pb, S)
ra, rb = R.reference_model(pa, S), R.reference_model(pb, S)
za = [torch.as_tensor(O.philox_normal(21, l, S, N, layer.D_out)) for l, layer in enumerate(oa.layers)]
zb = [torch.as_tensor(O.philox_normal(22, l, S, N, layer.D_out)) for l, layer in enumerate(ob.layers)]
y0 = np.linspace(0.05, 0.95, 6)
y1 = 1.0 - np.sqrt(y0)
order = list(np.argsort(-y0))
ynd = ns.EHVI.Y_ND([y0[:, None], y1[:, None]], order, nadir=[1.1, 1.1], ideal=[-0.1, -0.1])
with R.fixed_draws(za + zb):
    r_ehvi = ns.EHVI.EHVI([ra, rb], ns.tf.constant(pa["X"]), ynd, corr=False, approximation='None', S=S)
_, Fma, Fva = O.propagate(oa.layers, torch.as_tensor(pa["X"]), S, za)
_, Fmb, Fvb = O.propagate(ob.layers, torch.as_tensor(pa["X"]), S, zb)
m0, v0 = O.mixture_moments(Fma[-1], Fva[-1])
m1, v1 = O.mixture_moments(Fmb[-1], Fvb[-1])
check(O.ehvi_exact(m0, v0, m1, v1, ynd[0][:, 0], ynd[1][:, 0]), r_ehvi, "ehvi", tol=1e-10)

# ---- optimiser / search fixtures (aux_*.npz: not model cases) ----
# three iterations of the reference's DGP_Base.optimize_adam (models/dgp.py:132-154: Keras Adam on GPflow's unconstrained
# variables) on the c1_like problem, the draws of Philox seed 500 + step served to the reference's tf.random.normal
prob = _condition(O.synthetic_problem(2, [2], 50, 40))
om = O.model_from_problem(prob, 10)
rm = R.reference_model(prob, 10)
X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
draws = [[torch.as_tensor(O.philox_normal(500 + step, l, 10, 40, layer.D_out)) for l, layer in enumerate(om.layers)] for step in range(3)]
buf = io.StringIO()
with R.fixed_draws([z for d in draws for z in d]), contextlib.redirect_stdout(buf):
    rm.optimize_adam((ns.tf.constant(prob["X"]), ns.tf.constant(prob["Y"])), iterations=3, lr=0.01, messages=1)
printed = [float(line.split()[1]) for line in buf.getvalue().splitlines() if line.startswith("ELBO:")]
opt = O.AdamOracle(om, lr=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
m = om
aux = {"provenance": np.array(PROVENANCE + " (adam_*; de_* are the oracle's Philox choices)"), "adam_steps": np.int64(3),
       "adam_seed0": np.int64(500), "adam_lr": np.float64(0.01), "ehvi": plain(r_ehvi), "ehvi_ynd0": ynd[0][:, 0], "ehvi_ynd1": ynd[1][:, 0],
       "ehvi_S": np.int64(S), "ehvi_seed0": np.int64(21), "ehvi_seed1": np.int64(22)}
for step in range(3):
    val, m = opt.step(m, X, Y, draws[step])
    check(val, printed[step], f"adam elbo {step}")
    aux[f"adam_elbo{step}"] = np.float64(printed[step])
ref_params = {}
for i, rl in enumerate(rm.layers):
    for key, p in (("Z", rl.feature.Z), ("lengthscales", rl.kern.lengthscales), ("variance", rl.kern.variance), ("q_mu", rl.q_mu),
                   ("q_sqrt", rl.q_sqrt)):
        ref_params[f"layers.{i}.{key}"] = p.numpy()
ref_params["lik_var"] = rm.likelihood.likelihood.variance.numpy()
for k, v in m.named_params().items():
    check(v, ref_params[k], "adam " + k, tol=1e-10)
    aux["adam_" + k] = np.asarray(ref_params[k]).reshape(tuple(v.shape))
# random choices of differential-evolution generations (Philox key = seed, counter = (member, generation, slot, 0xDE))
for gen in (1, 7):
    a, b, c, forced, uni = O.de_choices(2 ** 63 + 5, gen, 10, 5)
    aux[f"de_g{gen}_a"], aux[f"de_g{gen}_b"], aux[f"de_g{gen}_c"], aux[f"de_g{gen}_forced"], aux[f"de_g{gen}_uni"] = a, b, c, forced, uni
np.savez_compressed(os.path.join(HERE, "aux_adam_de.npz"), **aux)
print("aux_adam_de", [float(aux[f"adam_elbo{i}"]) for i in range(3)])
