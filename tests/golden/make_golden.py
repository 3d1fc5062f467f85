"""Generates tests/golden/*.npz: small seeded problems with the CPU oracle's outputs (forward chain, ELBO, gradients, EI, EHVI).
The reference itself (TensorFlow/GPflow) cannot be imported in this image, so these vectors are ORACLE outputs — they pin the
CUDA path (and the oracle) against silent drift, not against the reference; the ties to reference-produced numbers are the
notebook known answers in tests/test_oracle_kat.py.
    python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dgp_oracle as O  # noqa: E402
from tests.helpers import _condition  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"c1_like": (2, [2], 50, 40, 10), "c2_like": (8, [8, 8], 64, 24, 4), "ragged": (5, [3, 6], 20, 17, 3),
         "white_mixed": (3, [3, 2], 40, 45, 4)}
WHITE = {"white_mixed": [True, False, True]}      # whitened layers (utils/layers.py:246,254-255,296-303), mixed with non-white

for name, (D0, units, M, N, S) in CASES.items():
    prob = _condition(O.synthetic_problem(D0, units, M, N))
    for layer, w in zip(prob["layers"], WHITE.get(name, [])):
        layer["white"] = w
    om = O.model_from_problem(prob, S)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    zs = [torch.as_tensor(O.philox_normal(4321, l, S, N, layer.D_out)) for l, layer in enumerate(om.layers)]
    Fs, Fm, Fv = O.propagate(om.layers, X, S, zs)
    val, g = O.elbo_and_grads(om, X, Y, zs)
    y_min = float(prob["Y"].min())
    out = {"X": prob["X"], "Y": prob["Y"], "lik_var": np.float64(prob["lik_var"]), "S": np.int64(S), "seed": np.int64(4321),
           "elbo": np.float64(val), "ei_analytic": O.ei_analytic(Fm[-1], Fv[-1], y_min).numpy(), "ei_mc": O.ei_mc(Fs[-1], y_min).numpy(),
           "y_min": np.float64(y_min)}
    pm, pv = O.predict(om, X, S, zs)
    out["predict_mean"], out["predict_var"] = pm.numpy(), pv.numpy()
    for l, layer in enumerate(prob["layers"]):
        for k in ("Z", "lengthscales", "q_mu", "q_sqrt"):
            out[f"layer{l}_{k}"] = layer[k]
        out[f"layer{l}_variance"] = np.float64(layer["variance"])
        out[f"layer{l}_mean_kind"] = np.array(layer["mean_kind"])
        out[f"layer{l}_white"] = np.int64(1 if layer.get("white", False) else 0)
        if layer["mf_W"] is not None:
            out[f"layer{l}_mf_W"], out[f"layer{l}_mf_b"] = layer["mf_W"], layer["mf_b"]
        out[f"F{l}"], out[f"Fmean{l}"], out[f"Fvar{l}"] = Fs[l].numpy(), Fm[l].numpy(), Fv[l].numpy()
    for k, v in g.items():
        out["grad_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, float(val), sum(v.nbytes for v in out.values()) // 1024, "KB")

# ---- optimiser / search fixtures (aux_*.npz: not model cases) ----
# three tf.optimizers.Adam steps (AdamOracle) on the c1_like problem, fresh Philox draws per step (seed 500 + step)
prob = _condition(O.synthetic_problem(2, [2], 50, 40))
om = O.model_from_problem(prob, 10)
X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
opt = O.AdamOracle(om, lr=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
m = om
aux = {"adam_steps": np.int64(3), "adam_seed0": np.int64(500), "adam_lr": np.float64(0.01)}
for step in range(3):
    zs = [torch.as_tensor(O.philox_normal(500 + step, l, 10, 40, layer.D_out)) for l, layer in enumerate(m.layers)]
    val, m = opt.step(m, X, Y, zs)
    aux[f"adam_elbo{step}"] = np.float64(val)
for k, v in m.named_params().items():
    aux["adam_" + k] = v.numpy()
# random choices of differential-evolution generations (Philox key = seed, counter = (member, generation, slot, 0xDE))
for gen in (1, 7):
    a, b, c, forced, uni = O.de_choices(2 ** 63 + 5, gen, 10, 5)
    aux[f"de_g{gen}_a"], aux[f"de_g{gen}_b"], aux[f"de_g{gen}_c"], aux[f"de_g{gen}_forced"], aux[f"de_g{gen}_uni"] = a, b, c, forced, uni
np.savez_compressed(os.path.join(HERE, "aux_adam_de.npz"), **aux)
print("aux_adam_de", [float(aux[f"adam_elbo{i}"]) for i in range(3)])
