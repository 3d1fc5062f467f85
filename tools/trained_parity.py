"""How far do CUDA and oracle drift on a TRAINED, ill-conditioned model (the notebook's 1-D step problem, cond(Kuu) limited by the
1e-6 jitter)? Prints relative errors of the ELBO and of every gradient after a short optimize_nat_adam run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgp_toolbox_b200 as D
from oracle import dgp_oracle as O

np.random.seed(0)
X = np.random.uniform(0, 1, 50)[:, None]
Z = np.random.uniform(0, 1, 25)[:, None]
Y = (X >= 0.5).astype(np.float64) + 1e-2 * np.random.randn(50, 1)
model = D.DGP(X, Y, Z, [D.RBF(lengthscales=[1.0], variance=1.0) for _ in range(3)], [1, 1], D.Gaussian(), num_samples=10, seed=0)
model.optimize_nat_adam(iterations1=200, iterations2=600, lr_adam=0.01, beta_1=0.8, beta_2=0.9, lr_gamma=0.01, ng_all=False, messages=10 ** 9)
layers = []
for l in model.layers:
    kind = {0: "zero", 1: "identity", 2: "linear"}[l.mean_function.mean_kind]
    layers.append(O.make_layer(l.feature.Z.numpy(), l.kern.lengthscales_vector(l.feature.Z.shape[1]).cpu().numpy(), float(l.kern.variance.value),
                               l.num_outputs, kind, q_mu=l.q_mu.numpy(), q_sqrt=l.q_sqrt.numpy()))
    print("cond(Kuu + jitter I) =", f"{float(torch.linalg.cond(O.kuu_chol(layers[-1])[0])):.3e}", "lengthscale", layers[-1].lengthscales.numpy())
om = O.OModel(layers=layers, lik_var=torch.tensor(float(model.likelihood.likelihood.variance.value), dtype=torch.float64), num_samples=10)
zs = [torch.randn(10, 50, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(5)) for _ in layers]
val_o, g_o = O.elbo_and_grads(om, torch.as_tensor(X), torch.as_tensor(Y), zs)
ctx = D._lib.get_context(0)
for name, vf in (("A-form adjoint", False), ("V-form adjoint", "always")):
    ctx.set_vform(True, vf)
    val, g = model.ELBO_and_grads((X, Y), zs=zs)
    errs = {k: float((g[k].cpu().reshape(v.shape) - v).abs().max() / max(float(v.abs().max()), 1e-300)) for k, v in g_o.items()}
    print(name, "ELBO", float(val), "oracle", float(val_o), "rel", abs(float(val) - float(val_o)) / abs(float(val_o)), "worst gradient rel err", max(errs.values()), max(errs, key=errs.get))
