"""Wall-clock cost of the reference's acquisition search with its default budget (Infill_criteria.py:61: population 300,
400 generations, then 1000 Adam steps; EI.run's default 1000 samples per candidate) on a BO-sized model (SO_BO: M = N points):
EI.optimize('DE+Adam'). Reports the two stages separately and the candidates evaluated per second."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgp_toolbox_b200 as D

rng = np.random.default_rng(0)
for name, (d, N, L, S) in {"bo_d2_n50_l2": (2, 50, 2, 1000), "bo_d6_n100_l3": (6, 100, 3, 1000)}.items():
    X = rng.uniform(-1, 1, (N, d))
    Y = np.sin(3 * X[:, :1]) * np.cos(2 * X[:, 1:2] if d > 1 else 1.0) + 0.05 * rng.standard_normal((N, 1))
    kernels = [D.RBF(lengthscales=[0.7] * d, variance=1.0) for _ in range(L)]
    model = D.DGP(X, Y, X.copy(), kernels, [d] * (L - 1), D.Gaussian(0.01), num_samples=10, seed=1)
    D.DGP_Base.optimize_adam(model, model.data, iterations=200, lr=0.01, messages=10 ** 9)
    crit = D.EI(float(Y.min()), d)
    bounds = (np.full(d, -1.0), np.full(d, 1.0))
    crit.optimize(model, bounds, popsize_DE=300, iterations_DE=5, iterations_adam=10, method='DE+Adam', num_samples=S)   # warm-up
    row = {"case": name, "d": d, "N": N, "layers": L, "samples_per_candidate": S}
    for method, kw in (("DE", dict(popsize_DE=300, iterations_DE=400)), ("Adam", dict(iterations_adam=1000))):
        crit.x_opt = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        crit.optimize(model, bounds, method=method, num_samples=S, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        evals = 300 * (crit.de_iterations + 1) if method == "DE" else 1000
        row[method] = {"seconds": round(dt, 3), "criterion_evaluations": evals, "candidates_per_s": round(evals / dt, 1),
                       "point_samples_per_s": round(evals * S / dt, 1), "neg_ei_at_optimum": float(crit.IC_optimized.sum())}
    print(json.dumps(row))
