"""Re-runs the training schedule of Notebooks_dgp/nb_DGP_regression.ipynb (cells 10-26) on the drop-in classes:
1-D step data (N=50, M=25), DGP with num_units=[1,1], S=10, optimize_nat_adam(iterations1=500, iterations2=5000, lr_adam=0.01,
beta_1=0.8, beta_2=0.9, lr_gamma=0.01, ng_all=False). The notebook's printed trajectory ends around ELBO = 104-109.
   python tools/notebook_regression.py [iterations1 iterations2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgp_toolbox_b200 as D

it1 = int(sys.argv[1]) if len(sys.argv) > 1 else 500
it2 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
np.random.seed(0)
X = np.random.uniform(0, 1, 50)[:, None]
Z = np.random.uniform(0, 1, 25)[:, None]
Y = (X >= 0.5).astype(np.float64) + 1e-2 * np.random.randn(50, 1)
kernels = [D.RBF(lengthscales=[1.0], variance=1.0) for _ in range(3)]
model = D.DGP(X, Y, Z, kernels, [1, 1], D.Gaussian(), num_samples=10, seed=0)
print("ELBO at construction:", float(model.ELBO((X, Y))))
t0 = time.time()
model.optimize_nat_adam(iterations1=it1, iterations2=it2, lr_adam=0.01, beta_1=0.8, beta_2=0.9, lr_gamma=0.01, ng_all=False, messages=500)
torch.cuda.synchronize()
dt = time.time() - t0
elbos = [float(model.ELBO((X, Y), seed=1000 + i)) for i in range(20)]
m, v = model.predict(np.linspace(0, 1, 11)[:, None], 100, seed=7)
print(f"trained in {dt:.1f} s ({(it1 + 2 * it2) / dt:.0f} ELBO+grad evaluations/s); final ELBO mean over 20 draws {np.mean(elbos):.2f} (std {np.std(elbos):.2f})")
print("predict mean on linspace(0,1,11):", np.round(m.cpu().numpy().ravel(), 3))
print("lik variance:", float(model.likelihood.likelihood.variance.value))
