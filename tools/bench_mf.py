"""BASELINE config 4 in the form this repo covers: multi-fidelity DGP (MF_DGP.py: composite kernels, augmented inducing inputs), 3
fidelities, M = 256 inducing points per layer, ELBO + every gradient per step. (The embedded-mapping variant MF_DGP_EM.py is not
implemented.) Prints one JSON line.
   python tools/bench_mf.py [--n 32768 8192 2048] [--samples 10] [--steps 3]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgp_toolbox_b200 as D
from dgp_toolbox_b200.models import MF_DGP

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs=3, default=[32768, 8192, 2048])
ap.add_argument("--m", type=int, default=256)
ap.add_argument("--din", type=int, default=4)
ap.add_argument("--samples", type=int, default=10)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--em", action="store_true", help="embedded-mapping variant (MF_DGP_EM.py): input spaces of din, din + 1, din + 2 dimensions")
args = ap.parse_args()
rng = np.random.default_rng(0)
f = lambda x: np.sin(3 * x[:, :1]) + 0.5 * x[:, 1:2]
X = [rng.uniform(0, 1, (n, args.din)) for n in args.n]
Y = [f(X[0]), 1.2 * f(X[1]) + 0.3 * X[1][:, :1] ** 2, 1.5 * f(X[2]) - 0.2 * X[2][:, 1:2]]
Z = [rng.uniform(0, 1, (args.m, args.din)) for _ in range(3)]
D._lib.get_context(0).set_workspace_limit(64 << 30)
if args.em:
    from dgp_toolbox_b200.models import MF_DGP_EM
    dims = [args.din, args.din + 1, args.din + 2]
    X = [rng.uniform(0, 1, (n, d)) for n, d in zip(args.n, dims)]
    Y = [f(X[0]), 1.2 * f(X[1]) + 0.3 * X[1][:, :1] ** 2, 1.5 * f(X[2]) - 0.2 * X[2][:, 1:2]]
    Z = [rng.uniform(0, 1, (args.m, d)) for d in dims]
    W = [rng.uniform(0, 1, (args.m, dims[2])), rng.uniform(0, 1, (args.m, dims[1]))]
    model = MF_DGP_EM.DGP_Base.make_mf_dgp(X, Z, W)
    X_red = [torch.as_tensor(X[1][:, :dims[0]].copy()).cuda(), torch.as_tensor(X[2][:, :dims[0]].copy()).cuda()]
else:
    model = MF_DGP.DGP_Base.make_mf_dgp(Z)
model.num_samples = args.samples
Xd, Yd = [torch.as_tensor(x).cuda() for x in X], [torch.as_tensor(y).cuda() for y in Y]
data = (Xd, Yd, X_red) if args.em else (Xd, Yd)
params = model.trainable_parameters
for _ in range(2):
    model.ELBO_and_grads(data, params)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    elbo, grads = model.ELBO_and_grads(data, params)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
ps = sum(args.n) * args.samples
name = "MF-DGP-EM" if args.em else "MF-DGP"
print(json.dumps({"metric": name + " ELBO+grad point-samples/s", "value": ps / (ms * 1e-3), "unit": "point-samples/s", "ms_per_step": ms,
                  "elbo": float(elbo),
                  "config": {"workload": f"{name} ({'MF_DGP_EM.py, embedded mapping' if args.em else 'MF_DGP.py'}), 3 fidelities, D_in={args.din}, M={args.m}, S={args.samples}, "
                                         f"N={args.n} points per fidelity, float64; every fidelity's data term propagates through all 3 layers",
                             "point_samples_per_step": ps}}))
