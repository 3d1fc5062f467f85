"""Multi-objective DGP (MO_DGP.py: two composite-kernel layers cycled 0 -> 1 -> 0 -> 1 ..., `loop` = 2) on BASELINE config 5's
acquisition workload: exact EHVI (EHVI.py:124-130,154-157, `mo_dgp` branch: both objectives from ONE chain) over chunks of candidate
points, and one ELBO + gradient step of the model (MO_DGP.py:187-216). Prints one JSON line.
   python tools/bench_mo.py [--n 16384] [--m 256] [--din 8] [--samples 32] [--steps 3]"""
import argparse, json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgp_toolbox_b200 as D
from dgp_toolbox_b200.models import MO_DGP

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384, help="candidates per chunk / training points per objective")
ap.add_argument("--m", type=int, default=256)
ap.add_argument("--din", type=int, default=8)
ap.add_argument("--samples", type=int, default=32)
ap.add_argument("--loop", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
rng = np.random.default_rng(0)
f0 = lambda x: np.sin(3 * x[:, :1]) + 0.5 * x[:, 1:2]
f1 = lambda x: np.cos(2 * x[:, :1]) * x[:, 1:2] - 0.3
Zx = rng.uniform(0, 1, (args.m, args.din))
Z = [np.concatenate([Zx, f1(Zx)], 1), Zx.copy()]
D._lib.get_context(0).set_workspace_limit(64 << 30)
model = MO_DGP.DGP_Base.make_mf_dgp(Z, loop=args.loop)
model.layers[0].kern.kernels[-1].variance.assign(1e-2)
for k, layer in enumerate(model.layers):
    layer.q_mu.assign((f0, f1)[k](Zx) + 0.05 * rng.standard_normal((args.m, 1)))
    layer.q_sqrt.assign(0.1 * layer.q_sqrt.value)
obj = types.SimpleNamespace(name="mo_dgp", model=model, _X=[Zx, Zx])
y0 = np.linspace(0.05, 0.95, 32)
YND = D.Y_ND([y0[:, None], (1 - np.sqrt(y0))[:, None]], list(range(32))[::-1], [1.1, 1.1], [-0.1, -0.1])
Xc = torch.as_tensor(rng.uniform(0, 1, (args.n, args.din))).cuda()


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps, out


ms_ehvi, ehvi = timed(lambda: D.EHVI(obj, Xc, YND, S=args.samples))
model.num_samples = args.samples
X = [Xc, Xc]
Y = [torch.as_tensor(f0(Xc.cpu().numpy())).cuda(), torch.as_tensor(f1(Xc.cpu().numpy())).cuda()]
params = model.trainable_parameters
ms_step, (elbo, _) = timed(lambda: model.ELBO_and_grads((X, Y), params))
apps = 2 + (1 if args.loop == 0 else 2 * args.loop)
print(json.dumps({"metric": "MO-DGP EHVI candidates/s", "value": args.n / (ms_ehvi * 1e-3), "unit": "candidates/s", "ms_per_chunk": ms_ehvi,
                  "ehvi_mean": float(ehvi.mean()), "elbo_grad_ms_per_step": ms_step,
                  "elbo_grad_point_samples_per_s": 2 * args.n * args.samples / (ms_step * 1e-3), "elbo": float(elbo),
                  "config": {"workload": f"MO-DGP (MO_DGP.py), 2 objectives, D_in={args.din}, M={args.m}, S={args.samples}, loop={args.loop} "
                                         f"({apps} layer applications per chain), chunks of {args.n} candidates, float64; exact EHVI on a 32-point front; "
                                         f"ELBO+grad: {args.n} points per objective, one chain per objective",
                             "layer_applications_per_chain": apps}}))
