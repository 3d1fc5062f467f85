"""Kernel-level breakdown of one multi-objective DGP ELBO + gradient step (tools/bench_mo.py's model: D_in = 8, M = 256, S = 32,
loop = 2, 16 384 points per objective) with torch.profiler (CUPTI sees the library's own launches). Evidence for DESIGN.md §6
"what the supplied-matrix path shares and where its time goes"; not a bench number (profiler overhead).
   python tools/prof_mo.py [--em]   (--em: the MF-DGP-EM step of bench.py's c4 entry instead)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgp_toolbox_b200 as D
from torch.profiler import ProfilerActivity, profile

rng = np.random.default_rng(0)
D._lib.get_context(0).set_workspace_limit(64 << 30)
if "--em" in sys.argv:
    from dgp_toolbox_b200.models import MF_DGP_EM
    dims, n4 = [2, 3, 4], [32768, 8192, 2048]
    f4 = lambda x: np.sin(3 * x[:, :1]) + 0.5 * x[:, 1:2]
    X = [rng.uniform(0, 1, (n, d)) for n, d in zip(n4, dims)]
    Y = [f4(X[0]), 1.2 * f4(X[1]) + 0.3 * X[1][:, :1] ** 2, 1.5 * f4(X[2]) - 0.2 * X[2][:, 1:2]]
    model = MF_DGP_EM.DGP_Base.make_mf_dgp(X, [rng.uniform(0, 1, (256, d)) for d in dims], [rng.uniform(0, 1, (256, 4)), rng.uniform(0, 1, (256, 3))])
    model.num_samples = 10
    d4 = [torch.from_numpy(x).cuda() for x in X + Y + [X[1][:, :2].copy(), X[2][:, :2].copy()]]
    data = (d4[0:3], d4[3:6], d4[6:8])
else:
    from dgp_toolbox_b200.models import MO_DGP
    n, m, din = 16384, 256, 8
    f0 = lambda x: np.sin(3 * x[:, :1]) + 0.5 * x[:, 1:2]
    f1 = lambda x: np.cos(2 * x[:, :1]) * x[:, 1:2] - 0.3
    Zx = rng.uniform(0, 1, (m, din))
    model = MO_DGP.DGP_Base.make_mf_dgp([np.concatenate([Zx, f1(Zx)], 1), Zx.copy()], loop=2)
    model.layers[0].kern.kernels[-1].variance.assign(1e-2)
    model.num_samples = 32
    Xc = torch.as_tensor(rng.uniform(0, 1, (n, din))).cuda()
    data = ([Xc, Xc], [torch.as_tensor(f0(Xc.cpu().numpy())).cuda(), torch.as_tensor(f1(Xc.cpu().numpy())).cuda()])
params = model.trainable_parameters
for _ in range(2):
    model.ELBO_and_grads(data, params)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    model.ELBO_and_grads(data, params)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=24, max_name_column_width=64))
