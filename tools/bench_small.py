"""Latency of one ELBO+grad step on small problems (BASELINE config 1: D=2, M=50, S=10, N=1000; and a BO-sized N=50 case)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_toolbox_b200 as D
from dgp_toolbox_b200 import synthetic

ctx = D._lib.get_context(0)
for name, (D0, units, M, S, N) in {"c1": (2, [2], 50, 10, 1000), "bo": (1, [1, 1], 25, 10, 50), "c2_small": (8, [8, 8, 8], 256, 32, 256)}.items():
    model = synthetic.model_from_problem(synthetic.synthetic_problem(D0, units, M, 8, ls_scale=0.3), S)
    X, Y = synthetic.minibatch(D0, N, 0)
    X, Y = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    for i in range(5):
        model.elbo_flat((X, Y), want_grad=True, seed=i)
    torch.cuda.synchronize()
    ctx.get_profile(reset=True); ctx.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 20
    for i in range(K):
        model.elbo_flat((X, Y), want_grad=True, seed=i)
    e1.record(); torch.cuda.synchronize()
    prof = ctx.get_profile(reset=True); ctx.set_profiling(False)
    print(json.dumps({"case": name, "ms_per_step": e0.elapsed_time(e1) / K, "launches_per_step": sum(v[1] for v in prof.values()) / K,
                      "categories_ms": {k: round(v[0] / K, 4) for k, v in prof.items() if v[0] > 0}}))
