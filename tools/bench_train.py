"""Wall-clock cost of one training iteration (ELBO gradient + Adam update) on launch-bound problems: BASELINE config 1
(D=2, M=50, S=10, N=1000), a BO-sized N=50 case and a small 4-layer config-2 shape. Three ways of driving the same kernels:
  stepwise       elbo_flat + dgp_adam_step from Python, one call each per iteration
  library        dgp_train_adam: the loop runs in the library
  library+graph  dgp_train_adam with dgp_set_graph(1): seed store + graph replay + Adam per iteration
Timed with CUDA events around K iterations after warm-up (host launch time is part of the figure: these steps are launch-bound)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_toolbox_b200 as D
from dgp_toolbox_b200 import synthetic

ctx = D._lib.get_context(0)
K = int(os.environ.get("K", "200"))
CASES = os.environ.get("CASES", "c1,bo,c2_small").split(",")
MODES = os.environ.get("MODES", "stepwise,library,library+graph").split(",")
for name, (D0, units, M, S, N) in {"c1": (2, [2], 50, 10, 1000), "bo": (1, [1, 1], 25, 10, 50), "c2_small": (8, [8, 8, 8], 256, 32, 256)}.items():
    if name not in CASES:
        continue
    row = {"case": name, "iterations": K}
    for mode in MODES:
        model = synthetic.model_from_problem(synthetic.synthetic_problem(D0, units, M, 8, ls_scale=0.3), S)
        X, Y = synthetic.minibatch(D0, N, 0)
        data = (torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
        params = model.trainable_parameters
        state = model._adam_state(params)
        ctx.set_graph(mode == "library+graph")

        def run(t0, n):
            if mode == "stepwise":
                for t in range(t0, t0 + n):
                    flat = model.elbo_flat(data)
                    model._adam_step(params, flat, state, t, 1e-3, 0.9, 0.999, 1e-7)
            else:
                model._train_adam(data, params, state, t0, n, 1e-3, 0.9, 0.999, 1e-7)

        run(1, 10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.launch_count(reset=True)
        e0.record()
        run(11, K)
        e1.record(); torch.cuda.synchronize()
        ctx.check()
        row[mode] = {"ms_per_iteration": round(e0.elapsed_time(e1) / K, 4), "kernels_per_iteration": ctx.launch_count() / K}
        ctx.set_graph(False)
    print(json.dumps(row))
