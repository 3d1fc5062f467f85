"""Wall-clock cost per iteration of DGP_Base.optimize_adam (the reference's training loop, models/dgp.py:132-154) on
launch-bound problems. Usage: python tools/bench_optimize_adam.py [repo_root]  (repo_root: which tree to import from)."""
import os, sys, json, time, io, contextlib
root = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import torch
import dgp_toolbox_b200 as D
from dgp_toolbox_b200 import synthetic

assert os.path.abspath(D.__file__).startswith(root), D.__file__
ITERS = int(os.environ.get("ITERS", "300"))
for name, (D0, units, M, S, N) in {"c1": (2, [2], 50, 10, 1000), "bo": (1, [1, 1], 25, 10, 50), "c2_small": (8, [8, 8, 8], 256, 32, 256)}.items():
    model = synthetic.model_from_problem(synthetic.synthetic_problem(D0, units, M, 8, ls_scale=0.3), S)
    X, Y = synthetic.minibatch(D0, N, 0)
    data = (torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    with contextlib.redirect_stdout(io.StringIO()):
        model.optimize_adam(data, iterations=20, lr=1e-3, messages=100)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.optimize_adam(data, iterations=ITERS, lr=1e-3, messages=100)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(json.dumps({"tree": os.path.basename(root), "case": name, "iterations": ITERS, "ms_per_iteration": round(1e3 * dt / ITERS, 4)}))
