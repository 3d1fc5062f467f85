"""Forward-only workloads of BASELINE.json (not the headline metric; bench.py is):
  config 3: 5-layer DGP (6 SVGP layers), D=20, M=512, S=64 — predict (predict_y + mixture moments) point-samples/s
  config 5: EI + exact 2-objective EHVI over candidate points with two config-2 shaped DGPs, S=32 — candidates/s
   python tools/bench_forward.py [--nb 8192] [--steps 5]
Prints one JSON line per workload."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgp_toolbox_b200 as D
from dgp_toolbox_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--nb", type=int, default=8192)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
ctx = D._lib.get_context(0)
ctx.set_workspace_limit(64 << 30)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_fp64_peaks.json")))["summary"]["fp64_dmma_tflops"]


def timed(fn, steps):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(3 + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# ---- config 3: predict ----
cfg = synthetic.CONFIGS["c3"]
model = synthetic.model_from_problem(synthetic.synthetic_problem(cfg["D0"], cfg["num_units"], cfg["M"], 8), cfg["S"])
pool = [torch.from_numpy(synthetic.minibatch(cfg["D0"], args.nb, i)[0]).cuda() for i in range(4)]
ms = timed(lambda i: model.predict(pool[i % 4], cfg["S"], seed=100 + i), args.steps)
f_fwd, _ = synthetic.flops_per_point_sample(cfg["D0"], cfg["num_units"], cfg["M"], S=cfg["S"])
ps = args.nb * cfg["S"]
print(json.dumps({"metric": "DGP predict point-samples/s", "value": ps / (ms * 1e-3), "unit": "point-samples/s", "ms_per_step": ms,
                  "config": {"workload": "5-layer DGP (6 SVGP layers), D=20, M=512, S=64, predict (mixture moments of predict_y), float64",
                             "points_per_step": args.nb, "samples": cfg["S"]},
                  "roofline": {"bound": "tensor", "achieved": f_fwd * ps / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                               "frac": f_fwd * ps / (ms * 1e-3) / 1e12 / peak, "algorithmic_flops_per_point_sample": f_fwd}}))

# ---- config 5: EI + EHVI ----
cfg = synthetic.CONFIGS["c2"]
m0 = synthetic.model_from_problem(synthetic.synthetic_problem(cfg["D0"], cfg["num_units"], cfg["M"], 8), cfg["S"])
m1 = synthetic.model_from_problem(synthetic.synthetic_problem(cfg["D0"], cfg["num_units"], cfg["M"], 8, seed_shift=100), cfg["S"])
nb = 2 * args.nb
pool = [torch.from_numpy(synthetic.minibatch(cfg["D0"], nb, 10 + i)[0]).cuda() for i in range(4)]
y0 = np.linspace(0.95, 0.05, 32)
y1 = 1.0 - np.sqrt(y0)
ynd = D.Y_ND([y0, y1], np.arange(32), nadir=(1.1, 1.1), ideal=(-0.1, -0.1))
ei = D.EI(0.0, cfg["D0"])


def acq(i):
    ei.run(m0, pool[i % 4], analytic=True, num_samples=cfg["S"], seed=i)
    D.EHVI([m0, m1], pool[i % 4], ynd, S=cfg["S"], seed=[i, 1000 + i])


ms = timed(acq, args.steps)
f_fwd, _ = synthetic.flops_per_point_sample(cfg["D0"], cfg["num_units"], cfg["M"], S=cfg["S"])
flops = 3 * f_fwd * nb * cfg["S"]   # three propagations per candidate batch: EI on model 0, EHVI on models 0 and 1
print(json.dumps({"metric": "DGP EI+EHVI candidates/s", "value": nb / (ms * 1e-3), "unit": "candidates/s", "ms_per_step": ms,
                  "config": {"workload": "EI (analytic) + exact 2-objective EHVI, two 3-layer DGPs (D=8, M=256), S=32, 32-point Pareto front, float64",
                             "candidates_per_step": nb, "samples": cfg["S"]},
                  "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                               "frac": flops / (ms * 1e-3) / 1e12 / peak}}))
