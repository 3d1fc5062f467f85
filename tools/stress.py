"""Stability checks that are too long for the test suite: many consecutive steps (workspace reuse, no growth, determinism),
a minibatch large enough to be split into chunks, and a wide/large layer shape through the gradient path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgp_toolbox_b200 as D
from dgp_toolbox_b200 import synthetic

ctx = D._lib.get_context(0)
cfg = synthetic.CONFIGS["c2"]
model = synthetic.model_from_problem(synthetic.synthetic_problem(cfg["D0"], cfg["num_units"], cfg["M"], 8), cfg["S"])
X, Y = synthetic.minibatch(cfg["D0"], 4096, 0)
X, Y = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
ref = model.elbo_flat((X, Y), seed=1).clone()
torch.cuda.synchronize()
m0 = torch.cuda.memory_allocated(); w0 = ctx.workspace_bytes()
t0 = time.time()
for i in range(300):
    out = model.elbo_flat((X, Y), seed=1)
torch.cuda.synchronize()
assert torch.equal(out, ref), "same seed must give bitwise identical results across 300 steps"
print(f"300 steps: {1e3 * (time.time() - t0) / 300:.2f} ms/step, bitwise stable, torch memory growth {torch.cuda.memory_allocated() - m0} B, workspace growth {ctx.workspace_bytes() - w0} B")

# chunked large minibatch: 100k points x 32 samples under a 4 GiB workspace limit
ctx.set_workspace_limit(4 << 30)
Xl, Yl = synthetic.minibatch(cfg["D0"], 100000, 1)
a = model.elbo_flat((Xl, Yl), seed=2).clone()
ctx.set_workspace_limit(40 << 30)
b = model.elbo_flat((Xl, Yl), seed=2).clone()
print("chunked (4 GiB) vs larger chunks (40 GiB): max rel diff", float((a - b).abs().max() / b.abs().max()), "workspace", ctx.workspace_bytes() >> 20, "MiB")
assert float((a - b).abs().max()) <= 1e-10 * float(b.abs().max())

# config-3 shape through the gradient path (D=20, M=512, S=64), finite and deterministic
cfg3 = synthetic.CONFIGS["c3"]
m3 = synthetic.model_from_problem(synthetic.synthetic_problem(cfg3["D0"], cfg3["num_units"], cfg3["M"], 8), cfg3["S"])
X3, Y3 = synthetic.minibatch(cfg3["D0"], 2048, 2)
g1 = m3.elbo_flat((X3, Y3), seed=3).clone()
g2 = m3.elbo_flat((X3, Y3), seed=3).clone()
assert torch.isfinite(g1).all() and torch.equal(g1, g2)
t0 = time.time()
for i in range(3):
    m3.elbo_flat((X3, Y3), seed=3)
torch.cuda.synchronize()
print(f"config-3 ELBO+grad, 2048 points x 64 samples: {(time.time() - t0) / 3 * 1e3:.1f} ms/step, {2048 * 64 * 3 / (time.time() - t0):.0f} point-samples/s")
ctx.check()

# launch-bound sizes: the side-stream overlap (deferred prep products, parameter contractions in flight behind the next layer's
# data path) and graph replay must not change a single bit from step to step
for name, (D0, units, M, S, N) in {"c1": (2, [2], 50, 10, 1000), "bo": (1, [1, 1], 25, 10, 50), "c2_small": (8, [8, 8, 8], 256, 32, 256)}.items():
    ms = synthetic.model_from_problem(synthetic.synthetic_problem(D0, units, M, 8, ls_scale=0.3), S)
    Xs, Ys = synthetic.minibatch(D0, N, 0)
    Xs, Ys = torch.from_numpy(Xs).cuda(), torch.from_numpy(Ys).cuda()
    buf = torch.empty(ms.grad_layout()[0], dtype=torch.float64, device="cuda")
    ctx.set_parallel_layers(False)
    serial = ms.elbo_flat((Xs, Ys), seed=5).clone()
    ctx.set_parallel_layers(True)
    for graph in (False, True):
        ctx.set_graph(graph)
        bad = 0
        for i in range(400):
            out = ms.elbo_flat((Xs, Ys), seed=5, out=buf)
            if i % 50 == 49:
                bad += int(not torch.equal(out, serial))
        ctx.set_graph(False)
        assert bad == 0, (name, graph, bad)
    print(f"{name}: 2 x 400 steps (direct / graph replay) bitwise equal to the single-stream result")
print("stress ok")
