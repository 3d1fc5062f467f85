"""Phase breakdown of the fused kernels (clock64 instrumentation, -DDGP_DEBUG_PHASECLK build selected with DGP_B200_LIB):
    DGP_B200_LIB=build/libdgp_b200_phase.so python tools/phase_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dgp_toolbox_b200 as D  # noqa: E402
from dgp_toolbox_b200 import synthetic  # noqa: E402

cfg = synthetic.CONFIGS["c2"]
prob = synthetic.synthetic_problem(cfg["D0"], cfg["num_units"], cfg["M"], 8)
model = synthetic.model_from_problem(prob, cfg["S"])
ctx = D._lib.get_context(0)
ctx.set_workspace_limit(64 << 30)
X, Y = synthetic.minibatch(cfg["D0"], 16384, 0)
X, Y = torch.as_tensor(X).cuda(), torch.as_tensor(Y).cuda()
for i in range(2):
    flat = model.elbo_flat((X, Y), want_grad=True, seed=1 + i)
    torch.cuda.synchronize()
    print("---- step", i, float(flat[0] - flat[1]), flush=True)
