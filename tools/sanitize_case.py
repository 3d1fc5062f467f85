"""Small end-to-end case for compute-sanitizer: fused + unfused conditional, ELBO+grad with chunking, predict, EI, EHVI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgp_toolbox_b200 as D
from dgp_toolbox_b200 import synthetic

ctx = D._lib.get_context(0)
for (D0, units, M, N, S) in [(3, [3, 2], 70, 150, 3), (8, [8], 128, 200, 4)]:
    model = synthetic.model_from_problem(synthetic.synthetic_problem(D0, units, M, 8), S)
    X, Y = synthetic.minibatch(D0, N, 0)
    for fused in (True, False):
        ctx.set_fused(fused)
        for share in (True, False):
            ctx.set_share_first_layer(share)
            flat = model.elbo_flat((X, Y), want_grad=True, seed=1)
            model.propagate(X, S=S, seed=2)
            model.predict(X, S, seed=3)
            D.EI(0.0, D0).run(model, X, analytic=False, num_samples=S, seed=4)
    ctx.set_fused(True); ctx.set_share_first_layer(True)
    ctx.set_vform(True, "always")                      # V-form adjoint (fused and, with DGP_B200_FUSED_BWD=0, unfused data path)
    model.elbo_flat((X, Y), want_grad=True, seed=1)
    torch.cuda.synchronize()
    ctx.set_vform(True, True)
    ctx.set_workspace_limit(64 << 20)
    model.elbo_flat((X, Y), want_grad=True, seed=1)
    ctx.set_workspace_limit(24 << 30)
    torch.cuda.synchronize()
    print("ok", D0, units, M, float(flat[0] - flat[1]))
