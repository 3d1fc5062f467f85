"""Expected-improvement evaluation with the reference's interface (dgp_dace/Infill_criteria.py:12-52). Only the
*evaluation* (EI.run on a DGP) is on the accelerated path; the DE/Adam search loops of the reference
(Infill_criteria.py:61-87) are host-side callers and out of scope."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class Infill_criteria(object):
    def __init__(self):
        self.name = 'Infill criteria'

    def run(self, x):
        raise NotImplementedError("method not implemented")

    def optimize(self):
        raise NotImplementedError("method not implemented")


class EI(Infill_criteria):
    """Infill_criteria.py:20-52."""

    def __init__(self, y_min, d):
        self.name = 'Expected Improvement'
        self.y_min = y_min
        self.d = d
        self.IC_optimized = None
        self.x_opt = None

    def run(self, model, x, analytic=True, num_samples=1000, zs=None, seed=None):
        """Returns -EI [N, D_L]. analytic: moment-match predict_f over the S samples (:39-41) then the closed form
        (:43-47); otherwise mean_s max(y_min - F, 0) on propagated samples (:49-51). One C-ABI call (dgp_ei)."""
        if getattr(model, "name", None) != 'dgp':
            raise NotImplementedError("only model.name == 'dgp' is on the accelerated path")
        X = model._check_X(_lib.as_device(x, model.device))
        N, D = X.shape[0], model.layers[-1].num_outputs
        out = torch.empty((N, D), dtype=torch.float64, device=X.device)
        if N == 0:
            return out
        m, keep = model._model_desc()
        zt, zp = model._zs(zs, num_samples, N)
        y_min = float(self.y_min.item() if hasattr(self.y_min, "item") else self.y_min)
        _lib.get_context(X.device).call("dgp_ei", C.byref(m), _lib.ptr(X), N, num_samples, zp, model._next_seed(seed), 0,
                                        y_min, 1 if analytic else 0, _lib.ptr(out))
        return out

    def run_with_grad(self, model, x, num_samples=1000, zs=None, seed=None):
        """(-EI [N, D_L], d sum(-EI) / dx [N, d]): the value and the gradient the reference's Adam-on-x loop takes with
        tape.gradient(loss, x) (Infill_criteria.py:79-84), analytic EI. One C-ABI call (dgp_ei_grad)."""
        if getattr(model, "name", None) != 'dgp':
            raise NotImplementedError("only model.name == 'dgp' is on the accelerated path")
        X = model._check_X(_lib.as_device(x, model.device))
        N, D = X.shape[0], model.layers[-1].num_outputs
        out = torch.empty((N, D), dtype=torch.float64, device=X.device)
        dx = torch.zeros_like(X)
        if N == 0:
            return out, dx
        m, keep = model._model_desc()
        zt, zp = model._zs(zs, num_samples, N)
        y_min = float(self.y_min.item() if hasattr(self.y_min, "item") else self.y_min)
        _lib.get_context(X.device).call("dgp_ei_grad", C.byref(m), _lib.ptr(X), N, num_samples, zp, model._next_seed(seed), 0,
                                        y_min, _lib.ptr(out), _lib.ptr(dx))
        return out, dx

    def loss(self, model, x, analytic):
        """Infill_criteria.py:53-60."""
        return self.run(model, x, analytic)
