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


def _scalar(v):
    return float(v.item() if hasattr(v, "item") else v)


def _moment_criterion(kind, model, x, y, num_samples, zs, seed, with_x=False):
    """predict_y mixture moments over `num_samples` propagated samples (the reference hard-codes 500), then one
    dgp_acq_moments launch."""
    if getattr(model, "name", None) != 'dgp':
        raise NotImplementedError("only model.name == 'dgp' is on the accelerated path")
    X = model._check_X(_lib.as_device(x, model.device))
    m, v = model.predict_moments(X, num_samples, add_lik_var=True, zs=zs, seed=seed)
    if m.shape[1] != 1:
        raise ValueError("the criterion expects a single-output model")
    N, d = X.shape
    out = torch.empty((N, d if with_x else 1), dtype=torch.float64, device=X.device)
    if N:
        _lib.get_context(X.device).call("dgp_acq_moments", kind, _lib.ptr(m), _lib.ptr(v), N, _scalar(y),
                                        _lib.ptr(X) if with_x else None, d if with_x else 0, _lib.ptr(out))
    return out


class WB2(Infill_criteria):
    """Infill_criteria.py:106-141: -(EI - mean) on predict_y moments."""

    def __init__(self, y_min, d):
        self.name = 'WB2 criterion'
        self.y_min = y_min
        self.d = d
        self.IC_optimized = None

    def run(self, model, x, num_samples=500, zs=None, seed=None):
        return _moment_criterion(1, model, x, self.y_min, num_samples, zs, seed)

    def loss(self, model, x):
        return self.run(model, x)


class WB2S(Infill_criteria):
    """Infill_criteria.py:169-206: -(S * EI - mean) with S = 1 / (1 + 1/exp(x)) elementwise in x -> [N, d]."""

    def __init__(self, y_min, d):
        self.name = 'WB2S criterion'
        self.y_min = y_min
        self.d = d
        self.IC_optimized = None

    def run(self, model, x, num_samples=500, zs=None, seed=None):
        return _moment_criterion(3, model, x, self.y_min, num_samples, zs, seed, with_x=True)

    def loss(self, model, x):
        return self.run(model, x)


class EV_one_constraint(Infill_criteria):
    """Infill_criteria.py:235-262: expected violation of one constraint model, analytic (moment-matched) or Monte-Carlo."""

    def __init__(self, zero_c, d):
        self.name = 'Expected Violation'
        self.zero_c = zero_c
        self.d = d
        self.IC_optimized = None

    def run(self, model, x, analytic=True, num_samples=100, zs=None, seed=None):
        if analytic:
            return _moment_criterion(2, model, x, self.zero_c, 500 if zs is None else num_samples, zs, seed)
        if getattr(model, "name", None) != 'dgp':
            raise NotImplementedError("only model.name == 'dgp' is on the accelerated path")
        F = model.propagate(x, S=num_samples, zs=zs, seed=seed)[0][-1]            # [S, N, D_L]
        S, N, D = F.shape
        out = torch.empty((N, D), dtype=torch.float64, device=F.device)
        if N:
            _lib.get_context(F.device).call("dgp_ev_mc", _lib.ptr(F), S, N * D, _scalar(self.zero_c), _lib.ptr(out))
        return out


class EV(Infill_criteria):
    """Infill_criteria.py:264-288: expected violations of a list of constraint models side by side, and the constrained
    criterion `where(max_j EV_j > threshold, sum_j EV_j + 10000, IC)` (the reference loops tf.cond per candidate, :284-288;
    here it is one vectorised select on the device, same values)."""

    def __init__(self, zero_c, d):
        self.name = 'Expected Violation'
        self.zero_c = zero_c
        self.d = d
        self.IC_optimized = None

    def run(self, model_C, x, analytic=True, num_samples=100, zs=None, seed=None):
        cols = [EV_one_constraint(self.zero_c[i], self.d).run(model_C[i], x, analytic=analytic, num_samples=num_samples,
                                                              zs=None if zs is None else zs[i], seed=seed)
                for i in range(len(model_C))]
        return torch.cat(cols, 1)

    def run_with_IC(self, IC, model_Y, model_C, x, threshold=0.1, analytic=True, num_samples=100, seed=None):
        ev = self.run(model_C, x, analytic=analytic, num_samples=num_samples, seed=seed)
        ic = IC.run(model_Y, x, seed=seed)
        return torch.where(ev.max(dim=1, keepdim=True).values > threshold, ev.sum(dim=1, keepdim=True) + 10000.0, ic)


class PoF(Infill_criteria):
    """Infill_criteria.py:318-345 computes the EI-style terms but never returns them and `run_with_IC` references an undefined
    name: there is no reference behaviour to reproduce."""

    def __init__(self, zero_c, d):
        self.name = 'Probability of feasability'
        self.zero_c = zero_c
        self.d = d
        self.IC_optimized = None

    def run(self, model_C, x):
        raise NotImplementedError("PoF.run has no return statement in the reference (Infill_criteria.py:325-341)")
