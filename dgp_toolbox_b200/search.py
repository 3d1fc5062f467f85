"""Acquisition search with the state on the device: the two stages of the reference's `optimize` methods
(dgp_dace/Infill_criteria.py:61-87,142-168,207-233,290-316) -- tfp.optimizer.differential_evolution_minimize, then
tf.optimizers.Adam -- on u with x = lw + (up - lw) / (1 + exp(u)). The criterion is evaluated by the library's model-level
entry points (one call per generation for the whole candidate population / one value+gradient call per Adam step); the
population update and the Adam step are one launch each (dgp_de_propose / dgp_de_select / dgp_adam_box_step), and with graph
replay on (dgp_set_graph) the criterion call is a seed store plus one graph launch because its buffers never move."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

DE_INIT_LAYER = 0xDE0   # Philox "layer" index of the initial population's normal draws


def _vec(v, d, device):
    t = _lib.as_device(np.broadcast_to(np.asarray(v, dtype=np.float64).reshape(-1), (d,)).copy(), device)
    return t


def initial_population(d, population_size, population_stddev, seed, device, initial_position=None):
    """TFP: the initial position itself plus population_size - 1 normal perturbations of it (u-space)."""
    ctx = _lib.get_context(device)
    init = torch.zeros(d, dtype=torch.float64, device=device) if initial_position is None else _vec(initial_position, d, device)
    z = torch.empty((1, population_size, d), dtype=torch.float64, device=device)
    ctx.call("dgp_philox_normal", int(seed), DE_INIT_LAYER, 1, population_size, d, 0, _lib.ptr(z))
    pop = init[None, :] + population_stddev * z[0]
    pop[0] = init
    return pop.contiguous()


def de_minimize(objective, lw, up, d, device, population_size=300, population_stddev=1.5, max_iterations=400, seed=0,
                differential_weight=0.5, crossover_prob=0.9, position_tolerance=1e-8, initial_population_u=None,
                check_every=50):
    """objective(X [pop, d] device tensor, out [pop, ncol] device tensor) fills `out` with the criterion (minimised; summed over
    the ncol columns); both buffers are the same objects in every generation. Returns a dict with the final population
    (u-space), its values, the best member (u and x) and the generations run. Stops early when the population has collapsed
    (max |pop - pop[0]| <= position_tolerance, checked every `check_every` generations -- the only host synchronisations)."""
    ctx = _lib.get_context(device)
    dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
    lw_t, up_t = _vec(lw, d, dev), _vec(up, d, dev)
    pop_u = initial_population(d, population_size, population_stddev, seed, dev) if initial_population_u is None \
        else _lib.as_device(initial_population_u, dev).clone()
    pop = pop_u.shape[0]
    cand_u = torch.empty_like(pop_u)
    cand_x = torch.empty_like(pop_u)
    ctx.call("dgp_box_from_u", _lib.ptr(pop_u), _lib.ptr(lw_t), _lib.ptr(up_t), pop, d, _lib.ptr(cand_x))
    vals = None
    pop_val = torch.empty(pop, dtype=torch.float64, device=dev)
    gen = 0
    for gen in range(0, max_iterations + 1):
        if gen > 0:
            ctx.call("dgp_de_propose", _lib.ptr(pop_u), pop, d, _lib.ptr(lw_t), _lib.ptr(up_t), int(seed), gen,
                     float(differential_weight), float(crossover_prob), _lib.ptr(cand_u), _lib.ptr(cand_x))
        if vals is None:
            vals = objective(cand_x, None)           # first call allocates the value buffer
            if not vals.is_contiguous():
                vals = vals.contiguous()
        else:
            objective(cand_x, vals)
        ncol = vals.shape[1] if vals.dim() > 1 else 1
        ctx.call("dgp_de_select", _lib.ptr(pop_u), _lib.ptr(pop_val), _lib.ptr(cand_u if gen > 0 else pop_u), _lib.ptr(vals), pop, d,
                 ncol, 1 if gen == 0 else 0)
        if gen > 0 and check_every and gen % check_every == 0:
            if float((pop_u - pop_u[0]).abs().max()) <= position_tolerance:
                break
    best = int(torch.argmin(pop_val))
    x = torch.empty_like(pop_u)
    ctx.call("dgp_box_from_u", _lib.ptr(pop_u), _lib.ptr(lw_t), _lib.ptr(up_t), pop, d, _lib.ptr(x))
    return {"population_u": pop_u, "values": pop_val, "best": best, "u": pop_u[best].clone(), "x": x[best].clone(),
            "value": pop_val[best].clone(), "iterations": gen}


def adam_box_minimize(value_and_grad, lw, up, u0, iterations=1000, lr=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    """value_and_grad(X [n, d]) -> (values [n, ncol], d sum(values) / dX [n, d]) device tensors; X is the same buffer in every
    step. u0 [n, d] (n independent searches side by side; the reference runs n = 1). Returns (u, x, last values)."""
    u = u0.clone().contiguous()
    n, d = u.shape
    ctx = _lib.get_context(u.device)
    lw_t, up_t = _vec(lw, d, u.device), _vec(up, d, u.device)
    X = torch.empty_like(u)
    ctx.call("dgp_box_from_u", _lib.ptr(u), _lib.ptr(lw_t), _lib.ptr(up_t), n, d, _lib.ptr(X))
    m, v = torch.zeros_like(u), torch.zeros_like(u)
    val = None
    for t in range(1, iterations + 1):
        val, dx = value_and_grad(X)
        ctx.call("dgp_adam_box_step", _lib.ptr(u), _lib.ptr(m), _lib.ptr(v), _lib.ptr(dx), _lib.ptr(lw_t), _lib.ptr(up_t), n, d, t,
                 float(lr), float(beta_1), float(beta_2), float(epsilon), _lib.ptr(X))
    return u, X, val


class GraphScope:
    """Graph replay for the duration of a search unless the caller already manages it."""

    def __init__(self, device):
        self.ctx = _lib.get_context(device)
        self.auto = not self.ctx.graph

    def __enter__(self):
        if self.auto:
            self.ctx.set_graph(True)
        return self.ctx

    def __exit__(self, *exc):
        if self.auto:
            self.ctx.set_graph(False)
        return False
