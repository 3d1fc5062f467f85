"""Minimal `tensorflow_probability` stand-in (TEST INFRASTRUCTURE, see ../README.md): Normal cdf/prob, the bijectors behind
GPflow's Parameter transforms, and a plain rand/1/bin differential evolution with TFP's defaults."""
import collections
import math
import types

import numpy as np
import tensorflow as tf
import torch

_c = tf.convert_to_tensor


class _Normal:
    """tfp.distributions.Normal(loc, scale): cdf = ndtr((x−loc)/scale), prob = exp(−½z²)/(scale·sqrt(2π))."""

    def __init__(self, loc, scale):
        self.loc, self.scale = _c(loc), _c(scale)

    def cdf(self, x):
        return torch.special.ndtr((_c(x) - self.loc) / self.scale)

    def prob(self, x):
        z = (_c(x) - self.loc) / self.scale
        return torch.exp(-0.5 * z * z) / (self.scale * math.sqrt(2.0 * math.pi))

    def log_prob(self, x):
        z = (_c(x) - self.loc) / self.scale
        return -0.5 * z * z - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi)


class _MVNFull:
    def __init__(self, loc, covariance_matrix):
        self.loc, self.cov = _c(loc), _c(covariance_matrix)

    def prob(self, x):
        d = _c(x) - self.loc
        L = torch.linalg.cholesky(self.cov)
        a = torch.linalg.solve_triangular(L, d.unsqueeze(-1), upper=False).squeeze(-1)
        k = d.shape[-1]
        logdet = 2.0 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
        return torch.exp(-0.5 * (a * a).sum(-1) - 0.5 * logdet - 0.5 * k * math.log(2.0 * math.pi))


distributions = types.SimpleNamespace(Normal=_Normal, MultivariateNormalFullCovariance=_MVNFull)


class Identity:
    def forward(self, x):
        return _c(x)

    def inverse(self, y):
        return _c(y)


class Softplus:
    def forward(self, x):
        x = _c(x)
        return torch.where(x > 0, x, torch.zeros_like(x)) + torch.log1p(torch.exp(-torch.abs(x)))

    def inverse(self, y):
        y = _c(y)
        return y + torch.log(-torch.expm1(-y))


class Shift:
    def __init__(self, shift):
        self.shift = shift

    def forward(self, x):
        return _c(x) + self.shift

    def inverse(self, y):
        return _c(y) - self.shift


class Chain:
    """Chain([b1, b2]).forward(x) = b1.forward(b2.forward(x)) (TFP applies the list right to left)."""

    def __init__(self, bijectors):
        self.bijectors = list(bijectors)

    def forward(self, x):
        for b in reversed(self.bijectors):
            x = b.forward(x)
        return x

    def inverse(self, y):
        for b in self.bijectors:
            y = b.inverse(y)
        return y


class FillTriangular:
    """tfp.bijectors.FillTriangular (lower): vector of n(n+1)/2 ↔ lower triangle, the 'clockwise spiral' packing:
    concat([x[..., n:], reverse(x)]) reshaped to [n, n] and masked to the lower band. [1..6] → [[4,0,0],[6,5,0],[3,2,1]]."""

    def forward(self, x):
        x = _c(x)
        m = x.shape[-1]
        n = int(round((math.sqrt(8 * m + 1) - 1) / 2))
        assert n * (n + 1) // 2 == m
        xc = torch.cat([x[..., n:], torch.flip(x, dims=[-1])], dim=-1)
        return torch.tril(xc.reshape(x.shape[:-1] + (n, n)))

    def inverse(self, y):
        y = _c(y)
        n = y.shape[-1]
        m = n * (n + 1) // 2
        idx = self.forward(torch.arange(1, m + 1, dtype=torch.float64)).to(torch.int64)   # position → source index + 1
        rows, cols = torch.tril_indices(n, n)
        out = torch.zeros(y.shape[:-2] + (m,), dtype=y.dtype)
        src = idx[rows, cols] - 1
        return out.index_copy(-1, src, y[..., rows, cols]) if not y.requires_grad else _scatter_last(out, src, y[..., rows, cols])


def _scatter_last(out, src, vals):
    res = out.clone()
    res[..., src] = vals
    return res


bijectors = types.SimpleNamespace(Identity=Identity, Softplus=Softplus, Shift=Shift, Chain=Chain, FillTriangular=FillTriangular)


DEResults = collections.namedtuple("DifferentialEvolutionOptimizerResults",
                                   ["converged", "num_objective_evaluations", "position", "objective_value",
                                    "final_population", "final_objective_values", "num_iterations"])


def _differential_evolution_minimize(objective_function, initial_population=None, initial_position=None, population_size=50,
                                     population_stddev=1.0, max_iterations=100, func_tolerance=0, position_tolerance=1e-8,
                                     differential_weight=0.5, crossover_prob=0.9, seed=None):
    """rand/1/bin with TFP's defaults. TFP's random stream is not reproducible here; the draws come from numpy's
    default_rng(seed), so only the converged optimum (not the trajectory) is comparable with the real library."""
    rng = np.random.default_rng(0 if seed is None else seed)
    if initial_population is None:
        x0 = np.asarray(_c(initial_position).numpy(), dtype=np.float64)
        pop = x0[None, :] + population_stddev * rng.standard_normal((population_size - 1, x0.size))
        pop = np.concatenate([x0[None, :], pop], 0)
    else:
        pop = np.asarray(_c(initial_population).numpy(), dtype=np.float64)
    n, d = pop.shape
    vals = np.asarray(_c(objective_function(_c(pop))).numpy()).reshape(n)
    evals, it = n, 0
    for it in range(1, max_iterations + 1):
        idx = np.array([rng.choice([j for j in range(n) if j != i], 3, replace=False) for i in range(n)])
        mutant = pop[idx[:, 0]] + differential_weight * (pop[idx[:, 1]] - pop[idx[:, 2]])
        cross = rng.random((n, d)) < crossover_prob
        cross[np.arange(n), rng.integers(0, d, n)] = True
        trial = np.where(cross, mutant, pop)
        tv = np.asarray(_c(objective_function(_c(trial))).numpy()).reshape(n)
        evals += n
        better = tv <= vals
        pop[better], vals[better] = trial[better], tv[better]
        if np.max(np.abs(pop - pop[np.argmin(vals)])) < position_tolerance:
            break
    b = int(np.argmin(vals))
    return DEResults(True, evals, pop[b].copy(), vals[b], pop, vals, it)


optimizer = types.SimpleNamespace(differential_evolution_minimize=_differential_evolution_minimize)
