"""Multi-objective DGP (MO cyclic loop MO_DGP.py:88-122, EHVI `mo_dgp` branch EHVI.py:124-130) against the reference's
own `MO_DGP.DGP_Base` / `EHVI.EHVI` executed under tests/ref_shim (tests/golden/mo_dgp.npz, tests/golden/make_golden_mo.py), plus
self-consistency checks of the two places where the reference cannot run as written (models/MO_DGP.py, DEVIATION notes)."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mo_dgp.npz")


def _rel(a, b):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a.reshape(b.shape) - b))) / max(float(np.max(np.abs(b))), 1e-300)


class Replay:
    """The reference's N(0,1) draws in the order it consumed them (the chain's start column is recorded as [N, 1])."""

    def __init__(self, arrays, device):
        self.arrays, self.i, self.device = arrays, 0, device

    def __call__(self, shape):
        z = self.arrays[self.i]
        self.i += 1
        assert int(np.prod(z.shape)) == int(np.prod(shape)) and z.shape[-2:] == tuple(shape)[-2:], (z.shape, shape, self.i)
        return torch.as_tensor(z, device=self.device).reshape(shape)


def _params(model):
    from dgp_toolbox_b200.composite import role_parameters
    out = {}
    for i, layer in enumerate(model.layers):
        for name, p in role_parameters(layer.eval.roles).items():
            if p is not None:
                out[f"layers.{i}.{name}"] = p
        out[f"layers.{i}.q_mu"], out[f"layers.{i}.q_sqrt"] = layer.q_mu, layer.q_sqrt
        out[f"layers.{i}.Z"] = layer.feature.Z_left if layer.augmented else layer.feature.Z
    out["lik_var"] = model.likelihood.likelihood.variance
    return out


def _golden_model(g):
    """The model of make_golden_mo.py: the reference's kernels (MO_DGP.py:262-283) on two NON-augmented layers."""
    from dgp_toolbox_b200.composite import RBF, LinearKernel, White
    from dgp_toolbox_b200.gpflow_shim import Gaussian
    from dgp_toolbox_b200.models import MO_DGP
    Din = int(g["Din"])
    kernels = []
    for l in range(2):
        r = list(range(Din + 1))
        kernels.append(RBF(active_dims=r[:Din], variance=1.0) * (RBF(active_dims=r[Din:], variance=1.0)
                                                                  + LinearKernel(active_dims=r[Din:], variance=1.0))
                       + RBF(active_dims=r[:Din], variance=1.0))
    kernels[0] = kernels[0] + White(variance=1e-6)
    layers = [MO_DGP.MFLayer(kernels[i], g[f"Zinit{i}"], 1, None) for i in range(2)]
    model = MO_DGP.DGP_Base(Gaussian(), layers, loop=int(g["loop"]), num_samples=int(g["S"]))
    params = _params(model)
    for k, p in params.items():
        p.assign(g["param_" + k])
    return model, params


def test_mo_dgp_elbo_and_gradients_match_the_reference_run():
    g = np.load(G, allow_pickle=False)
    assert "reference source" in str(g["provenance"])
    model, params = _golden_model(g)
    X, Y = [g["X0"], g["X1"]], [g["Y0"], g["Y1"]]
    model.draw = Replay([g[f"draw{j}"] for j in range(int(g["n_draws"]))], "cuda")
    elbo, grads = model.ELBO_and_grads((X, Y), list(params.values()), tf_sample_Z_right=False)
    assert model.draw.i == int(g["n_draws"])
    assert abs(float(elbo) - float(g["elbo"])) <= 1e-9 * abs(float(g["elbo"]))
    checked = 0
    for k, p in params.items():
        ref = g["grad_" + k]
        gp = torch.tril(grads[p]) if k.endswith("q_sqrt") else grads[p]
        assert _rel(gp, ref) < 1e-8, (k, _rel(gp, ref))
        checked += 1
    assert checked == 22


@pytest.mark.parametrize("loop", [2, 0, 1])
def test_mo_dgp_cyclic_propagate_matches_the_reference_run(loop):
    g = np.load(G, allow_pickle=False)
    model, _ = _golden_model(g)
    model.loop = loop
    model.draw = Replay([g[f"prop{loop}_start"]], "cuda")
    with torch.no_grad():
        Fs, Fm, Fv = model.propagate(g["Xt"], S=int(g["S"]), zs=[g["zt0"], g["zt1"]])
    assert model.draw.i == 1 and len(Fs) == 2
    for i in range(2):
        assert _rel(Fs[i], g[f"prop{loop}_F{i}"]) < 1e-9
        assert _rel(Fm[i], g[f"prop{loop}_Fmean{i}"]) < 1e-9
        assert _rel(Fv[i], g[f"prop{loop}_Fvar{i}"]) < 1e-9


def test_ehvi_mo_dgp_branch_matches_the_reference_run():
    import dgp_toolbox_b200 as D
    g = np.load(G, allow_pickle=False)
    model, _ = _golden_model(g)
    model.draw = Replay([g[f"ehvi_draw{j}"] for j in range(int(g["ehvi_n_draws"]))], "cuda")
    obj = types.SimpleNamespace(name="mo_dgp", model=model)
    YND = [g["ehvi_ynd0"], g["ehvi_ynd1"]]
    out = D.EHVI(obj, g["ehvi_X"], YND, corr=False, approximation='None', S=int(g["ehvi_S"]))
    assert model.draw.i == int(g["ehvi_n_draws"])
    assert out.shape == (7, 1) and _rel(out, g["ehvi"]) < 1e-8
    # value and input gradient of -EHVI with the same draws: value agrees, gradient agrees with central differences
    model.draw = Replay([g[f"ehvi_draw{j}"] for j in range(int(g["ehvi_n_draws"]))], "cuda")
    val, dx = D.EHVI_with_grad(obj, g["ehvi_X"], YND, S=int(g["ehvi_S"]))
    assert _rel(-val, g["ehvi"]) < 1e-8
    X = g["ehvi_X"].copy()
    h = 1e-5
    for (n, j) in [(0, 0), (3, 1), (6, 0)]:
        vals = []
        for sgn in (+1, -1):
            Xp = X.copy(); Xp[n, j] += sgn * h
            model.draw = Replay([g[f"ehvi_draw{j2}"] for j2 in range(int(g["ehvi_n_draws"]))], "cuda")
            vals.append(float(-D.EHVI(obj, Xp, YND, S=int(g["ehvi_S"]))[n, 0]))
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - float(dx[n, j])) <= 1e-4 * max(abs(fd), 1e-2), (n, j, fd, float(dx[n, j]))


def _toy():
    rng = np.random.default_rng(11)
    X = rng.uniform(0, 1, (16, 2))
    Y = [np.sin(3 * X[:, :1]) + 0.5 * X[:, 1:2], np.cos(2 * X[:, :1]) * X[:, 1:2] - 0.3]
    return [X, X.copy()], Y


def test_mo_constructor_and_augmented_elbo_are_consistent():
    """make_mf_dgp / sample_Z_right (the DEVIATION paths): shapes as the reference intends, the ELBO with re-sampled Z_right equals
    the ELBO evaluated on the same Z_right without re-sampling, and the gradient w.r.t. Z_left (which flows through Z_right) agrees
    with central differences under replayed draws."""
    import dgp_toolbox_b200 as D
    X, Y = _toy()
    mo = D.MultiObjDeepGP(X, Y, loop=1)
    m = mo.model
    m.num_samples = 3
    assert [l.augmented for l in m.layers] == [False, True]
    assert m.layers[0].feature.Z.shape == (16, 3) and m.layers[1].feature.Z_left.shape == (16, 2) and m.layers[1].feature.Z.shape == (16, 3)
    rng = np.random.default_rng(2)
    for layer in m.layers:
        layer.q_mu.assign(0.3 * rng.standard_normal((16, 1)))
    m.layers[0].kern.kernels[-1].variance.assign(1e-2)
    rec = []

    def recording(shape):
        z = torch.as_tensor(rng.standard_normal(shape), device="cuda")
        rec.append(z.cpu().numpy())
        return z
    m.draw = recording
    zl = m.layers[1].feature.Z_left
    elbo, grads = m.ELBO_and_grads((X, Y), [zl, m.layers[0].q_mu])
    n_refresh = 3                                    # start column + two applications of layer 0 (50 samples each)
    assert len(rec) == n_refresh + 2 * (1 + 4)       # + per objective: start column + 4 layer applications (loop = 1)
    m.draw = Replay(rec[n_refresh:], "cuda")
    with torch.no_grad():
        same = m.ELBO((X, Y), tf_sample_Z_right=False)
    assert abs(float(same) - float(elbo)) <= 1e-12 * abs(float(elbo))
    z0 = zl.value.clone()
    h = 1e-4                                         # the ELBO is ~4e4 here: smaller steps drown in its rounding error
    for (r, c) in [(0, 0), (7, 1)]:
        vals = []
        for sgn in (+1, -1):
            zp = z0.clone(); zp[r, c] += sgn * h
            zl.assign(zp)
            m.draw = Replay(rec, "cuda")
            with torch.no_grad():
                vals.append(float(m.ELBO((X, Y))))
        zl.assign(z0)
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - float(grads[zl][r, c])) <= 2e-5 * max(abs(fd), 1e-2), (r, c, fd, float(grads[zl][r, c]))


def test_multi_objective_model_trains_and_searches(capsys):
    """MultiObjDeepGP.optimize_adam (MO_DGP.py:344-417, shortened phases): which parameters each of the three phases moves (the
    ELBO itself is dominated by the 1e-6 White variance of objective 0 at the reference's start values, so its Monte-Carlo
    estimate is no test of progress); predict; optimize_EHVI on the object."""
    import dgp_toolbox_b200 as D
    X, Y = _toy()

    def run(i1, i2, i3):
        mo = D.MultiObjDeepGP(X, Y, loop=1)
        mo.model.num_samples = 4
        m = mo.model
        z0, zl0 = m.layers[0].feature.Z.value.clone(), m.layers[1].feature.Z_left.value.clone()
        ls0 = m.layers[1].eval.params["in_ls"].value.clone()
        mo.optimize_adam(lr=0.02, iterations1=i1, iterations2=i2, iterations3=i3, messages=1)
        trace = [float(l.split("ELBO:")[1]) for l in capsys.readouterr().out.splitlines() if l.startswith("ELBO:")]
        assert len(trace) == i1 + i2 + i3 and all(np.isfinite(trace)), trace
        moved = lambda a, b: float((a - b).abs().max()) > 0
        q_mu_is_y = all(_rel(m.layers[k].q_mu.value, Y[k]) == 0 for k in range(2))
        return mo, (moved(m.layers[1].eval.params["in_ls"].value, ls0), moved(m.layers[0].feature.Z.value, z0) and
                    moved(m.layers[1].feature.Z_left.value, zl0), not q_mu_is_y,
                    abs(float(m.likelihood.likelihood.variance.value) - 1e-2 * Y[1].var()) > 0)

    assert run(3, 0, 0)[1] == (True, False, False, False)         # kernel parameters only
    assert run(0, 3, 0)[1] == (True, True, False, False)          # + inducing inputs
    mo, flags = run(1, 1, 3)
    assert flags == (True, True, True, True)                      # + variational parameters and the likelihood variance
    assert np.isfinite(float(mo.objective()))
    mean, var = mo.predict(np.random.default_rng(0).uniform(0, 1, (6, 2)))
    assert mean.shape == (6, 1) and np.all(var > 0)
    nd = D.NDC(Y, np.full((16, 1), -1.0))
    YND = D.Y_ND(Y, nd, [2.0, 1.0], [-1.5, -1.5])
    x = D.optimize_EHVI(mo, YND, popsize_DE=12, popstd_DE=1.5, iterations_DE=3, iterations_adam=3, method='DE+Adam', S=8, seed=1)
    assert x.shape == (2, 1) and np.all((x >= 0) & (x <= 1))
    capsys.readouterr()


def test_mo_and_em_optimize_nat_adam_run(capsys):
    """optimize_nat_adam of the multi-objective (MO_DGP.py:418-494) and embedded-mapping (MF_DGP_EM.py:501-582) models: the
    natural-gradient step (dgp_natgrad_pairs; pinned to the reference by tests/test_gpu_mf.py's MF run) moves every (q_mu, q_sqrt)
    pair it is given, keeps q_sqrt lower-triangular with a positive diagonal, and leaves finite ELBO values."""
    import dgp_toolbox_b200 as D
    X, Y = _toy()
    gen = torch.Generator(device="cuda").manual_seed(3)          # fixed draws: whether a natural-gradient step of this size stays
    draw = lambda shape: torch.randn(shape, generator=gen, device="cuda", dtype=torch.float64)     # positive definite depends on them
    mo = D.MultiObjDeepGP(X, Y, loop=1, draw=draw)
    mo.model.num_samples = 3
    mo.model.layers[0].kern.kernels[-1].variance.assign(1e-1)      # (the reference's 1e-6 White variance makes the ELBO ~ -1e7 on this toy)
    mo.optimize_nat_adam(lr_adam=0.01, lr_gamma=1e-4, iterations1=1, iterations2=1, iterations3=2, messages=1)
    trace = [float(l.split("ELBO:")[1]) for l in capsys.readouterr().out.splitlines() if l.startswith("ELBO:")]
    assert len(trace) == 4 and all(np.isfinite(trace))
    for k, layer in enumerate(mo.model.layers):
        q = layer.q_sqrt.value
        assert float(torch.triu(q, 1).abs().max()) == 0 and float(torch.diagonal(q, dim1=1, dim2=2).min()) > 0
        assert _rel(layer.q_mu.value, Y[k]) > 0                      # moved away from the start value by the natural-gradient step
    rng = np.random.default_rng(5)
    Xe = [rng.uniform(0, 1, (12, 2)), rng.uniform(0, 1, (8, 3))]
    f = lambda x: np.sin(4 * x[:, :1]) + x[:, 1:2]
    Ye = [f(Xe[0]), 1.3 * f(Xe[1]) + 0.2 * Xe[1][:, 2:3]]
    em = D.MultiFidelityDeepGP_EM(Xe, Ye, [Xe[1][:, :2].copy()], draw=draw)
    em.model.num_samples = 3
    em.model.layers[0].kern.kernels[-1].variance.assign(1e-2)
    start = [l.q_sqrt.value.clone() for l in list(em.model.layers) + list(em.model.layers_red)]
    em.optimize_nat_adam(lr_adam=0.01, lr_gamma=1e-3, iterations1=1, iterations2=1, iterations3=2, messages=1)
    trace = [float(l.split("ELBO:")[1]) for l in capsys.readouterr().out.splitlines() if l.startswith("ELBO:")]
    assert len(trace) == 4 and all(np.isfinite(trace))
    for l, q0 in zip(list(em.model.layers) + list(em.model.layers_red), start):
        q = l.q_sqrt.value
        assert torch.isfinite(q).all() and float(torch.triu(q, 1).abs().max()) == 0 and float(torch.diagonal(q, dim1=1, dim2=2).min()) > 0
    assert np.isfinite(float(em.objective()))
