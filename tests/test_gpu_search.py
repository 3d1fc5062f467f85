"""Acquisition search on the device (reference Infill_criteria.py:61-87): differential evolution and Adam on the
sigmoid-reparameterised candidate, against the oracle's numpy restatement driven by the same Philox choices and draws."""
import numpy as np
import pytest
import torch

from oracle import dgp_oracle as O
from tests.helpers import both_models, rel_err

pytestmark = pytest.mark.gpu
GOLDEN = 0x9E3779B97F4A7C15


def _seeds(model, n):
    """The next n seeds the model's draw sequence will hand out."""
    return [(model.seed + GOLDEN * (model._draw + k)) & 0xFFFFFFFFFFFFFFFF for k in range(n)]


def _oracle_neg_ei(om, x, S, seed, y_min):
    zs = [torch.as_tensor(O.philox_normal(seed, l, S, x.shape[0], layer.D_out)) for l, layer in enumerate(om.layers)]
    _, Fm, Fv = O.propagate(om.layers, x, S, zs)
    return O.ei_analytic(Fm[-1], Fv[-1], y_min)


@pytest.mark.parametrize("pop,d", [(4, 1), (12, 3), (301, 6)])
def test_de_generation_matches_oracle_choices(pop, d):
    import dgp_toolbox_b200 as D
    ctx = D._lib.get_context(0)
    rng = np.random.default_rng(pop)
    pop_u = rng.standard_normal((pop, d))
    lw, up = -1.0 - rng.random(d), 1.0 + rng.random(d)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pu, lwt, upt = t(pop_u), t(lw), t(up)
    cu, cx = torch.empty_like(pu), torch.empty_like(pu)
    for gen in (1, 2, 77):
        ctx.call("dgp_de_propose", D._lib.ptr(pu), pop, d, D._lib.ptr(lwt), D._lib.ptr(upt), 2 ** 63 + 5, gen, 0.5, 0.9,
                 D._lib.ptr(cu), D._lib.ptr(cx))
        a, b, c, forced, uni = O.de_choices(2 ** 63 + 5, gen, pop, d)
        take = uni < 0.9
        take[np.arange(pop), forced] = True
        want = np.where(take, pop_u[a] + 0.5 * (pop_u[b] - pop_u[c]), pop_u)
        got = cu.cpu().numpy()
        assert np.array_equal(got == pop_u, ~take | (want == pop_u))          # the same dimensions crossed over
        assert np.max(np.abs(got - want)) <= 1e-15 * max(1.0, np.max(np.abs(want)))
        assert rel_err(cx, O.box_from_u(want, lw, up)) < 1e-14


def test_de_search_on_dgp_ei_matches_oracle():
    """Six generations of the whole loop (propose -> dgp_ei on the candidate population with fresh draws -> select) against the
    oracle's de_minimize fed with the oracle's EI on the same Philox draws: same survivors, same values."""
    import dgp_toolbox_b200 as D
    from dgp_toolbox_b200 import search
    prob, om, pm = both_models(3, [3], 24, 40, 8)
    pop, d, S, gens = 12, 3, 8, 6
    y_min = float(prob["Y"].min())
    lw, up = np.full(d, -2.0), np.full(d, 2.0)
    ei = D.EI(y_min, d)
    seeds = _seeds(pm, gens + 1)
    pop0 = search.initial_population(d, pop, 1.5, 99, 0)
    assert float(pop0[0].abs().max()) == 0.0
    z = O.philox_normal(99, search.DE_INIT_LAYER, 1, pop, d)[0]
    assert rel_err(pop0[1:], 1.5 * z[1:]) < 1e-14
    res = search.de_minimize(lambda X, out: ei.run(pm, X, True, S, out=out), lw, up, d, 0, pop, 1.5, gens, seed=4242,
                             initial_population_u=pop0)
    pu, vals = O.de_minimize(lambda x, g: _oracle_neg_ei(om, torch.as_tensor(x), S, seeds[g], y_min).numpy().reshape(-1),
                             lw, up, pop0.cpu().numpy(), gens, 4242)
    assert rel_err(res["population_u"], pu) < 1e-8
    assert rel_err(res["values"], vals) < 1e-8
    assert res["best"] == int(np.argmin(vals)) and res["iterations"] == gens
    assert rel_err(res["x"], O.box_from_u(pu[np.argmin(vals)], lw, up)) < 1e-8


def test_adam_on_box_matches_oracle_autograd():
    """Adam on u with d(-EI)/dx from dgp_ei_grad and the closed-form dx/du, against the oracle's Adam with autograd through
    x = lw + (up - lw) / (1 + exp(u)) and the oracle EI; two searches side by side."""
    import dgp_toolbox_b200 as D
    from dgp_toolbox_b200 import search
    prob, om, pm = both_models(3, [3], 24, 40, 8)
    d, S, steps, n = 3, 8, 5, 2
    y_min = float(prob["Y"].min())
    lw, up = np.array([-2.0, -1.5, -1.0]), np.array([2.0, 1.0, 3.0])
    ei = D.EI(y_min, d)
    seeds = _seeds(pm, steps)
    u0 = np.array([[0.3, -0.2, 0.1], [-0.5, 0.4, 0.0]])

    def vg(u, step):
        ut = torch.as_tensor(u).clone().requires_grad_(True)
        x = torch.as_tensor(lw) + torch.as_tensor(up - lw) / (1.0 + torch.exp(ut))
        val = _oracle_neg_ei(om, x, S, seeds[step], y_min).sum()
        val.backward()
        return float(val.detach()), ut.grad.numpy()

    u_o, _ = O.adam_box_minimize(vg, lw, up, u0, steps, lr=0.05)
    out = torch.empty((n, 1), dtype=torch.float64, device="cuda")
    dx = torch.zeros((n, d), dtype=torch.float64, device="cuda")
    u, X, val = search.adam_box_minimize(lambda X: ei.run_with_grad(pm, X, S, out=out, dx=dx), lw, up,
                                         torch.from_numpy(u0).cuda(), steps, lr=0.05)
    assert rel_err(u, u_o) < 1e-8
    assert rel_err(X, O.box_from_u(u_o, lw, up)) < 1e-8


def test_ei_optimize_de_then_adam_improves_on_a_dense_grid():
    """EI.optimize('DE+Adam') on a 1-D problem: the optimum it returns is inside the box and (on common draws) at least as good
    as the best point of a 400-point grid, up to the Monte-Carlo noise of the criterion."""
    import dgp_toolbox_b200 as D
    rng = np.random.default_rng(3)
    X = np.linspace(-1.0, 1.0, 14)[:, None]
    Y = np.sin(4.0 * X) + 0.3 * X + 0.02 * rng.standard_normal(X.shape)
    model = D.DGP(X, Y, X.copy(), [D.RBF(lengthscales=[0.4], variance=1.0) for _ in range(2)], [1], D.Gaussian(0.01),
                  num_samples=8, seed=5)
    D.DGP_Base.optimize_adam(model, model.data, iterations=300, lr=0.02, messages=10 ** 9)
    crit = D.EI(float(Y.min()), 1)
    x_opt = crit.optimize(model, (np.array([-1.0]), np.array([1.0])), popsize_DE=24, iterations_DE=25, iterations_adam=60,
                          method='DE+Adam', num_samples=64, seed=11)
    assert x_opt.shape == (1, 1) and -1.0 <= x_opt.item() <= 1.0
    grid = np.linspace(-1.0, 1.0, 400)[:, None]
    g = crit.run(model, grid, True, 256, seed=77).cpu().numpy().reshape(-1)
    at_opt = float(crit.run(model, x_opt.reshape(1, 1), True, 256, seed=77))
    # 10%: the criterion is nearly flat between its two best basins (-0.70 and -0.40) and is a Monte-Carlo estimate
    assert at_opt <= g.min() + 0.10 * abs(g.min()) + 1e-6, (at_opt, g.min(), x_opt.item(), grid[np.argmin(g)])
    x_multi = crit.optimize(model, (np.array([-1.0]), np.array([1.0])), popsize_DE=24, iterations_DE=25, iterations_adam=40,
                            method='DE+Adam', num_samples=64, seed=11, adam_starts=4)    # four refinements side by side
    assert x_multi.shape == (1, 1) and crit.IC_optimized.shape == (1, 1)
    assert -1.0 <= x_multi.item() <= 1.0 and float(crit.run(model, x_multi.reshape(1, 1), True, 256, seed=77)) <= 0.5 * (g.min() + g.max())
    wb2 = D.WB2(float(Y.min()), 1)
    xw = wb2.optimize(model, (np.array([-1.0]), np.array([1.0])), popsize_DE=12, iterations_DE=5, iterations_adam=20,
                      method='DE+Adam', seed=2)
    assert xw.shape == (1, 1) and -1.0 <= xw.item() <= 1.0
    gw = wb2.run(model, grid, num_samples=256, seed=78).cpu().numpy().reshape(-1)
    assert float(wb2.run(model, xw.reshape(1, 1), num_samples=256, seed=78)) <= gw.min() + 0.25 * (gw.max() - gw.min())


@pytest.mark.parametrize("D0,units,M,N,S", [(4, [4], 48, 60, 8), (3, [2, 3], 30, 25, 1), (2, [], 20, 30, 5)])
def test_wb2_and_ev_input_gradients_match_oracle_autograd(D0, units, M, N, S):
    """dgp_acq_grad: WB2 = -(EI - mean) and the analytic expected violation on predict_y moments, value and d/dx, against
    autograd through the oracle chain (Infill_criteria.py:124-133,249-257; the gradient of the Adam stage, :160-165)."""
    import dgp_toolbox_b200 as D
    prob, om, pm = both_models(D0, units, M, N, S)
    zs = [torch.randn(S, N, l.D_out, dtype=torch.float64, generator=torch.Generator().manual_seed(4 + i)) for i, l in enumerate(om.layers)]
    y = float(prob["Y"].min()) + 0.1
    for name, crit, ofun in (("wb2", D.WB2(y, D0), O.wb2), ("ev", D.EV_one_constraint(y, D0), O.ev_analytic)):
        X = torch.as_tensor(prob["X"]).clone().requires_grad_(True)
        ym, yv = O.predict_y(om, X, S, zs)
        val_o = ofun(ym, yv, y)
        val_o.sum().backward()
        val, dx = crit.run_with_grad(pm, prob["X"], num_samples=S, zs=zs)
        assert rel_err(val, val_o.detach()) < 1e-8, name
        assert rel_err(dx, X.grad) < 1e-8, name


def _exact_gp_model(X, T, noise, seed):
    """One SVGP layer with Z = X and q(u) set to the exact GP-regression posterior at the inducing points: the model's
    predictions are the GP posterior without any training."""
    import dgp_toolbox_b200 as D
    kern = D.RBF(lengthscales=[0.4], variance=1.0)
    model = D.DGP(X, T, X.copy(), [kern], [], D.Gaussian(noise), num_samples=4, seed=seed)
    K = kern.K(X).cpu().numpy()
    A = K @ np.linalg.inv(K + noise * np.eye(len(X)))
    Spost = K - A @ K
    Spost = 0.5 * (Spost + Spost.T) + 1e-9 * np.eye(len(X))
    model.layers[0].q_mu.assign(A @ T)
    model.layers[0].q_sqrt.assign(np.linalg.cholesky(Spost)[None])
    return model


def test_constrained_search_with_expected_violation():
    """EV.optimize_with_IC (Infill_criteria.py:290-316): EI on the objective model where the constraint model's expected violation
    stays under the threshold, violation + 10000 elsewhere. The DE stage must land in the feasible part of the box; the Adam
    stage follows the gradient of the active branch and stays there."""
    import dgp_toolbox_b200 as D
    rng = np.random.default_rng(5)
    X = np.linspace(-1.0, 1.0, 16)[:, None]
    Y = np.sin(4.0 * X) + 0.3 * X + 0.02 * rng.standard_normal(X.shape)
    Cc = X + 0.02 * rng.standard_normal(X.shape)            # constraint c(x) = x <= 0: the right half of the box is infeasible
    mY, mC = _exact_gp_model(X, Y, 1e-3, 5), _exact_gp_model(X, Cc, 1e-3, 6)
    ev = D.EV([0.0], 1)
    ei = D.EI(float(Y.min()), 1)
    grid = np.linspace(-1.0, 1.0, 101)[:, None]
    vals = ev.run_with_IC(ei, mY, [mC], grid, threshold=0.1, num_samples=64, seed=3).cpu().numpy().reshape(-1)
    assert (vals[grid[:, 0] > 0.5] > 9000.0).all() and (vals[grid[:, 0] < -0.5] < 1.0).all()
    x_opt = ev.optimize_with_IC(ei, mY, [mC], (np.array([-1.0]), np.array([1.0])), threshold=0.1, num_samples=64, popsize_DE=24,
                                iterations_DE=20, iterations_adam=30, method='DE+Adam', seed=9)
    assert x_opt.shape == (1, 1) and -1.0 <= x_opt.item() <= 0.3
    at = float(ev.run_with_IC(ei, mY, [mC], x_opt.reshape(1, 1), threshold=0.1, num_samples=256, seed=4))
    assert at < 1.0 and at <= vals[vals < 9000.0].min() + 0.1 * abs(vals[vals < 9000.0].min()) + 1e-3, (at, x_opt.item())


def test_wb2s_value_and_input_gradient_match_oracle_and_pof():
    """WB2S.run_with_grad (dgp_acq_grad kind 3: the chain's input gradient + the explicit -sig'(x) EI term, Infill_criteria.py:187-198)
    against autograd through the oracle; PoF = Phi((c - mean) / sigma) (the repaired Infill_criteria.py:318-345)."""
    import dgp_toolbox_b200 as D
    from tests.helpers import both_models
    prob, om, pm = both_models(3, [3], 20, 15, 6)
    S, N = 6, 15
    zs = [torch.as_tensor(O.philox_normal(8, l, S, N, layer.D_out)) for l, layer in enumerate(om.layers)]
    y_min = float(prob["Y"].min())
    X = torch.as_tensor(prob["X"]).clone().requires_grad_(True)
    ym, yv = O.predict_y(om, X, S, zs)
    val_o = O.wb2s(ym, yv, y_min, X)
    val_o.sum().backward()
    crit = D.WB2S(y_min, 3)
    val, dx = crit.run_with_grad(pm, prob["X"], num_samples=S, seed=8)
    assert tuple(val.shape) == (N, 3) and rel_err(val, val_o.detach()) < 1e-8 and rel_err(dx, X.grad) < 1e-8
    assert rel_err(crit.run(pm, prob["X"], num_samples=S, seed=8), val_o.detach()) < 1e-8
    with torch.no_grad():
        p_o = O.pof(ym.detach(), yv.detach(), 0.3)
    p = D.PoF(0.3, 3).run(pm, prob["X"], num_samples=S, seed=8)
    assert rel_err(p, p_o) < 1e-9 and float(p.min()) >= 0.0 and float(p.max()) <= 1.0


def test_wb2s_pof_and_ehvi_searches_run():
    """WB2S.optimize('DE+Adam'), PoF.optimize_with_IC and optimize_EHVI return points inside the box that are at least as good as
    the best of a coarse grid (up to Monte-Carlo noise)."""
    import dgp_toolbox_b200 as D
    rng = np.random.default_rng(3)
    X = np.linspace(-1.0, 1.0, 14)[:, None]
    Y = np.sin(4.0 * X) + 0.3 * X + 0.02 * rng.standard_normal(X.shape)
    mk = lambda Yv, seed: D.DGP(X, Yv, X.copy(), [D.RBF(lengthscales=[0.4], variance=1.0) for _ in range(2)], [1], D.Gaussian(0.01),
                                num_samples=8, seed=seed)
    model, cons, obj2 = mk(Y, 5), mk(np.cos(3.0 * X), 6), mk(np.cos(2.0 * X) + 0.5 * X, 7)
    for m in (model, cons, obj2):
        D.DGP_Base.optimize_adam(m, m.data, iterations=150, lr=0.02, messages=10 ** 9)
    box = (np.array([-1.0]), np.array([1.0]))
    w = D.WB2S(float(Y.min()), 1)
    x1 = w.optimize(model, box, popsize_DE=16, iterations_DE=10, iterations_adam=20, method='DE+Adam', seed=2)
    assert x1.shape == (1, 1) and -1.0 <= x1.item() <= 1.0 and tuple(w.IC_optimized.shape) == (1, 1)
    pof = D.PoF(0.0, 1)
    x2 = pof.optimize_with_IC(D.EI(float(Y.min()), 1), model, cons, box, popsize_DE=16, iterations_DE=10, seed=2)
    assert x2.shape == (1, 1) and -1.0 <= x2.item() <= 1.0 and float(pof.IC_optimized) <= 1e-12     # -EI * PoF <= 0
    y0 = np.linspace(-0.8, 0.6, 5)
    ynd = D.Y_ND([y0[:, None], (0.9 - y0)[:, None]], list(np.argsort(-y0)), nadir=[1.5, 1.5], ideal=[-1.5, -1.5])
    x3 = D.optimize_EHVI([model, obj2], ynd, popsize_DE=16, iterations_DE=8, S=32, seed=4, bounds=box)
    grid = np.linspace(-1.0, 1.0, 41)[:, None]
    g = D.EHVI([model, obj2], grid, ynd, S=256, seed=[9, 10]).cpu().numpy().reshape(-1)
    at = float(D.EHVI([model, obj2], x3.reshape(1, 1), ynd, S=256, seed=[9, 10]))
    assert x3.shape == (1, 1) and -1.0 <= x3.item() <= 1.0 and at >= 0.7 * g.max()


def test_ehvi_gradient_and_adam_stage():
    """dgp_ehvi2d_grad (value + partial derivatives w.r.t. the four moments) and EHVI_with_grad (the gradient w.r.t. the candidates
    through both models' adjoint chains, dgp_acq_grad kind 4) against autograd through the oracle; optimize_EHVI('DE+Adam') and
    ('Adam') run the stage the reference takes with tf.GradientTape (EHVI.py:218-234)."""
    import dgp_toolbox_b200 as D
    prob, om, pm = both_models(3, [3], 20, 15, 6)
    prob2, om2, pm2 = both_models(3, [3], 20, 15, 6, seed_shift=40)
    S, N = 6, 15
    zs = [[torch.as_tensor(O.philox_normal(8 + k, l, S, N, layer.D_out)) for l, layer in enumerate(m.layers)] for k, m in enumerate((om, om2))]
    y0 = np.linspace(-0.9, 0.7, 6)
    ynd = D.Y_ND([y0[:, None], (0.8 - y0)[:, None]], list(np.argsort(-y0)), nadir=[1.6, 1.6], ideal=[-1.6, -1.6])
    X = torch.as_tensor(prob["X"]).clone().requires_grad_(True)
    mom = [O.mixture_moments(*O.predict_f(m, X, S, z)) for m, z in zip((om, om2), zs)]
    y0t, y1t = [torch.as_tensor(np.asarray(a, dtype=np.float64).reshape(-1)) for a in (ynd[0], ynd[1])]
    e_o = O.ehvi_exact(mom[0][0], mom[0][1], mom[1][0], mom[1][1], y0t, y1t)
    (-e_o).sum().backward()
    val, dx = D.EHVI_with_grad([pm, pm2], prob["X"], ynd, S=S, seed=[8, 9])
    assert tuple(val.shape) == (N, 1) and rel_err(val, -e_o.detach()) < 1e-8
    assert rel_err(dx, X.grad) < 1e-7, rel_err(dx, X.grad)
    assert rel_err(D.EHVI([pm, pm2], prob["X"], ynd, S=S, seed=[8, 9]), e_o.detach()) < 1e-8
    # the moment partials alone
    mm = [t.detach().clone().requires_grad_(True) for t in (mom[0][0], mom[0][1], mom[1][0], mom[1][1])]
    O.ehvi_exact(*mm, y0t, y1t).sum().backward()
    ctx = D._lib.get_context(0)
    dv = [D._lib.as_device(t.detach().numpy()) for t in mm]
    out = torch.empty((N, 1), dtype=torch.float64, device="cuda")
    g = torch.empty((N, 4), dtype=torch.float64, device="cuda")
    y0d, y1d = D._lib.as_device(y0t.numpy()), D._lib.as_device(y1t.numpy())   # named: the pointers must outlive the call
    ctx.call("dgp_ehvi2d_grad", *[D._lib.ptr(t) for t in dv], N, D._lib.ptr(y0d), D._lib.ptr(y1d), int(y0t.numel()), D._lib.ptr(out),
             D._lib.ptr(g))
    torch.cuda.synchronize()
    g_o = torch.cat([t.grad for t in mm], 1)
    assert rel_err(g, g_o) < 1e-9
    # the search stages
    box = (np.full(3, -1.0), np.full(3, 1.0))
    base = float(D.EHVI([pm, pm2], np.zeros((1, 3)), ynd, S=64, seed=[3, 4]))
    for method in ("Adam", "DE+Adam"):
        x = D.optimize_EHVI([pm, pm2], ynd, popsize_DE=12, iterations_DE=4, iterations_adam=15, lr_adam=0.05, method=method, S=16, seed=5,
                            bounds=box, init_adam=np.zeros(3) if method == "Adam" else None)
        assert x.shape == (3, 1) and np.all(x >= -1.0) and np.all(x <= 1.0)
        if method == "Adam":   # 15 Adam steps from the origin do not make the criterion worse (up to Monte-Carlo noise)
            assert float(D.EHVI([pm, pm2], x.reshape(1, 3), ynd, S=64, seed=[3, 4])) >= base - 0.1 * abs(base) - 1e-6
