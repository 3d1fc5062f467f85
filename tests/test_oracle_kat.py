"""Pins the CPU oracle to the only numbers the reference itself printed,
plus self-consistency checks (finite differences, identities, Philox known-answer vectors)."""
import math

import numpy as np
import torch

from oracle import dgp_oracle as O


def _kat1_model(S=10):
    # Notebooks_dgp/nb_DGP_regression.ipynb cell 10 / cells 14-18
    np.random.seed(0)
    X = np.random.uniform(0, 1, 50)[:, None]
    Z = np.random.uniform(0, 1, 25)[:, None]
    Y = (X >= 0.5).astype(np.float64) + 1e-2 * np.random.randn(50, 1)
    kernels = [(np.array([1.0]), 1.0)] * 3
    model = O.make_dgp(X, Y, Z, kernels, [1, 1], lik_var=1.0, num_samples=S)
    return model, torch.as_tensor(X), torch.as_tensor(Y)


def _zs(model, N, S, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(S, N, l.D_out, dtype=torch.float64, generator=g) for l in model.layers]


def test_kat1_elbo_at_prior():
    model, X, Y = _kat1_model()
    for seed in (0, 1, 2):
        val = float(O.elbo(model, X, Y, _zs(model, 50, 10, seed)))
        assert abs(val - (-85.98812279560475)) <= 1e-10 * 85.98812279560475, val


def test_kat1b_after_qsqrt_scaling():
    model, X, Y = _kat1_model()
    for l in model.layers[:-1]:  # models/dgp.py:268-269
        l.q_sqrt = l.q_sqrt * 1e-3
    val = float(O.elbo(model, X, Y, _zs(model, 50, 10, 3)))
    expected = -85.98812279559426 - 2 * 0.5 * (25 * 1e-6 - 25 + 25 * math.log(1e6))
    assert abs(val - expected) <= 1e-10 * abs(expected), (val, expected)
    assert abs(expected - (-406.37591174470)) < 1e-8


def test_kat2_bo_constraint_model():
    # nb_dgp_BO cells 30/61: N = M = 5, Z = X, standardised targets (sum Y^2 = N), after q_sqrt *= 1e-3
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 1, (5, 1))
    Y = rng.standard_normal((5, 1))
    Y = (Y - Y.mean()) / Y.std()
    kernels = [(np.array([1.0]), 1.0)] * 3
    model = O.make_dgp(X, Y, X.copy(), kernels, [1, 1], lik_var=1.0, num_samples=10)
    for l in model.layers[:-1]:
        l.q_sqrt = l.q_sqrt * 1e-3
    val = float(O.elbo(model, torch.as_tensor(X), torch.as_tensor(Y), _zs(model, 5, 10, 0)))
    assert abs(val - (-73.6722504558447)) <= 1e-9 * 73.67, val


def test_kat3_number_parameters():
    model, _, _ = _kat1_model()
    assert O.number_parameters(model) == 2032


def test_prior_identities():
    model, X, Y = _kat1_model()
    for l in model.layers:
        assert abs(float(O.layer_KL(l))) < 1e-8
    m, v = O.conditional_ND(model.layers[-1], torch.rand(7, 1, dtype=torch.float64))
    assert torch.allclose(v, torch.ones_like(v), atol=1e-9)
    assert torch.allclose(m, torch.zeros_like(m), atol=1e-12)


def test_white_nonwhite_equivalence():
    prob = O.synthetic_problem(3, [3], 12, 20)
    lay = prob["layers"][0]
    nw = O.make_layer(lay["Z"], lay["lengthscales"], 1.0, 3, "identity", q_mu=lay["q_mu"], q_sqrt=lay["q_sqrt"])
    _, Lu = O.kuu_chol(nw)
    w = O.make_layer(lay["Z"], lay["lengthscales"], 1.0, 3, "identity", white=True, q_mu=lay["q_mu"], q_sqrt=lay["q_sqrt"])
    nw2 = O.make_layer(lay["Z"], lay["lengthscales"], 1.0, 3, "identity", q_mu=Lu @ w.q_mu, q_sqrt=Lu[None] @ w.q_sqrt)
    X = torch.as_tensor(prob["X"])
    m1, v1 = O.conditional_ND(w, X)
    m2, v2 = O.conditional_ND(nw2, X)
    assert torch.allclose(m1, m2, rtol=1e-9, atol=1e-11)
    assert torch.allclose(v1, v2, rtol=1e-9, atol=1e-11)
    assert abs(float(O.layer_KL(w) - O.layer_KL(nw2))) < 1e-8


def test_autograd_vs_finite_differences():
    prob = O.synthetic_problem(3, [3], 10, 16)
    model = O.model_from_problem(prob, num_samples=4)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    zs = _zs(model, 16, 4, 7)
    val, grads = O.elbo_and_grads(model, X, Y, zs)
    h = 1e-6
    rng = np.random.default_rng(0)
    for name in ["layers.0.Z", "layers.0.lengthscales", "layers.0.variance", "layers.0.q_mu", "layers.1.q_sqrt",
                 "layers.1.Z", "lik_var"]:
        p = model.named_params()[name]
        flat = p.view(-1)
        idxs = rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False)
        if name.endswith("q_sqrt"):
            idxs = [0, 10 + 1, 2 * 10 + 1]  # lower-triangular entries
        for i in idxs:
            old = float(flat[i])
            flat[i] = old + h
            fp = float(O.elbo(model, X, Y, zs))
            flat[i] = old - h
            fm = float(O.elbo(model, X, Y, zs))
            flat[i] = old
            fd = (fp - fm) / (2 * h)
            g = float(grads[name].view(-1)[i])
            assert abs(fd - g) <= 2e-5 * max(1.0, abs(g)), (name, i, fd, g)


def test_fill_triangular_matches_tfp_doc_example():
    # tfp FillTriangular docstring: [1..6] -> [[4,0,0],[6,5,0],[3,2,1]]
    out = O.fill_triangular(np.arange(1.0, 7.0))
    assert (out == np.array([[4, 0, 0], [6, 5, 0], [3, 2, 1]])).all()
    x = np.random.default_rng(0).standard_normal((2, 10))
    assert np.allclose(O.fill_triangular_inverse(O.fill_triangular(x)), x)


def test_softplus_roundtrip():
    th = np.array([1e-3, 0.1, 1.0, 30.0])
    assert np.allclose(O.softplus(O.softplus_inv(th)), th, rtol=1e-12)


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    r = O.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xFFFFFFFF
    r = O.philox4x32_10(f, f, f, f, f, f)
    assert [int(x) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = O.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normal_moments_and_shard_invariance():
    z = O.philox_normal(1234, 1, 8, 4096, 4)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    z2 = O.philox_normal(1234, 1, 8, 2048, 4, n_offset=2048)
    assert np.array_equal(z[:, 2048:], z2)


def test_ei_and_ehvi_basic():
    g = torch.Generator().manual_seed(0)
    Fm = torch.randn(16, 5, 1, dtype=torch.float64, generator=g)
    Fv = torch.rand(16, 5, 1, dtype=torch.float64, generator=g) + 0.1
    nei = O.ei_analytic(Fm, Fv, 0.3)
    assert (nei <= 0).all()
    # EHVI vs Monte-Carlo hypervolume improvement on one candidate
    y0 = np.array([0.8, 0.5, 0.2])  # descending objective 0
    y1 = np.array([0.2, 0.5, 0.8])
    a, b = O.Y_ND(y0, y1, nadir=(1.1, 1.1), ideal=(-10.0, -10.0))
    m0, v0, m1, v1 = [torch.tensor([x], dtype=torch.float64) for x in (0.4, 0.04, 0.3, 0.09)]
    val = float(O.ehvi_exact(m0, v0, m1, v1, a, b))
    rng = np.random.default_rng(0)
    f0 = 0.4 + 0.2 * rng.standard_normal(200000)
    f1 = 0.3 + 0.3 * rng.standard_normal(200000)

    def hv(front):
        front = sorted(front)
        tot, prev1 = 0.0, 1.1
        for p0, p1 in front:
            if p1 < prev1 and p0 < 1.1:
                tot += (1.1 - p0) * (prev1 - p1)
                prev1 = p1
        return tot
    base = hv(list(zip(y0, y1)))
    # improvement computed on a subsample (python loop)
    n = 4000
    imp = np.mean([max(hv(list(zip(y0, y1)) + [(f0[i], f1[i])]) - base, 0.0)
                   for i in range(n)])
    assert abs(val - imp) < 0.02, (val, imp)


def test_oracle_natural_gradient_step():
    """XiNat step: gamma = 0 is a no-op (theta -> (mu, S) round trip), a small gamma increases the ELBO."""
    prob = O.synthetic_problem(3, [3], 12, 30)
    om = O.model_from_problem(prob, 3)
    zs = [torch.randn(3, 30, l.D_out, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) for l in om.layers]
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    e0 = float(O.elbo(om, X, Y, zs))
    for (mu, R), l in zip(O.natgrad_step(om, X, Y, zs, 0.0, [0, 1]), om.layers):
        assert float((mu - l.q_mu).abs().max()) < 1e-12 and float((R - l.q_sqrt).abs().max()) < 1e-12
    for (mu, R), l in zip(O.natgrad_step(om, X, Y, zs, 1e-3, [0, 1]), om.layers):
        l.q_mu, l.q_sqrt = mu, R
    assert float(O.elbo(om, X, Y, zs)) > e0 + 100.0


def test_oracle_search_restatement_de_and_adam_on_a_quadratic():
    """The oracle's differential evolution / box-Adam (restating tfp.optimizer.differential_evolution_minimize and
    tf.optimizers.Adam as Infill_criteria.py:61-87 uses them): partner indices are distinct and never the member itself, one
    dimension always crosses over, and both stages find the minimum of a separable quadratic inside the box."""
    for pop, d in [(4, 1), (7, 2), (33, 5)]:
        for gen in (1, 9):
            a, b, c, forced, uni = O.de_choices(2 ** 40 + 3, gen, pop, d)
            ii = np.arange(pop)
            for v in (a, b, c):
                assert v.min() >= 0 and v.max() < pop and (v != ii).all()
            assert (a != b).all() and (a != c).all() and (b != c).all()
            assert forced.min() >= 0 and forced.max() < d and 0.0 < uni.min() and uni.max() < 1.0
    lw, up = np.array([-1.0, 0.0]), np.array([2.0, 1.0])
    target = np.array([0.5, 0.25])
    rng = np.random.default_rng(0)
    pop0 = np.r_[np.zeros((1, 2)), 1.5 * rng.standard_normal((19, 2))]
    pu, vals = O.de_minimize(lambda x, g: ((x - target) ** 2).sum(1), lw, up, pop0, 80, 11)
    assert np.allclose(O.box_from_u(pu[np.argmin(vals)], lw, up), target, atol=1e-6)
    assert (np.diff(np.sort(vals)) >= 0).all() and vals.max() < 1e-6      # selection only ever keeps improvements

    def vg(u, step):
        e = np.exp(u)
        x = lw + (up - lw) / (1.0 + e)
        return float(((x - target) ** 2).sum()), 2.0 * (x - target) * (-(up - lw) * e / (1.0 + e) ** 2)

    u, val = O.adam_box_minimize(vg, lw, up, np.zeros(2), 2000, lr=0.01)
    assert np.allclose(O.box_from_u(u, lw, up), target, atol=1e-3) and val < 1e-5


def test_adam_oracle_first_step_closed_form():
    """tf.optimizers.Adam's first step is u -= lr * g / (|g| + eps / sqrt(1 - beta_2)) for every variable, with g the gradient
    w.r.t. the UNCONSTRAINED variable. The oracle takes g by autograd through the bijectors; here it is rebuilt from the
    constrained-space gradients with the closed-form chain rules the product's kernel uses (softplus: d theta / d u =
    1 - exp(-(theta - lower)); FillTriangular: lower triangle only) -- the two routes must give the same update."""
    prob = O.synthetic_problem(3, [2], 12, 20, lik_var=0.3)
    om = O.model_from_problem(prob, 3)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    g0 = torch.Generator().manual_seed(0)
    zs = [torch.randn(3, 20, l.D_out, dtype=torch.float64, generator=g0) for l in om.layers]
    lr, b2, eps = 0.01, 0.999, 1e-7
    _, grads = O.elbo_and_grads(om, X, Y, zs)
    opt = O.AdamOracle(om, lr=lr, beta_2=b2, epsilon=eps)
    u_before = {k: v.clone() for k, v in opt.u.items()}
    opt.step(om, X, Y, zs)
    for name, theta in om.named_params().items():
        kind = O.AdamOracle._kind(name)
        gl = -grads[name]                                            # d(-ELBO) / d theta
        if kind in ("positive", "positive_shift"):
            gu = gl * (-torch.expm1(-(theta - (1e-6 if kind == "positive_shift" else 0.0))))
        elif kind == "triangular":
            gu = torch.tril(gl)
        else:
            gu = gl
        want = u_before[name] - lr * gu / (gu.abs() + eps / np.sqrt(1.0 - b2))
        assert torch.allclose(opt.u[name], want, rtol=1e-10, atol=1e-13), name


def test_oracle_full_cov_branches_are_consistent_with_the_diagonal_ones():
    """full_cov=True restatement (utils/layers.py:264-266, utils/utils.py:43-52): the diagonal of the full
    covariance is the diagonal-branch variance, the covariance is symmetric and positive semi-definite, one point reduces the
    full-covariance sample to the diagonal reparameterisation, and whitened / non-whitened layers agree as they do on the diagonal."""
    prob = O.synthetic_problem(3, [3], 14, 9)
    om = O.model_from_problem(prob, 2)
    X = torch.as_tensor(prob["X"])
    for layer in om.layers:
        Xl = torch.as_tensor(np.random.default_rng(0).standard_normal((9, layer.D_in)))
        m_d, v_d = O.conditional_ND(layer, Xl)
        m_f, v_f = O.conditional_ND_full(layer, Xl)
        assert v_f.shape == (9, 9, layer.D_out) and torch.allclose(m_f, m_d, rtol=1e-12, atol=1e-12)
        assert torch.allclose(torch.diagonal(v_f, dim1=0, dim2=1).T, v_d, rtol=1e-10, atol=1e-12)
        assert torch.allclose(v_f, v_f.transpose(0, 1), rtol=1e-12, atol=1e-13)
        assert float(torch.linalg.eigvalsh(v_f[:, :, 0]).min()) > -1e-9
    g = torch.Generator().manual_seed(1)
    zs = [torch.randn(2, 9, l.D_out, dtype=torch.float64, generator=g) for l in om.layers]
    Fs, Fm, Fv = O.propagate_full_cov(om.layers, X, 2, zs)
    assert Fs[-1].shape == (2, 9, 1) and Fv[-1].shape == (2, 9, 9, 1)
    _, Fm_d, Fv_d = O.propagate(om.layers, X, 2, zs)
    assert torch.allclose(Fm[0], Fm_d[0], rtol=1e-12, atol=1e-12)                    # first layer: same inputs either way
    assert torch.allclose(torch.diagonal(Fv[0], dim1=1, dim2=2).permute(0, 2, 1), Fv_d[0], rtol=1e-10, atol=1e-12)
    one = X[:1]
    z1 = [z[:, :1] for z in zs]
    Fs1, _, _ = O.propagate_full_cov(om.layers, one, 2, z1)
    Fs1_d, _, _ = O.propagate(om.layers, one, 2, z1)
    assert torch.allclose(Fs1[-1], Fs1_d[-1], rtol=1e-10, atol=1e-12)
