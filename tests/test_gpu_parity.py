"""GPU parity: the CUDA path (through the C ABI, via the drop-in classes) against the CPU oracle on identical seeded
inputs. Tolerance: 1e-9 relative (float64, north_star), relative to the un-cancelled scale of each quantity."""
import math

import numpy as np
import pytest
import torch

from oracle import dgp_oracle as O
from tests.helpers import both_models, oracle_zs, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(params=["fused", "fused-aform", "unfused"])
def conditional_path(request):
    """The implementations of the conditional: the fused kernel in V-form (default), the fused kernel in the reference's
    operation order (A = Ku^-1 Kuf explicit), and the unfused GEMM pipeline."""
    import dgp_toolbox_b200 as D
    ctx = D._lib.get_context(0)
    ctx.set_fused(request.param != "unfused")
    ctx.set_vform(request.param == "fused", "always" if request.param == "fused" else False)
    yield request.param
    ctx.set_fused(True)
    ctx.set_vform(True, True)


SHAPES = [
    # D0, num_units, M, N, S, condition (False = the SURVEY §8d inputs exactly as specified, l = sqrt(D_in))
    (2, [2], 50, 100, 10, True),          # config-1 shape (nb_DGP_regression-like: D=2, M=50, S=10)
    (8, [8, 8], 64, 48, 4, True),         # config-2 structure, small
    (8, [8, 8, 8], 256, 64, 8, False),    # config-2 exactly: 3 hidden layers (4 SVGP layers), M=256, cond(Ku) ~ 8e5
    (8, [8, 8, 8], 256, 40, 32, True),    # config-2 with S=32
    (5, [3, 6], 40, 37, 3, True),         # ragged: PCA-narrowing and zero-padding Linear mean functions, odd sizes
    (1, [1, 1], 25, 50, 10, True),        # KAT-1 shape
    (6, [6], 128, 150, 3, True),          # Mp = 128: BM=128 fused configuration with one row block
    (4, [4], 320, 70, 2, True),           # Mp = 320: BM=64 fused configuration, 5 row blocks
    (20, [20, 20], 512, 24, 2, True),     # config-3 layer shape: D=20, M=512 -> BM=128, PT=32 fused configuration, 4 row blocks
    (12, [16], 130, 40, 3, True),         # Mp = 192 (padding rows), D_out = 16 > 8, widening Linear mean function
    (5, [3, 6], 250, 77, 3, True),        # Mp = 256 with 6 padding rows, ragged widths, Linear mean functions, P = 231 (fused adjoint, BM = 256)
    (12, [16], 120, 90, 2, True),         # Mp = 128 with padding rows, D_in = 12 / 16 > 8 (fused adjoint, BM = 128, DMAX = 16)
]


@pytest.mark.parametrize("D0,num_units,M,N,S,cond", SHAPES)
def test_propagate_matches_oracle(D0, num_units, M, N, S, cond, conditional_path):
    prob, om, pm = both_models(D0, num_units, M, N, S, condition=cond)
    zs = oracle_zs(om, N, S, 7)
    X = torch.as_tensor(prob["X"])
    Fs_o, Fm_o, Fv_o = O.propagate(om.layers, X, S, zs)
    Fs, Fm, Fv = pm.propagate(prob["X"], S=S, zs=zs)
    for l in range(len(om.layers)):
        s2 = float(om.layers[l].variance)
        assert rel_err(Fm[l], Fm_o[l]) < TOL, ("mean", l)
        assert rel_err(Fv[l], Fv_o[l], scale=s2) < TOL, ("var", l)
        assert rel_err(Fs[l], Fs_o[l]) < TOL, ("sample", l)


@pytest.mark.parametrize("D0,num_units,M,N,S,cond", SHAPES)
def test_elbo_and_gradients_match_oracle(D0, num_units, M, N, S, cond, conditional_path):
    prob, om, pm = both_models(D0, num_units, M, N, S, condition=cond)
    zs = oracle_zs(om, N, S, 11)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    val_o, g_o = O.elbo_and_grads(om, X, Y, zs)
    val, g = pm.ELBO_and_grads((prob["X"], prob["Y"]), zs=zs)
    assert abs(float(val) - float(val_o)) <= TOL * abs(float(val_o))
    for k, go in g_o.items():
        gg = g[k]
        assert tuple(gg.shape) == tuple(go.shape) or gg.numel() == go.numel(), k
        assert rel_err(gg.reshape(go.shape), go) < TOL, k


def test_elbo_value_only_and_scale():
    prob, om, pm = both_models(3, [3], 30, 40, 5)
    zs = oracle_zs(om, 40, 5, 3)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    v_o = float(O.elbo(om, X, Y, zs))
    assert abs(float(pm.ELBO((prob["X"], prob["Y"]), zs=zs)) - v_o) <= TOL * abs(v_o)
    v_o3 = float(O.elbo(om, X, Y, zs, scale=3.0))
    flat = pm.elbo_flat((prob["X"], prob["Y"]), want_grad=False, scale=3.0, zs=zs)
    assert abs(float(flat[0] - flat[1]) - v_o3) <= TOL * abs(v_o3)


def test_layer_methods_match_oracle():
    prob, om, pm = both_models(4, [4], 33, 20, 2)
    X = torch.as_tensor(prob["X"])
    for lo, lp in zip(om.layers[:1], pm.layers[:1]):
        m_o, v_o = O.conditional_ND(lo, X)
        m, v = lp.conditional_ND(prob["X"])
        assert rel_err(m, m_o) < TOL and rel_err(v, v_o, scale=float(lo.variance)) < TOL
        assert abs(float(lp.KL()) - float(O.layer_KL(lo))) <= TOL * abs(float(O.layer_KL(lo)))
        Ku_o, Lu_o = O.kuu_chol(lo)
        lp.build_cholesky_if_needed()
        assert rel_err(lp.Ku, Ku_o) < TOL and rel_err(lp.Lu, Lu_o) < TOL
        K_o = O.rbf_K(lo.Z, X, lo.lengthscales, lo.variance)
        assert rel_err(lp.kern.K(lp.feature.Z.value, prob["X"]), K_o) < TOL
        X3 = torch.randn(3, 7, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
        z = torch.randn(3, 7, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
        s_o, m_o, v_o = O.sample_from_conditional(lo, X3, z)
        s, m, v = lp.sample_from_conditional(X3, z=z)
        assert rel_err(s, s_o) < TOL and rel_err(m, m_o) < TOL and rel_err(v, v_o, scale=float(lo.variance)) < TOL
        m2, v2 = lp.conditional_SND(X3)
        assert rel_err(m2, m_o) < TOL


def test_kat1_notebook_elbo_through_cuda():
    """SURVEY §8c KAT-1 / KAT-1b / KAT-3 through the drop-in DGP class (constructor path included)."""
    import dgp_toolbox_b200 as D
    np.random.seed(0)
    X = np.random.uniform(0, 1, 50)[:, None]
    Z = np.random.uniform(0, 1, 25)[:, None]
    Y = (X >= 0.5).astype(np.float64) + 1e-2 * np.random.randn(50, 1)
    kernels = [D.RBF(lengthscales=[1.0], variance=1.0) for _ in range(3)]
    model = D.DGP(X, Y, Z, kernels, [1, 1], D.Gaussian(), num_samples=10)
    assert model.number_parameters(trainable=False) == 2032
    val = float(model.ELBO((X, Y)))
    assert abs(val - (-85.98812279560475)) <= 1e-9 * 85.98812279560475, val
    for layer in model.layers[:-1]:
        layer.q_sqrt.assign(layer.q_sqrt.value * 1e-3)
    val = float(model.ELBO((X, Y)))
    assert abs(val - (-406.37591174470)) <= 1e-9 * 406.37591174470, val


def test_philox_words_are_bit_exact():
    """Seed / counter plumbing: the raw Philox-4x32-10 words of every (layer, s, n, d) equal the oracle's, bit for bit, including
    a shard offset and a seed with high bits."""
    import dgp_toolbox_b200 as D
    ctx = D._lib.get_context(0)
    for seed, layer, S, N, Dd, off in [(1234, 0, 3, 17, 2, 0), (0xDEADBEEFCAFEF00D, 3, 5, 9, 4, 1000003)]:
        w = torch.empty((S, N, Dd, 4), dtype=torch.int32, device="cuda")
        ctx.call("dgp_philox_raw", seed, layer, S, N, Dd, off, D._lib.ptr(w))
        got = w.cpu().numpy().view(np.uint32)
        assert np.array_equal(got, O.philox_uint32(seed, layer, S, N, Dd, n_offset=off))


def test_philox_stream_is_bit_exact_and_drives_the_chain(conditional_path):
    import dgp_toolbox_b200 as D
    prob, om, pm = both_models(3, [2], 20, 33, 6)
    S, N = 6, 33
    ctx = D._lib.get_context(0)
    zs = []
    for l, layer in enumerate(pm.layers):
        z = torch.empty((S, N, layer.num_outputs), dtype=torch.float64, device="cuda")
        ctx.call("dgp_philox_normal", 1234, l, S, N, layer.num_outputs, 5, D._lib.ptr(z))
        z_o = O.philox_normal(1234, l, S, N, layer.num_outputs, n_offset=5)
        # integer plumbing is bit exact; log/cos differ by <= a few ulp between libm and CUDA
        assert np.max(np.abs(z.cpu().numpy() - z_o)) < 1e-13
        zs.append(torch.as_tensor(z_o))
    X = torch.as_tensor(prob["X"])
    Fs_o, Fm_o, _ = O.propagate(om.layers, X, S, zs)
    Fs, Fm, _ = pm.propagate(prob["X"], S=S, zs=None, seed=1234, n_offset=5)   # in-kernel draws
    assert rel_err(Fs[-1], Fs_o[-1]) < TOL and rel_err(Fm[-1], Fm_o[-1]) < TOL


def test_predict_ei_ehvi_match_oracle():
    import dgp_toolbox_b200 as D
    prob, om, pm = both_models(4, [4], 48, 60, 8)
    S, N = 8, 60
    zs = oracle_zs(om, N, S, 5)
    X = torch.as_tensor(prob["X"])
    m_o, v_o = O.predict(om, X, S, zs)
    m, v = pm.predict_moments(prob["X"], S, add_lik_var=True, zs=zs)
    assert rel_err(m, m_o) < TOL and rel_err(v, v_o) < TOL
    ym, yv = pm.predict_y(prob["X"], S, zs=zs)
    ym_o, yv_o = O.predict_y(om, X, S, zs)
    assert rel_err(ym, ym_o) < TOL and rel_err(yv, yv_o) < TOL
    # EI analytic + MC
    Fs_o, Fm_o, Fv_o = O.propagate(om.layers, X, S, zs)
    y_min = float(prob["Y"].min())
    ei = D.EI(y_min, 4)
    assert rel_err(ei.run(pm, prob["X"], analytic=True, num_samples=S, zs=zs), O.ei_analytic(Fm_o[-1], Fv_o[-1], y_min)) < 1e-8
    assert rel_err(ei.run(pm, prob["X"], analytic=False, num_samples=S, zs=zs), O.ei_mc(Fs_o[-1], y_min)) < TOL
    # EHVI with two DGPs
    prob2, om2, pm2 = both_models(4, [4], 48, 60, 8, seed_shift=100)
    zs2 = oracle_zs(om2, N, S, 6)
    m0, v0 = O.mixture_moments(*O.predict_f(om, X, S, zs))
    m1, v1 = O.mixture_moments(*O.predict_f(om2, X, S, zs2))
    y0 = np.linspace(0.95, 0.05, 12)
    y1 = 1.0 - np.sqrt(y0)
    a, b = O.Y_ND(y0, y1, nadir=(1.1, 1.1), ideal=(-0.1, -0.1))
    e_o = O.ehvi_exact(m0, v0, m1, v1, a, b)
    e = D.EHVI([pm, pm2], prob["X"], [a, b], S=S, zs=[zs, zs2])
    assert rel_err(e, e_o) < 1e-8


def test_chunked_minibatch_equals_single_pass(conditional_path):
    """The workspace limit splits the minibatch into chunks of points; results must not depend on the split."""
    import dgp_toolbox_b200 as D
    prob, om, pm = both_models(4, [4], 64, 700, 4)
    ctx = D._lib.get_context(0)
    ref = pm.elbo_flat((prob["X"], prob["Y"]), want_grad=True, seed=99).clone()
    ctx.set_workspace_limit(64 << 20)
    try:
        chunked = pm.elbo_flat((prob["X"], prob["Y"]), want_grad=True, seed=99).clone()
    finally:
        ctx.set_workspace_limit(24 << 30)
    assert rel_err(chunked, ref) < 1e-11


def test_host_entry_point_matches_device_entry_point():
    prob, om, pm = both_models(3, [3], 32, 50, 4)
    dev = pm.elbo_flat((prob["X"], prob["Y"]), want_grad=True, seed=5).cpu().numpy()
    host = pm.elbo_flat_host(prob["X"], prob["Y"], want_grad=True, seed=5)
    assert np.array_equal(dev, host)


def test_gemm_engine_against_torch():
    import dgp_toolbox_b200 as D
    ctx = D._lib.get_context(0)
    g = torch.Generator(device="cuda").manual_seed(0)
    for (nt, M, N, K, tri, clow, batch, splitk) in [(0, 128, 256, 64, 0, 0, 1, 1), (0, 128, 128, 128, 1, 0, 2, 1),
                                                     (0, 128, 128, 128, 2, 0, 1, 1), (1, 128, 128, 512, 0, 1, 3, 4),
                                                     (0, 64, 32, 256, 0, 0, 1, 4), (1, 256, 256, 1024, 0, 0, 1, 8)]:
        A = torch.randn(batch, M, K, dtype=torch.float64, device="cuda", generator=g)
        if tri == 1:
            A = torch.tril(A)
        if tri == 2:
            A = torch.triu(A)
        B = torch.randn(batch, N, K, dtype=torch.float64, device="cuda", generator=g) if nt else \
            torch.randn(batch, K, N, dtype=torch.float64, device="cuda", generator=g)
        Cm = torch.randn(batch, M, N, dtype=torch.float64, device="cuda", generator=g)
        ks = torch.randn(batch, K, dtype=torch.float64, device="cuda", generator=g) if nt else None
        ref = 1.5 * (A * ks[:, None, :] if nt else A) @ (B.transpose(1, 2) if nt else B) + 0.5 * Cm
        out = Cm.clone()
        ctx.call("dgp_debug_gemm", nt, M, N, K, 1.5, D._lib.ptr(A), D._lib.ptr(B), 0.5, D._lib.ptr(out), tri, clow, batch, splitk,
                 D._lib.ptr(ks))
        if clow:
            ref, out = torch.tril(ref), torch.tril(out)
        assert rel_err(out, ref) < 1e-12, (nt, M, N, K, tri, clow, batch, splitk)


def test_natural_gradient_step_matches_oracle():
    """GPflow NaturalGradient (XiNat) step on every layer's (q_mu, q_sqrt): product (closed-form Cholesky adjoint on the CUDA
    gradients) against the oracle (autograd through the expectation parameters)."""
    prob, om, pm = both_models(3, [3], 24, 40, 4)
    zs = oracle_zs(om, 40, 4, 9)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    new = O.natgrad_step(om, X, Y, zs, 1e-3, [0, 1])
    pm.natgrad_step((prob["X"], prob["Y"]), 1e-3, [(l.q_mu, l.q_sqrt) for l in pm.layers], zs=zs)
    for (mu_o, R_o), layer in zip(new, pm.layers):
        assert rel_err(layer.q_mu.value, mu_o) < 1e-8
        assert rel_err(layer.q_sqrt.value, R_o) < 1e-8


@pytest.mark.parametrize("shape", [(5, [3, 6], 70, 33, 3, 0.05), (2, [2], 130, 21, 2, 1e-3)])
def test_natural_gradient_step_in_library_wide_and_padded(shape):
    """dgp_natgrad_step on layers with several outputs, M not a multiple of the 64-row padding, a larger step, and a subset of the
    layers (ng_all=False of models/dgp.py:198-201 updates the last layer only)."""
    D0, units, M, N, S, gamma = shape
    prob, om, pm = both_models(D0, units, M, N, S)
    zs = oracle_zs(om, N, S, 5)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    last = len(om.layers) - 1
    before = [l.q_sqrt.value.clone() for l in pm.layers]
    (mu_o, R_o), = O.natgrad_step(om, X, Y, zs, gamma, [last])
    pm.natgrad_step((prob["X"], prob["Y"]), gamma, [(pm.layers[-1].q_mu, pm.layers[-1].q_sqrt)], zs=zs)
    assert rel_err(pm.layers[-1].q_mu.value, mu_o) < 1e-8 and rel_err(pm.layers[-1].q_sqrt.value, R_o) < 1e-8
    for l, b in zip(pm.layers[:-1], before[:-1]):
        assert torch.equal(l.q_sqrt.value, b)                      # the other layers are untouched
    new = O.natgrad_step(om, X, Y, zs, gamma, list(range(last + 1)))
    prob2, om2, pm2 = both_models(D0, units, M, N, S)
    pm2.natgrad_step((prob["X"], prob["Y"]), gamma, [(l.q_mu, l.q_sqrt) for l in pm2.layers], zs=zs)
    for (mu_o, R_o), layer in zip(new, pm2.layers):
        assert rel_err(layer.q_mu.value, mu_o) < 1e-8 and rel_err(layer.q_sqrt.value, R_o) < 1e-8


@pytest.mark.parametrize("dims", [(3, [3], 24, 40, 4), (1, [1, 1], 25, 50, 10)])
def test_nat_adam_loop_in_library_matches_stepwise_calls(dims):
    """dgp_train_nat_adam (part 2 of optimize_nat_adam, models/dgp.py:331-345, in one call) == the same iterations issued one
    C-ABI call at a time (elbo_flat + dgp_adam_step, then elbo_flat + dgp_natgrad_step), bit for bit, with and without graph replay.
    The 1-D case is the notebook's shape (one lengthscale per kernel: the descriptor must point at the parameter itself, not at
    a broadcast copy that an in-library loop would leave stale)."""
    import dgp_toolbox_b200 as D
    results = []
    gamma = 0.05 if dims[0] == 3 else 1e-3      # the 1-D problem's gradients are large: a bigger step leaves the PD cone
    for mode in ("stepwise", "library", "library+graph"):
        prob, om, pm = both_models(*dims)
        data = (torch.as_tensor(prob["X"]).cuda(), torch.as_tensor(prob["Y"]).cuda())
        for l in pm.layers:
            D.gpflow.set_trainable(l.q_mu, False)
            D.gpflow.set_trainable(l.q_sqrt, False)
        vp = [(l.q_mu, l.q_sqrt) for l in pm.layers]
        params = pm.trainable_parameters
        state = pm._adam_state(params)
        ctx = D._lib.get_context(0)
        if mode == "stepwise":
            for k in range(3):
                flat = pm.elbo_flat(data, want_grad=True)
                pm._adam_step(params, flat, state, k + 1, 0.01, 0.9, 0.999, 1e-7)
                pm.natgrad_step(data, gamma, vp)
        else:
            ctx.set_graph(mode.endswith("graph"))
            try:
                trace = pm._train_nat_adam(data, params, state, 1, 3, 0.01, 0.9, 0.999, 1e-7, vp, gamma)
                assert bool(torch.isfinite(trace).all())
            finally:
                ctx.set_graph(False)
        ctx.check()
        results.append([p.value.clone() for p in pm.parameters])
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert torch.equal(a, b)


def test_e_log_p_y_matches_oracle():
    """DGP_Base.E_log_p_Y (models/dgp.py:79-87) through dgp_e_log_p_y."""
    prob, om, pm = both_models(5, [3, 6], 40, 37, 3)
    zs = oracle_zs(om, 37, 3, 2)
    ref = O.E_log_p_Y(om, torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"]), zs)
    got = pm.E_log_p_Y(prob["X"], prob["Y"], zs=zs)
    assert tuple(got.shape) == tuple(ref.shape) and rel_err(got, ref) < TOL


@pytest.mark.parametrize("isotropic", [False, True])
def test_adam_steps_match_oracle_tf_adam(isotropic):
    """dgp_adam_step (one launch: bijector inverse, chain rule, Adam, bijector) against the oracle's tf.optimizers.Adam on
    persistent unconstrained variables with autograd through the bijectors (reference models/dgp.py:132-154). Four steps on
    fixed draws; `isotropic` uses one shared lengthscale per kernel (summed gradient, broadcast value)."""
    from oracle.dgp_oracle import synthetic_problem, model_from_problem
    from tests.helpers import product_model_from_problem, _condition
    prob = _condition(synthetic_problem(3, [3], 24, 40, lik_var=0.5))
    if isotropic:
        for l in prob["layers"]:
            l["lengthscales"] = np.asarray(float(np.min(l["lengthscales"])))
    om, pm = model_from_problem(prob, 4), product_model_from_problem(prob, 4)
    zs = oracle_zs(om, 40, 4, 9)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    opt = O.AdamOracle(om, lr=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
    params = pm.trainable_parameters
    state = pm._adam_state(params)
    m = om
    for t in range(1, 5):
        v_o, m = opt.step(m, X, Y, zs)
        flat = pm.elbo_flat((prob["X"], prob["Y"]), zs=zs)
        assert rel_err(flat[0] - flat[1], v_o) < 1e-8
        pm._adam_step(params, flat, state, t, 0.01, 0.9, 0.999, 1e-7)
    for lo, lp in zip(m.layers, pm.layers):
        assert rel_err(lp.feature.Z.value, lo.Z) < 1e-8
        assert rel_err(lp.kern.lengthscales.value, lo.lengthscales) < 1e-8
        assert rel_err(lp.kern.variance.value, lo.variance) < 1e-8
        assert rel_err(lp.q_mu.value, lo.q_mu, scale=1e-2) < 1e-8
        assert rel_err(lp.q_sqrt.value, lo.q_sqrt) < 1e-8
        assert float(torch.triu(lp.q_sqrt.value, 1).abs().max()) == 0.0
    assert rel_err(pm.likelihood.likelihood.variance.value, m.lik_var) < 1e-8


@pytest.mark.parametrize("D0,units,M,N,S", [(4, [4], 48, 60, 8), (3, [2, 3], 30, 25, 1), (2, [], 20, 30, 5)])
def test_ei_input_gradient_matches_oracle_autograd(D0, units, M, N, S):
    """d sum(-EI) / dx (the gradient of the reference's Adam-on-x acquisition search, Infill_criteria.py:79-84) against
    autograd through the oracle chain; covers first-layer sharing (S > 1, L >= 2), S = 1 and a single-layer model."""
    import dgp_toolbox_b200 as D
    prob, om, pm = both_models(D0, units, M, N, S)
    zs = oracle_zs(om, N, S, 4)
    X = torch.as_tensor(prob["X"]).clone().requires_grad_(True)
    _, Fm_o, Fv_o = O.propagate(om.layers, X, S, zs)
    y_min = float(prob["Y"].min())
    neg_ei_o = O.ei_analytic(Fm_o[-1], Fv_o[-1], y_min)
    neg_ei_o.sum().backward()
    neg_ei, dx = D.EI(y_min, D0).run_with_grad(pm, prob["X"], num_samples=S, zs=zs)
    assert rel_err(neg_ei, neg_ei_o.detach()) < 1e-8
    assert rel_err(dx, X.grad) < 1e-8


@pytest.mark.parametrize("kernels", [["matern32", "matern52", "rbf"], ["matern52", "matern32", "matern32"]])
def test_matern_kernels_match_oracle(kernels, conditional_path):
    """Matern32 / Matern52 layers (the kernel options of BO/SO_BO.py:190-197,237-244): chain, ELBO, every gradient, K()."""
    prob, om, pm = both_models(3, [3, 2], 40, 45, 4, kernels=kernels)
    zs = oracle_zs(om, 45, 4, 8)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    Fs_o, Fm_o, Fv_o = O.propagate(om.layers, X, 4, zs)
    Fs, Fm, Fv = pm.propagate(prob["X"], S=4, zs=zs)
    for l in range(3):
        assert rel_err(Fm[l], Fm_o[l]) < TOL and rel_err(Fv[l], Fv_o[l], scale=1.0) < TOL and rel_err(Fs[l], Fs_o[l]) < TOL
    val_o, g_o = O.elbo_and_grads(om, X, Y, zs)
    val, g = pm.ELBO_and_grads((prob["X"], prob["Y"]), zs=zs)
    assert abs(float(val) - float(val_o)) <= TOL * abs(float(val_o))
    for k, go in g_o.items():
        assert rel_err(g[k].reshape(go.shape), go) < TOL, k
    lo, lp = om.layers[0], pm.layers[0]
    assert rel_err(lp.kern.K(lp.feature.Z.value, prob["X"]), O.kernel_K(lo.Z, X, lo.lengthscales, lo.variance, lo.kernel_kind)) < TOL


@pytest.mark.parametrize("white", [[True, True, True], [True, False, True], [False, True, False]])
@pytest.mark.parametrize("S", [4, 1])
def test_whitened_layers_match_oracle(white, S):
    """white=True layers (utils/layers.py:246,254-255,296-303): q(v) = N(q_mu, q_sqrt q_sqrt^T) with u = Lu v, so the
    conditional stops after the first triangular solve and the KL is taken against N(0, I). Chain, layer methods, ELBO, every
    gradient (the V-form adjoint with C_d = q_sqrt_d^T, beta = q_mu), EI input gradient; mixed with non-white layers."""
    import dgp_toolbox_b200 as D
    prob, om, pm = both_models(3, [3, 2], 40, 45, S, white=white)
    assert [l.white for l in om.layers] == white and [l.white for l in pm.layers] == white
    zs = oracle_zs(om, 45, S, 8)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    Fs_o, Fm_o, Fv_o = O.propagate(om.layers, X, S, zs)
    Fs, Fm, Fv = pm.propagate(prob["X"], S=S, zs=zs)
    for l in range(3):
        assert rel_err(Fm[l], Fm_o[l]) < TOL and rel_err(Fv[l], Fv_o[l], scale=1.0) < TOL and rel_err(Fs[l], Fs_o[l]) < TOL
    for lo, lp in zip(om.layers, pm.layers):
        Xl = torch.as_tensor(np.random.default_rng(1).standard_normal((17, lo.D_in)))
        m_o, v_o = O.conditional_ND(lo, Xl)
        m, v = lp.conditional_ND(Xl)
        assert rel_err(m, m_o) < TOL and rel_err(v, v_o, scale=1.0) < TOL
        assert rel_err(lp.KL(), O.layer_KL(lo)) < TOL
    val_o, g_o = O.elbo_and_grads(om, X, Y, zs)
    val, g = pm.ELBO_and_grads((prob["X"], prob["Y"]), zs=zs)
    assert abs(float(val) - float(val_o)) <= TOL * abs(float(val_o))
    for k, go in g_o.items():
        assert rel_err(g[k].reshape(go.shape), go) < TOL, k
    assert abs(float(pm.ELBO((prob["X"], prob["Y"]), zs=zs)) - float(val_o)) <= TOL * abs(float(val_o))
    Xg = X.clone().requires_grad_(True)
    _, Fm2, Fv2 = O.propagate(om.layers, Xg, S, zs)
    y_min = float(prob["Y"].min())
    O.ei_analytic(Fm2[-1], Fv2[-1], y_min).sum().backward()
    _, dx = D.EI(y_min, 3).run_with_grad(pm, prob["X"], num_samples=S, zs=zs)
    assert rel_err(dx, Xg.grad) < 1e-8


def test_whitened_dgp_constructor_and_training_step():
    """DGP(..., white=True) (models/dgp.py:248-252): q_sqrt starts at the identity, the initial ELBO equals the oracle's, an Adam
    run raises it, and the unfused debug pipeline refuses whitened layers instead of computing the non-white formulas."""
    import dgp_toolbox_b200 as D
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (60, 2))
    Y = np.sin(3 * X[:, :1]) * np.cos(2 * X[:, 1:]) + 0.05 * rng.standard_normal((60, 1))
    Z = X[:20].copy()
    model = D.DGP(X, Y, Z, [D.RBF(lengthscales=[0.7, 0.7], variance=1.0) for _ in range(2)], [2], D.Gaussian(0.1), num_samples=4,
                  seed=1, white=True)
    om = O.make_dgp(X, Y, Z, [(np.array([0.7, 0.7]), 1.0) for _ in range(2)], [2], lik_var=0.1, num_samples=4, white=True)
    for l in model.layers:
        assert torch.equal(l.q_sqrt.value, torch.eye(20, dtype=torch.float64, device="cuda")[None].repeat(l.num_outputs, 1, 1))
    zs = oracle_zs(om, 60, 4, 3)
    val_o = O.elbo(om, torch.as_tensor(X), torch.as_tensor(Y), zs)
    assert abs(float(model.ELBO((X, Y), zs=zs)) - float(val_o)) <= TOL * abs(float(val_o))
    before = float(model.ELBO((X, Y), seed=123))
    D.DGP_Base.optimize_adam(model, model.data, iterations=150, lr=0.01, messages=10 ** 9)
    assert float(model.ELBO((X, Y), seed=123)) > before + 50.0
    ctx = D._lib.get_context(0)
    ctx.set_fused(False)
    try:
        with pytest.raises(D._lib.DGPError, match="white"):
            model.ELBO((X, Y), seed=1)
    finally:
        ctx.set_fused(True)


def test_wb2_wb2s_ev_match_oracle():
    """The other moment-based criteria of Infill_criteria.py (WB2, WB2S, EV analytic / Monte-Carlo, EV.run_with_IC)."""
    import dgp_toolbox_b200 as D
    prob, om, pm = both_models(3, [3], 32, 50, 6)
    prob2, om2, pm2 = both_models(3, [3], 32, 50, 6, seed_shift=50)
    S, N = 6, 50
    zs, zs2 = oracle_zs(om, N, S, 1), oracle_zs(om2, N, S, 2)
    X = torch.as_tensor(prob["X"])
    ym, yv = O.predict_y(om, X, S, zs)
    y_min = float(prob["Y"].min())
    assert rel_err(D.WB2(y_min, 3).run(pm, prob["X"], num_samples=S, zs=zs), O.wb2(ym, yv, y_min)) < 1e-8
    assert rel_err(D.WB2S(y_min, 3).run(pm, prob["X"], num_samples=S, zs=zs), O.wb2s(ym, yv, y_min, X)) < 1e-8
    c = 0.2
    assert rel_err(D.EV_one_constraint(c, 3).run(pm, prob["X"], analytic=True, num_samples=S, zs=zs), O.ev_analytic(ym, yv, c)) < 1e-8
    Fs_o, _, _ = O.propagate(om.layers, X, S, zs)
    assert rel_err(D.EV_one_constraint(c, 3).run(pm, prob["X"], analytic=False, num_samples=S, zs=zs), O.ev_mc(Fs_o[-1], c)) < TOL
    ym2, yv2 = O.predict_y(om2, X, S, zs2)
    ev = D.EV([c, -0.1], 3).run([pm, pm2], prob["X"], analytic=True, num_samples=S, zs=[zs, zs2])
    ev_o = torch.cat([O.ev_analytic(ym, yv, c), O.ev_analytic(ym2, yv2, -0.1)], 1)
    assert rel_err(ev, ev_o) < 1e-8
    assert rel_err(D.PoF(c, 3).run(pm, prob["X"], num_samples=S, zs=zs), O.pof(ym, yv, c)) < 1e-8


@pytest.mark.parametrize("dims", [(3, [3], 14, 11, 3, None), (5, [3, 6], 70, 77, 2, None), (8, [8], 256, 512, 2, None),
                                  (3, [3, 2], 40, 45, 2, [True, False, True])])
def test_full_cov_propagate_matches_oracle(dims):
    """propagate(full_cov=True) (models/dgp.py:34-63): per-sample conditionals with the full N x N covariance
    (utils/layers.py:76-80,264-268,276) and the Cholesky reparameterisation (utils/utils.py:43-52), up to N = 512."""
    D0, units, M, N, S, white = dims
    prob, om, pm = both_models(D0, units, M, N, S, white=white)
    zs = oracle_zs(om, N, S, 3)
    with torch.no_grad():
        Fs_o, Fm_o, Fv_o = O.propagate_full_cov(om.layers, torch.as_tensor(prob["X"]), S, zs)
    Fs, Fm, Fv = pm.propagate(prob["X"], full_cov=True, S=S, zs=zs)
    for l in range(len(om.layers)):
        assert tuple(Fv[l].shape) == (S, N, N, om.layers[l].D_out)
        assert rel_err(Fm[l], Fm_o[l]) < TOL and rel_err(Fv[l], Fv_o[l], scale=float(om.layers[l].variance)) < TOL
        assert rel_err(Fs[l], Fs_o[l]) < 1e-8      # through an N x N Cholesky of a covariance with jitter-level eigenvalues
    m, v = pm.predict_f(prob["X"], full_cov=True, S=S, zs=zs)
    assert rel_err(m, Fm_o[-1]) < TOL and rel_err(v, Fv_o[-1], scale=1.0) < TOL


def test_full_cov_layer_methods_match_oracle():
    """SVGP_Layer.conditional_ND / conditional_SND / sample_from_conditional with full_cov=True; the diagonal of the covariance is
    the full_cov=False variance."""
    prob, om, pm = both_models(4, [4], 30, 25, 3)
    ol, pl = om.layers[0], pm.layers[0]
    X = torch.as_tensor(prob["X"])
    m_o, v_o = O.conditional_ND_full(ol, X)
    m, v = pl.conditional_ND(prob["X"], full_cov=True)
    assert tuple(v.shape) == (25, 25, 4) and rel_err(m, m_o) < TOL and rel_err(v, v_o, scale=1.0) < TOL
    md, vd = pl.conditional_ND(prob["X"])
    assert rel_err(torch.diagonal(v, dim1=0, dim2=1).T, vd, scale=1.0) < TOL
    Xs = torch.stack([X, X * 0.9 + 0.1, X - 0.2])
    ms, vs = pl.conditional_SND(Xs.numpy(), full_cov=True)
    for s in range(3):
        m_s, v_s = O.conditional_ND_full(ol, Xs[s])
        assert rel_err(ms[s], m_s) < TOL and rel_err(vs[s], v_s, scale=1.0) < TOL
    z = torch.randn(3, 25, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(4))
    F, mean, var = pl.sample_from_conditional(Xs.numpy(), z=z, full_cov=True)
    F_o = O.reparameterize_full(ms.cpu(), vs.cpu(), z)
    assert rel_err(F, F_o) < 1e-8 and rel_err(mean, ms) < 1e-12 and rel_err(var, vs, scale=1.0) < 1e-12
