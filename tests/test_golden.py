"""Committed golden vectors (tests/golden/*.npz): outputs of the REFERENCE'S OWN SOURCE, executed unmodified under the
stand-in tensorflow / gpflow of tests/ref_shim by tests/golden/make_golden.py (see the `provenance` key of every file).
CPU: the oracle reproduces them and the Philox draws they were made with are bit-stable.
GPU: the CUDA path, driven by its own in-kernel Philox draws (same seed), reproduces them to 1e-9."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dgp_oracle as O

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(f for f in glob.glob(os.path.join(HERE, "*.npz")) if not os.path.basename(f).startswith(("aux_", "mf_", "mo_")))   # mf_dgp.npz: tests/test_gpu_mf.py, mo_dgp.npz: tests/test_gpu_mo.py
AUX = os.path.join(HERE, "aux_adam_de.npz")


def _load(path):
    g = np.load(path, allow_pickle=False)
    nl = len([k for k in g.files if k.endswith("_Z")])
    layers = []
    for l in range(nl):
        kind = str(g[f"layer{l}_mean_kind"])
        layers.append(dict(Z=g[f"layer{l}_Z"], lengthscales=g[f"layer{l}_lengthscales"], variance=float(g[f"layer{l}_variance"]),
                           q_mu=g[f"layer{l}_q_mu"], q_sqrt=g[f"layer{l}_q_sqrt"], mean_kind=kind,
                           mf_W=g[f"layer{l}_mf_W"] if kind == "linear" else None, mf_b=g[f"layer{l}_mf_b"] if kind == "linear" else None,
                           white=bool(g[f"layer{l}_white"]) if f"layer{l}_white" in g.files else False,
                           kernel=str(g[f"layer{l}_kernel"]) if f"layer{l}_kernel" in g.files else "rbf"))
    return g, dict(X=g["X"], Y=g["Y"], layers=layers, lik_var=float(g["lik_var"])), nl


def test_golden_files_exist():
    assert len(FILES) >= 5 and os.path.exists(AUX)
    for f in FILES + [AUX]:
        assert "reference source" in str(np.load(f, allow_pickle=False)["provenance"]), f


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_golden(path):
    g, prob, nl = _load(path)
    S, N, seed = int(g["S"]), g["X"].shape[0], int(g["seed"])
    om = O.model_from_problem(prob, S)
    zs = [torch.as_tensor(O.philox_normal(seed, l, S, N, layer.D_out)) for l, layer in enumerate(om.layers)]
    val, grads = O.elbo_and_grads(om, torch.as_tensor(g["X"]), torch.as_tensor(g["Y"]), zs)
    assert abs(float(val) - float(g["elbo"])) <= 1e-11 * abs(float(g["elbo"]))
    for k, v in grads.items():
        ref = g["grad_" + k]
        assert np.max(np.abs(v.numpy() - ref)) <= 1e-9 * max(np.max(np.abs(ref)), 1e-300), k


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_path_reproduces_golden(path):
    import dgp_toolbox_b200 as D
    from tests.helpers import product_model_from_problem, rel_err
    g, prob, nl = _load(path)
    S, seed = int(g["S"]), int(g["seed"])
    pm = product_model_from_problem(prob, S)
    Fs, Fm, Fv = pm.propagate(g["X"], S=S, seed=seed)               # in-kernel Philox draws, same stream as the fixtures
    for l in range(nl):
        assert rel_err(Fm[l], g[f"Fmean{l}"]) < 1e-9 and rel_err(Fv[l], g[f"Fvar{l}"], scale=1.0) < 1e-9 and rel_err(Fs[l], g[f"F{l}"]) < 1e-9
    val, grads = pm.ELBO_and_grads((g["X"], g["Y"]), seed=seed)
    assert abs(float(val) - float(g["elbo"])) <= 1e-9 * abs(float(g["elbo"]))
    for k, v in grads.items():
        ref = g["grad_" + k]
        assert rel_err(v.reshape(ref.shape), ref) < 1e-9, k
    m, v = pm.predict(g["X"], S, seed=seed)
    assert rel_err(m, g["predict_mean"]) < 1e-9 and rel_err(v, g["predict_var"]) < 1e-9
    ei = D.EI(float(g["y_min"]), g["X"].shape[1])
    assert rel_err(ei.run(pm, g["X"], analytic=True, num_samples=S, seed=seed), g["ei_analytic"]) < 1e-8
    assert rel_err(ei.run(pm, g["X"], analytic=False, num_samples=S, seed=seed), g["ei_mc"]) < 1e-9


def _ehvi_problems():
    from tests.helpers import _condition
    return _condition(O.synthetic_problem(3, [3], 16, 14)), _condition(O.synthetic_problem(3, [3], 16, 14, seed_shift=5))


def test_oracle_reproduces_golden_ehvi():
    g = np.load(AUX, allow_pickle=False)
    pa, pb = _ehvi_problems()
    S, N = int(g["ehvi_S"]), 14
    moments = []
    for prob, seed in ((pa, int(g["ehvi_seed0"])), (pb, int(g["ehvi_seed1"]))):
        om = O.model_from_problem(prob, S)
        zs = [torch.as_tensor(O.philox_normal(seed, l, S, N, layer.D_out)) for l, layer in enumerate(om.layers)]
        _, Fm, Fv = O.propagate(om.layers, torch.as_tensor(pa["X"]), S, zs)
        moments += list(O.mixture_moments(Fm[-1], Fv[-1]))
    val = O.ehvi_exact(*moments, g["ehvi_ynd0"], g["ehvi_ynd1"])
    assert np.max(np.abs(val.numpy() - g["ehvi"])) <= 1e-10 * np.max(np.abs(g["ehvi"]))


@pytest.mark.gpu
def test_cuda_ehvi_reproduces_golden():
    import dgp_toolbox_b200 as D
    from tests.helpers import product_model_from_problem, rel_err
    g = np.load(AUX, allow_pickle=False)
    pa, pb = _ehvi_problems()
    S = int(g["ehvi_S"])
    ma, mb = product_model_from_problem(pa, S), product_model_from_problem(pb, S)
    out = D.EHVI([ma, mb], pa["X"], [g["ehvi_ynd0"][:, None], g["ehvi_ynd1"][:, None]], corr=False, S=S,
                 seed=[int(g["ehvi_seed0"]), int(g["ehvi_seed1"])])
    assert rel_err(out, g["ehvi"]) < 1e-8


def _adam_problem():
    from tests.helpers import _condition
    return _condition(O.synthetic_problem(2, [2], 50, 40))


def test_oracle_reproduces_golden_adam_and_de_choices():
    g = np.load(AUX, allow_pickle=False)
    prob = _adam_problem()
    om = O.model_from_problem(prob, 10)
    X, Y = torch.as_tensor(prob["X"]), torch.as_tensor(prob["Y"])
    opt = O.AdamOracle(om, lr=float(g["adam_lr"]))
    m = om
    for step in range(int(g["adam_steps"])):
        zs = [torch.as_tensor(O.philox_normal(int(g["adam_seed0"]) + step, l, 10, 40, layer.D_out)) for l, layer in enumerate(m.layers)]
        val, m = opt.step(m, X, Y, zs)
        assert abs(float(val) - float(g[f"adam_elbo{step}"])) <= 1e-11 * abs(float(g[f"adam_elbo{step}"]))
    for k, v in m.named_params().items():
        ref = g["adam_" + k]
        # Adam's first steps are ±lr·g/|g|: entries whose gradient is at rounding level amplify the 1e-14 differences between
        # the reference's and the oracle's summation order, hence 1e-10 here (measured 2e-11)
        assert np.max(np.abs(v.numpy() - ref)) <= 1e-10 * max(np.max(np.abs(ref)), 1e-2), k
    for gen in (1, 7):
        a, b, c, forced, uni = O.de_choices(2 ** 63 + 5, gen, 10, 5)
        for name, arr in (("a", a), ("b", b), ("c", c), ("forced", forced), ("uni", uni)):
            assert np.array_equal(arr, g[f"de_g{gen}_{name}"]), (gen, name)       # integer plumbing: bit-exact


@pytest.mark.gpu
def test_cuda_adam_steps_reproduce_golden():
    """Three training iterations (in-kernel Philox draws of seed 500 + step, dgp_adam_step) land on the fixture's parameters."""
    from tests.helpers import product_model_from_problem, rel_err
    g = np.load(AUX, allow_pickle=False)
    prob = _adam_problem()
    pm = product_model_from_problem(prob, 10)
    params = pm.trainable_parameters
    state = pm._adam_state(params)
    for step in range(int(g["adam_steps"])):
        flat = pm.elbo_flat((prob["X"], prob["Y"]), seed=int(g["adam_seed0"]) + step)
        assert abs(float(flat[0] - flat[1]) - float(g[f"adam_elbo{step}"])) <= 1e-8 * abs(float(g[f"adam_elbo{step}"]))
        pm._adam_step(params, flat, state, step + 1, float(g["adam_lr"]), 0.9, 0.999, 1e-7)
    for i, l in enumerate(pm.layers):
        for key, p in (("Z", l.feature.Z), ("lengthscales", l.kern.lengthscales), ("variance", l.kern.variance), ("q_mu", l.q_mu),
                       ("q_sqrt", l.q_sqrt)):
            ref = g[f"adam_layers.{i}.{key}"]
            assert rel_err(p.value.reshape(ref.shape), ref, scale=1e-2) < 1e-8, (i, key)
    assert rel_err(pm.likelihood.likelihood.variance.value.reshape(()), g["adam_lik_var"]) < 1e-8
