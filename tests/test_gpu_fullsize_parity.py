"""Oracle parity AT THE BENCHMARKED LAUNCH SIZES (BASELINE config 2: 4 SVGP layers, D = 8, M = 256, S = 32): the default code
path of a large call -- V-form adjoint (>= 32 768 point-samples), 128x128 GEMM tiles, wave-aware split-K, side-stream overlap
of the parameter contractions, fused resident-tile kernels, >= 2 workspace chunks at 16 384 points -- against the CPU oracle
accumulated over 256-point chunks (oracle.elbo_and_grads_chunked; both sides draw the same Philox stream at the global point
index). Also one config-3-shaped `predict` chunk and one config-5 EI + exact-EHVI chunk. Tolerance 1e-9 relative (float64), as
BASELINE.json's north_star states. ~1 minute of host time on the GPU box."""
import numpy as np
import pytest
import torch

from oracle import dgp_oracle as O
from tests.helpers import product_model_from_problem, rel_err

pytestmark = pytest.mark.gpu


def _c2_problem(seed_shift=0):
    return O.synthetic_problem(8, [8, 8, 8], 256, 8, seed_shift=seed_shift)


def _minibatch(D0, N, index):
    from dgp_toolbox_b200 import synthetic
    return synthetic.minibatch(D0, N, index)


@pytest.mark.parametrize("N", [4096, 16384])
def test_config2_elbo_and_gradients_match_oracle_at_bench_size(N):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    prob = _c2_problem()
    S, seed = 32, 20260
    om = O.model_from_problem(prob, S)
    pm = product_model_from_problem(prob, S)
    X, Y = _minibatch(8, N, 3)
    scale = 1.0e6 / N
    val_o, g_o = O.elbo_and_grads_chunked(om, torch.as_tensor(X), torch.as_tensor(Y), seed, chunk=256, scale=scale)
    val, g = pm.ELBO_and_grads((X, Y), seed=seed, scale=scale)
    assert abs(float(val) - val_o) <= 1e-9 * abs(val_o), (float(val), val_o)
    for k, v in g_o.items():
        assert rel_err(g[k].reshape(v.shape), v) < 1e-9, (k, rel_err(g[k].reshape(v.shape), v))


def test_config3_shaped_predict_chunk_matches_oracle():
    """predict_y / predict (models/dgp.py:113-124,362-366) on the config-3 layer shape: 6 SVGP layers, D = 20, M = 512, S = 64."""
    prob = O.synthetic_problem(20, [20] * 5, 512, 8)
    S, N, seed = 64, 256, 77
    om = O.model_from_problem(prob, S)
    pm = product_model_from_problem(prob, S)
    X, _ = _minibatch(20, N, 5)
    zs = [torch.as_tensor(O.philox_normal(seed, l, S, N, layer.D_out)) for l, layer in enumerate(om.layers)]
    with torch.no_grad():
        m_o, v_o = O.predict(om, torch.as_tensor(X), S, zs)
    m, v = pm.predict(X, S, seed=seed)
    assert rel_err(m, m_o) < 1e-9 and rel_err(v, v_o) < 1e-9


def test_config5_ei_and_ehvi_chunk_matches_oracle():
    """EI.run (Infill_criteria.py:28-52) and the exact 2-objective EHVI over two config-2 DGPs (EHVI.py:107-119,150-157) on a
    2048-candidate chunk with the SURVEY §8d Pareto front (32 points on y1 = 1 - sqrt(y0))."""
    import dgp_toolbox_b200 as D
    S, N = 32, 2048
    pa, pb = _c2_problem(), _c2_problem(seed_shift=7)
    oa, ob = O.model_from_problem(pa, S), O.model_from_problem(pb, S)
    ma, mb = product_model_from_problem(pa, S), product_model_from_problem(pb, S)
    X, _ = _minibatch(8, N, 9)
    Xt = torch.as_tensor(X)
    y0 = np.linspace(0.05, 0.95, 32)
    y1 = 1.0 - np.sqrt(y0)
    order = np.argsort(-y0)
    ynd0, ynd1 = O.Y_ND(y0[order], y1[order], (1.1, 1.1), (-0.1, -0.1))
    moments = []
    with torch.no_grad():
        for om, seed in ((oa, 31), (ob, 32)):
            m_acc, v_acc, ei_acc = [], [], []
            for lo in range(0, N, 256):
                zs = [torch.as_tensor(O.philox_normal(seed, l, S, 256, layer.D_out, n_offset=lo)) for l, layer in enumerate(om.layers)]
                _, Fm, Fv = O.propagate(om.layers, Xt[lo:lo + 256], S, zs)
                mm, vv = O.mixture_moments(Fm[-1], Fv[-1])
                m_acc.append(mm); v_acc.append(vv); ei_acc.append(O.ei_analytic(Fm[-1], Fv[-1], 0.2))
            moments += [torch.cat(m_acc), torch.cat(v_acc)]
            if om is oa:
                ei_o = torch.cat(ei_acc)
        ehvi_o = O.ehvi_exact(*moments, ynd0, ynd1)
    ei = D.EI(0.2, 8).run(ma, X, analytic=True, num_samples=S, seed=31)
    assert rel_err(ei, ei_o) < 1e-8
    ehvi = D.EHVI([ma, mb], X, [ynd0[:, None], ynd1[:, None]], corr=False, S=S, seed=[31, 32])
    assert rel_err(ehvi, ehvi_o) < 1e-8
