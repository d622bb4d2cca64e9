"""Edge cases of the device path: tiny / ragged / degenerate shapes, checked against the CPU oracle."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import grief_oracle as orc

pytestmark = pytest.mark.gpu


def _model(n, d, m, p, names=None, ls=None, seed=0, type2=False, w=None, noise=0.2):
    import gp_grief_b200 as gp
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    y = np.sin(3 * x.sum(axis=1, keepdims=True)) + 0.1 * rng.standard_normal((n, 1))
    ms = m if isinstance(m, (list, tuple)) else [m] * d
    xg = [np.linspace(0, 1, mi) if mi > 1 else np.array([0.5]) for mi in ms]
    names = names or ["RBF"] * d
    ls = ls or [0.3 + 0.07 * i for i in range(d)]
    grid = gp.grid.InducingGrid(xg=[g.reshape(-1, 1) for g in xg])
    kl = [getattr(gp.kern, names[i])(1, lengthscale=ls[i]) for i in range(d)]
    if type2:
        kern = gp.kern.GriefKernel(kl, grid, n_eigs=p, reweight_eig_funs=False, opt_kernel_params=True)
    else:
        kern = gp.kern.GriefKernel(kl, grid, n_eigs=p, w=1. if w is None else w)
    mdl = gp.models.GPGriefModel(x, y, kern, noise_var=noise)
    p_eff = kern.n_eigs
    wv = np.ones(p_eff) if w is None else np.asarray(w, float)
    fit, basis = orc.lml_full(names, [1.0] * d, ls, xg, p_eff, x, y, wv, noise)
    return mdl, fit, basis, x


@pytest.mark.parametrize("n,d,m,p", [(1, 1, 3, 2), (2, 1, 5, 5), (7, 2, 3, 9), (17, 3, [4, 1, 3], 12), (127, 2, 6, 36),
                                     (128, 2, 6, 30), (129, 3, 5, 125), (1000, 1, 40, 40), (513, 2, 12, 129), (300, 4, 3, 81)])
def test_small_and_ragged_shapes(n, d, m, p):
    mdl, fit, basis, x = _model(n, d, m, p)
    ll, g = mdl.log_likelihood(return_gradient=True)
    assert_allclose(float(np.asarray(ll).squeeze()), fit.lml, rtol=1e-9, atol=1e-9)
    gs, gw = orc.adjoint_gradient(fit)
    assert_allclose(g[0], gs, rtol=1e-8, atol=1e-8 * max(1.0, abs(gs)))
    assert_allclose(g[-mdl.kern.n_eigs:], gw, rtol=1e-8, atol=1e-8 * max(1.0, np.abs(gw).max()))
    lam = np.sort(basis.log_lam)
    if np.all(np.diff(lam) > 0):            # no ties: index parity is defined
        assert_array_equal(np.asarray(mdl.kern._eig_pos), basis.eig_loc)
    yhat, var = mdl.predict(x[:3], compute_var='diag')
    Phin = orc.grief_phi(basis, ["RBF"] * d, [1.0] * d, [0.3 + 0.07 * i for i in range(d)],
                         [np.linspace(0, 1, mi) if mi > 1 else np.array([0.5]) for mi in (m if isinstance(m, list) else [m] * d)], x[:3])
    yh_o, yv_o = orc.predict(fit, Phin)
    assert_allclose(yhat.squeeze(), yh_o, rtol=1e-8, atol=1e-10)
    assert_allclose(var.squeeze(), np.diag(yv_o), rtol=1e-8)


def test_n_eigs_larger_than_grid_is_clipped():
    mdl, fit, basis, _ = _model(50, 2, 3, 1000)
    assert mdl.kern.n_eigs == 9
    assert_allclose(float(np.asarray(mdl.log_likelihood()).squeeze()), fit.lml, rtol=1e-9)


def test_many_dimensions_more_groups():
    """d = 12 with 3 grid points: the table needs several groups; also exercises the G > 4 kernels via width_cap."""
    import torch
    from gp_grief_b200 import device
    mdl, fit, basis, x = _model(400, 12, 3, 700, seed=3)
    assert_allclose(float(np.asarray(mdl.log_likelihood()).squeeze()), fit.lml, rtol=1e-9)
    d = 12
    names, ls = ["RBF"] * d, [0.3 + 0.07 * i for i in range(d)]
    xg = [np.linspace(0, 1, 3)] * d
    Q = [basis.Q[d - 1 - i] for i in range(d)]
    eig = [basis.eig[d - 1 - i] for i in range(d)]
    Phi_o = orc.grief_phi(basis, names, [1.0] * d, ls, xg, x)
    groups = []
    for cap in (20, 24, 32, 40):
        try:
            plan = device.DevicePlan(names, [1.0] * d, ls, xg, Q, eig, basis.eig_loc[:, ::-1], width_cap=cap)
        except NotImplementedError:      # cap too small for 8 groups: the library must say so, not miscompute
            continue
        groups.append(plan.n_groups)
        T = plan.build_tables(torch.from_numpy(x).cuda())
        A = plan.gram(T, x.shape[0]).cpu().numpy()
        assert_allclose(A, Phi_o.T.dot(Phi_o), rtol=0, atol=1e-12 * np.abs(A).max())
        B = np.random.default_rng(1).standard_normal((700, 700)); B = B + B.T
        q = plan.quadform_rows(T, x.shape[0], torch.from_numpy(B).cuda()).cpu().numpy()
        q_o = np.einsum("ij,jk,ik->i", Phi_o, B, Phi_o)
        assert_allclose(q, q_o, rtol=0, atol=1e-11 * np.abs(q_o).max())
    assert max(groups) >= 5, groups


def test_type2_gradient_all_kernels_vs_finite_differences():
    for name in ("RBF", "Exponential", "Matern32", "Matern52"):
        mdl, fit, basis, x = _model(600, 3, 7, 40, names=[name] * 3, type2=True, seed=5)
        ll, g = mdl.log_likelihood(return_gradient=True)
        free = np.nonzero(~np.isnan(g))[0]
        params = mdl.parameters
        num = np.zeros(free.size)
        for i, idx in enumerate(free):              # central differences of OUR LML on one selection branch
            h = 1e-5 * max(1.0, abs(params[idx]))
            pp, pm = params.copy(), params.copy()
            pp[idx] += h; pm[idx] -= h
            num[i] = (float(mdl._compute_log_likelihood(pp).squeeze()) - float(mdl._compute_log_likelihood(pm).squeeze())) / (2 * h)
        assert_allclose(g[free], num, rtol=2e-5, atol=2e-5 * np.abs(num).max(), err_msg=name)


def test_zero_rows_of_phi_and_far_points():
    """Inputs far outside the grid give Phi rows that underflow towards zero; nothing may turn into NaN."""
    import gp_grief_b200 as gp
    mdl, fit, basis, x = _model(64, 2, 5, 10)
    xfar = np.array([[50.0, -40.0], [0.5, 1e3]])
    yhat, var = mdl.predict(xfar, compute_var='diag')
    assert np.all(np.isfinite(yhat)) and np.all(np.isfinite(var))
    assert_allclose(var.squeeze(), float(mdl.noise_var), rtol=1e-6)     # no signal => prior noise only
