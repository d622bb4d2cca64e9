"""Kernels that exist only as host code: kernels with children, GPyKernel, user subclasses of BaseKernel.

Their K_xu columns are evaluated on the host per row chunk and uploaded (grief_build_tables_kxu); everything after that is the
device path.  Goldens `host_*.npz` come from the unmodified reference with composite in-house kernels (oracle/gen_golden.py,
composite_cases)."""
import sys
import types

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _pkg():
    import gp_grief_b200 as gp
    return gp


def build_composite_model(g):
    """As oracle/gen_golden.py:composite_kernel_case built the reference model."""
    gp = _pkg()
    d = int(g["n_grid_dims"])
    kern_list = []
    for i in range(d):
        k1 = getattr(gp.kern, str(g["parent_name"]))(1, variance=float(g["parent_variance"]), lengthscale=g["parent_lengthscales"][i])
        k2 = getattr(gp.kern, str(g["child_name"]))(1, variance=float(g["child_variance"]), lengthscale=g["child_lengthscales"][i])
        kern_list.append(k1 * k2 if str(g["op"]) == "mul" else k1 + k2)
    grid = gp.grid.InducingGrid(xg=[g["xg_%d" % i].reshape(-1, 1) for i in range(d)])
    if bool(g["type2"]):
        kern = gp.kern.GriefKernel(kern_list, grid, n_eigs=int(g["n_eigs"]), reweight_eig_funs=False, opt_kernel_params=True)
    else:
        kern = gp.kern.GriefKernel(kern_list, grid, n_eigs=int(g["n_eigs"]))
        for k in kern.kern_list:
            for _, ch in k._children:
                for key in ch.constraint_map:
                    ch.constraint_map[key] = np.tile('fixed', np.shape(ch.constraint_map[key]))
    return gp.models.GPGriefModel(g["x"], g["y"], kern, noise_var=float(g["noise_var"]))


@pytest.mark.parametrize("name", ["host_t1_sum_n800_d3_m9_p36", "host_t1_prod_n800_d3_m9_p36"])
def test_kernels_with_children_match_reference(name):
    g = load_golden(name)
    m = build_composite_model(g)
    assert m.kern.has_host_kernels() and m.kern.device_plan().host_kernels
    assert_array_equal(m.parameters, g["parameters"])
    assert_array_equal(np.asarray(m.constraints, dtype=str), g["constraints"])
    ll, grad = m.log_likelihood(return_gradient=True)
    assert_allclose(float(ll), float(g["lml"]), rtol=1e-9)
    ref = g["grad_adjoint"]
    free = ~np.isnan(ref)
    assert_array_equal(np.isnan(grad), np.isnan(ref))
    assert_allclose(grad[free], ref[free], rtol=1e-9, atol=1e-9 * np.abs(ref[free]).max())
    for k in range(int(g["n_grid_dims"])):
        assert_array_equal(m.kern._Sp[k].indicies, g["sel_%d" % k])
    assert_array_equal(m.kern._log_lam, g["log_lam"])
    assert_allclose(m._A, g["A"], rtol=0, atol=1e-12 * np.abs(g["A"]).max())
    yhat, vdiag = m.predict(g["xnew"], compute_var='diag')
    assert_allclose(yhat.squeeze(), g["yhat"], rtol=1e-9, atol=1e-9 * np.abs(g["yhat"]).max())
    assert_allclose(vdiag.squeeze(), g["yvar_diag"], rtol=1e-9)


def test_type2_with_host_kernels_uses_finite_differences_like_the_reference():
    g = load_golden("host_t2_sum_n500_d2_m8_p24")
    m = build_composite_model(g)
    assert m.grad_method == 'finite_difference'           # no analytic d/d(theta) for a kernel the device cannot differentiate
    assert_array_equal(m.parameters, g["parameters"])
    ll, grad = m.log_likelihood(return_gradient=True)
    assert_allclose(float(ll), float(g["lml"]), rtol=1e-9)
    ref = g["grad_fd"]
    free = ~np.isnan(ref)
    assert_array_equal(np.isnan(grad), np.isnan(ref))
    assert_allclose(grad[free], ref[free], rtol=1e-3, atol=1e-3 * np.abs(ref[free]).max())      # forward differences, h = 1e-6
    m.grad_method = 'adjoint'
    m._gradient = None
    with pytest.raises(NotImplementedError):
        m.log_likelihood(return_gradient=True)


def test_host_chunks_and_mixed_device_host_dimensions():
    """Several host chunks (ragged last chunk), host kernels in some dimensions only: the tables equal the all-device tables."""
    import torch
    gp = _pkg()
    rng = np.random.default_rng(3)
    d, m_, n = 4, 7, 1000
    xg = [np.linspace(0, 1, m_).reshape(-1, 1) for _ in range(d)]
    ls = [0.3, 0.45, 0.6, 0.5]

    class PlainRBF(gp.kern.BaseKernel):                    # a user kernel: same formula as RBF, but only host code
        def __init__(self, lengthscale):
            super(PlainRBF, self).__init__(1, None, "PlainRBF")
            self.lengthscale, self.variance = np.float64(lengthscale), np.float64(1.0)
            self.parameter_list = ['variance', 'lengthscale']
            self.constraint_map = {'variance': '+ve', 'lengthscale': '+ve'}

        def cov(self, x, z=None):
            x, z = self._process_cov_inputs(x, z)
            return self.variance * np.exp(-0.5 * (x - z.T) ** 2 / self.lengthscale ** 2)

        def grad_x(self, x, z):
            return -self.cov(x, z) * (x - z.T) / self.lengthscale ** 2

    def grief(kerns):
        return gp.kern.GriefKernel(kerns, gp.grid.InducingGrid(xg=xg), n_eigs=60)

    k_dev = grief([gp.kern.RBF(1, lengthscale=l) for l in ls])
    k_mix = grief([gp.kern.RBF(1, lengthscale=ls[0]), PlainRBF(ls[1]), gp.kern.RBF(1, lengthscale=ls[2]), PlainRBF(ls[3])])
    assert not k_dev.has_host_kernels() and k_mix.has_host_kernels()
    x = rng.random((n, d))
    y = rng.standard_normal((n, 1))
    plan = k_mix.device_plan()
    assert sorted(plan.host_kernels) == [1, 3]
    plan.host_chunk_rows = 256                             # 4 chunks, the last one ragged
    xd = torch.from_numpy(x).cuda()
    T_mix, T_dev = plan.build_tables(xd), k_dev.device_plan().build_tables(xd)
    assert_allclose(T_mix.cpu().numpy(), T_dev.cpu().numpy(), rtol=0, atol=1e-14)
    T_mix_dx, T_dev_dx = plan.build_tables(xd, deriv_dim=3), k_dev.device_plan().build_tables(xd, deriv_dim=3)
    assert_allclose(T_mix_dx.cpu().numpy(), T_dev_dx.cpu().numpy(), rtol=0, atol=1e-13)
    m_dev = gp.models.GPGriefModel(x, y, k_dev, noise_var=0.2)
    m_mix = gp.models.GPGriefModel(x, y, k_mix, noise_var=0.2)
    assert_allclose(float(m_mix.log_likelihood()), float(m_dev.log_likelihood()), rtol=1e-12)
    xq = rng.random((33, d))
    assert_allclose(m_mix.d_Yhat_d_x(xq, 1), m_dev.d_Yhat_d_x(xq, 1), rtol=1e-9, atol=1e-12)
    # the C ABI refuses a host-kernel plan without K_xu
    from gp_grief_b200 import _native as nat
    with pytest.raises(ValueError):
        nat.check(nat.lib().grief_build_tables(plan._h, nat.dev_ptr(xd), d, n, nat.dev_ptr(T_mix), nat.stream_ptr()))


class _FakeParam(np.ndarray):
    """What GPyKernel touches of a GPy parameter: .values, ._name, slice assignment, np.size / np.shape."""

    def __new__(cls, value, name):
        obj = np.asarray([value], dtype=float).view(cls)
        obj._name = name
        return obj

    def __array_finalize__(self, obj):
        self._name = getattr(obj, "_name", None)

    @property
    def values(self):
        return np.asarray(self)


def _install_fake_gpy(monkeypatch):
    """A stand-in for the GPy package (not installable here): GPy.kern.Kern and a GPy.kern.RBF with K / gradients_X."""
    GPy = types.ModuleType("GPy")
    kern = types.ModuleType("GPy.kern")

    class Kern(object):
        pass

    class RBF(Kern):
        def __init__(self, input_dim, variance=1., lengthscale=1.):
            self.input_dim = input_dim
            self.variance = _FakeParam(variance, "variance")
            self.lengthscale = _FakeParam(lengthscale, "lengthscale")
            self.flattened_parameters = [self.variance, self.lengthscale]

        def K(self, x, z=None):
            z = x if z is None else z
            return float(self.variance[0]) * np.exp(-0.5 * (x - z.T) ** 2 / float(self.lengthscale[0]) ** 2)

        def gradients_X(self, dL_dK, x, z):
            return dL_dK * (-self.K(x, z) * (x - z.T) / float(self.lengthscale[0]) ** 2).sum(1, keepdims=True)

    kern.Kern, kern.RBF = Kern, RBF
    GPy.kern = kern
    monkeypatch.setitem(sys.modules, "GPy", GPy)
    monkeypatch.setitem(sys.modules, "GPy.kern", kern)
    return GPy


def test_gpy_kernel_wrapper_runs_on_the_device_path(monkeypatch):
    """GPyKernel (kern/gpy_kernel.py) against the in-house RBF with the same formula: parameters / constraints plumbing,
    LML, predictions and d Yhat / d x all agree; kernel parameters are differentiated by finite differences."""
    GPy = _install_fake_gpy(monkeypatch)
    gp = _pkg()
    rng = np.random.default_rng(11)
    d, n = 3, 600
    xg = [np.linspace(0, 1, 8).reshape(-1, 1) for _ in range(d)]
    ls = [0.4, 0.55, 0.7]
    x, y = rng.random((n, d)), rng.standard_normal((n, 1))

    def model(kerns, **kw):
        kern = gp.kern.GriefKernel(kerns, gp.grid.InducingGrid(xg=xg), n_eigs=40, **kw)
        return gp.models.GPGriefModel(x, y, kern, noise_var=0.3)

    gk = [gp.kern.GPyKernel(1, kernel="RBF", lengthscale=l) for l in ls]
    assert gk[0].name == "GPy - RBF" and gp.kern.GPyKernel(1, kernel=GPy.kern.RBF(1), name="mine").name == "mine"
    with pytest.raises(TypeError):
        gp.kern.GPyKernel(1, kernel=3.0)
    m_gpy = model(gk)
    m_own = model([gp.kern.RBF(1, lengthscale=l) for l in ls])
    assert_array_equal(m_gpy.parameters, m_own.parameters)
    assert_array_equal(np.asarray(m_gpy.constraints, dtype=str), np.asarray(m_own.constraints, dtype=str))
    assert_allclose(float(m_gpy.log_likelihood()), float(m_own.log_likelihood()), rtol=1e-12)
    xq = rng.random((20, d))
    for a, b in zip(m_gpy.predict(xq, compute_var='diag'), m_own.predict(xq, compute_var='diag')):
        assert_allclose(a, b, rtol=1e-10)
    assert_allclose(m_gpy.d_Yhat_d_x(xq, 2), m_own.d_Yhat_d_x(xq, 2), rtol=1e-9, atol=1e-12)
    # Type-II: parameters are written through to the GPy objects; gradient by finite differences, equal to the analytic one
    gk2 = [gp.kern.GPyKernel(1, kernel="RBF", lengthscale=l) for l in ls]
    t_gpy = model(gk2, reweight_eig_funs=False, opt_kernel_params=True)
    t_own = model([gp.kern.RBF(1, lengthscale=l) for l in ls], reweight_eig_funs=False, opt_kernel_params=True)
    assert t_gpy.grad_method == 'finite_difference' and t_own.grad_method == 'adjoint'
    assert_array_equal(np.asarray(t_gpy.constraints, dtype=str), np.asarray(t_own.constraints, dtype=str))
    prm = t_gpy.parameters.copy()
    prm[2] = 0.47
    t_gpy.parameters = prm
    t_own.parameters = prm
    assert float(gk2[0].kern.lengthscale[0]) == 0.47
    l1, g1 = t_gpy.log_likelihood(return_gradient=True)
    l2, g2 = t_own.log_likelihood(return_gradient=True)
    assert_allclose(float(l1), float(l2), rtol=1e-12)
    free = ~np.isnan(g2)                                  # (the finite-difference gradient carries zeros, not NaN, in fixed slots)
    assert_allclose(g1[free], g2[free], rtol=2e-3, atol=2e-3 * np.abs(g2[free]).max())
