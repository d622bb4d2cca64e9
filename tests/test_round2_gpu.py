"""Round-2 GPU tests: digit count of the INT8 arithmetic, the Phi^T y fused into the pass-1 builder, per-plan options,
the eigen-gap guard of the analytic gradient, empty row shards."""
import logging

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import load_golden, model_cases, case_inputs
from oracle import grief_oracle as orc
from gp_grief_b200 import _native as nat
import test_native_gpu as tn
import test_api_gpu as ta

pytestmark = pytest.mark.gpu


def _random_plan(d=4, m=6, p=200, seed=5, kernel="RBF"):
    xg = [np.linspace(0, 1, m) for _ in range(d)]
    names, var, ls = [kernel] * d, [1.0] * d, [0.3 + 0.07 * i for i in range(d)]
    basis = orc.setup_inducing_cov(names, var, ls, xg, p)
    c = dict(d=d, names=names, variances=var, lengthscales=ls, xg=xg)
    return tn._plan_from_basis(c, basis), basis, c


@pytest.mark.parametrize("mode", [1, 0], ids=["int8", "fp64"])
@pytest.mark.parametrize("name", model_cases())
def test_fused_phi_t_y_matches_reference(name, mode):
    """r = Phi^T y out of the Gram builder's sweep (grief_gram_ry) equals the stand-alone kernel and the reference's Phi.T.dot(y)
    (models/gp_grief_model.py:234)."""
    import torch
    g = load_golden(name)
    c = case_inputs(g)
    basis = orc.setup_inducing_cov(c["names"], c["variances"], c["lengthscales"], c["xg"], c["n_eigs"])
    plan = tn._plan_from_basis(c, basis)
    plan.set_option(nat.OPT_GEMM_MODE, mode)
    Phi_o = orc.grief_phi(basis, c["names"], c["variances"], c["lengthscales"], c["xg"], c["x"])
    n = c["x"].shape[0]
    T = plan.build_tables(torch.from_numpy(np.ascontiguousarray(c["x"])).cuda())
    y = torch.from_numpy(np.ascontiguousarray(c["y"].reshape(-1))).cuda()
    r = torch.full((plan.p,), np.nan, dtype=torch.float64, device="cuda")
    A = plan.gram(T, n, y=y, r_out=r)
    r_scale = float(np.abs(Phi_o).T.dot(np.abs(c["y"])).max())
    assert_allclose(r.cpu().numpy(), Phi_o.T.dot(c["y"]).reshape(-1), rtol=0, atol=1e-13 * r_scale)
    assert_allclose(r.cpu().numpy(), plan.phi_t_vec(T, n, y).cpu().numpy(), rtol=0, atol=1e-13 * r_scale)
    A_o = Phi_o.T.dot(Phi_o)
    assert_allclose(A.cpu().numpy(), A_o, rtol=0, atol=1e-12 * np.abs(A_o).max())


def test_fused_phi_t_y_over_several_slabs_and_ragged_rows():
    import torch
    rng = np.random.default_rng(21)
    plan, basis, c = _random_plan(p=300)
    n = 20011
    X = torch.from_numpy(rng.random((n, c["d"]))).cuda()
    y = torch.from_numpy(rng.standard_normal(n)).cuda()
    T = plan.build_tables(X)
    Phi = plan.phi_rows(T, n)
    r_ref = (Phi.T @ y).cpu().numpy()
    scale = float((Phi.abs().T @ y.abs()).max())
    for budget in (0, 384 * 8 * 3000):                      # one slab / seven slabs of 2944 rows
        plan.set_option(nat.OPT_SLAB_BUDGET, budget)
        r = torch.empty((plan.p,), dtype=torch.float64, device="cuda")
        plan.gram(T, n, y=y, r_out=r)
        assert_allclose(r.cpu().numpy(), r_ref, rtol=0, atol=1e-13 * scale)


@pytest.mark.parametrize("kernel", ["RBF", "Matern32"])
def test_digit_count_controls_the_error_of_both_products(kernel):
    """8 D - 2 bits per operand below its row maximum: the error of A = Phi^T Phi and of q = diag(Phi B Phi^T) must shrink by
    ~2^8 per digit and stay under the documented bound; D = 7 is the DGEMM class, the defaults are 6 / 4 / 6 (Gram / gradient pass /
    predictive variance)."""
    import torch
    rng = np.random.default_rng(8)
    plan, basis, c = _random_plan(d=4, m=7, p=260, kernel=kernel)
    n = 30011
    X = torch.from_numpy(rng.random((n, c["d"]))).cuda()
    T = plan.build_tables(X)
    Phi = plan.phi_rows(T, n)
    A_ref = (Phi.T @ Phi).cpu().numpy()
    B = rng.standard_normal((plan.p, plan.p))
    B = torch.from_numpy(B + B.T).cuda()
    q_ref = ((Phi @ B) * Phi).sum(1).cpu().numpy()
    assert (plan.get_option(nat.OPT_DIGITS_GRAM), plan.get_option(nat.OPT_DIGITS_Z), plan.get_option(nat.OPT_DIGITS_VAR)) == (6, 4, 6)
    errs_a, errs_q = [], []
    for D in (3, 4, 5, 6, 7):
        plan.set_option(nat.OPT_DIGITS_GRAM, D)
        plan.set_option(nat.OPT_DIGITS_VAR, D)
        A = plan.gram(T, n).cpu().numpy()
        assert_array_equal(A, A.T)
        q = plan.quadform_rows(T, n, B).cpu().numpy()
        ea = np.abs(A - A_ref).max() / np.abs(A_ref).max()
        eq = np.abs(q - q_ref).max() / np.abs(q_ref).max()
        errs_a.append(ea)
        errs_q.append(eq)
        bound = max(2.0 ** -(8 * D - 10), 2e-13)            # fp64 reference itself is ~1e-14 here
        assert ea < bound and eq < 64 * bound, (D, ea, eq, bound)
    assert all(errs_a[i] > 10 * errs_a[i + 1] for i in range(3)) and errs_a[3] >= errs_a[4], errs_a
    assert all(errs_q[i] > 10 * errs_q[i + 1] for i in range(3)), errs_q
    with pytest.raises(ValueError):
        plan.set_option(nat.OPT_DIGITS_Z, 8)
    with pytest.raises(ValueError):
        plan.set_option(nat.OPT_DIGITS_VAR, 8)
    with pytest.raises(ValueError):
        plan.set_option(nat.OPT_DIGITS_GRAM, 2)


@pytest.mark.parametrize("digits", [(7, 7), (6, 5), (6, 4), (5, 4)])
@pytest.mark.parametrize("name", ["syn_t2_n2000_d4_m8_p64", "syn_t2_matern52_n1500_d3_m10_p48", "syn_t1_n3000_d6_m10_p256_w",
                                  "c1_automobile"])
def test_reduced_digit_models_stay_within_the_north_star_tolerance(name, digits):
    """The reference goldens (LML, gradients) at <= 1e-9 relative with fewer int8 digit products than the 28 + 28 of the DGEMM
    class (default (6, 4): 22 + 10)."""
    g = load_golden(name)
    m = ta.build_model(g)
    m.gemm_digits = digits
    ll, grad = m.log_likelihood(return_gradient=True)
    assert_allclose(float(np.asarray(ll).squeeze()), float(g["lml"]), rtol=1e-9)
    if "grad_adjoint" in g and not bool(g["type2"]):
        ref = g["grad_adjoint"]
        ok = ~np.isnan(ref) & ~np.isnan(grad)
        assert_allclose(grad[ok], ref[ok], rtol=1e-9, atol=1e-9 * np.abs(ref[ok]).max())
    if bool(g["type2"]):
        m7 = ta.build_model(g)
        m7.gemm_digits = (7, 7)
        _, g7 = m7.log_likelihood(return_gradient=True)
        ok = ~np.isnan(g7)
        assert_allclose(grad[ok], g7[ok], rtol=0, atol=1e-9 * np.abs(g7[ok]).max())


def test_options_are_per_plan_not_process_wide():
    import torch
    rng = np.random.default_rng(2)
    plan_a, _, c = _random_plan(p=150, seed=1)
    plan_b, _, _ = _random_plan(p=150, seed=1)
    before = nat.lib().grief_get_default_option(nat.OPT_GEMM_MODE)
    plan_a.set_option(nat.OPT_GEMM_MODE, 0)
    plan_a.set_option(nat.OPT_DIGITS_GRAM, 5)
    assert plan_b.get_option(nat.OPT_GEMM_MODE) == before == 3          # neither the other plan nor the defaults moved
    assert plan_b.get_option(nat.OPT_DIGITS_GRAM) == 6
    assert nat.lib().grief_get_default_option(nat.OPT_DIGITS_GRAM) == 6
    n = 5000
    X = torch.from_numpy(rng.random((n, c["d"]))).cuda()
    Ta, Tb = plan_a.build_tables(X), plan_b.build_tables(X)
    Aa, Ab = plan_a.gram(Ta, n).cpu().numpy(), plan_b.gram(Tb, n).cpu().numpy()
    assert_allclose(Aa, Ab, rtol=0, atol=1e-12 * np.abs(Ab).max())
    # defaults are copied when a plan is created
    try:
        nat.check(nat.lib().grief_set_default_option(nat.OPT_DIGITS_Z, 6))
        plan_c, _, _ = _random_plan(p=150, seed=1)
        assert plan_c.get_option(nat.OPT_DIGITS_Z) == 6 and plan_b.get_option(nat.OPT_DIGITS_Z) == 4
    finally:
        nat.check(nat.lib().grief_set_default_option(nat.OPT_DIGITS_Z, 4))
    with pytest.raises(ValueError):
        nat.check(nat.lib().grief_set_default_option(99, 1))


def test_repeated_grid_eigenvalue_falls_back_to_finite_differences(caplog):
    """d = 2 with n_eigs = m^d selects every grid eigenpair, including the cluster at dim_noise_var where the eigenvector
    derivative does not exist: the model must say so and return the finite-difference gradient, not inf / NaN."""
    import gp_grief_b200 as gp
    from gp_grief_b200.kern.grief_kernel import DegenerateEigenpairError
    rng = np.random.default_rng(0)
    d, m, n = 2, 12, 400
    x = rng.random((n, d))
    y = np.sin(3 * x.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((n, 1))
    grid = gp.grid.InducingGrid(xg=[np.linspace(0, 1, m).reshape(-1, 1)] * d)
    kern = gp.kern.GriefKernel([gp.kern.RBF(1, lengthscale=0.6 + 0.1 * i) for i in range(d)], grid, n_eigs=m ** d,
                               reweight_eig_funs=False, opt_kernel_params=True, dim_noise_var=1e-12)
    model = gp.models.GPGriefModel(x, y, kern, noise_var=0.1)
    assert model.grad_method == 'adjoint'
    pmap = kern.base_parameter_map()
    with pytest.raises(DegenerateEigenpairError):
        kern.scaled_eigvec_derivatives([pmap[1]])
    with caplog.at_level(logging.WARNING):
        ll, grad = model.log_likelihood(return_gradient=True)
    assert "finite differences" in caplog.text
    free = np.logical_not(model._fixed_indicies)
    assert np.all(np.isfinite(grad[free]))
    ll_fd, grad_fd = model._finite_diff_gradient(model.parameters)
    assert_allclose(grad[free], grad_fd[free], rtol=1e-12, atol=0)


def test_empty_row_shard_evaluates_to_the_prior_terms():
    """A rank whose shard is empty (sharding.row_shard tail) must contribute zeros, not raise before the all-reduce."""
    import gp_grief_b200 as gp
    d, m, p = 3, 6, 40
    grid = gp.grid.InducingGrid(xg=[np.linspace(0, 1, m).reshape(-1, 1)] * d)
    kern = gp.kern.GriefKernel([gp.kern.RBF(1, lengthscale=0.3 + 0.1 * i) for i in range(d)], grid, n_eigs=p,
                               reweight_eig_funs=False, opt_kernel_params=True)
    model = gp.models.GPGriefModel(np.zeros((0, d)), np.zeros((0, 1)), kern, noise_var=0.1)
    st = model._stats()
    assert float(st['A'].abs().max()) == 0.0 and float(st['r'].abs().max()) == 0.0 and float(st['s']) == 0.0
    out = model._cov_setup(want_grad=True)
    g = model._theta_gradient([(0, 'lengthscale'), (1, 'variance')], out)
    assert_array_equal(g, np.zeros(2))


def test_arithmetic_audit_accepts_the_defaults_and_escalates_when_asked(caplog):
    """GPGriefModel audits the INT8 digit counts a posteriori against the FP64 arithmetic on a row sample (first evaluation):
    the defaults pass with a wide margin; an impossible tolerance drives both products to 7 digits (DGEMM class) with a warning,
    and the results still match the reference."""
    g = load_golden("syn_t2_n2000_d4_m8_p64")
    m = ta.build_model(g)
    ll, grad = m.log_likelihood(return_gradient=True)
    au = m.arithmetic_audit
    assert au["gram"]["digits"] == 6 and au["grad"]["digits"] == 4 and au["gram"]["rows"] == 2000
    assert au["gram"]["estimated_lml_rel_error"] < 1e-12 and au["grad"]["estimated_grad_error_over_max_abs"] < 2.5e-10
    assert_allclose(float(np.asarray(ll).squeeze()), float(g["lml"]), rtol=1e-9)
    # second evaluation at new parameters: no second audit
    prm = m.parameters.copy()
    prm[2] *= 1.01
    m.parameters = prm
    m.arithmetic_audit = None
    m.log_likelihood(return_gradient=True)
    assert m.arithmetic_audit is None
    # impossible tolerance: escalate to 7 digits
    m2 = ta.build_model(g)
    m2.audit_tol = m2.audit_tol_grad = 1e-300
    with caplog.at_level(logging.WARNING):
        ll2, grad2 = m2.log_likelihood(return_gradient=True)
    assert "recomputing with 7 digits" in caplog.text
    plan = m2._plan()
    assert plan.get_option(nat.OPT_DIGITS_GRAM) == 7 and plan.get_option(nat.OPT_DIGITS_Z) == 7
    assert m2.arithmetic_audit["gram"]["digits"] == 7 and m2.arithmetic_audit["grad"]["digits"] == 7
    assert_allclose(float(np.asarray(ll2).squeeze()), float(g["lml"]), rtol=1e-9)
    ok = ~np.isnan(grad)
    assert_allclose(grad2[ok], grad[ok], rtol=0, atol=1e-9 * np.abs(grad[ok]).max())
    # the audit can be switched off, and is skipped in the FP64 arithmetic
    m3 = ta.build_model(g)
    m3.audit_rows = 0
    m3.log_likelihood(return_gradient=True)
    assert m3.arithmetic_audit is None


@pytest.mark.parametrize("mode", [5, 1, 0], ids=["int8-pairs", "int8", "fp64"])
def test_row_maxima_from_pass_1_give_identical_gradient(mode):
    """grief_gram_ry can record max_j |Phi[row, j]| (every element passes through its registers); grief_grad_theta then skips its
    own maximum sweep.  The digit planes, hence the gradient, must be bit-identical either way."""
    import torch
    g = load_golden("syn_t2_n2000_d4_m8_p64")
    m = ta.build_model(g)
    m.audit_rows = 0
    before = nat.lib().grief_get_default_option(nat.OPT_GEMM_MODE)
    nat.check(nat.lib().grief_set_default_option(nat.OPT_GEMM_MODE, mode))
    try:
        m.kern._plan = None
        out = m._cov_setup(want_grad=True)
        plan = m._plan()
        assert plan.get_option(nat.OPT_GEMM_MODE) == mode
        T, rowmax = m._dev['tables'], m._dev['rowmax']
        n = m.num_local
        Phi = plan.phi_rows(T, n)
        hi_ref = (Phi.abs().max(dim=1).values.view(torch.int64) >> 32).to(torch.int32)
        assert int((rowmax[:n] - hi_ref).abs().max()) <= 1          # same maximum up to the rounding of a different product order
        assert int(rowmax[n:].abs().max()) == 0 if rowmax.numel() > n else True
        pmap = m.kern.base_parameter_map()
        active = [pmap[i] for i in range(len(pmap)) if pmap[i][1] == 'lengthscale' or pmap[i][0] == 0]
        dqs = m.kern.scaled_eigvec_derivatives(active)
        plan.grad_setup([a[0] for a in active], [0 if a[1] == 'variance' else 1 for a in active], dqs)
        args = (T, m._X_dev, m._y_dev, n, out['Pinv'], out['b'], float(m.noise_var))
        g_two_sweeps = plan.grad_theta(*args).cpu().numpy()
        g_one_sweep = plan.grad_theta(*args, rowmax=rowmax).cpu().numpy()
        assert_array_equal(g_one_sweep, g_two_sweeps)
    finally:
        nat.check(nat.lib().grief_set_default_option(nat.OPT_GEMM_MODE, before))
