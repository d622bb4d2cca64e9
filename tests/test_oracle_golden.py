"""Pins the CPU oracle (oracle/grief_oracle.py) to the reference's own outputs (tests/golden/)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import load_golden, model_cases, case_inputs
from oracle import grief_oracle as orc


def test_kron_eigenvalues_reference_test():
    """Inputs/outputs of the reference's tests/test_tensors/test_kron_eigenvalues.py:23-92."""
    g = load_golden("kron_eigs_d10_m3_p5")
    eigs = [g["eigs_%d" % i] for i in range(10)]
    p = int(g["n_eigs"])
    for mode in ("largest", "smallest"):
        for log_expand in (False, True):
            tag = "%s_%s" % (mode, "log" if log_expand else "lin")
            loc, vals = orc.find_extremum_eigs(eigs, p, mode=mode, log_expand=log_expand)
            assert_array_equal(loc, g["loc_" + tag])
            assert_array_equal(vals, g["vals_" + tag])          # bit-exact: same NumPy calls
            lin = np.exp(vals) if log_expand else vals
            ref = g["all_sorted_top"] if mode == "largest" else g["all_sorted_bottom"]
            assert_allclose(np.sort(lin), np.sort(ref), rtol=0, atol=1e-15)   # the reference's own assertion (decimal=15)


@pytest.mark.parametrize("name", ["topk_c2_d6_m10_p1024", "topk_c3_d10_m20_p4096",
                                  "topk_c4_d32_m8_p2048", "topk_c5_d8_m16_p8192"])
def test_topk_at_bench_configs(name):
    g = load_golden(name)
    d, p = int(g["d"]), int(g["p"])
    eigs = [g["eigs_%d" % k] for k in range(d)]
    loc, vals = orc.find_extremum_eigs(eigs, p)
    assert int(g["n_ties"]) == 0
    assert_array_equal(loc, g["eig_loc"].astype(np.int64))
    assert_array_equal(vals, g["log_lam"])


@pytest.mark.parametrize("name", model_cases())
def test_model_case(name):
    g = load_golden(name)
    c = case_inputs(g)
    alias_ties = name.startswith("ref_test_gp_grief")       # [kern,]*d => exact ties
    basis = orc.setup_inducing_cov(c["names"], c["variances"], c["lengthscales"], c["xg"], c["n_eigs"])
    assert_allclose(basis.log_lam, g["log_lam"], rtol=0, atol=0 if not alias_ties else 1e-12)
    for k in range(c["d"]):
        assert_array_equal(basis.Q[k], g["Q_%d" % k])
        if not alias_ties:
            assert_array_equal(basis.eig_loc[:, k], g["sel_%d" % k])
    Phi = orc.grief_phi(basis, c["names"], c["variances"], c["lengthscales"], c["xg"], c["x"])
    if "Phi" in g and not alias_ties:
        assert_array_equal(Phi, g["Phi"])
    f = orc.fit_from_phi(Phi, c["y"], c["w"], c["noise_var"])
    assert_allclose(f.lml, float(g["lml"]), rtol=1e-13)
    assert_allclose(f.A, g["A"], rtol=1e-12, atol=1e-12 * np.abs(g["A"]).max())
    assert_allclose(f.alpha.squeeze(), g["alpha"], rtol=1e-9, atol=1e-12)
    assert_allclose(f.log_det, float(g["log_det"]), rtol=1e-13)
    if "grad_adjoint" in g:
        gs, gw = orc.adjoint_gradient(f)
        ref = g["grad_adjoint"]
        assert_allclose(gs, ref[0], rtol=1e-11)
        assert_allclose(gw, ref[-c["n_eigs"]:], rtol=1e-10, atol=1e-12)
    if "xnew" in g:
        Phin = orc.grief_phi(basis, c["names"], c["variances"], c["lengthscales"], c["xg"], g["xnew"])
        yhat, yvar = orc.predict(f, Phin)
        assert_allclose(yhat, g["yhat"], rtol=1e-10, atol=1e-12)
        assert_allclose(np.diag(yvar), g["yvar_diag"], rtol=1e-11)


def test_type2_finite_difference_gradient():
    """models/basemodel.py:328-361 restated; parameters = [noise, (var_i, ls_i)..., w...]."""
    g = load_golden("syn_t2_n2000_d4_m8_p64")
    c = case_inputs(g)
    d, p = c["d"], c["n_eigs"]

    def lml_of(params):
        var = params[1:1 + 2 * d:2]
        ls = params[2:2 + 2 * d:2]
        f, _ = orc.lml_full(c["names"], list(var), list(ls), c["xg"], p, c["x"], c["y"], params[1 + 2 * d:], params[0])
        return f.lml

    free = g["free"]
    base, grad = orc.finite_diff_gradient(lml_of, g["parameters"].copy(), free)
    assert_allclose(base, float(g["lml"]), rtol=1e-13)
    assert_allclose(grad[free], g["grad_fd"][free], rtol=1e-4, atol=1e-3)   # FD noise: h=1e-6 on |LML|~1e3


def test_gpweb_reference_test():
    g = load_golden("gpweb_n100_p4")
    Phi, y = g["Phi"], g["y"].reshape(-1, 1)
    params = g["parameters"]
    lml, grad = orc.gpweb_lml_grad(Phi.T.dot(Phi), Phi.T.dot(y), float(y.T.dot(y)), y.shape[0], params[1:], params[0])
    assert_allclose(lml, float(g["lml"]), rtol=1e-13)
    assert_allclose(grad, g["grad"], rtol=1e-11)


def test_log_kron_reference_test():
    """tests/test_linalg.py:8-12."""
    np.random.seed(0)
    a, b = np.random.rand(5), np.random.rand(7)
    assert_allclose(orc.log_kron(np.log(a), b), np.log(np.kron(a, b)), rtol=0, atol=1e-14)


@pytest.mark.parametrize("name", ["host_t1_sum_n800_d3_m9_p36", "host_t1_prod_n800_d3_m9_p36", "host_t2_sum_n500_d2_m8_p24"])
def test_oracle_with_composite_kernels(name):
    """Kernels with children (kern/basekernel.py:131-190) given to the oracle as callables: pins `kernel_cov`'s callable branch
    to the reference's output for `k1 + k2` / `k1 * k2` kernels (oracle/gen_golden.py:composite_cases)."""
    g = load_golden(name)
    d = int(g["n_grid_dims"])

    def composite(i):
        def cov(x, z):
            kp = orc.kernel_cov(str(g["parent_name"]), x, z, float(g["parent_variance"]), g["parent_lengthscales"][i])
            kc = orc.kernel_cov(str(g["child_name"]), x, z, float(g["child_variance"]), g["child_lengthscales"][i])
            return kp * kc if str(g["op"]) == "mul" else kp + kc
        return cov

    names = [composite(i) for i in range(d)]
    xg = [g["xg_%d" % i] for i in range(d)]
    p = int(g["n_eigs"])
    basis = orc.setup_inducing_cov(names, [1.0] * d, [1.0] * d, xg, p)
    for k in range(d):
        assert_array_equal(basis.Q[k], g["Q_%d" % k])
        assert_array_equal(basis.eig_loc[:, k], g["sel_%d" % k])
    assert_array_equal(basis.log_lam, g["log_lam"])
    Phi = orc.grief_phi(basis, names, [1.0] * d, [1.0] * d, xg, g["x"])
    f = orc.fit_from_phi(Phi, g["y"], np.ones(p), float(g["noise_var"]))
    assert_allclose(f.lml, float(g["lml"]), rtol=1e-13)
    assert_allclose(f.A, g["A"], rtol=1e-12, atol=1e-12 * np.abs(g["A"]).max())
    Phin = orc.grief_phi(basis, names, [1.0] * d, [1.0] * d, xg, g["xnew"])
    yhat, yvar = orc.predict(f, Phin)
    assert_allclose(yhat, g["yhat"], rtol=1e-10, atol=1e-12)
    assert_allclose(np.diag(yvar), g["yvar_diag"], rtol=1e-11)
