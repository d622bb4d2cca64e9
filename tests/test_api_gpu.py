"""GPU tests through the public Python API (the drop-in boundary): they read like the reference's own tests."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_almost_equal, assert_array_almost_equal, assert_array_equal
from scipy.stats import multivariate_normal as mvn

from conftest import load_golden, model_cases

pytestmark = pytest.mark.gpu


def _pkg():
    import gp_grief_b200 as gp_grief
    return gp_grief


def build_model(g, alias=None, type2=None):
    """Model from a golden fixture, built exactly as oracle/gen_golden.py built the reference model."""
    gp = _pkg()
    d = int(g["n_grid_dims"])
    cls = getattr(gp.kern, str(g["kernel_name"]))
    alias = bool(g["alias"]) if alias is None else alias
    type2 = bool(g["type2"]) if type2 is None else type2
    if alias:
        k0 = cls(1, variance=g["variances"][0], lengthscale=g["lengthscales"][0])
        kern_list = [k0, ] * d
    else:
        kern_list = [cls(1, variance=g["variances"][i], lengthscale=g["lengthscales"][i]) for i in range(d)]
    grid = gp.grid.InducingGrid(xg=[g["xg_%d" % i].reshape(-1, 1) for i in range(d)])
    if type2:
        kern = gp.kern.GriefKernel(kern_list, grid, n_eigs=int(g["n_eigs"]), reweight_eig_funs=False, opt_kernel_params=True)
    else:
        kern = gp.kern.GriefKernel(kern_list, grid, n_eigs=int(g["n_eigs"]), w=g["w"].copy())
    return gp.models.GPGriefModel(g["x"], g["y"], kern, noise_var=float(g["noise_var"]))


def test_reference_test_gp_grief_model():
    """tests/test_models/test_gp_grief_model.py:14-40 of the reference, with the in-house RBF."""
    gp = _pkg()
    np.random.seed(0)
    d, n = 5, 100
    x = np.random.rand(n, d)
    y = np.random.rand(n, 1)
    grid = gp.grid.InducingGrid(x)
    kern = gp.kern.RBF(1, lengthscale=0.5)
    kern = gp.kern.GriefKernel(kern_list=[kern, ] * d, grid=grid, n_eigs=50)
    m = gp.models.GPGriefModel(x, y, kern, noise_var=0.1)
    lml = m._compute_log_likelihood(m.parameters)
    K = m._mv_cov(np.identity(n))
    alp = m._mv_cov_inv(y)
    assert_array_almost_equal(alp, np.linalg.solve(K, y), decimal=6)
    assert_almost_equal(m._cov_log_det(), np.linalg.slogdet(K)[1], decimal=6)
    lml_exact = mvn.logpdf(x=y.squeeze(), mean=np.zeros(n), cov=K)
    assert_almost_equal(lml, lml_exact, decimal=6)
    # and against the number the reference itself produced on these inputs
    assert_allclose(float(lml), float(load_golden("ref_test_gp_grief_model")["lml"]), rtol=1e-9)


@pytest.mark.parametrize("name", [n for n in model_cases() if not n.startswith("syn_t2")])
def test_type1_model_matches_reference(name):
    g = load_golden(name)
    m = build_model(g)
    params = m.parameters
    assert_array_equal(params, g["parameters"])
    ll, grad = m.log_likelihood(return_gradient=True)
    assert_allclose(float(ll), float(g["lml"]), rtol=1e-9)
    ref = g["grad_adjoint"]
    free = ~np.isnan(ref)
    assert_array_equal(np.isnan(grad), np.isnan(ref))
    assert_allclose(grad[free], ref[free], rtol=1e-9, atol=1e-9 * np.abs(ref[free]).max())
    tie_free = not bool(g["alias"])
    if tie_free:     # eigen-index parity is only defined without ties
        d = int(g["n_grid_dims"])
        for k in range(d):
            assert_array_equal(m.kern._Sp[k].indicies, g["sel_%d" % k])
        assert_array_equal(m.kern._log_lam, g["log_lam"])
        assert_allclose(m._A, g["A"], rtol=0, atol=1e-12 * np.abs(g["A"]).max())
        assert_allclose(m._alpha.squeeze(), g["alpha"], rtol=1e-8, atol=1e-9 * np.abs(g["alpha"]).max())
        if "Phi" in g:
            Phi = m.kern.cov(g["x"])[0]
            assert_allclose(Phi, g["Phi"], rtol=1e-11, atol=1e-13 * np.abs(g["Phi"]).max())
    if "xnew" in g:
        yhat, yvar = m.predict(g["xnew"])
        assert yhat.shape == (g["xnew"].shape[0], 1) and yvar.shape == (g["xnew"].shape[0],) * 2
        assert_allclose(yhat.squeeze(), g["yhat"], rtol=1e-9, atol=1e-9 * np.abs(g["yhat"]).max())
        assert_allclose(np.diag(yvar), g["yvar_diag"], rtol=1e-9)
        if "yvar_full" in g:
            assert_allclose(yvar, g["yvar_full"], rtol=1e-8, atol=1e-9 * np.abs(g["yvar_full"]).max())
        yhat2, vdiag = m.predict(g["xnew"], compute_var='diag')
        assert_allclose(vdiag.squeeze(), g["yvar_diag"], rtol=1e-9)
        assert_allclose(yhat2, yhat, rtol=0, atol=0)


@pytest.mark.parametrize("name", ["syn_t2_n2000_d4_m8_p64", "syn_t2_matern52_n1500_d3_m10_p48"])
def test_type2_analytic_gradient(name):
    """Analytic kernel-parameter gradient vs (a) Richardson central differences of the REFERENCE's LML (tight)
    and (b) the reference's own forward-difference gradient at its checkgrad tolerance."""
    g = load_golden(name)
    m = build_model(g)
    assert m.grad_method == 'adjoint'
    ll, grad = m.log_likelihood(return_gradient=True)
    assert_allclose(float(ll), float(g["lml"]), rtol=1e-9)
    free = g["free"]
    assert_array_equal(np.nonzero(~np.isnan(grad))[0], free)
    gc, gfd = g["grad_central"][free], g["grad_fd"][free]
    assert_allclose(grad[free], gc, rtol=1e-7, atol=1e-8 * np.abs(gc).max())   # Richardson error, not ours, limits this
    ratio = grad[free] / gfd
    big = np.abs(gfd) > 1e-2 * np.abs(gfd).max()
    assert_array_almost_equal(ratio[big], np.ones(big.sum()), decimal=3)
    # the reference's method, through our kernels
    m.grad_method = 'finite_difference'
    m._gradient = None
    _, gfd_ours = m.log_likelihood(return_gradient=True)
    assert_allclose(gfd_ours[free], gfd, rtol=1e-3, atol=1e-3 * np.abs(gfd).max())


def test_checkgrad_and_optimize_type2():
    g = load_golden("syn_t2_n2000_d4_m8_p64")
    m = build_model(g)
    assert m.checkgrad(decimal=3)
    ll0 = float(m.log_likelihood())
    opt = m.optimize(max_iters=5)
    assert opt is not None
    assert float(m.log_likelihood()) > ll0


def test_optimize_type1_weights():
    g = load_golden("syn_t1_n2000_d4_m8_p64")
    m = build_model(g)
    ll0 = float(m.log_likelihood())
    m.optimize(max_iters=10)
    assert float(m.log_likelihood()) > ll0
    assert np.all(m.kern.w > 0)


def test_parameter_cache_semantics():
    """Type-I keeps A across parameter changes, Type-II rebuilds it (reference gp_grief_model.py:53-68)."""
    g = load_golden("syn_t1_n2000_d4_m8_p64")
    m = build_model(g)
    m.fit()
    A0 = m._A.copy()
    p = m.parameters
    p[0] *= 1.5
    m.parameters = p
    assert m._P is None and m._A is not None
    assert_array_equal(m._A, A0)
    m2 = build_model(load_golden("syn_t2_n2000_d4_m8_p64"))
    m2.fit()
    p = m2.parameters
    p[2] *= 1.1
    m2.parameters = p
    assert m2._A is None


def test_error_conventions():
    gp = _pkg()
    g = load_golden("syn_t1_n2000_d4_m8_p64")
    m = build_model(g)
    with pytest.raises(ValueError):
        gp.models.GPGriefModel(g["x"], g["y"][:-1], m.kern)
    with pytest.raises(RuntimeError):
        gp.models.GPGriefModel(g["x"], np.hstack([g["y"], g["y"]]), m.kern)
    with pytest.raises(AssertionError):
        gp.kern.GriefKernel(m.kern.kern_list, "not a grid")
    with pytest.raises(NotImplementedError):
        m.kern.diag_val


def test_kron_matrix_host_methods():
    """tests/test_tensors/test_kron_matrix_sym.py + test_kron_eigenvalues.py flavoured checks of the host type."""
    gp = _pkg()
    np.random.seed(0)
    A = [np.random.rand(n, n) + np.eye(n) for n in (3, 4, 2)]
    A = [a.dot(a.T) + 1e-6 * np.eye(a.shape[0]) for a in A]
    K = gp.tensors.KronMatrix([a.copy() for a in A], sym=True)
    big = np.kron(np.kron(A[0], A[1]), A[2])
    assert_allclose(K.expand(), big, rtol=1e-12)
    x = np.random.rand(big.shape[0], 1)
    assert_allclose(K * x, big.dot(x), rtol=1e-10)
    assert_allclose(K.kronvec_div(x), np.linalg.solve(big, x), rtol=1e-8)
    U = K.chol()
    assert_allclose(U.solve_chol(x), np.linalg.solve(big, x), rtol=1e-8)
    Q, T = K.schur()
    assert_allclose(Q.solve_schur(T, x), np.linalg.solve(big, x), rtol=1e-8)
    assert_allclose(K.eig_vals().log_det(), np.linalg.slogdet(big)[1], rtol=1e-10)
    # top-p through the KronMatrix method (device) against the brute-force expansion
    eigs = gp.tensors.KronMatrix([np.random.rand(3) for _ in range(10)])
    loc, vals, gl = eigs.find_extremum_eigs(5, mode='largest', log_expand=True, sort=True, compute_global_loc=True)
    all_eigs = eigs.expand()
    assert_array_almost_equal(np.exp(vals), all_eigs[gl], decimal=15)
    assert_array_almost_equal(np.sort(np.exp(vals)), np.sort(all_eigs)[-5:], decimal=15)
    for mode in ('largest', 'smallest'):
        loc, vals, gl = eigs.find_extremum_eigs(5, mode=mode, log_expand=False, sort=True, compute_global_loc=True)
        ref = np.sort(all_eigs)[-5:] if mode == 'largest' else np.sort(all_eigs)[:5]
        assert_array_almost_equal(np.sort(vals), ref, decimal=15)


def test_gpweb_reference_test():
    """tests/test_models/test_gp_web_model.py:13-34 of the reference + its outputs on the same inputs (golden)."""
    gp = _pkg()
    g = load_golden("gpweb_n100_p4")
    m = gp.models.GPwebModel(Phi=g["Phi"], y=g["y"])
    m.parameters = g["parameters"].copy()
    ll = m.log_likelihood()
    w = m.kern.parameters.reshape((1, -1))
    K = g["Phi"].dot(np.diag(w.squeeze()).dot(g["Phi"].T)) + m.noise_var * np.identity(g["Phi"].shape[0])
    assert_array_almost_equal(ll, mvn.logpdf(g["y"].squeeze(), mean=np.zeros(g["Phi"].shape[0]), cov=K))
    assert_allclose(float(ll), float(g["lml"]), rtol=1e-9)
    assert m.checkgrad()
    _, grad = m.log_likelihood(return_gradient=True)
    assert_allclose(grad, g["grad"], rtol=1e-9)
    yhat, yvar = m.predict(g["Phi_new"])
    assert_allclose(yhat.squeeze(), g["yhat"], rtol=1e-9)
    assert_allclose(yvar, g["yvar"], rtol=1e-9)


def test_grief_to_web_model_shares_statistics():
    g = load_golden("syn_t1_n3000_d6_m10_p256_w")
    m = build_model(g)
    web = m.to_web_model()
    assert_allclose(float(web.log_likelihood()), float(g["lml"]), rtol=1e-9)
    _, gw = web.log_likelihood(return_gradient=True)
    ref = g["grad_adjoint"]
    assert_allclose(gw[0], ref[0], rtol=1e-9)
    assert_allclose(gw[1:], ref[-256:], rtol=1e-9, atol=1e-9 * np.abs(ref[-256:]).max())


@pytest.mark.gpu
def test_kron_matvec_on_device_matches_host():
    """Kronecker mat-vec through per-factor GEMMs on the device (reference tensors/kron_matrix.py:52-97)."""
    import gp_grief_b200 as gp
    rng = np.random.RandomState(4)
    shapes = [(7, 5), (3, 3), (20, 20), (4, 9)]
    K = gp.tensors.KronMatrix([rng.randn(*s) for s in shapes])
    x = rng.randn(int(K.shape[1]), 1)
    y_host = K.kronvec_prod(x)
    y_dev = K.kronvec_prod(x, device=True)
    assert y_dev.shape == y_host.shape
    np.testing.assert_allclose(y_dev, y_host, rtol=0, atol=1e-12 * np.abs(y_host).max())
    dense = np.kron(np.kron(np.kron(K.K[0], K.K[1]), K.K[2]), K.K[3])
    np.testing.assert_allclose(y_dev, dense.dot(x), rtol=0, atol=1e-11 * np.abs(y_host).max())


@pytest.mark.gpu
@pytest.mark.parametrize("kname", ["RBF", "Matern32", "Matern52", "Exponential"])
def test_d_yhat_d_x_on_device(kname):
    """input gradient of the predictive mean (reference models/gp_grief_model.py:127-134) -- device tables with
    one kernel replaced by its x-derivative, against the host evaluation of d Phi / d x and against central differences."""
    import gp_grief_b200 as gp
    rng = np.random.RandomState(11)
    n, d, m, p = 400, 3, 7, 40
    X = rng.rand(n, d)
    Y = np.sin(3 * X[:, :1]) + X[:, 1:2] ** 2 + 0.05 * rng.randn(n, 1)
    grid = gp.grid.InducingGrid(xg=[np.linspace(0, 1, m).reshape(-1, 1)] * d)
    K = getattr(gp.kern, kname)
    kern = gp.kern.GriefKernel([K(1, lengthscale=0.35 + 0.1 * i) for i in range(d)], grid, n_eigs=p, reweight_eig_funs=False)
    mdl = gp.models.GPGriefModel(X, Y, kern, noise_var=0.05)
    Xnew = 0.1 + 0.8 * rng.rand(25, d) + 0.0137                 # away from the grid points (Exponential has a kink there)
    for dim in range(d):
        g_dev = mdl.d_Yhat_d_x(Xnew, dim)
        assert g_dev.shape == (25, 1)
        g_host = mdl.kern.cov_grad(Xnew, dim).dot(mdl._alpha_p)
        np.testing.assert_allclose(g_dev, g_host.reshape(-1, 1), rtol=0, atol=1e-10 * max(1.0, np.abs(g_host).max()))
        h = 1e-6
        Xp, Xm = Xnew.copy(), Xnew.copy()
        Xp[:, dim] += h
        Xm[:, dim] -= h
        fd = (mdl.predict(Xp, compute_var=None) - mdl.predict(Xm, compute_var=None)) / (2 * h)
        np.testing.assert_allclose(g_dev, fd, rtol=0, atol=2e-5 * max(1.0, np.abs(fd).max()))


@pytest.mark.gpu
def test_rowcol_khatri_rao_matvec_on_device_reference_test():
    """tests/test_tensors/test_RowColKhatriRaoMatrix.py:9-48 of the reference with device=True: A x, A^T x and the
    RowColKhatriRaoMatrixTransposed form through grief_rowcol_kr_matvec."""
    import gp_grief_b200 as gp
    from gp_grief_b200.tensors import KronMatrix, KhatriRaoMatrix, RowColKhatriRaoMatrix, RowColKhatriRaoMatrixTransposed
    np.random.seed(0)
    N, p, d = 5, 6, 3
    grid_shape = np.random.randint(low=2, high=15, size=d)
    R = np.empty(d, dtype=object)
    R[:] = [np.random.rand(p, m) - 0.5 for m in grid_shape]
    K = np.empty(d, dtype=object)
    K[:] = [np.random.rand(m, m) - 0.5 for m in grid_shape]
    C = np.empty(d, dtype=object)
    C[:] = [np.random.rand(m, N) - 0.5 for m in grid_shape]
    for i in range(d):
        R[i][0, :] = 0.
    vec = np.random.rand(N, 1) - 0.5
    vecT = np.random.rand(p, 1) - 0.5
    A = RowColKhatriRaoMatrix(R=R, K=K, C=C, device=True)
    AT = RowColKhatriRaoMatrixTransposed(R=R, K=K, C=C, device=True)
    Rk, Ck, Kk = KhatriRaoMatrix(R, partition=0), KhatriRaoMatrix(C, partition=1), KronMatrix(K)
    assert_array_almost_equal(A * vec, Rk * (Kk * (Ck * vec)))
    assert_array_almost_equal(A.T * vecT, Ck.T * (Kk.T * (Rk.T * vecT)))
    assert_array_almost_equal(AT * vecT, Ck.T * (Kk.T * (Rk.T * vecT)))
    assert A.T.device and AT.T.device
    host = RowColKhatriRaoMatrix(R=R, K=K, C=C)
    assert_allclose(A * vec, host * vec, rtol=0, atol=1e-14)


@pytest.mark.gpu
def test_rowcol_khatri_rao_matvec_on_device_large_and_selection_rows():
    """Ragged sizes (rows not a multiple of 64, columns not a multiple of 64), dense and selection-matrix row factors mixed, the
    GRIEF use (kern/grief_kernel.py:96-104 through expand_SKC: R = selection of eigenvector rows, C = Q^T K_ux)."""
    import gp_grief_b200 as gp
    from gp_grief_b200.tensors import RowColKhatriRaoMatrix, RowColKhatriRaoMatrixTransposed, SelectionMatrixSparse
    rng = np.random.default_rng(12)
    rows, cols = 1000 + 37, 333
    ms = [5, 17, 1, 30]
    C = [rng.standard_normal((m, cols)) for m in ms]
    x = rng.standard_normal((cols, 1))
    R_dense = [rng.standard_normal((rows, m)) for m in ms]
    A = RowColKhatriRaoMatrix(R=R_dense, K=None, C=C, device=True)
    ref = np.prod([r.dot(c) for r, c in zip(R_dense, C)], axis=0).dot(x)
    assert_allclose(A * x, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    v = rng.standard_normal((rows, 1))
    refT = np.prod([r.dot(c) for r, c in zip(R_dense, C)], axis=0).T.dot(v)
    assert_allclose(A.T * v, refT, rtol=0, atol=1e-12 * np.abs(refT).max())
    # selection-matrix rows (index vectors on the device), mixed with one dense factor
    sel = [SelectionMatrixSparse((rng.integers(0, m, size=rows), m)) for m in ms]
    R_mix = [sel[0], R_dense[1], sel[2], sel[3]]
    G = np.prod([C[t][sel[t].indicies, :] if t != 1 else R_dense[1].dot(C[1]) for t in range(4)], axis=0)
    Am = RowColKhatriRaoMatrix(R=R_mix, K=None, C=C, device=True)
    assert_allclose(Am * x, G.dot(x), rtol=0, atol=1e-12 * np.abs(G.dot(x)).max())
    R_sel = sel
    Gs = np.prod([C[t][sel[t].indicies, :] for t in range(4)], axis=0)
    As = RowColKhatriRaoMatrix(R=R_sel, K=None, C=C, device=True)
    assert_allclose(As * x, Gs.dot(x), rtol=0, atol=1e-12 * np.abs(Gs.dot(x)).max())
    assert isinstance(As.T, RowColKhatriRaoMatrixTransposed)
    assert_allclose(As.T * v, Gs.T.dot(v), rtol=0, atol=1e-12 * np.abs(Gs.T.dot(v)).max())
