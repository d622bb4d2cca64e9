"""GPU parity tests of the C-ABI kernels against the CPU oracle and the reference goldens."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import load_golden, model_cases, case_inputs
from oracle import grief_oracle as orc
from gp_grief_b200 import _native as nat

pytestmark = pytest.mark.gpu


def _dev():
    from gp_grief_b200 import device
    return device


def _plan_from_basis(c, basis, width_cap=0):
    """Oracle basis (KronMatrix order) -> DevicePlan (input order)."""
    d = c["d"]
    Q = [basis.Q[d - 1 - i] for i in range(d)]
    eig = [basis.eig[d - 1 - i] for i in range(d)]
    loc = basis.eig_loc[:, ::-1]
    return _dev().DevicePlan(c["names"], c["variances"], c["lengthscales"], c["xg"], Q, eig, loc, width_cap=width_cap)


@pytest.mark.parametrize("name", ["topk_c2_d6_m10_p1024", "topk_c3_d10_m20_p4096",
                                  "topk_c4_d32_m8_p2048", "topk_c5_d8_m16_p8192"])
def test_topk_bit_exact_at_bench_configs(name):
    g = load_golden(name)
    d, p = int(g["d"]), int(g["p"])
    eigs = [g["eigs_%d" % k] for k in range(d)]
    loc, lam = _dev().topk_kron(eigs, p)
    assert_array_equal(loc, g["eig_loc"].astype(np.int64))      # bit-exact indices (tie-free configs)
    assert_array_equal(lam, g["log_lam"])                      # bit-exact values


def test_topk_reference_kron_eigenvalue_test():
    """tests/test_tensors/test_kron_eigenvalues.py:64-78 (log_expand=True, largest) through the device path."""
    g = load_golden("kron_eigs_d10_m3_p5")
    eigs = [g["eigs_%d" % i] for i in range(10)]
    loc, lam = _dev().topk_kron(eigs, 5)
    assert_array_equal(loc, g["loc_largest_log"])
    assert_array_equal(lam, g["vals_largest_log"])
    assert_allclose(np.sort(np.exp(lam)), np.sort(g["all_sorted_top"]), rtol=0, atol=1e-15)


@pytest.mark.parametrize("d,m,p,seed", [(1, 7, 5, 0), (1, 5, 5, 1), (3, 4, 64, 2), (3, 4, 63, 3), (6, 1, 1, 4),
                                        (5, 3, 200, 5), (4, 9, 1000, 6), (2, 60, 3000, 7), (12, 2, 16384 // 4, 8)])
def test_topk_random_vs_oracle(d, m, p, seed):
    rng = np.random.default_rng(seed)
    eigs = [rng.random(m) + 1e-3 for _ in range(d)]
    p = int(min(p, float(m) ** d))
    loc_o, lam_o = orc.find_extremum_eigs(eigs, p)
    loc, lam = _dev().topk_kron(eigs, p)
    assert_array_equal(lam, lam_o)
    if np.all(np.diff(lam_o) != 0):
        assert_array_equal(loc, loc_o)
    # every returned tuple reproduces its value
    chk = np.sum([np.log(eigs[k])[loc[:, k]] for k in range(d)], axis=0)
    assert_allclose(chk, lam, rtol=1e-13, atol=1e-13)


def test_topk_with_ties_is_a_valid_top_set():
    """Aliased kernels => exact ties: multiset of values must match, indices must be valid/distinct."""
    rng = np.random.default_rng(11)
    e = rng.random(6) + 0.1
    eigs = [e.copy() for _ in range(5)]
    p = 300
    loc_o, lam_o = orc.find_extremum_eigs(eigs, p)
    loc, lam = _dev().topk_kron(eigs, p)
    assert_allclose(np.sort(lam), np.sort(lam_o), rtol=0, atol=1e-13)
    assert len({tuple(r) for r in loc}) == p


@pytest.mark.parametrize("name", model_cases())
@pytest.mark.parametrize("width_cap", [0, 24])
def test_phi_and_gram_match_reference(name, width_cap):
    import torch
    g = load_golden(name)
    c = case_inputs(g)
    basis = orc.setup_inducing_cov(c["names"], c["variances"], c["lengthscales"], c["xg"], c["n_eigs"])
    try:
        plan = _plan_from_basis(c, basis, width_cap)
    except NotImplementedError:
        pytest.skip("table does not fit width_cap=%d" % width_cap)
    Phi_o = orc.grief_phi(basis, c["names"], c["variances"], c["lengthscales"], c["xg"], c["x"])
    n = c["x"].shape[0]
    X = torch.from_numpy(np.ascontiguousarray(c["x"])).cuda()
    T = plan.build_tables(X)
    Phi = plan.phi_rows(T, n).cpu().numpy()
    scale = np.abs(Phi_o).max()
    assert_allclose(Phi, Phi_o, rtol=1e-11, atol=1e-13 * scale)
    A = plan.gram(T, n).cpu().numpy()
    A_o = Phi_o.T.dot(Phi_o)
    assert_allclose(A, A_o, rtol=0, atol=1e-12 * np.abs(A_o).max())
    assert_array_equal(A, A.T)
    y = torch.from_numpy(np.ascontiguousarray(c["y"].reshape(-1))).cuda()
    r = plan.phi_t_vec(T, n, y).cpu().numpy()
    r_scale = float(np.abs(Phi_o).T.dot(np.abs(c["y"])).max())
    assert_allclose(r, Phi_o.T.dot(c["y"]).reshape(-1), rtol=0, atol=1e-13 * r_scale)
    v = np.random.default_rng(0).standard_normal(plan.p)
    pv = plan.phi_vec(T, n, torch.from_numpy(v).cuda()).cpu().numpy()
    assert_allclose(pv, Phi_o.dot(v), rtol=0, atol=1e-11 * np.abs(Phi_o.dot(v)).max())
    s = _dev().sumsq(y).cpu().numpy()[0]
    assert_allclose(s, float((c["y"] ** 2).sum()), rtol=1e-14)


@pytest.mark.parametrize("name", model_cases())
def test_lml_and_adjoint_gradient_match_reference(name):
    """LML, d/dw, d/dnoise within 1e-9 relative of the reference (BASELINE north_star tolerance)."""
    import torch
    g = load_golden(name)
    c = case_inputs(g)
    basis = orc.setup_inducing_cov(c["names"], c["variances"], c["lengthscales"], c["xg"], c["n_eigs"])
    plan = _plan_from_basis(c, basis)
    n = c["x"].shape[0]
    X = torch.from_numpy(np.ascontiguousarray(c["x"])).cuda()
    y = torch.from_numpy(np.ascontiguousarray(c["y"].reshape(-1))).cuda()
    T = plan.build_tables(X)
    A = plan.gram(T, n)
    r = plan.phi_t_vec(T, n, y)
    s = _dev().sumsq(y)
    w = torch.from_numpy(np.ascontiguousarray(c["w"])).cuda()
    out = _dev().DeviceSolver().solve(A, r, s, w, c["noise_var"], n, want_grad=True, want_G2=True)
    assert_allclose(out["lml"], float(g["lml"]), rtol=1e-9)
    assert_allclose(out["logdet"], float(g["log_det"]), rtol=1e-9)
    if "grad_adjoint" in g:
        ref = g["grad_adjoint"]
        assert_allclose(out["grad_noise"], ref[0], rtol=1e-9, atol=1e-9 * abs(float(g["lml"])))
        gw = out["grad_w"].cpu().numpy()
        rw = ref[-c["n_eigs"]:]
        assert_allclose(gw, rw, rtol=1e-9, atol=1e-9 * np.abs(rw).max())
    # alpha_p == b  (models/gp_grief_model.py:97) -> predictive mean
    if "xnew" in g:
        Xn = torch.from_numpy(np.ascontiguousarray(g["xnew"])).cuda()
        Tn = plan.build_tables(Xn)
        mean = plan.phi_vec(Tn, Xn.shape[0], out["b"]).cpu().numpy()
        assert_allclose(mean, g["yhat"], rtol=1e-9, atol=1e-9 * np.abs(g["yhat"]).max())


def test_not_positive_definite_raises_linalgerror():
    import torch
    p = 8
    A = -torch.eye(p, dtype=torch.float64, device="cuda")
    r = torch.ones(p, dtype=torch.float64, device="cuda")
    s = torch.ones(1, dtype=torch.float64, device="cuda")
    w = torch.ones(p, dtype=torch.float64, device="cuda")
    with pytest.raises(np.linalg.LinAlgError):
        _dev().DeviceSolver().solve(A, r, s, w, 0.5, 100)


def test_gram_larger_random_shape_vs_materialised_phi():
    """n not a multiple of the chunk, p not a multiple of the tile, several row splits."""
    import torch
    rng = np.random.default_rng(5)
    d, m, p, n = 5, 6, 300, 70001
    xg = [np.linspace(0, 1, m) for _ in range(d)]
    names = ["RBF"] * d
    var, ls = [1.0] * d, [0.3 + 0.07 * i for i in range(d)]
    basis = orc.setup_inducing_cov(names, var, ls, xg, p)
    c = dict(d=d, names=names, variances=var, lengthscales=ls, xg=xg)
    plan = _plan_from_basis(c, basis)
    X = torch.from_numpy(rng.random((n, d))).cuda()
    T = plan.build_tables(X)
    Phi = plan.phi_rows(T, n)
    A = plan.gram(T, n)
    A_ref = Phi.T @ Phi
    assert_allclose(A.cpu().numpy(), A_ref.cpu().numpy(), rtol=0, atol=1e-12 * float(A_ref.abs().max()))
    Phi_o = orc.grief_phi(basis, names, var, ls, xg, X[:2000].cpu().numpy())
    assert_allclose(Phi[:2000].cpu().numpy(), Phi_o, rtol=1e-11, atol=1e-13 * np.abs(Phi_o).max())


@pytest.mark.gpu
@pytest.mark.parametrize("p", [1, 5, 127, 128, 129, 300, 640, 1000])
def test_dense_stage_against_scipy(p):
    """The hand-written Cholesky / solve / inverse (dense.cu) against scipy (reference gp_grief_model.py:152-153,171-191)."""
    import scipy.linalg as sl
    import torch
    rng = np.random.RandomState(100 + p)
    n = 4 * p + 7
    Phi = rng.randn(n, p) / np.sqrt(n)
    y = rng.randn(n)
    w = rng.rand(p) + 0.5
    noise = 0.3
    A = Phi.T.dot(Phi)
    r = Phi.T.dot(y)
    P = A + np.diag(noise / w)
    U = sl.cho_factor(P)[0]
    b_ref = sl.cho_solve((np.triu(U), False), r)
    Pinv_ref = np.linalg.inv(P)
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = _dev().DeviceSolver().solve(dv(A), dv(r), dv(np.array([y.dot(y)])), dv(w), noise, n, want_grad=True, want_G2=True)
    L = out["L"].cpu().numpy()
    assert np.all(np.tril(L, -1) == 0.0)
    np.testing.assert_allclose(L, np.triu(U), rtol=1e-10, atol=1e-12 * np.abs(U).max())
    np.testing.assert_allclose(out["b"].cpu().numpy(), b_ref, rtol=1e-9, atol=1e-11 * np.abs(b_ref).max())
    Pinv = out["Pinv"].cpu().numpy()
    assert np.array_equal(Pinv, Pinv.T)
    np.testing.assert_allclose(Pinv, Pinv_ref, rtol=1e-8, atol=1e-10 * np.abs(Pinv_ref).max())
    logdet = 2.0 * np.log(np.diag(U)).sum()
    logdet_full = logdet + np.log(w).sum() + (n - p) * np.log(noise)       # log|Phi W Phi^T + noise I|
    assert abs(out["logdet"] - logdet_full) <= 1e-10 * max(1.0, abs(logdet_full))
    G2_ref = -(Pinv_ref + np.outer(b_ref, b_ref) / noise)
    np.testing.assert_allclose(out["G2"].cpu().numpy(), G2_ref, rtol=1e-8, atol=1e-10 * np.abs(G2_ref).max())


@pytest.mark.gpu
def test_dense_stage_not_pd_reports_minor():
    import torch
    p = 200
    A = np.eye(p)
    A[150, 150] = -5.0
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    with pytest.raises(np.linalg.LinAlgError, match="151"):
        _dev().DeviceSolver().solve(dv(A), dv(np.ones(p)), dv(np.array([1.0])), dv(np.full(p, 1e30)), 0.5, 100)


@pytest.mark.gpu
def test_gram_and_quadform_over_several_slabs():
    """Pass 1 with a tiny Phi^T budget (many slabs accumulated into the split partials) and the Z = Phi B product
    over more rows than one slab (148 * 256) must match the materialised Phi."""
    import torch
    rng = np.random.default_rng(11)
    d, m, p, n = 4, 5, 200, 40003
    xg = [np.linspace(0, 1, m) for _ in range(d)]
    names = ["Matern52"] * d
    var, ls = [1.0] * d, [0.4 + 0.05 * i for i in range(d)]
    basis = orc.setup_inducing_cov(names, var, ls, xg, p)
    plan = _plan_from_basis(dict(d=d, names=names, variances=var, lengthscales=ls, xg=xg), basis)
    X = torch.from_numpy(rng.random((n, d))).cuda()
    T = plan.build_tables(X)
    Phi = plan.phi_rows(T, n)
    A_ref = (Phi.T @ Phi).cpu().numpy()
    A_one = plan.gram(T, n).cpu().numpy()
    plan.set_option(nat.OPT_SLAB_BUDGET, 256 * 8 * 3000)        # 3000-row slabs (rounded down to 2944) -> 14 slabs
    A_many = plan.gram(T, n).cpu().numpy()
    plan.set_option(nat.OPT_SLAB_BUDGET, 0)
    tol = 1e-12 * np.abs(A_ref).max()
    assert_allclose(A_one, A_ref, rtol=0, atol=tol)
    assert_allclose(A_many, A_ref, rtol=0, atol=tol)
    assert np.array_equal(A_many, A_many.T)
    B = rng.standard_normal((p, p))
    B = torch.from_numpy(B + B.T).cuda()
    q = plan.quadform_rows(T, n, B).cpu().numpy()
    q_ref = ((Phi @ B) * Phi).sum(1).cpu().numpy()
    assert_allclose(q, q_ref, rtol=0, atol=1e-11 * np.abs(q_ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(128, 128, 2), (256, 384, 50), (1280, 256, 1000)])
def test_gemm_nt_entry(M, N, K):
    import torch
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((M, K), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((N, K), dtype=torch.float64, device="cuda", generator=g)
    C0 = torch.randn((M, N), dtype=torch.float64, device="cuda", generator=g)
    C = C0.clone()
    nat.check(nat.lib().grief_gemm_nt(nat.dev_ptr(A), K, nat.dev_ptr(B), K, nat.dev_ptr(C), N, M, N, K, -0.5, 2.0,
                                      nat.stream_ptr()))
    ref = 2.0 * C0 - 0.5 * (A @ B.T)
    assert_allclose(C.cpu().numpy(), ref.cpu().numpy(), rtol=0, atol=1e-12 * float(ref.abs().max()) * max(1, K) ** 0.5)
    with pytest.raises(ValueError):                        # lda < K
        nat.check(nat.lib().grief_gemm_nt(nat.dev_ptr(A), K - 1, nat.dev_ptr(B), K, nat.dev_ptr(C), N, M, N, K, 1.0, 0.0,
                                          nat.stream_ptr()))


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 3, 7), (130, 129, 33), (300, 1, 257)])
def test_gemm_nt_helper_ragged_shapes(M, N, K):
    """Partial tiles, odd K and odd leading dimensions go through padded operand copies and guarded stores."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(7 * M + N + K)
    A = torch.randn((M, K), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((N, K), dtype=torch.float64, device="cuda", generator=g)
    C = torch.full((M + 2, N + 3), 7.0, dtype=torch.float64, device="cuda")
    out = _dev().gemm_nt(A, B, alpha=2.0, beta=1.0, out=C[:M, :N])
    ref = 7.0 + 2.0 * (A @ B.T)
    assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), rtol=0, atol=1e-12 * max(1.0, float(ref.abs().max())))
    assert float(C[M:].min()) == 7.0 and float(C[:, N:].min()) == 7.0 and float(C[M:].max()) == 7.0      # nothing outside written
