"""CPU-only tests: host logic of the package, the C-ABI export surface, and the no-fallback guarantee."""
import ctypes
import os
import re

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import load_golden, case_inputs, ROOT
from oracle import grief_oracle as orc

import gp_grief_b200 as gp
from gp_grief_b200 import _native


def test_library_exports_every_symbol_of_the_header():
    """The shared library loads without a GPU and exports exactly what include/grief_b200.h declares."""
    assert os.path.exists(_native.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_native.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "grief_b200.h")).read()
    declared = set(re.findall(r"\b(grief_[a-z0-9_]+)\s*\(", header))
    declared -= {"grief_plan", "grief_ctx"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), "symbol %s declared in the header but not exported" % name
    assert declared == set(_native.SIGNATURES), (declared ^ set(_native.SIGNATURES))
    assert lib.grief_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = load_golden("syn_t1_n2000_d4_m8_p64")
    d = 4
    grid = gp.grid.InducingGrid(xg=[g["xg_%d" % i].reshape(-1, 1) for i in range(d)])
    kern = gp.kern.GriefKernel([gp.kern.RBF(1, lengthscale=0.3 + 0.05 * i) for i in range(d)], grid, n_eigs=64)
    with pytest.raises(RuntimeError):
        gp.models.GPGriefModel(g["x"], g["y"], kern, noise_var=0.1)
    with pytest.raises(RuntimeError):
        kern.cov(g["x"][:5])
    assert not any("oracle" in m for m in __import__("sys").modules if m.startswith("gp_grief_b200"))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gp_grief_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)


@pytest.mark.parametrize("name", ["RBF", "Exponential", "Matern32", "Matern52"])
def test_kernels_match_oracle_and_derivatives(name):
    rng = np.random.default_rng(0)
    x, z = rng.random((7, 1)), rng.random((5, 1))
    k = getattr(gp.kern, name)(1, variance=1.3, lengthscale=0.45)
    assert_array_equal(k.cov(x, z), orc.kernel_cov(name, x[:, 0], z[:, 0], 1.3, 0.45))
    h = 1e-6
    kp = getattr(gp.kern, name)(1, variance=1.3, lengthscale=0.45 + h)
    km = getattr(gp.kern, name)(1, variance=1.3, lengthscale=0.45 - h)
    assert_allclose(k.grad_lengthscale(x, z), (kp.cov(x, z) - km.cov(x, z)) / (2 * h), rtol=1e-6, atol=1e-8)
    assert_allclose(k.grad_variance(x, z), k.cov(x, z) / 1.3, rtol=1e-14)
    assert_allclose(k.grad_x(x, z), (k.cov(x + h, z) - k.cov(x - h, z)) / (2 * h), rtol=1e-5, atol=1e-7)
    assert_array_equal(k.parameters, [1.3, 0.45])
    assert list(k.constraints) == ['+ve', '+ve']


def test_grid_kernel_factors_bit_identical_to_reference():
    """cov_grid + schur reproduce the reference's Schur vectors bit for bit (inputs of the top-p selection)."""
    g = load_golden("syn_t1_ragged_n700_d5_p40")
    c = case_inputs(g)
    d = c["d"]
    grid = gp.grid.InducingGrid(xg=[x.reshape(-1, 1) for x in c["xg"]])
    assert_array_equal(grid.grid_shape, [7, 1, 2, 12, 5])
    gk = gp.kern.GridKernel([gp.kern.RBF(1, variance=c["variances"][i], lengthscale=c["lengthscales"][i]) for i in range(d)])
    Kuu = gk.cov_grid(grid.xg, dim_noise_var=1e-12)
    Q, T = Kuu.schur()
    for k in range(d):
        assert_array_equal(Q.K[k], g["Q_%d" % k])
    assert list(gk.constraints) == ['+ve', '+ve'] + ['fixed', '+ve'] * (d - 1)


def test_inducing_grid_from_data():
    np.random.seed(0)
    x = np.random.rand(100, 3)
    x[:, 2] = np.round(x[:, 2] * 3) / 3
    grid = gp.grid.InducingGrid(x)
    assert_array_equal(grid.grid_shape, [10, 10, 4])
    assert grid.num_data == 400.0 and grid.input_dim == 3
    assert_allclose(grid.xg[0][:, 0], np.linspace(x[:, 0].min(), x[:, 0].max(), 10))
    assert_array_equal(grid.xg[2][:, 0], np.unique(x[:, 2]))
    g2 = gp.grid.InducingGrid(x, mbar=5, beyond_domain=0.1)
    assert_array_equal(g2.grid_shape, [5, 5, 4])
    assert g2.xg[0][0, 0] < x[:, 0].min() and g2.xg[0][-1, 0] > x[:, 0].max()
    assert gp.grid.grid2mat(np.arange(2), np.arange(3)).shape == (6, 2)


def test_tensor_types_host():
    """tests/test_tensors/* of the reference, restated: KR mat-vecs, logged expansion with zeros, selections."""
    np.random.seed(0)
    T = gp.tensors
    A = [np.random.rand(5, 3), np.random.rand(5, 2), np.random.rand(5, 4)]
    KR = T.KhatriRaoMatrix(A, partition=0)
    dense = np.vstack([np.kron(np.kron(A[0][i], A[1][i]), A[2][i]) for i in range(5)])
    assert_allclose(KR.expand(), dense)
    x = np.random.rand(dense.shape[1], 1)
    assert_allclose(KR * x, dense.dot(x))
    assert_allclose(KR.T * np.ones((5, 1)), dense.T.dot(np.ones((5, 1))))
    R = [np.random.rand(6, 3) for _ in range(2)]
    K = [np.random.rand(3, 3) for _ in range(2)]
    C = [np.random.rand(3, 4) for _ in range(2)]
    R[0][1, :] = 0.
    rkc = T.RowColKhatriRaoMatrix(R, K, C)
    full = (R[0].dot(K[0]).dot(C[0])) * (R[1].dot(K[1]).dot(C[1]))
    assert_allclose(rkc.expand(), full)
    lg, sg = rkc.expand(logged=True)
    assert_allclose(sg * np.exp(lg), full, atol=1e-15)
    v = np.random.rand(4, 1)
    assert_allclose(rkc * v, full.dot(v))
    assert_allclose(rkc.T * np.ones((6, 1)), full.T.dot(np.ones((6, 1))))
    idx = np.array([2, 0, 2, 1])
    S = T.SelectionMatrixSparse((idx, 3))
    M = np.random.rand(3, 5)
    assert_array_equal(S.mul(M), M[idx])
    assert_array_equal(S.mul_unique(M)[S.unique_inverse], M[idx])
    assert_array_equal(T.SelectionMatrix((idx, 3)).mul(M), M[idx])
    lp, sn = T.expand_SKC([S], [M.T[:3, :3]], [np.random.rand(3, 6)])
    assert lp.shape == (4, 6)
    assert_allclose(gp.linalg.log_kron(np.array([1., 2.]), np.array([3., 4., 5.])), np.log(np.kron([1., 2.], [3., 4., 5.])))


def test_grief_kernel_parameter_plumbing():
    g = load_golden("syn_t2_n2000_d4_m8_p64")
    d = 4
    grid = gp.grid.InducingGrid(xg=[g["xg_%d" % i].reshape(-1, 1) for i in range(d)])
    kl = [gp.kern.RBF(1, lengthscale=0.3 + 0.05 * i) for i in range(d)]
    kern = gp.kern.GriefKernel(kl, grid, n_eigs=64, reweight_eig_funs=False, opt_kernel_params=True)
    assert_array_equal(kern.parameters, g["parameters"][1:])
    assert list(kern.constraints) == [str(c) for c in g["constraints"][1:]]
    p2 = kern.parameters
    p2[1] = 0.77
    kern.parameters = p2
    assert kl[0].lengthscale == 0.77
    assert kern.base_parameter_map()[:3] == [(0, 'variance'), (0, 'lengthscale'), (1, 'variance')]
    assert not kern.has_aliased_kernels()
    k1 = gp.kern.GriefKernel([gp.kern.RBF(1), ] * 3, gp.grid.InducingGrid(xg=[np.linspace(0, 1, 4).reshape(-1, 1)] * 3), n_eigs=1000)
    assert k1.n_eigs == 64 and k1.has_aliased_kernels()
    assert list(k1.constraints[:6]) == ['fixed'] * 6 and list(k1.constraints[6:8]) == ['+ve'] * 2


def test_logexp_transform_roundtrip():
    t = gp.linalg.LogexpTransformation()
    f = np.array([1e-3, 0.5, 3.0, 40.0])
    assert_allclose(t.inverse_transform(t.transform(f)), f, rtol=1e-12)
    h = 1e-6
    assert_allclose(t.transform_grad(f, np.ones(4)), (t.inverse_transform(t.transform(f) + h) - f) / h, rtol=1e-4)


def test_sharding_and_layout():
    from gp_grief_b200.sharding import row_shard, stats_layout
    for n, w in [(10, 3), (7, 8), (10_000_000, 8), (5, 1)]:
        cover = []
        for r in range(w):
            a, b = row_shard(n, w, r)
            assert 0 <= a <= b <= n
            cover += list(range(a, b)) if n < 100 else []
        if n < 100:
            assert cover == list(range(n))
    lay = stats_layout(4)
    assert lay["A"] == (0, 16) and lay["r"] == (16, 20) and lay["s"] == (20, 21) and lay["size"] == 21


def test_int8_digit_scheme_is_double_precision_class():
    """NumPy restatement of the arithmetic of csrc/ozaki.cu + phi_stage.cu:store_digits (no GPU involved).

    Each operand row is scaled by 2^-e (e = frexp exponent of the row maximum), truncated to 54 bits and cut into 7 balanced
    base-256 digits by the carry-free bias trick; A B^T is rebuilt from the exact integer products of the digit planes with
    a + b <= 6.  The result must agree with the float64 product to a few ulps of the row-maxima product times K.
    """
    rng = np.random.RandomState(7)
    M, N, K = 24, 40, 300
    A = (rng.rand(M, K) - 0.5) * np.exp(4 * (rng.rand(M, K) - 0.5))
    B = (rng.rand(N, K) - 0.5) * np.exp(4 * (rng.rand(N, K) - 0.5))
    A[3] = 0.0                                        # an all-zero row keeps exponent 0 and all-zero digits

    def digits(X):
        amax = np.abs(X).max(axis=1)
        e = np.where(amax > 0, np.frexp(amax)[1], 0).astype(np.int64)
        q = np.trunc(np.ldexp(X, (54 - e)[:, None])).astype(np.int64)               # |q| < 2^54, exact
        bias = np.int64(0x0080808080808080)
        w = (q + bias) ^ bias                                                       # byte k of w, read as int8, is the digit of 256^k
        d = np.stack([((w >> np.int64(8 * (6 - s))) & np.int64(0xFF)).astype(np.uint8).view(np.int8).astype(np.int64)
                      for s in range(7)])                                           # digit 0 = most significant
        recon = sum(d[s] * np.int64(256) ** (6 - s) for s in range(7))
        assert np.array_equal(recon, q)                                             # the digits represent q exactly
        assert d.min() >= -128 and d.max() <= 127
        return d, e

    dA, eA = digits(A)
    dB, eB = digits(B)
    C = np.zeros((M, N))
    for g in range(6, -1, -1):                                                      # least significant group first, as the drain does
        acc = np.zeros((M, N), dtype=np.int64)
        for a in range(g + 1):
            acc += dA[a].dot(dB[g - a].T)                                           # exact integer GEMM (int32 on the tensor cores)
        assert np.abs(acc).max() < 2 ** 31
        C += np.ldexp(acc.astype(np.float64), -12 - 8 * g)
    C = np.ldexp(C, eA[:, None] + eB[None, :])
    ref = A.dot(B.T)
    scale = np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=1)[None, :] * K
    scale[scale == 0] = 1.0
    assert np.abs(C - ref).max() / np.abs(ref).max() < 5e-15
    assert (np.abs(C - ref) / scale).max() < 2.0 ** -52
    assert np.all(C[3] == 0.0)


def test_rowcol_khatri_rao_device_factors_are_plain_products():
    """Host side of RowColKhatriRaoMatrix(device=True): the factor lists handed to grief_rowcol_kr_matvec must describe the same
    matrix as expand() -- selection matrices as index vectors, the transposed form as C^T / R^T (no GPU involved)."""
    from gp_grief_b200.tensors import RowColKhatriRaoMatrix, RowColKhatriRaoMatrixTransposed, SelectionMatrixSparse
    rng = np.random.default_rng(5)
    rows, cols, ms = 11, 7, [3, 5, 2]
    C = [rng.standard_normal((m, cols)) for m in ms]
    K = [rng.standard_normal((m, m)) for m in ms]
    Rd = [rng.standard_normal((rows, m)) for m in ms]
    sel = [SelectionMatrixSparse((rng.integers(0, m, size=rows), m)) for m in ms]

    def dense_from(R, Cf):
        out = 1.0
        for r, c in zip(R, Cf):
            r = np.asarray(r)
            out = out * (c[r, :] if r.ndim == 1 else r.dot(c))
        return out

    A = RowColKhatriRaoMatrix(R=Rd, K=K, C=C, device=True)
    R, Cf = A._device_factors()
    np.testing.assert_allclose(dense_from(R, Cf), A.expand(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(A.expand(), np.prod([r.dot(k.dot(c)) for r, k, c in zip(Rd, K, C)], axis=0), rtol=0, atol=1e-13)
    As = RowColKhatriRaoMatrix(R=sel, K=None, C=C, device=True)
    R, Cf = As._device_factors()
    assert all(np.asarray(r).ndim == 1 for r in R)
    G = np.prod([C[t][sel[t].indicies, :] for t in range(3)], axis=0)
    np.testing.assert_allclose(dense_from(R, Cf), G, rtol=0, atol=0)
    AT = As.T
    assert isinstance(AT, RowColKhatriRaoMatrixTransposed) and AT.device and tuple(AT.shape) == (cols, rows)
    R, Cf = AT._device_factors()
    np.testing.assert_allclose(dense_from(R, Cf), G.T, rtol=0, atol=0)
    v = rng.standard_normal((rows, 1))
    host = RowColKhatriRaoMatrixTransposed(R=sel, K=None, C=C)
    np.testing.assert_allclose(host * v, G.T.dot(v), rtol=0, atol=1e-13)


def test_gpy_kernel_plumbing_without_gpy(monkeypatch):
    """GPyKernel (kern/gpy_kernel.py of the reference): constructor forms, cov with children, parameter / constraint vectors,
    fix_variance -- against a stand-in for the GPy module; without GPy the constructor raises ImportError (no GPU involved)."""
    import sys
    import types
    import gp_grief_b200 as gp
    monkeypatch.setitem(sys.modules, "GPy", None)
    with pytest.raises(ImportError):
        gp.kern.GPyKernel(1, kernel="RBF")

    class Param(np.ndarray):
        def __new__(cls, value, name):
            obj = np.asarray([value], dtype=float).view(cls)
            obj._name = name
            return obj

        def __array_finalize__(self, obj):
            self._name = getattr(obj, "_name", None)

        @property
        def values(self):
            return np.asarray(self)

    GPy, kmod = types.ModuleType("GPy"), types.ModuleType("GPy.kern")

    class Kern(object):
        pass

    class RBF(Kern):
        def __init__(self, input_dim, variance=1., lengthscale=1.):
            self.variance, self.lengthscale = Param(variance, "variance"), Param(lengthscale, "lengthscale")
            self.flattened_parameters = [self.variance, self.lengthscale]

        def K(self, x, z=None):
            z = x if z is None else z
            return float(self.variance[0]) * np.exp(-0.5 * (x - z.T) ** 2 / float(self.lengthscale[0]) ** 2)

    kmod.Kern, kmod.RBF = Kern, RBF
    GPy.kern = kmod
    monkeypatch.setitem(sys.modules, "GPy", GPy)
    monkeypatch.setitem(sys.modules, "GPy.kern", kmod)
    k = gp.kern.GPyKernel(1, kernel="RBF", lengthscale=0.4)
    own = gp.kern.RBF(1, lengthscale=0.4)
    x = np.linspace(0, 1, 6).reshape(-1, 1)
    np.testing.assert_allclose(k.cov(x), own.cov(x), rtol=1e-15)
    np.testing.assert_array_equal(k.parameters, own.parameters)
    assert list(k.constraints) == ['+ve', '+ve'] and k.device_id is None and k.name == "GPy - RBF"
    k.parameters = np.array([2.0, 0.7])
    assert float(k.kern.variance[0]) == 2.0 and float(k.kern.lengthscale[0]) == 0.7
    k.fix_variance()
    assert list(k.constraints) == ['fixed', '+ve']
    both = k + gp.kern.RBF(1, lengthscale=0.2)                 # a child kernel is folded into cov and the parameter vector
    np.testing.assert_allclose(both.cov(x), k.cov(x) + gp.kern.RBF(1, lengthscale=0.2).cov(x), rtol=1e-15)
    assert both.parameters.size == 4 and both.constraints.size == 4
    with pytest.raises(TypeError):
        gp.kern.GPyKernel(1, kernel=None)
