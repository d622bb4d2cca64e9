"""CPU oracle for the GP-GRIEF hot path: a NumPy restatement of the reference's algorithm.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module.  The product
(`gp_grief_b200`) never imports it and has no CPU fallback.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function below against the
fixtures in `tests/golden/`, which were produced by running the unmodified reference
(`/root/reference/gp_grief`, float64 NumPy/SciPy, in-house kernels) through
`oracle/gen_golden.py`.  Those fixtures include the inputs and the reference's outputs of its own
hot-path tests (tests/test_tensors/test_kron_eigenvalues.py:12-92,
tests/test_models/test_gp_grief_model.py:14-40 restated with the in-house RBF,
tests/test_models/test_gp_web_model.py:13-34) and BASELINE config C1 (automobile).
Third-party arithmetic outside /root/reference: GPy (un-pinned, requirements.txt:6) is only reached
through `GPyKernel` (kern/gpy_kernel.py:56); it is absent here, so parity for GPy-backed kernels is
unpinned and out of scope.  NumPy/SciPy (LAPACK dgees via scipy.linalg.schur, dpotrf/dpotrs) are
called directly, exactly like the reference does.

All citations are relative to /root/reference/gp_grief/.
"""
import numpy as np
import scipy.linalg as la
from scipy.linalg import cho_factor, cho_solve

# --------------------------------------------------------------------------------------------
# 1-D stationary kernels  (kern/stationary.py)
# --------------------------------------------------------------------------------------------

def _dist2(x, z):
    """Squared distances of two column vectors, kern/stationary.py:12-43 (one active dim, l=1)."""
    x = np.asarray(x, float).reshape(-1, 1, 1)
    z = np.asarray(z, float).reshape(1, -1, 1)
    return np.sum(np.power((x - z) / np.ones(1).reshape((1, 1, 1)), 2), axis=2, keepdims=False)


def kernel_cov(name, x, z, variance, lengthscale):
    """k(x, z) for the in-house kernels (name: str), or for any kernel given as a callable cov(x, z) (GPyKernel,
    kernels with children: kern/gpy_kernel.py:45-58, kern/basekernel.py:131-146).

    RBF          kern/stationary.py:121-127 (incl. the lengthscale < 1e-6 guard at :121-122)
    Exponential  kern/stationary.py:172-173
    Matern32     kern/stationary.py:213-214
    Matern52     kern/stationary.py:254-256
    """
    if callable(name):                       # any other kernel: kern.cov(x, z) as kern/grid_kernel.py:99,171 would call it
        return np.asarray(name(np.asarray(x, float).reshape(-1, 1), np.asarray(z, float).reshape(-1, 1)), dtype=float)
    d2 = _dist2(x, z)
    if name == "RBF":
        if lengthscale < 1e-6:
            return variance * (d2 == 0)
        return variance * np.exp(-0.5 * d2 / lengthscale ** 2)
    if name == "Exponential":
        r = np.sqrt(d2) / lengthscale
        return variance * np.exp(-r)
    if name == "Matern32":
        r = np.sqrt(d2) / lengthscale
        return variance * (1. + np.sqrt(3.) * r) * np.exp(-np.sqrt(3.) * r)
    if name == "Matern52":
        r2 = d2 / lengthscale ** 2
        r = np.sqrt(r2)
        return variance * (1. + np.sqrt(5.) * r + (5. / 3) * r2) * np.exp(-np.sqrt(5.) * r)
    raise ValueError("unknown kernel %r" % (name,))


# --------------------------------------------------------------------------------------------
# top-p Kronecker eigenvalue selection  (tensors/kron_matrix.py:369-446, linalg.py:74-89)
# --------------------------------------------------------------------------------------------

def log_kron(a_logged, b):
    """linalg.py:74-89 with a_logged=True: outer sum of log a and log b, row-major ravel."""
    return (a_logged.reshape((-1, 1)) + np.log(b).reshape((1, -1))).reshape(-1)


def find_extremum_eigs(eig_list, n_eigs, mode="largest", log_expand=True, sort=True):
    """Beam search for the n_eigs extreme Kronecker-product eigenvalues.

    Follows tensors/kron_matrix.py:396-425 step for step (same NumPy calls, hence the same
    implementation-defined behaviour of argpartition/argsort under ties) and the final sort at
    :439-443.  Returns (eig_loc (n_eigs, d) int64, eig_vals (n_eigs,)).
    """
    def get_extremum(vec):                                   # kron_matrix.py:396-404
        if np.size(vec) <= n_eigs:
            return np.arange(n_eigs), vec
        if mode == "largest":
            ind = np.argpartition(vec, -n_eigs)[-n_eigs:]
        else:
            ind = np.argpartition(vec, n_eigs)[:n_eigs]
        return ind, vec[ind]

    eig_loc, eig_vals = get_extremum(np.asarray(eig_list[0]))  # :407
    eig_loc = eig_loc.reshape((-1, 1))
    if log_expand:
        eig_vals = np.log(eig_vals)                          # :410
    for i in range(1, len(eig_list)):                        # :411
        Ki = np.asarray(eig_list[i])
        if log_expand:
            inds, eig_vals = get_extremum(log_kron(eig_vals, Ki))
        else:
            inds, eig_vals = get_extremum(np.kron(eig_vals, Ki))
        el1 = eig_loc[np.int32(np.floor_divide(inds, Ki.size)), :]   # :419
        el2 = np.int32(np.mod(inds, Ki.size))                        # :420
        eig_loc = np.hstack([el1.reshape((inds.size, -1)), el2.reshape((-1, 1))])
    if sort:                                                 # :440-443
        order = np.argsort(eig_vals)[::-1]
        eig_vals = eig_vals[order]
        eig_loc = eig_loc[order]
    return eig_loc.astype(np.int64), eig_vals


# --------------------------------------------------------------------------------------------
# inducing-grid eigen setup  (kern/grief_kernel.py:168-190)
# --------------------------------------------------------------------------------------------

class Basis(object):
    """What `_setup_inducing_cov` leaves on the kernel object, in KronMatrix order.

    Index k of every list refers to KronMatrix.K[k], i.e. input dimension d-1-k
    (kern/grid_kernel.py:109 and :174 reverse the dimension list).
    """
    def __init__(self, Q, eig, eig_loc, log_lam):
        self.Q = Q                  # list of (m_k, m_k) Schur vectors        (_Quu.K)
        self.eig = eig              # list of (m_k,) Schur diagonals (unsorted) (T.diag().K)
        self.eig_loc = eig_loc      # (p, d) selected eigen-indices            (_Sp[k].indicies)
        self.log_lam = log_lam      # (p,) log eigenvalue products, descending (_log_lam)
        self.unique = [np.unique(eig_loc[:, k], return_inverse=True) for k in range(eig_loc.shape[1])]


def grid_gram_factors(kernel_names, variances, lengthscales, xg, dim_noise_var=1e-12):
    """Per-dimension K_uu + dim_noise_var*I in KronMatrix (reversed) order.

    kern/grid_kernel.py:101-114 (`cov_grid`) and tensors/kron_matrix.py:276-294 (`sub_shift`,
    Fortran order).
    """
    d = len(xg)
    K = [kernel_cov(kernel_names[i], xg[i], xg[i], variances[i], lengthscales[i]) for i in range(d)]
    K = K[::-1]
    return [np.asarray(Ki + dim_noise_var * np.identity(Ki.shape[0]), order="F") for Ki in K]


def setup_inducing_cov(kernel_names, variances, lengthscales, xg, n_eigs, dim_noise_var=1e-12):
    """kern/grief_kernel.py:168-190: Schur of every grid factor, top-p selection, selection lists."""
    Kuu = grid_gram_factors(kernel_names, variances, lengthscales, xg, dim_noise_var)
    Q, eig = [], []
    for Ki in Kuu:                                           # tensors/kron_matrix.py:161-171
        T, Z = la.schur(Ki)
        Q.append(Z)
        eig.append(np.diag(T))                               # tensors/kron_matrix.py:254-265
    total = np.prod([float(e.size) for e in eig])
    n_eigs = int(min(n_eigs, total))                         # grief_kernel.py:36,183
    eig_loc, log_lam = find_extremum_eigs(eig, n_eigs, mode="largest", log_expand=True)
    return Basis(Q, eig, eig_loc, log_lam)


# --------------------------------------------------------------------------------------------
# basis matrix Phi  (kern/grief_kernel.py:68-111, tensors/tensors.py:97-128)
# --------------------------------------------------------------------------------------------

def expand_SKC_logged(basis, Kux):
    """tensors/tensors.py:110-126 with logged=True.  Kux[k] is (m_k, n) in KronMatrix order."""
    log_prod = 0.
    sign = 1.
    for k, c in enumerate(Kux):
        uniq, inv = basis.unique[k]
        x_unique = basis.Q[k].T[uniq, :].dot(c)              # s.mul_unique(k).dot(c), tensors.py:118
        sign = sign * np.int32(np.sign(x_unique))[inv]
        x_unique[x_unique == 0] = 1.
        log_prod = log_prod + np.log(np.abs(x_unique))[inv]
    return log_prod, sign


def grief_phi(basis, kernel_names, variances, lengthscales, xg, x):
    """Phi (n, p) = sign.T * exp(log_matrix.T - 0.5*log_lam)  (kern/grief_kernel.py:96-104)."""
    d = len(xg)
    Kxu = [kernel_cov(kernel_names[i], x[:, i], xg[i], variances[i], lengthscales[i]) for i in range(d)]
    Kxu = Kxu[::-1]                                          # kern/grid_kernel.py:171-174
    Kux = [k.T for k in Kxu]
    log_matrix, sign = expand_SKC_logged(basis, Kux)
    return sign.T * np.exp(log_matrix.T - 0.5 * basis.log_lam.reshape((1, -1)))


# --------------------------------------------------------------------------------------------
# GP-GRIEF model  (models/gp_grief_model.py)
# --------------------------------------------------------------------------------------------

class Fit(object):
    pass


def fit_from_phi(Phi, y, w, noise_var):
    """models/gp_grief_model.py:137-153 (`_cov_setup`), :78-87 (`fit`), :228-235, :238-245, :203-214."""
    n, p = Phi.shape
    f = Fit()
    f.Phi, f.w, f.noise_var, f.n, f.p = Phi, w, noise_var, n, p
    f.A = Phi.T.dot(Phi)                                               # :149
    f.P = f.A + np.diag(noise_var / w)                                 # :152
    f.Pchol = cho_factor(f.P)                                          # :153
    f.alpha = (y - Phi.dot(cho_solve(f.Pchol, Phi.T.dot(y)))) / noise_var   # :234
    f.log_det = (2. * np.sum(np.log(np.diag(f.Pchol[0]))) + np.sum(np.log(w))
                 + float(n - p) * np.log(noise_var))                   # :243-245
    f.lml = float((-0.5 * (y.T.dot(f.alpha) + f.log_det + n * np.log(np.pi * 2))).squeeze())  # :212-213
    return f


def adjoint_gradient(f, reweight=True, noise_free=True):
    """d LML / d w  and  d LML / d noise_var   (models/gp_grief_model.py:171-191)."""
    Pinv_A = cho_solve(f.Pchol, f.A)
    gw = None
    if reweight:
        data_fit = 0.5 * np.power(f.Phi.T.dot(f.alpha), 2).squeeze()
        complexity = -0.5 * (f.A.diagonal() - (f.A * Pinv_A).sum(axis=0)) / f.noise_var
        gw = data_fit + complexity
    gs = None
    if noise_free:
        gs = float((0.5 * f.alpha.T.dot(f.alpha) - 0.5 * (float(f.n) - np.trace(Pinv_A)) / f.noise_var).squeeze())
    return gs, gw


def predict(f, Phi_new):
    """models/gp_grief_model.py:89-125: mean (M,), full covariance (M, M)."""
    alpha_p = f.Phi.T.dot(f.alpha) * f.w.reshape((-1, 1))             # :97
    yhat = Phi_new.dot(alpha_p)                                       # :119
    yvar = f.noise_var * Phi_new.dot(cho_solve(f.Pchol, Phi_new.T)) + f.noise_var * np.eye(Phi_new.shape[0])
    return yhat.squeeze(axis=1), yvar


def lml_full(kernel_names, variances, lengthscales, xg, n_eigs, x, y, w, noise_var, dim_noise_var=1e-12):
    """One complete reference evaluation: setup + Phi + fit.  Returns (Fit, Basis)."""
    basis = setup_inducing_cov(kernel_names, variances, lengthscales, xg, n_eigs, dim_noise_var)
    Phi = grief_phi(basis, kernel_names, variances, lengthscales, xg, x)
    if np.ndim(w) == 0:
        w = np.full(Phi.shape[1], float(w))
    return fit_from_phi(Phi, y, w, noise_var), basis


def finite_diff_gradient(lml_of_params, params, free_inds, step=1e-6):
    """models/basemodel.py:328-361: forward differences, one extra evaluation per free parameter."""
    fs = np.zeros(len(free_inds))
    for i, idx in enumerate(free_inds):
        p_fs = params.copy()
        p_fs[idx] += step
        fs[i] = lml_of_params(p_fs)
    base = lml_of_params(params)
    grad = np.zeros(params.shape)
    grad[free_inds] = (fs - base) / step
    return base, grad


# --------------------------------------------------------------------------------------------
# GPweb (reduced statistics) model  (models/gp_web_model.py)
# --------------------------------------------------------------------------------------------

def gpweb_lml_grad(A, r, yty, n, w, noise_var):
    """models/gp_web_model.py:51-69 and :72-106: LML and [d/d noise_var, d/d w...]."""
    p = A.shape[0]
    r = r.reshape(-1, 1)
    P = A + np.diag(noise_var / w)
    Pchol = cho_factor(P)
    Pinv_r = cho_solve(Pchol, r)
    datafit = (yty - r.T.dot(Pinv_r)) / noise_var
    complexity = 2. * np.sum(np.log(np.diag(Pchol[0]))) + np.sum(np.log(w)) + float(n - p) * np.log(noise_var)
    lml = float((-0.5 * (complexity + datafit + n * np.log(2. * np.pi))).squeeze())
    Pinv_A = cho_solve(Pchol, A)
    grad = np.zeros(p + 1)
    dfit = -np.power((r - A.dot(Pinv_r)) / noise_var, 2)
    cplx = (A.diagonal() - (A * Pinv_A).sum(axis=0)) / noise_var
    grad[1:] = -0.5 * dfit.squeeze() - 0.5 * cplx.squeeze()
    dfit = -(yty - 2. * r.T.dot(Pinv_r) + Pinv_r.T.dot(A.dot(Pinv_r))) / (noise_var ** 2)
    cplx = (float(n) - np.trace(Pinv_A)) / noise_var
    grad[0] = float((-0.5 * (dfit + cplx)).squeeze())
    return lml, grad
