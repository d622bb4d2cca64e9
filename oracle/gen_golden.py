#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (where /root/reference exists); the
GPU box never executes this file.  The reference package (`/root/reference/gp_grief`) imports
`GPy` at kern/basekernel.py:3 and kern/gpy_kernel.py:3, which is not installed and cannot be
installed (no network).  We therefore write a 4-file *import stub* for GPy into
`oracle/_ref/gpy_stub/` (git-ignored) that defines exactly the names those two import sites touch
(`GPy.kern.Kern`, `GPy.kern.src.stationary.Stationary`) and drive the reference with its in-house
kernels (`gp_grief.kern.RBF` etc., kern/stationary.py:79-258), which have no GPy dependency.

Usage:  python oracle/gen_golden.py            (writes tests/golden/*.npz)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GP_GRIEF_REFERENCE", "/root/reference")
STUB = os.path.join(HERE, "_ref", "gpy_stub")
GOLD = os.path.join(ROOT, "tests", "golden")


def write_gpy_stub():
    files = {
        "GPy/__init__.py": "from . import kern\n",
        "GPy/kern/__init__.py": "class Kern(object):\n    pass\nfrom . import src\n",
        "GPy/kern/src/__init__.py": "from . import stationary\n",
        "GPy/kern/src/stationary.py": "class Stationary(object):\n    pass\n",
    }
    for rel, body in files.items():
        path = os.path.join(STUB, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(body)


write_gpy_stub()
sys.path.insert(0, REF)
sys.path.insert(0, STUB)

import warnings  # noqa: E402

warnings.filterwarnings("ignore")
import numpy as np  # noqa: E402
import gp_grief  # noqa: E402  (the reference)
from gp_grief.grid import InducingGrid  # noqa: E402
from gp_grief.kern import RBF, Exponential, Matern32, Matern52, GriefKernel  # noqa: E402
from gp_grief.models import GPGriefModel, GPwebModel  # noqa: E402
from gp_grief.tensors import KronMatrix  # noqa: E402

assert os.path.realpath(gp_grief.__file__).startswith(os.path.realpath(REF)), gp_grief.__file__
sys.path.insert(0, ROOT)
from oracle.synthetic import synthetic_xy, linspace_grid  # noqa: E402

KERNELS = {"RBF": RBF, "Exponential": Exponential, "Matern32": Matern32, "Matern52": Matern52}


def kernel_state(kern):
    """Everything the new implementation needs to reproduce the reference's gauge."""
    kern._setup_inducing_cov()
    d = kern.grid_dim
    out = {}
    for k in range(d):  # k indexes KronMatrix.K, i.e. input dimension d-1-k (grid_kernel.py:109)
        out["Q_%d" % k] = np.ascontiguousarray(kern._Quu.K[k])
        out["sel_%d" % k] = np.asarray(kern._Sp[k].indicies, dtype=np.int64)
    out["log_lam"] = np.asarray(kern._log_lam)
    return out


def run_model_case(name, x, y, xg, kernel_name, variances, lengthscales, n_eigs, noise_var, w=None,
                   xnew=None, type2=False, alias=False, keep_phi=True, grid_from_x=False):
    d = x.shape[1]
    cls = KERNELS[kernel_name]
    if alias:
        k0 = cls(1, variance=variances[0], lengthscale=lengthscales[0])
        kern_list = [k0, ] * d
    else:
        kern_list = [cls(1, variance=variances[i], lengthscale=lengthscales[i]) for i in range(d)]
    if grid_from_x:
        grid = InducingGrid(x=x)
    else:  # object array: NumPy >= 1.24 refuses to build a ragged array implicitly (grid.py:145)
        xg_obj = np.empty(len(xg), dtype=object)
        for i, g in enumerate(xg):
            xg_obj[i] = np.asarray(g, float).reshape(-1, 1)
        grid = InducingGrid(xg=xg_obj)
    if type2:
        kern = GriefKernel(kern_list, grid, n_eigs=n_eigs, reweight_eig_funs=False, opt_kernel_params=True)
    else:
        kern = GriefKernel(kern_list, grid, n_eigs=n_eigs)
    if w is not None:
        kern.w = np.asarray(w, dtype=float)
    m = GPGriefModel(x, y, kern, noise_var=noise_var)
    out = {"x": x, "y": y, "kernel_name": np.array(kernel_name), "variances": np.asarray(variances, float),
           "lengthscales": np.asarray(lengthscales, float), "n_eigs": np.int64(kern.n_eigs),
           "noise_var": np.float64(noise_var), "type2": np.bool_(type2), "alias": np.bool_(alias),
           "w": np.asarray(kern.w, float), "n_grid_dims": np.int64(d)}
    for i in range(d):
        out["xg_%d" % i] = np.asarray(grid.xg[i], float).reshape(-1)
    params = m.parameters
    out["parameters"] = params
    out["constraints"] = np.array([c.decode() if isinstance(c, bytes) else str(c) for c in m.constraints])
    lml = m._compute_log_likelihood(params)
    out["lml"] = np.float64(np.asarray(lml).squeeze())
    out.update(kernel_state(kern))
    Phi = m._Phi
    if keep_phi:
        out["Phi"] = np.ascontiguousarray(Phi)
    out["A"] = m._A
    out["r"] = Phi.T.dot(y).squeeze()
    out["alpha"] = m._alpha.squeeze()
    out["log_det"] = np.float64(m._cov_log_det())
    out["chol_diag"] = np.diag(m._Pchol[0]).copy()
    # gradient the reference's own way
    if type2:
        ll, g = m._finite_diff_gradient(params.copy())
        out["grad_fd"] = g
        # central differences with Richardson extrapolation of the REFERENCE's LML.  The LML jumps wherever the
        # selected eigen-index set changes, so the step is shrunk until all four evaluation points keep the
        # base point's selection (then the difference quotient is taken on one smooth branch).
        free = np.nonzero(np.logical_not(m._fixed_indicies))[0]
        base_sel = [np.sort(np.ravel_multi_index(np.array([s_.indicies for s_ in kern._Sp]),
                                                 [int(s_.shape[1]) for s_ in kern._Sp]))]
        gc = np.zeros(params.shape)
        steps = np.zeros(params.shape)
        for idx in free:
            def f(hh):
                vals = []
                for sgn in (+1, -1):
                    pp = params.copy(); pp[idx] += sgn * hh
                    vals.append(float(np.asarray(m._compute_log_likelihood(pp)).squeeze()))
                    sel = np.sort(np.ravel_multi_index(np.array([s_.indicies for s_ in kern._Sp]),
                                                       [int(s_.shape[1]) for s_ in kern._Sp]))
                    if not np.array_equal(sel, base_sel[0]):
                        return None
                return (vals[0] - vals[1]) / (2 * hh)
            h = 2e-4 * max(1.0, abs(params[idx]))
            while True:
                d1, d2 = f(h), f(h / 2)
                if d1 is not None and d2 is not None:
                    break
                h /= 4
                assert h > 1e-9
            gc[idx] = (4 * d2 - d1) / 3
            steps[idx] = h
        out["grad_central_step"] = steps
        out["grad_central"] = gc
        out["free"] = free
        m.parameters = params
    else:
        m.parameters = params
        m.grad_method = "adjoint"
        ll, g = m._adjoint_gradient(params.copy())
        out["grad_adjoint"] = g
    if xnew is not None:
        m.parameters = params
        yhat, yvar = m.predict(xnew)
        out["xnew"] = xnew
        out["yhat"] = yhat.squeeze()
        out["yvar_diag"] = np.diag(yvar).copy()
        if xnew.shape[0] <= 64:
            out["yvar_full"] = yvar
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("%-28s n=%d d=%d p=%d lml=%.15e" % (name, x.shape[0], d, kern.n_eigs, out["lml"]))
    return out


def topk_case(name, d, m, p, lengthscales=None):
    """Top-p selection at a benchmark configuration (SURVEY 8(d) kernels), reference output."""
    xg = linspace_grid(d, m)
    if lengthscales is None:
        lengthscales = [0.3 + 0.05 * i for i in range(d)]
    kern_list = [RBF(1, variance=1., lengthscale=lengthscales[i]) for i in range(d)]
    grid = InducingGrid(xg=[g.reshape(-1, 1) for g in xg])
    kern = GriefKernel(kern_list, grid, n_eigs=p)
    kern._setup_inducing_cov()
    T_diag = []
    Kuu = kern.cov_grid(grid.xg, dim_noise_var=kern.dim_noise_var)
    (_, T) = Kuu.schur()
    eigs = T.diag()
    loc, logl = eigs.find_extremum_eigs(n_eigs=int(kern.n_eigs), mode="largest", log_expand=True)[:2]
    srt = np.sort(logl)
    out = {"d": np.int64(d), "m": np.int64(m), "p": np.int64(p), "lengthscales": np.asarray(lengthscales),
           "eig_loc": loc.astype(np.int16), "log_lam": logl,
           "n_ties": np.int64(np.sum(np.diff(srt) == 0))}
    for k in range(d):
        out["eigs_%d" % k] = np.asarray(eigs.K[k])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("%-28s d=%d m=%d p=%d ties=%d min_gap=%.3e" % (name, d, m, p, out["n_ties"], np.min(np.diff(srt))))


def kron_eigs_case():
    """tests/test_tensors/test_kron_eigenvalues.py:12-92 inputs and the reference's outputs."""
    np.random.seed(1)
    d, n, n_eigs = 10, 3, 5
    eigs = KronMatrix([np.random.rand(n) for _ in range(d)])
    out = {"n_eigs": np.int64(n_eigs)}
    for i in range(d):
        out["eigs_%d" % i] = eigs.K[i]
    out["all_sorted_top"] = np.sort(eigs.expand())[::-1][:n_eigs]
    out["all_sorted_bottom"] = np.sort(eigs.expand())[:n_eigs]
    for mode in ("largest", "smallest"):
        for log_expand in (False, True):
            loc, vals, gl = eigs.find_extremum_eigs(n_eigs, mode=mode, log_expand=log_expand, sort=True,
                                                    compute_global_loc=True)
            tag = "%s_%s" % (mode, "log" if log_expand else "lin")
            out["loc_" + tag] = loc.astype(np.int64)
            out["vals_" + tag] = vals
            out["gloc_" + tag] = gl.astype(np.int64)
    np.savez_compressed(os.path.join(GOLD, "kron_eigs_d10_m3_p5.npz"), **out)
    print("kron_eigs_d10_m3_p5 done")


def web_case():
    """tests/test_models/test_gp_web_model.py:13-34 — reduced-statistics LML and adjoint gradient."""
    np.random.seed(0)
    X = np.random.randn(100, 4)
    X[:, 0] = 1.
    Y = np.dot(X, [0.5, 0.1, 0.25, 1.]) + 0.1 * np.random.randn(X.shape[0])
    m = GPwebModel(Phi=X, y=Y)
    params = np.random.rand(*m.parameters.shape) + 1e-6
    m.parameters = params
    ll, g = m._adjoint_gradient(m.parameters)
    Xnew = np.random.randn(6, 4)
    yhat, yvar = m.predict(Xnew)
    out = {"Phi": X, "y": Y, "parameters": params, "lml": np.float64(np.asarray(ll).squeeze()), "grad": g,
           "Phi_new": Xnew, "yhat": yhat.squeeze(), "yvar": yvar}
    np.savez_compressed(os.path.join(GOLD, "gpweb_n100_p4.npz"), **out)
    print("gpweb_n100_p4 lml=%.15e" % out["lml"])


def automobile_case():
    """BASELINE config C1: Type-II notebook cell 9 data preparation (lines 193-199)."""
    from sklearn.preprocessing import StandardScaler
    np.random.seed(0)
    data = np.loadtxt(os.path.join(REF, "tutorials", "automobile.csv"), delimiter=",")
    i_train = np.random.rand(data.shape[0]) < 0.9
    x = StandardScaler().fit_transform(data[:, :-1])
    y = StandardScaler().fit_transform(data[:, (-1,)])
    d = x.shape[1]
    # grid from ALL rows (as in the notebook), model on the training rows
    grid = InducingGrid(x=x)
    kern_list = [RBF(1, lengthscale=1.) for _ in range(d)]
    kern = GriefKernel(kern_list, grid, n_eigs=100)
    m = GPGriefModel(x[i_train], y[i_train], kern, noise_var=1.)
    params = m.parameters
    ll, g = m._adjoint_gradient(params.copy())
    out = {"x": x[i_train], "y": y[i_train], "xnew": x[~i_train], "ynew": y[~i_train],
           "kernel_name": np.array("RBF"), "variances": np.ones(d), "lengthscales": np.ones(d),
           "n_eigs": np.int64(100), "noise_var": np.float64(1.), "w": np.ones(100), "n_grid_dims": np.int64(d),
           "type2": np.bool_(False), "alias": np.bool_(False),
           "lml": np.float64(np.asarray(ll).squeeze()), "grad_adjoint": g, "parameters": params,
           "A": m._A, "r": m._Phi.T.dot(m.Y).squeeze(), "alpha": m._alpha.squeeze(),
           "log_det": np.float64(m._cov_log_det()), "Phi": np.ascontiguousarray(m._Phi)}
    for i in range(d):
        out["xg_%d" % i] = np.asarray(grid.xg[i], float).reshape(-1)
    out.update(kernel_state(kern))
    yhat, yvar = m.predict(x[~i_train])
    out["yhat"] = yhat.squeeze()
    out["yvar_diag"] = np.diag(yvar).copy()
    out["yvar_full"] = yvar
    srt = np.sort(out["log_lam"])
    out["n_ties"] = np.int64(np.sum(np.diff(srt) == 0))
    np.savez_compressed(os.path.join(GOLD, "c1_automobile.npz"), **out)
    print("c1_automobile n=%d d=%d lml=%.15e dsig=%.15e ties=%d" % (x[i_train].shape[0], d, out["lml"], g[0], out["n_ties"]))


def composite_kernel_case(name, x, y, xg, op, parent, child, n_eigs, noise_var, xnew, type2=False):
    """Kernels WITH CHILDREN (kern/basekernel.py:131-190: `k1 * k2`, `k1 + k2`): no closed formula inside libgrief_b200, so the
    new implementation evaluates their K_xu on the host and uploads it (grief_build_tables_kxu).  parent / child:
    (kernel class name, variance, [lengthscale per dimension])."""
    d = x.shape[1]
    kern_list = []
    for i in range(d):
        k1 = KERNELS[parent[0]](1, variance=parent[1], lengthscale=parent[2][i])
        k2 = KERNELS[child[0]](1, variance=child[1], lengthscale=child[2][i])
        kern_list.append(k1 * k2 if op == "mul" else k1 + k2)
    xg_obj = np.empty(len(xg), dtype=object)
    for i, g in enumerate(xg):
        xg_obj[i] = np.asarray(g, float).reshape(-1, 1)
    grid = InducingGrid(xg=xg_obj)
    if type2:
        kern = GriefKernel(kern_list, grid, n_eigs=n_eigs, reweight_eig_funs=False, opt_kernel_params=True)
    else:
        kern = GriefKernel(kern_list, grid, n_eigs=n_eigs)
        for k in kern.kern_list:      # GriefKernel fixes the parents' parameters only (grief_kernel.py:44-52): fix the children's too
            for _, ch in k._children:
                for key in ch.constraint_map:
                    ch.constraint_map[key] = np.tile('fixed', np.shape(ch.constraint_map[key]))
    m = GPGriefModel(x, y, kern, noise_var=noise_var)
    params = m.parameters
    out = {"x": x, "y": y, "op": np.array(op), "parent_name": np.array(parent[0]), "parent_variance": np.float64(parent[1]),
           "parent_lengthscales": np.asarray(parent[2], float), "child_name": np.array(child[0]),
           "child_variance": np.float64(child[1]), "child_lengthscales": np.asarray(child[2], float),
           "n_eigs": np.int64(kern.n_eigs), "noise_var": np.float64(noise_var), "type2": np.bool_(type2),
           "n_grid_dims": np.int64(d), "parameters": params,
           "constraints": np.array([c.decode() if isinstance(c, bytes) else str(c) for c in m.constraints])}
    for i in range(d):
        out["xg_%d" % i] = np.asarray(grid.xg[i], float).reshape(-1)
    out["lml"] = np.float64(np.asarray(m._compute_log_likelihood(params)).squeeze())
    out.update(kernel_state(kern))
    out["A"] = m._A
    out["log_det"] = np.float64(m._cov_log_det())
    if type2:
        ll, g = m._finite_diff_gradient(params.copy())
        out["grad_fd"] = g
    else:
        m.grad_method = "adjoint"
        ll, g = m._adjoint_gradient(params.copy())
        out["grad_adjoint"] = g
    m.parameters = params
    yhat, yvar = m.predict(xnew)
    out["xnew"] = xnew
    out["yhat"] = yhat.squeeze()
    out["yvar_diag"] = np.diag(yvar).copy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("%-28s n=%d d=%d p=%d lml=%.15e" % (name, x.shape[0], d, kern.n_eigs, out["lml"]))


def composite_cases():
    x, y = synthetic_xy(800, 3, chunk=1 << 10)
    xnew = synthetic_xy(12, 3, chunk_id0=10 ** 6)[0]
    composite_kernel_case("host_t1_sum_n800_d3_m9_p36", x, y, linspace_grid(3, 9), "add", ("RBF", 1.0, [0.35, 0.5, 0.65]),
                          ("Matern32", 0.4, [0.8, 0.6, 0.9]), 36, 0.15, xnew)
    composite_kernel_case("host_t1_prod_n800_d3_m9_p36", x, y, linspace_grid(3, 9), "mul", ("RBF", 1.1, [0.45, 0.5, 0.7]),
                          ("Exponential", 1.0, [1.5, 2.0, 1.2]), 36, 0.15, xnew)
    composite_kernel_case("host_t2_sum_n500_d2_m8_p24", x[:500, :2], y[:500], linspace_grid(2, 8), "add", ("RBF", 1.0, [0.35, 0.5]),
                          ("Matern52", 0.5, [0.7, 0.9]), 24, 0.2, xnew[:, :2], type2=True)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--composite-only" in sys.argv:
        composite_cases()
        return
    kron_eigs_case()
    web_case()

    # restated tests/test_models/test_gp_grief_model.py:14-40 (in-house RBF instead of GPyKernel)
    np.random.seed(0)
    d, n = 5, 100
    x = np.random.rand(n, d)
    y = np.random.rand(n, 1)
    run_model_case("ref_test_gp_grief_model", x, y, None, "RBF", [1.] * d, [0.5] * d, 50, 0.1,
                   alias=True, grid_from_x=True, xnew=np.random.rand(7, d))

    # tie-free synthetic Type-I cases (SURVEY 8(d) generator, distinct lengthscales)
    x, y = synthetic_xy(2000, 4, chunk=1 << 11)
    run_model_case("syn_t1_n2000_d4_m8_p64", x, y, linspace_grid(4, 8), "RBF", [1.] * 4,
                   [0.3 + 0.05 * i for i in range(4)], 64, 0.1, xnew=synthetic_xy(50, 4, chunk_id0=10 ** 6)[0],
                   keep_phi=False)
    x, y = synthetic_xy(3000, 6, chunk=1 << 11)
    rng = np.random.default_rng(7)
    run_model_case("syn_t1_n3000_d6_m10_p256_w", x, y, linspace_grid(6, 10), "RBF",
                   [1.3, 1., 1., 1., 1., 1.], [0.3 + 0.05 * i for i in range(6)], 256, 0.1,
                   w=rng.random(256) + 0.5, xnew=synthetic_xy(40, 6, chunk_id0=10 ** 6)[0], keep_phi=False)
    # ragged grid with m_i = 1 and m_i = 2 dimensions
    x, y = synthetic_xy(700, 5, chunk=1 << 10)
    xg = [np.linspace(0, 1, 7), np.array([0.5]), np.linspace(0, 1, 2), np.linspace(-0.1, 1.1, 12), np.linspace(0, 1, 5)]
    run_model_case("syn_t1_ragged_n700_d5_p40", x, y, xg, "RBF", [1.] * 5, [0.4, 0.7, 0.55, 0.3, 0.6], 40, 0.2,
                   xnew=synthetic_xy(20, 5, chunk_id0=10 ** 6)[0])
    # the other in-house kernels
    for kname in ("Exponential", "Matern32", "Matern52"):
        x, y = synthetic_xy(600, 3, chunk=1 << 10)
        run_model_case("syn_t1_%s_n600_d3_m9_p30" % kname.lower(), x, y, linspace_grid(3, 9), kname, [1.] * 3,
                       [0.35, 0.5, 0.65], 30, 0.15, xnew=synthetic_xy(10, 3, chunk_id0=10 ** 6)[0])
    # Type-II: reference finite-difference gradient + Richardson central differences of the reference LML
    x, y = synthetic_xy(2000, 4, chunk=1 << 11)
    run_model_case("syn_t2_n2000_d4_m8_p64", x, y, linspace_grid(4, 8), "RBF", [1.] * 4,
                   [0.3 + 0.05 * i for i in range(4)], 64, 0.1, type2=True, keep_phi=False)
    x, y = synthetic_xy(1500, 3, chunk=1 << 11)
    run_model_case("syn_t2_matern52_n1500_d3_m10_p48", x, y, linspace_grid(3, 10), "Matern52", [1.2, 1., 1.],
                   [0.4, 0.5, 0.6], 48, 0.1, type2=True, keep_phi=False)

    automobile_case()
    composite_cases()

    # top-p selection at the benchmark configurations
    topk_case("topk_c2_d6_m10_p1024", 6, 10, 1024)
    topk_case("topk_c3_d10_m20_p4096", 10, 20, 4096)
    topk_case("topk_c4_d32_m8_p2048", 32, 8, 2048)
    topk_case("topk_c5_d8_m16_p8192", 8, 16, 8192)


if __name__ == "__main__":
    main()
