"""GP-GRIEF regression model on the B200 (reference: gp_grief/models/gp_grief_model.py).

Same class, constructor and methods as the reference.  Data rows live on the GPU; one evaluation is
    prepass tables  ->  Gram A = Phi^T Phi, r = Phi^T y, s = y^T y   (csrc/rows.cu, csrc/phi_stage.cu, csrc/ozaki.cu | dense.cu)
    [all-reduce of (A | r | s) when the rows are sharded over ranks]
    Cholesky / solve / LML / d/dw / d/dnoise_var                           (csrc/solve.cu, csrc/dense.cu)
    analytic d/d(kernel hyper-parameters), reverse mode                    (csrc/phi_stage.cu, csrc/ozaki.cu | dense.cu, csrc/grad.cu)
and Phi (n x p) is never formed.  Differences from the reference that a caller can observe:
  * `_Phi`, `_alpha` are computed on demand (they are n-sized); `_A`, `_P`, `_Pchol` are host copies.
  * with `opt_kernel_params=True` and distinct in-house kernels the default `grad_method` is the analytic
    'adjoint' path; the reference can only finite-difference there (gp_grief_model.py:71-74,194-196).
    Set `m.grad_method = 'finite_difference'` to reproduce the reference's gradient bit-for-bit in method; kernels that exist only
    as host code (GPyKernel, kernels with children) always take that route, as in the reference.
  * `predict(Xnew, compute_var='diag')` returns the marginal variances without the M x M covariance.
  * the two O(n p^2) products run as exact int8 digit GEMMs on the tensor cores (`gemm_digits`); the first evaluation of a model
    audits the digit counts against the FP64 arithmetic on a row sample (`arithmetic_audit`, `audit_rows`, `audit_tol*`).
There is no CPU path: constructing the model without a CUDA device raises.
"""
from logging import getLogger

import numpy as np

from ..kern import BaseKernel, GriefKernel
from .basemodel import BaseModel

logger = getLogger(__name__)


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


class GPGriefModel(BaseModel):
    """GP with GRId-structured Eigen Functions."""

    _CACHE_NAMES = ('_A', '_P', '_Pchol', '_alpha', '_alpha_p', '_Phi', '_X_last_pred', '_Phi_last_pred')

    def __init__(self, X, Y, kern, noise_var=1., distributed=False):
        """X (n, d), Y (n, 1), kern a GriefKernel.  With distributed=True, X and Y are this rank's row shard of a
        torch.distributed job (NCCL); the statistics are all-reduced and every rank gets identical results."""
        super(GPGriefModel, self).__init__()
        assert X.ndim == 2
        assert Y.ndim == 2
        self.X = np.asarray(X)
        self.Y = np.asarray(Y)
        assert not np.any(np.isnan(Y))
        self.num_local, self.input_dim = self.X.shape
        if Y.shape[0] != self.num_local:
            raise ValueError('X and Y sizes are inconsistent')
        self.output_dim = self.Y.shape[1]
        if self.output_dim != 1:
            raise RuntimeError('this only deals with 1 response for now')
        assert isinstance(kern, GriefKernel)
        assert np.ndim(kern.kern_list) == 1
        for ki in kern.kern_list:
            assert isinstance(ki, BaseKernel)
            assert ki.n_dims == 1, "currently only 1-dimensional grids allowed"
        self.kern = kern
        self.noise_var = np.float64(noise_var)

        import torch
        from .. import device
        self._torch = torch
        self._device_mod = device
        device._torch()                                  # raises if there is no CUDA device
        self._dist = _dist() if distributed else None
        self.num_data = self.num_local
        if self._dist is not None:
            cnt = torch.tensor([self.num_local], dtype=torch.int64, device="cuda")
            self._dist.all_reduce(cnt)
            self.num_data = int(cnt.item())
        self._X_dev = torch.as_tensor(np.ascontiguousarray(self.X, dtype=np.float64)).cuda()
        self._y_dev = torch.as_tensor(np.ascontiguousarray(self.Y[:, 0], dtype=np.float64)).cuda()
        self._dev = {}                                   # device-resident caches (tensors)
        self._host = {}                                  # lazily copied host views of the caches

        self.dependent_attributes = np.unique(np.concatenate(
            (self.dependent_attributes, ['_P', '_Pchol', '_alpha_p'])))
        if self.kern.opt_kernel_params:
            self.dependent_attributes = np.unique(np.concatenate(
                (self.dependent_attributes, ['_A', '_Phi', '_X_last_pred', '_Phi_last_pred'])))
            analytic_ok = not self.kern.has_aliased_kernels() and not self.kern.has_host_kernels()
            self.grad_method = 'adjoint' if analytic_ok else 'finite_difference'
        else:
            self.grad_method = 'adjoint'

    # ------------------------------------------------------------------ cache attributes of the reference
    # The reference invalidates by `setattr(self, name, None)`; these properties route that to the device caches.
    def __getattr__(self, name):
        if name in GPGriefModel._CACHE_NAMES:
            return self._get_cache(name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in GPGriefModel._CACHE_NAMES:
            self._set_cache(name, value)
        else:
            object.__setattr__(self, name, value)

    def _set_cache(self, name, value):
        host, dev = self.__dict__.get('_host'), self.__dict__.get('_dev')
        if host is None:
            return
        if value is None:
            host.pop(name, None)
            if name == '_A':
                for k in ('stats', 'tables', 'plan_id', 'rowmax'):
                    dev.pop(k, None)
            if name in ('_P', '_Pchol'):
                for k in ('solve',):
                    dev.pop(k, None)
            if name == '_Phi_last_pred':
                dev.pop('pred', None)
        else:
            host[name] = value

    def _get_cache(self, name):
        host, dev = self._host, self._dev
        if name in host:
            return host[name]
        t = self._torch
        if name == '_A':
            if 'stats' not in dev:
                return None
            host[name] = dev['stats']['A'].cpu().numpy()
        elif name == '_P':
            if 'solve' not in dev:
                return None
            host[name] = self._get_cache('_A') + np.diag(self.noise_var / self._w)
        elif name == '_Pchol':
            if 'solve' not in dev:
                return None
            host[name] = (dev['solve']['L'].cpu().numpy(), False)      # upper factor, scipy cho_factor convention
        elif name == '_alpha_p':
            if 'solve' not in dev:
                return None
            host[name] = dev['solve']['b'].cpu().numpy().reshape((-1, 1))
        elif name == '_alpha':
            if 'solve' not in dev:
                return None
            plan = self.kern.device_plan()
            fitted = plan.phi_vec(dev['tables'], self.num_local, dev['solve']['b'])
            host[name] = ((self._y_dev - fitted) / float(self.noise_var)).cpu().numpy().reshape((-1, 1))
        elif name == '_Phi':
            if 'tables' not in dev:
                return None
            plan = self.kern.device_plan()
            host[name] = plan.phi_rows(dev['tables'], self.num_local).cpu().numpy()
        else:
            return None
        return host[name]

    # ------------------------------------------------------------------ fit
    def fit(self, **kwargs):
        """Reduced statistics, Cholesky of P and alpha_p = P^-1 Phi^T y (reference :78-87)."""
        self.parameters
        self._cov_setup()

    #: (digits of A = Phi^T Phi, digits of Zp = Phi P^-1 in the gradient pass[, digits of the predictive-variance product]) for the
    #: INT8 tensor-core arithmetic, or None for the library defaults 6 / 4 / 6 (include/grief_b200.h, GRIEF_OPT_DIGITS_*).
    #: 8 D - 2 bits per operand below its row maximum.
    gemm_digits = None

    #: A-posteriori audit of the INT8 digit counts (first evaluation of a model, and again after `reset_audit()`): the two O(n p^2)
    #: products are recomputed on `audit_rows` rows in the FP64 DMMA arithmetic, the difference is pushed through the first-order
    #: sensitivity of the FULL problem (its P^-1 and b) and extrapolated linearly to all rows -- the worst case of a systematic
    #: truncation bias, random errors only grow like sqrt(n).  If the estimated relative effect on the LML (or on the largest
    #: gradient component) exceeds `audit_tol`, the digit count of that product is raised and the product recomputed.  The result
    #: is kept in `arithmetic_audit`.  audit_rows = 0 switches the audit off.
    #: Measured at C3 (n = 10^6, p = 4096, tools/audit_scaling.py): with 4 digits the gradient pass differs from FP64 by 1.6e-11 of
    #: the largest kernel-parameter gradient component over all rows and by 6e-11 on 16384 rows (the part that comes from rounding
    #: P^-1 to 30 bits is the same for every row and does not average out); 5 digits: 7e-14; the 6-digit Gram moves the LML by 1e-15.
    audit_rows = 16384
    audit_tol = 1e-10           # LML
    audit_tol_grad = 2.5e-10    # kernel-parameter gradient, relative to its largest component (north star: 1e-9)
    arithmetic_audit = None

    def reset_audit(self):
        """Audit the arithmetic again at the next evaluation (e.g. after moving far in hyper-parameter space)."""
        self._audit_done = {}

    def _plan(self):
        """The kernel's device plan with this model's arithmetic options applied."""
        from .. import _native as nat
        plan = self.kern.device_plan()
        digits = self.__dict__.get('_audited_digits') or self.gemm_digits
        if digits is not None:
            plan.set_option(nat.OPT_DIGITS_GRAM, digits[0])
            plan.set_option(nat.OPT_DIGITS_Z, digits[1])
            if len(digits) > 2:
                plan.set_option(nat.OPT_DIGITS_VAR, digits[2])
        return plan

    # ------------------------------------------------------------------ a-posteriori audit of the INT8 arithmetic
    def _audit_pending(self, what):
        from .. import _native as nat
        if not self.audit_rows or self.__dict__.setdefault('_audit_done', {}).get(what):
            return False
        return (self._plan().get_option(nat.OPT_GEMM_MODE) & 1) == 1

    def _audit_max_over_ranks(self, value):
        if self._dist is None:
            return float(value)
        t = self._torch.tensor([float(value)], dtype=self._torch.float64, device="cuda")
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def _raise_digits(self, which):
        """One more digit for product `which` (0 Gram, 1 gradient pass); False when it is already at the DGEMM class."""
        from .. import _native as nat
        plan = self._plan()
        cur = [plan.get_option(nat.OPT_DIGITS_GRAM), plan.get_option(nat.OPT_DIGITS_Z), plan.get_option(nat.OPT_DIGITS_VAR)]
        if cur[which] >= 7:
            return False
        cur[which] += 1
        self._audited_digits = tuple(cur)
        return True

    def _audit_gram(self, solve_out):
        """Estimated relative LML error caused by the digit truncation of A = Phi^T Phi.
        d LML = -1/2 [ tr(P^-1 dA) + b^T dA b / noise_var ] with dA measured on the audit rows (INT8 digits vs FP64 DMMA)."""
        from .. import _native as nat
        t = self._torch
        plan = self._plan()
        ns = int(min(self.num_local, self.audit_rows))
        est = 0.0
        if ns > 0:
            T = self._dev['tables']
            A8 = plan.gram(T, ns)
            mode = plan.get_option(nat.OPT_GEMM_MODE)
            plan.set_option(nat.OPT_GEMM_MODE, 0)
            try:
                A64 = plan.gram(T, ns)
            finally:
                plan.set_option(nat.OPT_GEMM_MODE, mode)
            dA = A8 - A64
            b = solve_out['b']
            d_lml = -0.5 * ((solve_out['Pinv'] * dA).sum() + (dA * t.outer(b, b)).sum() / float(self.noise_var))
            est = abs(float(d_lml)) * (self.num_local / float(ns))
        est = self._audit_max_over_ranks(est) * (1 if self._dist is None else self._dist.get_world_size())
        return est / max(abs(float(solve_out['lml'])), 1e-300), ns

    def _audit_grad(self, solve_out, g_full):
        """Estimated error of the kernel-parameter gradient caused by the digit truncation of Zp = Phi P^-1, relative to the
        largest gradient component: pass 2 on the audit rows in both arithmetics, difference extrapolated linearly."""
        from .. import _native as nat
        plan = self._plan()
        ns = int(min(self.num_local, self.audit_rows))
        est = 0.0
        if ns > 0:
            args = (self._dev['tables'], self._X_dev, self._y_dev, ns, solve_out['Pinv'], solve_out['b'], float(self.noise_var))
            g8 = plan.grad_theta(*args)
            mode = plan.get_option(nat.OPT_GEMM_MODE)
            plan.set_option(nat.OPT_GEMM_MODE, 0)
            try:
                g64 = plan.grad_theta(*args)
            finally:
                plan.set_option(nat.OPT_GEMM_MODE, mode)
            est = float((g8 - g64).abs().max()) * (self.num_local / float(ns))
        est = self._audit_max_over_ranks(est) * (1 if self._dist is None else self._dist.get_world_size())
        return est / max(float(np.abs(g_full).max()), 1e-300), ns

    def _stats(self):
        """A = Phi^T Phi, r = Phi^T y, s = y^T y on the device (all-reduced over ranks), cached like `_A`."""
        dev = self._dev
        if 'stats' in dev:
            return dev['stats']
        t = self._torch
        plan = self._plan()
        p = plan.p
        T = plan.build_tables(self._X_dev)
        from ..sharding import stats_layout
        lay = stats_layout(p)
        buf = t.zeros((lay["size"],), dtype=t.float64, device="cuda")
        A = buf[lay["A"][0]:lay["A"][1]].view(p, p)
        r = buf[lay["r"][0]:lay["r"][1]]
        s = buf[lay["s"][0]:lay["s"][1]]
        ws = dev.get('gram_ws')
        need = plan.gram_workspace_bytes(self.num_local)
        if ws is None or ws.numel() < need:
            ws = t.empty((need,), dtype=t.uint8, device="cuda")
            dev['gram_ws'] = ws
        # A and r = Phi^T y from one sweep; Type-II also keeps the row maxima of |Phi| for the gradient pass over the same tables
        rowmax = t.empty((T.shape[0],), dtype=t.int32, device="cuda") if self.kern.opt_kernel_params else None
        plan.gram(T, self.num_local, out=A, workspace=ws, y=self._y_dev, r_out=r, rowmax_out=rowmax)
        dev['rowmax'] = rowmax
        s.copy_(self._device_mod.sumsq(self._y_dev))
        if self._dist is not None:
            self._dist.all_reduce(buf)
        dev['tables'] = T
        dev['stats'] = dict(A=A, r=r, s=s, buf=buf)
        return dev['stats']

    def _cov_setup(self, want_grad=False, want_G2=False):
        dev = self._dev
        have = dev.get('solve')
        if have is not None and (have['Pinv'] is not None or not (want_grad or want_G2)) and \
                (have['G2'] is not None or not want_G2):
            return have
        self._w = self.kern.w
        from .. import _native as nat
        while True:
            audit = self._audit_pending('gram')
            st = self._stats()
            w_dev = self._torch.as_tensor(np.ascontiguousarray(self._w, dtype=np.float64)).cuda()
            out = self._device_mod.shared_solver().solve(st['A'], st['r'], st['s'], w_dev, float(self.noise_var), self.num_data,
                                     want_grad=want_grad or want_G2 or audit, want_G2=want_G2)
            if not audit:
                break
            rel, ns = self._audit_gram(out)
            digits = self._plan().get_option(nat.OPT_DIGITS_GRAM)
            rec = dict(self.arithmetic_audit or {})
            rec['gram'] = {"digits": digits, "rows": ns, "estimated_lml_rel_error": rel, "tol": float(self.audit_tol)}
            self.arithmetic_audit = rec
            if rel <= self.audit_tol or not self._raise_digits(0):
                self._audit_done['gram'] = True
                break
            logger.warning("INT8 Gram with %d digits: estimated relative LML error %.2e > %.1e on %d audit rows; recomputing with %d digits",
                           digits, rel, self.audit_tol, ns, digits + 1)
            for k in ('stats', 'plan_id'):          # the tables stay valid: only the product is redone
                dev.pop(k, None)
            self._host.pop('_A', None)
        dev['solve'] = out
        for k in ('_P', '_Pchol', '_alpha', '_alpha_p'):
            self._host.pop(k, None)
        return out

    def to_web_model(self):
        """GPwebModel sharing this model's device-resident statistics (Type-I inner loop / MCMC: O(p^3) per evaluation)."""
        from .gp_web_model import GPwebModel
        self.parameters
        st = self._stats()
        web = GPwebModel.from_statistics(st['A'], st['r'], st['s'], self.num_data, noise_var=self.noise_var)
        web.kern.parameters = np.array(self.kern.w, dtype=float)
        return web

    # ------------------------------------------------------------------ likelihood and gradients
    def _compute_log_likelihood(self, parameters):
        """log N(y | 0, Phi W Phi^T + noise_var I), returned as a (1, 1) array like the reference (:203-214)."""
        self.parameters = parameters
        out = self._cov_setup()
        return np.array([[out['lml']]])

    def _adjoint_gradient(self, parameters):
        """(log likelihood, gradient); NaN in the slots of fixed parameters (reference :156-200)."""
        assert isinstance(parameters, np.ndarray)
        self.parameters = parameters
        free = np.logical_not(self._fixed_indicies)
        gradient = np.zeros(parameters.shape) + np.nan
        n_base = parameters.size - 1 - self.kern.n_eigs
        theta_free = np.nonzero(free[1:1 + n_base])[0]
        need_theta = self.kern.opt_kernel_params and theta_free.size > 0
        out = self._cov_setup(want_grad=True)
        log_like = np.array([[out['lml']]])
        if self.kern.reweight_eig_funs:
            gradient[-self.kern.n_eigs:] = out['grad_w'].cpu().numpy()
        if self.noise_var_constraint != 'fixed':
            gradient[0] = out['grad_noise']
        if need_theta:
            if self.kern.has_aliased_kernels() or self.kern.has_host_kernels():
                raise NotImplementedError("analytic kernel-parameter gradient needs distinct in-house kernel objects (RBF, Exponential, "
                                          "Matern32, Matern52, no children) per dimension; use grad_method='finite_difference'")
            pmap = self.kern.base_parameter_map()
            active = [pmap[i] for i in theta_free]
            from ..kern.grief_kernel import DegenerateEigenpairError
            try:
                gradient[1 + theta_free] = self._theta_gradient(active, out)
            except DegenerateEigenpairError as e:      # repeated grid eigenvalue among the selected ones: no analytic derivative
                logger.warning("analytic kernel-parameter gradient unavailable (%s); using finite differences for this evaluation", e)
                return self._finite_diff_gradient(parameters)
        assert not np.any(np.isnan(gradient[free])), "gradient missed!"
        if not np.all(np.isfinite(gradient[free])):
            raise FloatingPointError("non-finite entries in the log-likelihood gradient at free parameters %s"
                                     % np.nonzero(free & ~np.isfinite(gradient))[0].tolist())
        return log_like, gradient

    def _theta_gradient(self, active, solve_out):
        """d LML / d theta for the active base-kernel parameters [(dim, kind)] (pass 2 on the device)."""
        t = self._torch
        from .. import _native as nat
        plan = self._plan()
        dqs = self.kern.scaled_eigvec_derivatives(active)
        plan.grad_setup([a[0] for a in active], [0 if a[1] == 'variance' else 1 for a in active], dqs)
        while True:
            g = plan.grad_theta(self._dev['tables'], self._X_dev, self._y_dev, self.num_local, solve_out['Pinv'],
                                solve_out['b'], float(self.noise_var), rowmax=self._dev.get('rowmax'))
            if self._dist is not None:
                self._dist.all_reduce(g)
            g = g.cpu().numpy()
            if not self._audit_pending('grad') or len(active) == 0:
                return g
            rel, ns = self._audit_grad(solve_out, g)
            digits = plan.get_option(nat.OPT_DIGITS_Z)
            rec = dict(self.arithmetic_audit or {})
            rec['grad'] = {"digits": digits, "rows": ns, "estimated_grad_error_over_max_abs": rel, "tol": float(self.audit_tol_grad)}
            self.arithmetic_audit = rec
            if rel <= self.audit_tol_grad or not self._raise_digits(1):
                self._audit_done['grad'] = True
                return g
            logger.warning("INT8 gradient pass with %d digits: estimated gradient error %.2e of the largest component > %.1e on %d audit "
                           "rows; recomputing with %d digits", digits, rel, self.audit_tol_grad, ns, digits + 1)
            plan = self._plan()

    # ------------------------------------------------------------------ prediction
    def predict_precompute(self, Xnew):
        logger.debug('Predicting model at new points.')
        assert Xnew.ndim == 2
        assert Xnew.shape[1] == self.input_dim
        self.parameters
        self._cov_setup()

    def predict(self, Xnew, compute_var='full'):
        """Posterior mean (M, 1) and covariance at Xnew (reference :99-125).

        compute_var: 'full' -> (M, M) covariance as the reference returns (small M only);
                     'diag' -> (M, 1) marginal variances, computed without the M x M matrix; None -> mean only.
        """
        self.predict_precompute(Xnew)
        t = self._torch
        plan = self._plan()
        out = self._cov_setup(want_grad=compute_var is not None)
        M = Xnew.shape[0]
        Xd = t.as_tensor(np.ascontiguousarray(Xnew, dtype=np.float64)).cuda()
        Tn = plan.build_tables(Xd)
        Yhat = plan.phi_vec(Tn, M, out['b']).cpu().numpy().reshape((-1, 1))       # alpha_p == b
        if compute_var is None:
            return Yhat
        nv = float(self.noise_var)
        if compute_var == 'diag':
            q = plan.quadform_rows(Tn, M, out['Pinv'])
            return Yhat, (nv * (q + 1.0)).cpu().numpy().reshape((-1, 1))
        from ..device import gemm_nt
        Phi_new = plan.phi_rows(Tn, M)
        Yhatvar = nv * t.eye(M, dtype=t.float64, device="cuda")
        gemm_nt(gemm_nt(Phi_new, out['Pinv']), Phi_new, alpha=nv, beta=1.0, out=Yhatvar)     # P^-1 is symmetric
        return Yhat, Yhatvar.cpu().numpy()

    def d_Yhat_d_x(self, Xnew, dim):
        """d Yhat / d x[:, dim] (reference :127-134) on the device: the basis tables of Xnew with the kernel of dimension `dim`
        replaced by its x-derivative, times alpha_p.  (`kern.cov_grad` keeps the host evaluation of d Phi / d x.)"""
        self.predict_precompute(Xnew)
        t = self._torch
        plan = self.kern.device_plan()
        out = self._cov_setup(want_grad=False)
        Xd = t.as_tensor(np.ascontiguousarray(Xnew, dtype=np.float64)).cuda()
        Tn = plan.build_tables(Xd, deriv_dim=int(dim))
        return plan.phi_vec(Tn, Xnew.shape[0], out['b']).cpu().numpy().reshape((-1, 1))      # alpha_p == b

    # ------------------------------------------------------------------ covariance operators (inspection-sized n)
    def _mv_cov(self, x):
        """(Phi W Phi^T + noise_var I) x."""
        assert x.shape[0] == self.num_local
        Phi = self._Phi
        assert Phi is not None, "cov has not been setup"
        return Phi.dot(Phi.T.dot(x) * self._w.reshape((-1, 1))) + x * self.noise_var

    def _mv_cov_inv(self, x):
        """(Phi W Phi^T + noise_var I)^-1 x through the matrix inversion lemma."""
        from scipy.linalg import cho_solve
        assert x.shape[0] == self.num_local
        assert self._Pchol is not None, "cov has not been setup"
        Phi = self._Phi
        return (x - Phi.dot(cho_solve(self._Pchol, Phi.T.dot(x)))) / self.noise_var

    def _cov_log_det(self):
        assert self._dev.get('solve') is not None, "cov has not been setup"
        return self._dev['solve']['logdet']
