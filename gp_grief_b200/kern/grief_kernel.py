"""GRIEF kernel: grid-structured eigenfunction basis (reference: gp_grief/kern/grief_kernel.py).

Host side of the drop-in boundary.  What stays on the host (and why): the d tiny grid covariance
matrices and their Schur factorisations -- O(d m^3), microseconds, and they must be the SAME SciPy/LAPACK
call as the reference (kron_matrix.py:170) so that the eigenvalue inputs of the top-p selection are
bit-identical.  Everything that scales with p or n runs on the GPU:
  * top-p selection over the Kronecker eigenvalues          -> csrc/topk.cu      (KronMatrix.find_extremum_eigs)
  * per-row basis factors k(x, U) Q lambda^-1/2 and their group products -> csrc/rows.cu  (DevicePlan.build_tables)
  * Phi rows (only when user code asks for the matrix itself) -> csrc/rows.cu      (DevicePlan.phi_rows)
"""
import logging

import numpy as np

from ..grid import InducingGrid
from ..tensors import SelectionMatrixSparse
from .grid_kernel import GridKernel

logger = logging.getLogger(__name__)


class DegenerateEigenpairError(ArithmeticError):
    """A selected grid eigenpair has a (numerically) repeated eigenvalue: its eigenvector derivative does not exist."""


class GriefKernel(GridKernel):
    """Kernel  k(x, z) = Phi(x) diag(w) Phi(z)^T  with Phi the p leading grid eigenfunctions.

    Constructor and attributes as in the reference (grief_kernel.py:17-65): kern_list, grid, n_eigs,
    reweight_eig_funs, opt_kernel_params, w, w_constraints, dim_noise_var, log_KRrowcol, _Quu, _log_lam,
    _Sp, _old_base_kern_params.
    """

    def __init__(self, kern_list, grid, n_eigs=1000, reweight_eig_funs=True, opt_kernel_params=False, w=1.,
                 dim_noise_var=1e-12, log_KRrowcol=True, **kwargs):
        self.reweight_eig_funs = bool(reweight_eig_funs)
        self.opt_kernel_params = bool(opt_kernel_params)
        super(GriefKernel, self).__init__(kern_list=kern_list, **kwargs)
        assert isinstance(grid, InducingGrid), "must be an InducingGrid"
        assert grid.input_dim == self.n_dims, "number of dimensions do not match"
        self.grid = grid
        self.dim_noise_var = float(dim_noise_var)
        self.n_eigs = int(min(n_eigs, self.grid.num_data))
        if not self.opt_kernel_params:       # the base-kernel hyper-parameters are not optimised: fix them all
            for kern in self.kern_list:
                if hasattr(kern, "constraint_list"):
                    kern.constraint_list = np.tile('fixed', np.shape(kern.constraint_list))
                else:
                    for key in kern.constraint_map:
                        kern.constraint_map[key] = np.tile('fixed', np.shape(kern.constraint_map[key]))
        self.w_constraints = np.array(['+ve' if self.reweight_eig_funs else 'fixed', ] * self.n_eigs)
        if np.ndim(w) == 0:                  # scalar initial weight (the reference's `w == 1.` test breaks for arrays)
            self.w = np.full(self.n_eigs, float(w))
        else:
            w = np.asarray(w, dtype=float)
            assert w.shape == (self.n_eigs,)
            assert np.all(w > 0.), "w's must be positive"
            self.w = w
        assert np.all(self.w > 0.), "w's must be positive"
        self._old_base_kern_params = None
        self.log_KRrowcol = log_KRrowcol
        self._plan = None
        self._eigs = None

    # ------------------------------------------------------------------ basis evaluation
    def cov(self, x, z=None):
        """(Phi_L, w, Phi_R) with Phi evaluated at x (and z); K = Phi_L diag(w) Phi_R^T.

        The matrices are materialised on the GPU and copied to the host: use this for inspection and small
        problems.  GPGriefModel never materialises Phi.
        """
        assert x.shape[1] == self.n_dims
        if z is not None:
            return self.cov(x=x)[0], self.w, self.cov(x=z)[0]
        import torch
        plan = self.device_plan()
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).cuda()
        T = plan.build_tables(xd)
        Phi = plan.phi_rows(T, x.shape[0]).cpu().numpy()
        return Phi, self.w, Phi

    def cov_grad(self, x, grad_dim):
        """d Phi / d x[:, grad_dim]  (reference :113-126; host evaluation, inspection-sized inputs)."""
        from ..tensors import expand_SKC
        self._setup_inducing_cov()
        dKxu = self.cov_kr_grad(x, self.grid.xg, grad_dim)
        dKux = [k.T for k in dKxu]
        log_matrix, sign = expand_SKC(S=self._Sp, K=self._Quu.T.K, C=dKux)
        return sign.T * np.exp(log_matrix.T - 0.5 * self._log_lam.reshape((1, -1)))

    # ------------------------------------------------------------------ parameters: [base kernels..., w]
    @property
    def parameters(self):
        return np.concatenate([super(GriefKernel, self).parameters, self.w], axis=0)

    @parameters.setter
    def parameters(self, value):
        n_theta = value.size - self.n_eigs
        GridKernel.parameters.fset(self, value[:n_theta])
        self.w = value[n_theta:]

    @property
    def constraints(self):
        return np.concatenate([super(GriefKernel, self).constraints, self.w_constraints], axis=0)

    @property
    def diag_val(self):
        raise NotImplementedError('')

    # ------------------------------------------------------------------ eigen setup on the inducing grid
    def _setup_inducing_cov(self):
        """Grid covariance factors, their Schur forms and the p leading Kronecker eigenpairs.

        Cached on the base-kernel parameters exactly like the reference (:171-173).
        """
        base = super(GriefKernel, self).parameters
        if self._old_base_kern_params is not None and np.array_equal(self._old_base_kern_params, base):
            return
        Kuu = self.cov_grid(self.grid.xg, dim_noise_var=self.dim_noise_var)
        self._Quu, T = Kuu.schur()
        eigs = T.diag()
        n_eigs = int(min(self.n_eigs, eigs.shape[0]))
        eig_pos, self._log_lam = eigs.find_extremum_eigs(n_eigs=n_eigs, mode='largest', log_expand=True)[:2]   # GPU
        self._Sp = [SelectionMatrixSparse((np.ascontiguousarray(col), Kuu.K[i].shape[0])) for i, col in enumerate(eig_pos.T)]
        self._eigs = eigs
        self._eig_pos = eig_pos
        self._plan = None
        self._old_base_kern_params = base

    # ------------------------------------------------------------------ device side
    def _device_kernel_spec(self):
        """(names, variances, lengthscales, host_kernels): 1-d RBF / Exponential / Matern32 / Matern52 kernels without children are
        evaluated inside libgrief_b200; every other BaseKernel (GPyKernel, kernels with children, user subclasses) keeps its host
        `cov` -- its K_xu columns are evaluated per row chunk on the host and uploaded (DevicePlan._build_tables_host_kxu), the rest
        of the path is the same device code (reference: kern/grid_kernel.py:171 calls kern.cov for whatever the kernel is)."""
        names, var, ls, host = [], [], [], {}
        for i, k in enumerate(self.kern_list):
            if k.n_dims != 1:
                raise NotImplementedError("currently only 1-dimensional grids allowed (kernel %s has n_dims=%d)" % (k.name, k.n_dims))
            if getattr(k, "device_id", None) is None or len(k._children) > 0:
                names.append("host"); var.append(1.0); ls.append(1.0)
                host[i] = k
            else:
                names.append(k.device_id)
                var.append(float(k.variance))
                ls.append(float(k.lengthscale))
        return names, var, ls, host

    def has_host_kernels(self):
        """True when a dimension's kernel is evaluated on the host (no analytic kernel-parameter gradient then)."""
        return any(getattr(k, "device_id", None) is None or len(k._children) > 0 for k in self.kern_list)

    def device_plan(self):
        """The DevicePlan (basis description resident on the GPU) for the current hyper-parameters."""
        self._setup_inducing_cov()
        if self._plan is None:
            from ..device import DevicePlan
            d = self.grid_dim
            names, var, ls, host = self._device_kernel_spec()
            xg = [np.asarray(self.grid.xg[i], dtype=float).reshape(-1) for i in range(d)]
            Q = [np.asarray(self._Quu.K[d - 1 - i]) for i in range(d)]          # factor k <-> input dimension d-1-k
            eig = [np.asarray(self._eigs.K[d - 1 - i]) for i in range(d)]
            loc = np.asarray(self._eig_pos)[:, ::-1]
            self._plan = DevicePlan(names, var, ls, xg, Q, eig, loc, host_kernels=host)
        return self._plan

    def base_parameter_map(self):
        """[(input dimension, 'variance' | 'lengthscale')] for every entry of the base-kernel parameter vector."""
        out = []
        for i, k in enumerate(self.kern_list):
            for name in k.parameter_list:
                for _ in range(np.size(getattr(k, name))):
                    out.append((i, name))
        return out

    def has_aliased_kernels(self):
        return len({id(k) for k in self.kern_list}) != len(self.kern_list)

    def scaled_eigvec_derivatives(self, active):
        """d(qs_i)/d(theta) for the active parameters [(dim, kind)], kind in {'variance','lengthscale'}.

        qs_i[:, k] = q_k / sqrt(lambda_k) are the scaled Schur vectors handed to the device plan.  First-order
        perturbation of the symmetric m_i x m_i eigenproblem K_uu,i = Q diag(lambda) Q^T:
            d lambda_k = q_k^T dK q_k,     d q_k = sum_{l != k} q_l (q_l^T dK q_k) / (lambda_k - lambda_l).
        The expansion needs simple eigenvalues.  Grid eigenvalues saturate at dim_noise_var after a few modes, so a selected
        index can sit in a cluster whose gaps are below the rounding noise of the eigen-solver (eps * max lambda); there the
        eigenvectors are an arbitrary rotation and the derivative is meaningless: DegenerateEigenpairError is raised (the model
        then falls back to finite differences) instead of returning inf / NaN.
        """
        plan = self.device_plan()
        d = self.grid_dim
        out = []
        for (i, kind) in active:
            kern = self.kern_list[i]
            xg = np.asarray(self.grid.xg[i], dtype=float).reshape(-1, 1)
            dK = kern.grad_variance(xg, xg) if kind == 'variance' else kern.grad_lengthscale(xg, xg)
            Q = np.asarray(self._Quu.K[d - 1 - i])
            lam = np.asarray(self._eigs.K[d - 1 - i])
            M = Q.T.dot(dK).dot(Q)
            uniq = plan.unique[i]
            dqs = np.zeros((Q.shape[0], uniq.size))
            gap_floor = 64.0 * np.finfo(float).eps * float(np.max(np.abs(lam)))
            for c, k in enumerate(uniq):
                gap = lam[k] - lam
                gap[k] = 1.0
                close = np.nonzero(np.abs(gap) < gap_floor)[0]
                if close.size:
                    raise DegenerateEigenpairError(
                        "input dimension %d: selected grid eigenvalue %d (%.3e) is within %.1e of eigenvalue %d; the analytic "
                        "kernel-parameter gradient needs simple eigenvalues" % (i, int(k), lam[k], gap_floor, int(close[0])))
                coef = M[:, k] / gap
                coef[k] = 0.0
                dq = Q.dot(coef)
                dqs[:, c] = dq / np.sqrt(lam[k]) - 0.5 * Q[:, k] * lam[k] ** (-1.5) * M[k, k]
            out.append(np.ascontiguousarray(dqs))
        return out
