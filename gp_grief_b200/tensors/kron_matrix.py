"""KronMatrix: a Kronecker product of small dense factors, never expanded.

API of gp_grief/tensors/kron_matrix.py (reference :12-474).  The factors are tiny (m_i x m_i, m_i <= 64),
so everything here is host NumPy EXCEPT the one method on the GP-GRIEF hot path,
`find_extremum_eigs(mode='largest', log_expand=True)`, which runs on the GPU
(csrc/topk.cu through gp_grief_b200.device.topk_kron) and reproduces the reference bit for bit.
"""
import logging
from warnings import warn

import numpy as np
import scipy.linalg as la

from ..linalg import log_kron

logger = logging.getLogger(__name__)


def _apply_factorwise(K_list, x, op_cols):
    """Shared engine of the Kronecker mat-vec / solves.

    Walk the factors from the last to the first (mat-vec) keeping the running vector as a matrix whose
    leading axis is the factor's column space; `op_cols(i, Y)` applies factor i to the columns of Y.
    """
    y = x
    for i in reversed(range(len(K_list))):
        Y = np.reshape(y, (np.shape(K_list[i])[1], -1), order='F')
        y = op_cols(i, Y).T
    return np.reshape(y, (-1, 1), order='F')


class KronMatrix(object):
    def __init__(self, K, sym=False):
        self._K = K                                     # shallow reference, like the reference implementation
        self.n = len(self.K)
        self.sshape = np.vstack([np.shape(Ki) for Ki in self.K])
        self.shape = np.atleast_1d(np.prod(np.float64(self.sshape), axis=0))
        if np.all(self.shape < np.iinfo(np.uint64).max):
            self.shape = np.uint64(self.shape)          # else keep the float shape (overflowing grids)
        self.ndim = self.shape.size
        assert self.ndim <= 2, "kron matrix cannot be more than 2d"
        self.square = self.ndim == 2 and self.shape[0] == self.shape[1]
        self.sym = sym
        if sym:
            assert np.array_equal(self.sshape[:, 0], self.sshape[:, 1]), \
                'this matrix cannot be symmetric: it is not square'
            self.ensure_fortran()

    @property
    def K(self):
        return self._K

    @K.setter
    def K(self, K):
        raise AttributeError("Attribute is Read only.")

    # ---------------------------------------------------------------- products
    def kronvec_prod(self, x, device=False):
        """K x for a column vector x (M,1) -> (N,1).  device=True runs the per-factor products on the GPU (dense factors only)."""
        if x.shape != (self.shape[1], 1):
            raise ValueError('x is the wrong shape, must be (%d,1), not %s' % (self.shape[1], repr(x.shape)))
        if device:
            assert all(isinstance(Ki, np.ndarray) and Ki.ndim == 2 for Ki in self.K), "device mat-vec needs dense 2-D factors"
            from ..device import kron_matvec
            return kron_matvec(list(self.K), x)

        def apply(i, Y):
            Ki = self.K[i]
            return np.asarray(Ki).dot(Y) if isinstance(Ki, np.ndarray) else Ki * Y
        return _apply_factorwise(self.K, x, apply)

    def __mul__(self, x):
        return self.kronvec_prod(x)

    def kronkron_prod(self, X):
        if not isinstance(X, KronMatrix):
            raise TypeError("X is not a KronMatrix")
        if X.n != self.n:
            raise TypeError('inconsistent kron structure')
        if not np.array_equal(X.sshape[:, 0], self.sshape[:, 1]):
            raise TypeError("Dimensions of X submatricies are not consistent")
        return KronMatrix([self.K[i].dot(X.K[i]) for i in range(self.n)])

    def kronvec_div(self, x):
        """K^-1 x, factor by factor."""
        assert self.ndim == 2
        if x.shape != (self.shape[0], 1):
            raise ValueError('x wrong shape, must be (%d,1)' % self.shape[0])
        y = x
        for i, Ki in enumerate(self.K):
            Y = np.reshape(y, (-1, self.sshape[i, 0]), order='F')
            if hasattr(Ki, "solve"):
                y = Ki.solve(b=Y.T)
            else:
                y = la.solve(Ki, Y.T, assume_a='pos' if self.sym else 'gen')
        return y.reshape((-1, 1), order='F')

    # ---------------------------------------------------------------- per-factor decompositions
    def _map(self, method, fallback):
        out = np.empty(self.n, dtype=object)
        for i, Ki in enumerate(self.K):
            out[i] = getattr(Ki, method)() if hasattr(Ki, method) else fallback(Ki)
        return out

    def chol(self):
        """Upper-triangular Cholesky factor of every factor."""
        assert self.square
        return KronMatrix(self._map("chol", lambda Ki: np.linalg.cholesky(Ki).T))

    def schur(self):
        """(Q, T) with K = Q T Q^T per factor (LAPACK dgees through scipy.linalg.schur, as the reference)."""
        assert self.square
        T = np.empty(self.n, dtype=object)
        Q = np.empty(self.n, dtype=object)
        for i, Ki in enumerate(self.K):
            T[i], Q[i] = Ki.schur() if hasattr(Ki, "schur") else la.schur(Ki)
        return KronMatrix(Q), KronMatrix(T)

    def svd(self):
        assert self.square, "matrix must be square for current implementation"
        Q = np.empty(self.n, dtype=object)
        s = np.empty(self.n, dtype=object)
        for i, Ki in enumerate(self.K):
            try:
                Q[i], s[i] = Ki.svd() if hasattr(Ki, "svd") else np.linalg.svd(Ki, full_matrices=0, compute_uv=1)[:2]
            except np.linalg.LinAlgError:
                logger.error('SVD failed on dimension %d.' % i)
                raise
        return KronMatrix(Q), KronMatrix(s)

    def transpose(self):
        assert self.ndim == 2
        return self if self.sym else KronMatrix([Ki.T for Ki in self.K])
    T = property(transpose)

    def expand(self, log_expansion=False):
        """Dense matrix (or vector).  Expensive; `log_expansion` returns the log of a 1-D product."""
        if log_expansion:
            Kb = np.array([0.])
            for Ki in self.K:
                Kb = log_kron(a=Kb, b=Ki.expand() if hasattr(Ki, "expand") else Ki, a_logged=True)
        else:
            Kb = 1.
            if self.ndim == 1 and self.n > 10:
                warn('consider using numerically more stable log_expansion')
            for Ki in self.K:
                Kb = np.kron(Kb, Ki.expand() if hasattr(Ki, "expand") else Ki)
        return Kb.reshape(np.int32(self.shape))

    def inv(self):
        assert self.square
        return KronMatrix(self._map("inv", np.linalg.inv))

    def diag(self):
        """Diagonal as a 1-D KronMatrix."""
        assert self.ndim == 2
        return KronMatrix(self._map("diag", np.diag))

    def sub_cond(self):
        assert self.square
        return [np.linalg.cond(Ki) for Ki in self.K]

    def sub_shift(self, shift=1e-6):
        """Add shift*I to every factor, in place (conditioning)."""
        if not np.array_equal(self.sshape[:, 0], self.sshape[:, 1]):
            raise RuntimeError('can only apply sub_shift for square matricies')
        for i, Ki in enumerate(self.K):
            self.K[i] = Ki + shift * np.identity(self.sshape[i, 0])
        if self.sym:
            self.ensure_fortran()
        return self

    def ensure_fortran(self):
        for i, Ki in enumerate(self.K):
            if isinstance(Ki, np.ndarray):
                self.K[i] = np.asarray(Ki, order='F')
        return self

    # ---------------------------------------------------------------- solves
    def solve_chol(U, x):
        """U \\ (U' \\ x) for an upper-triangular Cholesky KronMatrix U."""
        if x.shape != (U.shape[0], 1):
            raise ValueError('x wrong shape, must be (%d,1)' % U.shape[0])
        y = x
        for i, Ui in enumerate(U.K):
            Y = np.reshape(y, (-1, U.sshape[i, 0]), order='F')
            if hasattr(Ui, "solve_chol"):
                y = Ui.solve_chol(Y.T)
            else:
                y = la.solve_triangular(Ui, Y.T, trans='T', lower=False)
                y = la.solve_triangular(Ui, y, trans='N', lower=False)
        return y.reshape((-1, 1), order='F')

    def solve_schur(Q, t, x, shift=0.0):
        """(K + shift I)^-1 x from K = Q diag(t) Q^T."""
        if x.shape != (Q.shape[0], 1):
            raise ValueError('x wrong shape, must be (%d,1)' % Q.shape[0])
        if isinstance(t, KronMatrix):
            t = t.diag().expand()
        y = (Q.T) * x
        y = y / np.reshape(t + shift, y.shape)
        return Q * y

    def eig_vals(self):
        assert self.ndim == 2
        fb = np.linalg.eigvalsh if self.sym else np.linalg.eigvals
        return KronMatrix(self._map("eig_vals", fb))

    # ---------------------------------------------------------------- extreme eigenvalues
    def find_extremum_eigs(eigs, n_eigs, mode='largest', log_expand=False, sort=True, compute_global_loc=False):
        """Positions / values of the n_eigs extreme entries of a 1-D KronMatrix (reference :369-446).

        mode='largest' with log_expand=True and sort=True -- the call GriefKernel makes -- runs on the GPU.
        The other three combinations are API surface only (never on the GRIEF path) and use host NumPy.
        Returns (eig_loc (n_eigs, d), eig_vals (n_eigs,), global_loc or None).
        """
        assert eigs.ndim == 1, "eigs must be a 1D KronMatrix"
        assert isinstance(n_eigs, (int, np.integer)), "n_eigs=%s must be an integer" % repr(n_eigs)
        assert n_eigs >= 1, "must use at least 1 eigenvalue"
        assert n_eigs <= eigs.shape[0], "n_eigs > number of eigenvalues"
        assert mode == 'largest' or mode == 'smallest'
        if not log_expand and eigs.n > 10:
            warn('should use log option which will be more numerically stable')
        if mode == 'largest' and log_expand and sort:
            from ..device import topk_kron
            eig_loc, eig_vals = topk_kron([np.asarray(Ki) for Ki in eigs.K], int(n_eigs))
        else:
            eig_loc, eig_vals = _beam_search_host([np.asarray(Ki) for Ki in eigs.K], int(n_eigs), mode, log_expand, sort)
        global_loc = None
        if compute_global_loc:
            global_loc = np.zeros(eig_loc.shape[0], dtype=int)
            span = 1
            for i in reversed(range(eigs.n)):
                global_loc = span * eig_loc[:, i] + global_loc
                span *= eigs.K[i].size
        return eig_loc, eig_vals, global_loc

    def get_col(self, pos):
        """Column `pos` (one index per factor) as a KronMatrix of column vectors."""
        assert len(pos) == self.n
        assert np.size(pos[0]) == 1
        assert isinstance(pos[0], (int, np.integer))
        assert self.ndim == 2
        return KronMatrix([self.K[i][:, j].reshape((-1, 1)) for i, j in enumerate(pos)])

    def log_det(eig_vals):
        """log-determinant from a 1-D KronMatrix of per-factor eigenvalues."""
        assert eig_vals.ndim == 1
        sizes = eig_vals.sshape.reshape(-1).astype(float)
        total = np.prod(sizes)
        return float(sum((total / sizes[i]) * np.sum(np.log(e)) for i, e in enumerate(eig_vals.K)))


def _beam_search_host(factors, n_eigs, mode, log_expand, sort):
    """Host beam search for the non-hot-path modes of find_extremum_eigs (smallest / non-log)."""
    largest = (mode == 'largest')

    def keep(vec):
        if vec.size <= n_eigs:
            return np.arange(vec.size), vec
        idx = np.argpartition(vec, -n_eigs)[-n_eigs:] if largest else np.argpartition(vec, n_eigs)[:n_eigs]
        return idx, vec[idx]

    idx, vals = keep(factors[0])
    loc = idx.reshape((-1, 1))
    if log_expand:
        vals = np.log(vals)
    for f in factors[1:]:
        cand = log_kron(vals, f, a_logged=True) if log_expand else np.kron(vals, f)
        idx, vals = keep(cand)
        loc = np.hstack([loc[idx // f.size], (idx % f.size).reshape((-1, 1))])
    if sort:
        order = np.argsort(vals)[::-1]
        vals, loc = vals[order], loc[order]
    return loc.astype(np.int64), vals
