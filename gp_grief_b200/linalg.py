"""Host-side numerical helpers of the gp_grief API (reference: gp_grief/linalg.py).

Only tiny, O(m) / O(#parameters) work lives here; nothing on the n- or p^2-sized path.
"""
import logging
import sys

import numpy as np

logger = logging.getLogger(__name__)


def log_kron(a, b, a_logged=False, b_logged=False):
    """log(kron(a, b)) of two 1-D arrays as an outer SUM of logs (reference linalg.py:74-89).

    The add order (row-major ravel of log a[:,None] + log b[None,:]) is the one the device top-p
    kernel reproduces bit for bit.
    """
    a = np.asarray(a)
    b = np.asarray(b)
    if a.ndim != 1 or b.ndim != 1:
        raise AssertionError("currently only working for 1d arrays")
    la = a if a_logged else np.log(a)
    lb = b if b_logged else np.log(b)
    return np.add.outer(la, lb).ravel()


class LogexpTransformation(object):
    """softplus re-parametrisation of positive parameters (reference linalg.py:107-125)."""
    _lim_val = 36.
    _log_lim_val = np.log(np.finfo(np.float64).max)

    def inverse_transform(self, x):
        x = np.asarray(x, dtype=float)
        safe = np.clip(x, -self._log_lim_val, self._lim_val)
        return np.where(x > self._lim_val, x, np.log1p(np.exp(safe)))

    def transform(self, f):
        f = np.asarray(f, dtype=float)
        with np.errstate(over="ignore"):
            return np.where(f > self._lim_val, f, np.log(np.expm1(f)))

    def transform_grad(self, f, grad_f):
        f = np.asarray(f, dtype=float)
        return grad_f * np.where(f > self._lim_val, 1., -np.expm1(-f))


class solver_counter(object):
    """Iteration counter / best-so-far backup used by BaseModel.optimize (reference linalg.py:53-71)."""

    def __init__(self, disp=True):
        self._disp = disp
        self.niter = 0
        self.backup = None

    def __call__(self, rk=None, msg='', store=None):
        self.niter += 1
        if self._disp:
            logger.info('iter %3i. %s' % (self.niter, msg))
            sys.stdout.flush()
        if store is not None:
            self.backup = store


def uniquetol(x, tol=1e-6, relative=False):
    """Unique values of a 1-D array up to a tolerance (reference linalg.py:92-104)."""
    x = np.asarray(x)
    assert x.ndim == 1
    if relative:
        tol = np.float64(tol) * np.ptp(x)
    close = np.abs(x[:, None] - x[None, :]) <= tol
    return x[~np.triu(close, 1).any(axis=0)]
