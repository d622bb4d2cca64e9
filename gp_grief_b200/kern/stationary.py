"""In-house stationary kernels: RBF, Exponential, Matern32, Matern52.

Same formulas, constructor arguments and attributes as gp_grief/kern/stationary.py of the reference
(RBF :108-134, Exponential :161-175, Matern32 :202-216, Matern52 :243-258).  `cov` is host NumPy and
is only used on the tiny per-dimension grid matrices (m_i x m_i) and by user code; the n-sized
evaluations k(x_n, U) run inside the CUDA prepass (csrc/rows.cu), which implements the same formulas.
`grad_lengthscale` / `grad_variance` are additions (the reference has no analytic kernel-parameter
derivatives): they feed the eigen-perturbation used by the analytic hyper-parameter gradient.
"""
import logging

import numpy as np

from .basekernel import BaseKernel

logger = logging.getLogger(__name__)


class Stationary(BaseKernel):
    """Base class: distance helpers shared by the stationary kernels."""

    def _scaled_diff(self, x, z, lengthscale):
        x, z = self._process_cov_inputs(x, z)
        d = self.active_dims.size
        xa = np.asarray(x)[:, self.active_dims].reshape((x.shape[0], 1, d))
        za = np.asarray(z)[:, self.active_dims].reshape((1, z.shape[0], d))
        if lengthscale is None:
            ls = np.ones(d, dtype='d')
        elif isinstance(lengthscale, float):
            ls = lengthscale * np.ones(d, dtype='d')
        else:
            ls = np.asarray(lengthscale).flatten()
            assert len(ls) == d
        return (xa - za) / ls.reshape((1, 1, d))

    def distances_squared(self, x, z=None, lengthscale=None):
        """(N, M) squared distances (reference stationary.py:12-43)."""
        return np.sum(np.power(self._scaled_diff(x, z, lengthscale), 2), axis=2, keepdims=False)

    def distances(self, x, z=None, lengthscale=None):
        """(N, M, d) signed per-dimension distances (reference stationary.py:46-76)."""
        return self._scaled_diff(x, z, lengthscale)

    def _init_params(self, variance, lengthscale):
        assert np.size(variance) == 1
        assert np.size(lengthscale) == 1
        self.variance = np.float64(variance)
        self.lengthscale = np.float64(lengthscale)
        self.parameter_list = ['variance', 'lengthscale']
        self.constraint_map = {'variance': '+ve', 'lengthscale': '+ve'}

    device_id = None   # id of the kernel formula inside libgrief_b200 (None: not available on the device)

    def grad_variance(self, x, z=None):
        """d cov / d variance (no children)."""
        return self._base_cov(x, z) / self.variance

    def grad_lengthscale(self, x, z=None):
        raise NotImplementedError

    def grad_x(self, x, z):
        """d k(x, z) / d x for one-dimensional inputs, shape (N, M)."""
        raise NotImplementedError

    def _signed_diff_1d(self, x, z):
        assert self.active_dims.size == 1, "grad_x is implemented for 1-d kernels"
        return self.distances(x, z)[:, :, 0]


class RBF(Stationary):
    """Squared exponential kernel, one lengthscale shared by the active dimensions."""
    device_id = "RBF"

    def __init__(self, n_dims, variance=1., lengthscale=1., active_dims=None, name=None):
        super(RBF, self).__init__(n_dims=n_dims, active_dims=active_dims, name=name)
        logger.debug('Initializing %s kernel.' % self.name)
        self._init_params(variance, lengthscale)

    def _base_cov(self, x, z=None, lengthscale=None):
        if self.lengthscale < 1e-6:   # guard against division by ~0 (reference :121-122)
            logger.debug('protected RBF against zero-division since lengthscale too small (%s).' % repr(self.lengthscale))
            return self.variance * (self.distances_squared(x=x, z=z) == 0)
        if lengthscale is None:
            return self.variance * np.exp(-0.5 * self.distances_squared(x=x, z=z) / self.lengthscale ** 2)
        lengthscale = np.asarray(lengthscale).flatten()
        assert len(lengthscale) == self.active_dims.size
        return self.variance * np.exp(-0.5 * self.distances_squared(x=x, z=z, lengthscale=lengthscale))

    def cov(self, x, z=None, lengthscale=None):
        return self._apply_children(self._base_cov(x, z, lengthscale), x, z)

    def grad_lengthscale(self, x, z=None):
        if self.lengthscale < 1e-6:
            return np.zeros((x.shape[0], (x if z is None else z).shape[0]))
        return self._base_cov(x, z) * self.distances_squared(x=x, z=z) / self.lengthscale ** 3

    def grad_x(self, x, z):
        return -self._base_cov(x, z) * self._signed_diff_1d(x, z) / self.lengthscale ** 2


class Exponential(Stationary):
    device_id = "Exponential"

    def __init__(self, n_dims, variance=1., lengthscale=1., active_dims=None, name=None):
        super(Exponential, self).__init__(n_dims=n_dims, active_dims=active_dims, name=name)
        logger.debug('Initializing %s kernel.' % self.name)
        self._init_params(variance, lengthscale)

    def _base_cov(self, x, z=None):
        r = np.sqrt(self.distances_squared(x=x, z=z)) / self.lengthscale
        return self.variance * np.exp(-r)

    def cov(self, x, z=None):
        return self._apply_children(self._base_cov(x, z), x, z)

    def grad_lengthscale(self, x, z=None):
        r = np.sqrt(self.distances_squared(x=x, z=z)) / self.lengthscale
        return self.variance * np.exp(-r) * r / self.lengthscale

    def grad_x(self, x, z):
        diff = self._signed_diff_1d(x, z)
        return -self._base_cov(x, z) * np.sign(diff) / self.lengthscale


class Matern32(Stationary):
    device_id = "Matern32"

    def __init__(self, n_dims, variance=1., lengthscale=1., active_dims=None, name=None):
        super(Matern32, self).__init__(n_dims=n_dims, active_dims=active_dims, name=name)
        logger.debug('Initializing %s kernel.' % self.name)
        self._init_params(variance, lengthscale)

    def _base_cov(self, x, z=None):
        r = np.sqrt(self.distances_squared(x=x, z=z)) / self.lengthscale
        return self.variance * (1. + np.sqrt(3.) * r) * np.exp(-np.sqrt(3.) * r)

    def cov(self, x, z=None):
        return self._apply_children(self._base_cov(x, z), x, z)

    def grad_lengthscale(self, x, z=None):
        r = np.sqrt(self.distances_squared(x=x, z=z)) / self.lengthscale
        return self.variance * 3. * r * r * np.exp(-np.sqrt(3.) * r) / self.lengthscale

    def grad_x(self, x, z):
        diff = self._signed_diff_1d(x, z)
        r = np.abs(diff) / self.lengthscale
        return -self.variance * 3. * diff * np.exp(-np.sqrt(3.) * r) / self.lengthscale ** 2


class Matern52(Stationary):
    device_id = "Matern52"

    def __init__(self, n_dims, variance=1., lengthscale=1., active_dims=None, name=None):
        super(Matern52, self).__init__(n_dims=n_dims, active_dims=active_dims, name=name)
        logger.debug('Initializing %s kernel.' % self.name)
        self._init_params(variance, lengthscale)

    def _base_cov(self, x, z=None):
        r2 = self.distances_squared(x=x, z=z) / self.lengthscale ** 2
        r = np.sqrt(r2)
        return self.variance * (1. + np.sqrt(5.) * r + (5. / 3) * r2) * np.exp(-np.sqrt(5.) * r)

    def cov(self, x, z=None):
        return self._apply_children(self._base_cov(x, z), x, z)

    def grad_lengthscale(self, x, z=None):
        r2 = self.distances_squared(x=x, z=z) / self.lengthscale ** 2
        r = np.sqrt(r2)
        return self.variance * np.exp(-np.sqrt(5.) * r) * (5. / 3) * r2 * (1. + np.sqrt(5.) * r) / self.lengthscale

    def grad_x(self, x, z):
        diff = self._signed_diff_1d(x, z)
        r = np.abs(diff) / self.lengthscale
        return -self.variance * (5. / 3) * diff * (1. + np.sqrt(5.) * r) * np.exp(-np.sqrt(5.) * r) / self.lengthscale ** 2
