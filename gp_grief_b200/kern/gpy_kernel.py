"""Kernels borrowed from the GPy library (reference: gp_grief/kern/gpy_kernel.py).

GPy is an optional dependency: it is imported when a GPyKernel is constructed, not when this package is.  A GPy kernel has no
formula inside libgrief_b200, so a GriefKernel built on it takes the host-evaluated route: `cov` below is called per chunk of data
rows for K_xu,i, the (rows, m_i) block is uploaded and every later stage (projection on the grid eigenvectors, Gram, solve,
prediction) is the same device code as for the in-house kernels (DevicePlan._build_tables_host_kxu, grief_build_tables_kxu).
Kernel hyper-parameters of a GPy kernel are differentiated by finite differences, as in the reference (models/basemodel.py:328-361).
"""
import logging

import numpy as np

from .basekernel import BaseKernel

logger = logging.getLogger(__name__)


def _gpy():
    try:
        import GPy
    except ImportError as e:                                       # pragma: no cover - depends on the environment
        raise ImportError("GPyKernel needs the GPy package (pip install GPy); the in-house kernels RBF, Exponential, Matern32 and "
                          "Matern52 do not") from e
    return GPy


class GPyKernel(BaseKernel):
    """A GPy covariance function behind the BaseKernel interface (cov, parameters, constraints, fix_variance)."""

    device_id = None              # evaluated on the host

    def __init__(self, n_dims, kernel=None, name=None, **kwargs):
        """kernel: the name of a class in GPy.kern (constructed with input_dim=n_dims and **kwargs) or a GPy.kern.Kern instance."""
        GPy = _gpy()
        if isinstance(kernel, str):
            label = "GPy - " + kernel
            gpy_kern = getattr(GPy.kern, kernel)(input_dim=n_dims, **kwargs)
        elif isinstance(kernel, GPy.kern.Kern):
            label = "GPy - " + repr(kernel)
            gpy_kern = kernel
        else:
            raise TypeError("must specify kernel as str or a GPy kernel object")
        super(GPyKernel, self).__init__(n_dims=n_dims, active_dims=None, name=label if name is None else name)
        logger.debug('Using the %s kernel.', self.name)
        self.kern = gpy_kern
        self.constraint_list = [['+ve', ] * np.size(prm.values) for prm in self.kern.flattened_parameters]

    def cov(self, x, z=None):
        """(N, M) covariance matrix between the rows of x (N, d) and z (M, d); z = x when omitted."""
        return self._apply_children(self.kern.K(x, z), x, z)

    def grad_x(self, x, z):
        """d cov(x, z) / d x for 1-d inputs, column by column through GPy's gradients_X (kern/grid_kernel.py:193-199)."""
        assert self.n_dims == 1 and not self._children
        out = np.zeros((x.shape[0], z.shape[0]))
        for j in range(z.shape[0]):
            out[:, [j]] = self.kern.gradients_X(1, x, z[[j]])
        return out

    # ---- flat parameter vector: the GPy parameters in GPy's order, then the children's ----
    @property
    def parameters(self):
        parts = [np.ravel(prm.values) for prm in self.kern.flattened_parameters]
        parts += [child.parameters for _, child in self._children]
        return np.concatenate(parts, axis=0) if parts else np.array([])

    @parameters.setter
    def parameters(self, value):
        assert isinstance(value, np.ndarray)
        assert value.ndim == 1
        pos = 0
        for prm in self.kern.flattened_parameters:
            size = np.size(prm)
            prm[:] = value[pos:pos + size].reshape(np.shape(prm))
            pos += size
        for _, child in self._children:
            size = np.size(child.parameters)
            child.parameters = value[pos:pos + size]
            pos += size

    @property
    def constraints(self):
        parts = [np.ravel(c) for c in self.constraint_list]
        parts += [child.constraints for _, child in self._children]
        return np.concatenate(parts, axis=0) if parts else np.array([])

    def fix_variance(self):
        """Fix the (first) variance parameter: a product of kernels needs one free amplitude only."""
        hits = [i for i, prm in enumerate(self.kern.flattened_parameters) if 'variance' in prm._name.lower()]
        if not hits:
            raise RuntimeError("No variance parameter found")
        if len(hits) > 1 or np.size(self.constraint_list[hits[0]]) > 1:
            logger.info("Multiple variance parameters found in the GPy kernel, will only fix the first")
        self.constraint_list[hits[0]][0] = 'fixed'
