"""Product of 1-D kernels on a Cartesian grid (reference: gp_grief/kern/grid_kernel.py)."""
import logging

import numpy as np

from ..tensors import KronMatrix, KhatriRaoMatrix

logger = logging.getLogger(__name__)


class GridKernel(object):
    """One kernel per grid dimension; the covariance is their product.

    Same constructor, attributes (kern_list, grid_dim, radial_kernel, n_dims) and methods as the reference.
    Dimension-order convention kept from the reference: `cov_grid` and `cov_kr` return their factors in
    REVERSED dimension order (grid_kernel.py:109,174), so factor k belongs to input dimension d-1-k.
    """

    def __init__(self, kern_list, radial_kernel=False):
        self.kern_list = kern_list
        self.grid_dim = len(kern_list)
        assert isinstance(radial_kernel, bool)
        self.radial_kernel = radial_kernel
        if radial_kernel:
            for kern in self.kern_list:
                assert kern.n_dims == self.kern_list[0].n_dims, "number of grid dims must be equal for all slices"
            self.kern_list = [self.kern_list[0], ] * np.size(kern_list)
        else:
            # one free amplitude is enough for a product: fix the variance of every kernel but the first
            for kern in self.kern_list[1:]:
                if hasattr(kern, 'fix_variance'):
                    kern.fix_variance()
                elif np.size(kern.constraint_map['variance']) > 1:
                    logger.info("Multiple variance parameters found in the kernel, will only fix the first")
                    kern.constraint_map['variance'][0] = 'fixed'
                else:
                    kern.constraint_map['variance'] = 'fixed'
        self.n_dims = np.sum([kern.n_dims for kern in self.kern_list])

    def cov_grid(self, x, z=None, dim_noise_var=None, use_toeplitz=False):
        """Covariance between two grids as a KronMatrix (factors reversed), diagonal shifted by dim_noise_var."""
        assert dim_noise_var is not None, "dim_noise_var must be specified"
        if isinstance(use_toeplitz, bool):
            use_toeplitz = [use_toeplitz, ] * self.grid_dim
        else:
            assert np.size(use_toeplitz) == self.grid_dim
        if np.any(use_toeplitz):
            assert z is None, "toeplitz can only be used where the (square) covariance matrix is being computed"
        assert len(x) == self.grid_dim
        cross = z is not None
        if cross:
            assert len(z) == self.grid_dim
        factors = []
        for i, (kern, toep) in enumerate(zip(self.kern_list, use_toeplitz)):
            zi = z[i] if cross else None
            if toep and zi is None:
                factors.append(kern.cov_toeplitz(x=x[i]))
            else:
                factors.append(kern.cov(x=x[i], z=zi))
        K = KronMatrix(factors[::-1], sym=not cross)
        if dim_noise_var != 0.:
            assert not cross, "not implemented for cross covariances yet"
            K = K.sub_shift(shift=dim_noise_var)
        return K

    def cov(self, x, z=None, dim_noise_var=None):
        """Dense (N, M) covariance: Hadamard product over the dimensions."""
        assert dim_noise_var is None, "currenly no way to add dim_noise_var"
        K = None
        lo = 0
        for kern in self.kern_list:
            hi = lo + kern.n_dims
            Ki = kern.cov(x=x[:, lo:hi], z=None if z is None else z[:, lo:hi])
            K = Ki if K is None else np.multiply(K, Ki)
            lo = hi
        return K

    def cov_kr(self, x, z, dim_noise_var=None, form_kr=True):
        """Cross covariance between scattered x (N, d) and a grid z as a row-partitioned Khatri-Rao matrix."""
        assert dim_noise_var is None, "currenly no way to add dim_noise_var"
        N, d = x.shape
        assert self.grid_dim == d, "currently only works for 1-dimensional grids"
        Kxz = [kern.cov(x=x[:, (i,)], z=z[i]) for i, kern in enumerate(self.kern_list)][::-1]
        return KhatriRaoMatrix(A=Kxz, partition=0) if form_kr else Kxz

    def cov_kr_grad(self, x, z, grad_dim):
        """d cov_kr / d x[:, grad_dim] (list of per-dimension matrices, reversed order).

        The reference can only do this through GPy (`gradients_X`, grid_kernel.py:196-199); here the
        in-house kernels provide `grad_x`.
        """
        N, d = x.shape
        assert self.grid_dim == d
        out = []
        for i, k in enumerate(self.kern_list):
            if i == grad_dim:
                if not hasattr(k, "grad_x"):
                    raise NotImplementedError
                out.append(k.grad_x(x[:, (i,)], z[i]))
            else:
                out.append(k.cov(x=x[:, (i,)], z=z[i]))
        return out[::-1]

    @property
    def parameters(self):
        if self.radial_kernel:
            return np.ravel(self.kern_list[0].parameters)
        return np.concatenate([np.ravel(kern.parameters) for kern in self.kern_list], axis=0)

    @parameters.setter
    def parameters(self, value):
        assert isinstance(value, np.ndarray)
        assert value.ndim == 1
        if self.radial_kernel:
            self.kern_list[0].parameters = value
            self.kern_list = [self.kern_list[0], ] * np.size(self.kern_list)
            return
        pos = 0
        for kern in self.kern_list:      # aliased kernel objects: the last slice written wins, as in the reference
            old = kern.parameters
            kern.parameters = value[pos:pos + np.size(old)].reshape(np.shape(old))
            pos += np.size(old)

    @property
    def constraints(self):
        if self.radial_kernel:
            return np.ravel(self.kern_list[0].constraints)
        return np.concatenate([np.ravel(kern.constraints) for kern in self.kern_list], axis=0)

    @property
    def diag_val(self):
        """k(x, x) of the (stationary) product kernel."""
        return self.cov(np.zeros((1, self.n_dims))).squeeze()
