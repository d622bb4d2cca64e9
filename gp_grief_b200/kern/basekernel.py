"""Kernel base class: parameter / constraint plumbing and `*`, `+` composition.

Mirrors gp_grief/kern/basekernel.py of the reference (same attributes: n_dims, active_dims, name,
parameter_list, constraint_map, _children), without its GPy import.  Host Python only.
"""
import copy as _copy
import logging

import numpy as np

logger = logging.getLogger(__name__)


class BaseKernel(object):
    def __init__(self, n_dims, active_dims, name):
        self.n_dims = n_dims
        if active_dims is None:
            active_dims = np.arange(self.n_dims)
        else:
            active_dims = np.ravel(active_dims)
            assert 'int' in active_dims.dtype.type.__name__
            assert active_dims.min() >= 0
            assert active_dims.max() < self.n_dims
        self.active_dims = active_dims
        self.name = self.__class__.__name__ if name is None else name
        self.parameter_list = None   # names of the parameter attributes
        self.constraint_map = None   # name -> constraint string(s)
        self._children = []          # [(operation, kernel)] from * and +

    def cov(self, x, z=None):
        raise NotImplementedError('Not implemented')

    # ---- flat parameter vector: own parameters first, then the children's ----
    @property
    def parameters(self):
        if self.parameter_list is None:
            raise NotImplementedError('Need to specify kern.parameter_list')
        parts = [np.ravel(getattr(self, name)) for name in self.parameter_list]
        parts += [child.parameters for _, child in self._children]
        return np.concatenate(parts, axis=0) if parts else np.array([])

    @parameters.setter
    def parameters(self, value):
        assert isinstance(value, np.ndarray)
        assert value.ndim == 1
        pos = 0
        for name in self.parameter_list:
            old = getattr(self, name)
            size = np.size(old)
            setattr(self, name, value[pos:pos + size].reshape(np.shape(old)))
            pos += size
        for _, child in self._children:
            size = np.size(child.parameters)
            child.parameters = value[pos:pos + size]
            pos += size

    @property
    def constraints(self):
        if self.constraint_map is None:
            raise NotImplementedError('Need to specify kern.constraint_map')
        parts = [np.ravel(self.constraint_map[name]) for name in self.parameter_list]
        parts += [child.constraints for _, child in self._children]
        return np.concatenate(parts, axis=0) if parts else np.array([])

    def is_stationary(self):
        from .stationary import Stationary
        return isinstance(self, Stationary)

    def _process_cov_inputs(self, x, z):
        assert x.ndim == 2
        assert x.shape[1] == self.n_dims
        if z is None:
            z = x
        else:
            assert z.ndim == 2
            assert z.shape[1] == self.n_dims, "should be %d dims, not %d" % (self.n_dims, z.shape[1])
        return x, z

    def _apply_children(self, K, x, z=None):
        """Fold the child kernels into the parent's covariance (last step of every cov())."""
        for operation, child in self._children:
            if operation == 'mul':
                K = np.multiply(K, child.cov(x, z))
            elif operation == 'add':
                K = np.add(K, child.cov(x, z))
            else:
                raise ValueError('Unknown kernel operation %s' % repr(operation))
        return K

    def __mul__(self, other):
        assert isinstance(other, BaseKernel)
        assert other.n_dims == self.n_dims
        parent, child = self.copy(), other.copy()
        if np.size(child.constraint_map['variance']) > 1:   # the product needs one free amplitude only
            child.constraint_map['variance'][0] = 'fixed'
        else:
            child.constraint_map['variance'] = 'fixed'
        parent._children.append(('mul', child))
        return parent

    def __add__(self, other):
        assert isinstance(other, BaseKernel), 'k2 must be a kernel'
        parent, child = self.copy(), other.copy()
        parent._children.append(('add', child))
        return parent

    def copy(self):
        dup = _copy.deepcopy(self)
        dup._children = [(_copy.deepcopy(op), child.copy()) for op, child in dup._children]
        return dup
