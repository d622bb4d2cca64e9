"""Lazy tensor wrappers and the host form of expand_SKC (reference: gp_grief/tensors/tensors.py)."""
import numpy as np

from .selection_matrix import SelectionMatrixSparse


class TensorProduct(object):
    """Product of several 2-D tensors applied to a vector without forming the product."""

    def __init__(self, tensor_list):
        self.tensors = tensor_list
        self.n_tensors = len(tensor_list)
        self.shape = (self.tensors[0].shape[0], self.tensors[-1].shape[1])
        for a, b in zip(self.tensors[:-1], self.tensors[1:]):
            assert a.shape[1] == b.shape[0]

    @property
    def T(self):
        raise NotImplementedError('easy to do this')

    def __mul__(self, x):
        assert x.shape == (self.shape[1], 1), "vector is wrong shape"
        y = x
        for t in reversed(self.tensors):
            y = t * y
        return y


class TensorSum(object):
    """Sum of several equally shaped 2-D tensors applied to a vector."""

    def __init__(self, tensor_list):
        self.tensors = tensor_list
        self.n_tensors = len(tensor_list)
        self.shape = self.tensors[0].shape
        for a, b in zip(self.tensors[:-1], self.tensors[1:]):
            assert a.shape == b.shape

    @property
    def T(self):
        raise NotImplementedError('easy to do this')

    def __mul__(self, x):
        assert x.shape == (self.shape[1], 1), "vector is wrong shape"
        y = np.zeros((self.shape[0], 1))
        for t in self.tensors:
            y += t * x
        return y


class Array(object):
    """numpy.ndarray wearing the tensor interface."""

    def __init__(self, A):
        self.A = A
        self.shape = A.shape

    def __mul__(self, x):
        return self.A.dot(x)

    @property
    def T(self):
        return Array(self.A.T)

    def expand(self):
        return self.A


def expand_SKC(S, K, C, logged=True):
    """Rows of (selection KR) * (Kronecker) * (column KR) on the HOST, for user code and small checks.

    Same contract as the reference (tensors/tensors.py:97-128): returns `prod` (p, n), or with
    logged=True the pair (log|prod|, sign).  The GP-GRIEF model does NOT call this: its Phi is produced
    tile by tile on the GPU (csrc/rows.cu + csrc/phi_stage.cu) from the same three ingredients.
    """
    assert isinstance(S, (list, np.ndarray))
    assert isinstance(S[0], SelectionMatrixSparse)
    assert isinstance(K, (list, np.ndarray))
    assert isinstance(C, (list, np.ndarray))
    log_prod, sign, prod = 0., 1., 1.
    for s, k, c in zip(S, K, C):
        x_unique = np.array(s.mul_unique(k).dot(c), dtype=float)     # only the distinct rows are formed
        if logged:
            sign = sign * np.int32(np.sign(x_unique))[s.unique_inverse]
            x_unique[x_unique == 0] = 1.
            log_prod = log_prod + np.log(np.abs(x_unique))[s.unique_inverse]
        else:
            prod = prod * x_unique[s.unique_inverse]
    return (log_prod, sign) if logged else prod
