"""Row-selection 'matrices' (reference: gp_grief/tensors/selection_matrix.py)."""
import numpy as np
import scipy.sparse as sparse


class SelectionMatrix(object):
    """One non-zero per row, stored as a boolean CSR matrix (+ its transpose)."""
    ndim = 2

    def __init__(self, indicies):
        if isinstance(indicies, tuple):
            assert len(indicies) == 2
            assert indicies[0].ndim == 1
            idx = indicies[0]
            self.shape = [idx.size, indicies[1]]
        else:
            assert indicies.ndim == 1
            assert indicies.dtype == bool
            idx = np.nonzero(indicies)[0]
            self.shape = [idx.size, indicies.size]
        nnz = self.shape[0]
        self.sel = sparse.csr_matrix((np.ones(nnz, dtype=bool), (np.arange(nnz), idx)), shape=self.shape, dtype=bool)
        self.sel_T = self.sel.T

    def mul(self, x):
        return self.sel * x

    def mul_T(self, x):
        return self.sel_T * x


class SelectionMatrixSparse(object):
    """Index-vector form: the per-dimension eigen-index table of the GRIEF basis.

    `indicies` (p,), `unique`, `unique_inverse` are what the device plan consumes
    (gp_grief_b200.device.DevicePlan): the unique list selects the eigenvectors that are evaluated,
    the inverse map routes every basis column to its entry.
    """
    ndim = 2

    def __init__(self, indicies):
        assert isinstance(indicies, tuple)
        assert len(indicies) == 2
        assert indicies[0].ndim == 1
        self.shape = [indicies[0].size, indicies[1]]
        self.indicies = indicies[0]
        self.unique, self.unique_inverse = np.unique(self.indicies, return_inverse=True)

    def mul(self, x):
        assert x.ndim == 2
        return x[self.indicies, :]
    dot = __mul__ = mul

    def mul_unique(self, x):
        """Rows of x at the unique indices; `result[self.unique_inverse]` recovers `mul(x)`."""
        assert x.ndim == 2
        return x[self.unique, :]

    def mul_T(self, x):
        raise NotImplementedError('Not finished')

    def __getitem__(self, key):
        if isinstance(key, tuple):
            key = key[0]
        return SelectionMatrixSparse(indicies=(np.atleast_1d(self.indicies[key]), self.shape[1]))
