"""Khatri-Rao matrix types (reference: gp_grief/tensors/khatri_rao_matrix.py).

`KhatriRaoMatrix` is the row/column-partitioned block-Kronecker container GridKernel.cov_kr returns;
`RowColKhatriRaoMatrix` is the R*K*C product with memory-bounded row chunking.  Both are host NumPy
containers: the GRIEF hot path never goes through them (the device builds Phi tiles directly from the
per-dimension factors, csrc/phi_stage.cu) -- exactly as in the reference, whose GriefKernel imports
RowColKhatriRaoMatrix but calls expand_SKC instead.  `RowColKhatriRaoMatrix(..., device=True)` runs its mat-vec
(and that of its transpose) on the GPU through grief_rowcol_kr_matvec (csrc/krmatvec.cu), without forming any
block of rows; `device=True` without a CUDA device raises.
"""
import numpy as np
import scipy.sparse as sparse

from .block_matrix import BlockMatrix
from .kron_matrix import KronMatrix
from .selection_matrix import SelectionMatrix, SelectionMatrixSparse


class KhatriRaoMatrix(BlockMatrix):
    def __init__(self, A, partition=None):
        """A: list of equal-height (partition=0) or equal-width (partition=1) matrices, or a 2-D array of KronMatrix."""
        if isinstance(A, np.ndarray) and A.ndim == 2 and isinstance(A[0, 0], KronMatrix):   # (np.ndim of a ragged list raises)
            super(KhatriRaoMatrix, self).__init__(A)
            return
        assert partition in range(2)
        count = A[0].shape[0] if partition == 0 else A[0].shape[1]
        blocks = np.empty(count, dtype=object)
        for i in range(count):
            if partition == 0:
                blocks[i] = KronMatrix([Aj[(i,), :] for Aj in A], sym=False)
            else:
                blocks[i] = KronMatrix([Aj[:, (i,)] for Aj in A], sym=False)
        super(KhatriRaoMatrix, self).__init__(blocks.reshape((count, 1) if partition == 0 else (1, count)))


class RowColKhatriRaoMatrix(object):
    """R K C with R a row-partitioned and C a column-partitioned Khatri-Rao matrix, K a Kronecker matrix."""

    def __init__(self, R, K, C, nGb=1., device=False):
        self.device = bool(device)
        self.shape = (R[0].shape[0], C[0].shape[1])
        self.d = len(R)
        self.R = R
        if K is not None:
            if not (isinstance(K, np.ndarray) and K.dtype == object):     # a plain list of differently sized factors
                factors = list(K)
                K = np.empty(len(factors), dtype=object)
                for i, Ki in enumerate(factors):
                    K[i] = Ki
            assert len(K) == len(C) == self.d, "number of dims inconsistent"
            self.C = np.empty(self.d, dtype=object)
            for i in range(self.d):
                assert K[i].shape[0] == K[i].shape[1] == R[i].shape[1], \
                    "K must be a square Kronecker product matrix, and must be consistent with R"
                self.C[i] = K[i].dot(C[i])
        else:
            self.C = C
        self.nGb = nGb
        self._set_chunk()

    def _set_chunk(self):
        self.n_rows_at_once = 1
        if self.nGb is not None:
            self.n_rows_at_once = max(1, np.int32(np.floor(self.nGb * 1e9 / (8 * self.shape[1]))))

    @property
    def T(self):
        if isinstance(self.R[0], (SelectionMatrix, SelectionMatrixSparse)):
            return RowColKhatriRaoMatrixTransposed(R=self.R, K=None, C=self.C, nGb=self.nGb, device=self.device)
        return RowColKhatriRaoMatrix(R=[Ci.T for Ci in self.C], K=None, C=[Ri.T for Ri in self.R], nGb=self.nGb, device=self.device)

    @staticmethod
    def _dense(M):
        """Dense array of a factor (sparse matrices and boolean selection matrices are expanded)."""
        if isinstance(M, SelectionMatrix):
            M = M.sel
        return np.asarray(M.toarray() if sparse.issparse(M) else M, dtype=float)

    def _device_factors(self):
        """(R, C) for device.rowcol_kr_matvec: selection matrices become index vectors, everything else dense."""
        R = []
        for Ri in self.R:
            if isinstance(Ri, SelectionMatrixSparse):
                R.append(np.asarray(Ri.indicies, dtype=np.int64))
            elif isinstance(Ri, SelectionMatrix):
                R.append(np.asarray(Ri.sel.indices, dtype=np.int64))        # CSR with one entry per row
            else:
                R.append(self._dense(Ri))
        return R, [self._dense(Ci) for Ci in self.C]

    def _rows_1d(self, i_d, i_rows):
        if sparse.issparse(self.C[i_d]):
            return self.R[i_d][i_rows, :] * self.C[i_d]
        return self.R[i_d][i_rows, :].dot(self.C[i_d])

    def get_rows(self, i_rows, logged=False):
        """Rows `i_rows` of the product; with logged=True returns (log|rows|, sign)."""
        if not logged:
            rows = 1.
            for i_d in range(self.d):
                rows = rows * self._rows_1d(i_d, i_rows)
            return rows
        log_rows, sign = 0., 1.
        for i_d in range(self.d):
            r1 = np.array(self._rows_1d(i_d, i_rows), dtype=float)
            sign = sign * np.int32(np.sign(r1))
            r1[sign == 0] = 1.
            log_rows = log_rows + np.log(np.abs(r1))
        return log_rows, sign

    def expand(self, logged=False):
        return self.get_rows(i_rows=slice(None), logged=logged)

    def __mul__(self, x):
        """Memory-bounded mat-vec: at most n_rows_at_once rows of the product exist at a time (host), or none at all (device)."""
        assert x.shape == (self.shape[1], 1)
        if self.device:
            from ..device import rowcol_kr_matvec
            R, C = self._device_factors()
            return rowcol_kr_matvec(R, C, x)
        y = np.zeros((self.shape[0], 1))
        for start in range(0, self.shape[0], int(self.n_rows_at_once)):
            i_rows = np.arange(start, min(start + self.n_rows_at_once, self.shape[0]))
            y[i_rows, :] = self.get_rows(i_rows).dot(x)
        return y


class RowColKhatriRaoMatrixTransposed(RowColKhatriRaoMatrix):
    """Transpose of R K C when R is a selection matrix (never transposed explicitly)."""

    def __init__(self, *args, **kwargs):
        super(RowColKhatriRaoMatrixTransposed, self).__init__(*args, **kwargs)
        self.shape = self.shape[::-1]
        self._set_chunk()

    def _device_factors(self):
        """The transpose as a plain R' C' product: R'_t = C_t^T (dense), C'_t = R_t^T (dense; a selection matrix is expanded)."""
        return [self._dense(Ci).T for Ci in self.C], [self._dense(Ri.mul(np.identity(Ri.shape[1])) if isinstance(Ri, SelectionMatrixSparse)
                                                                  else Ri).T for Ri in self.R]

    def get_rows(self, i_rows):
        cols = None
        for i_d in range(self.d):
            Ci = self.C[i_d][:, i_rows]
            piece = self.R[i_d] * Ci if sparse.issparse(self.C[i_d]) else self.R[i_d].dot(Ci)
            cols = piece if cols is None else cols * piece
        return cols.T

    @property
    def T(self):
        return RowColKhatriRaoMatrix(R=self.R, K=None, C=self.C, nGb=self.nGb, device=self.device)
