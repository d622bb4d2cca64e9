"""Generic block matrix with mat-vec, transpose and expand (reference: gp_grief/tensors/block_matrix.py).

Host container only -- base class of KhatriRaoMatrix.
"""
import numpy as np


class BlockMatrix(object):
    def __init__(self, A):
        """A: 2-D object array of blocks; each block needs `.shape`, `*` (mat-vec), `.T` and (optionally) `.expand()`."""
        assert A.ndim == 2, 'A must be 2d'
        self.A = A
        self.block_shape = self.A.shape
        heights = [blk.shape[0] for blk in self.A[:, 0]]
        widths = [blk.shape[1] for blk in self.A[0, :]]
        self._partition_shape = (heights, widths)
        self.shape = (np.sum(heights), np.sum(widths))
        for i in range(self.block_shape[0]):
            for j in range(self.block_shape[1]):
                assert np.all(A[i, j].shape == self.partition_shape(i, j)), \
                    "A[%d,%d].shape should be %s, not %s" % (i, j, repr(self.partition_shape(i, j)), repr(A[i, j].shape))
        self.vec_split = np.cumsum([0, ] + widths, dtype='i')

    def partition_shape(self, i, j):
        return (self._partition_shape[0][i], self._partition_shape[1][j])

    def __mul__(self, x):
        assert x.shape == (self.shape[1], 1)
        pieces = [x[self.vec_split[j]:self.vec_split[j + 1], :] for j in range(self.block_shape[1])]
        rows = []
        for i in range(self.block_shape[0]):
            acc = 0
            for j, piece in enumerate(pieces):
                acc = acc + self.A[i, j] * piece
            rows.append(acc)
        return np.concatenate(rows, axis=0)

    def transpose(self):
        At = np.empty(self.block_shape[::-1], dtype=object)
        for i in range(self.block_shape[0]):
            for j in range(self.block_shape[1]):
                At[j, i] = self.A[i, j].T
        return self.__class__(A=At)
    T = property(transpose)

    def expand(self):
        out = np.zeros(np.asarray(self.shape, dtype='i'))
        r = np.cumsum([0, ] + self._partition_shape[0], dtype='i')
        c = np.cumsum([0, ] + self._partition_shape[1], dtype='i')
        for i in range(self.block_shape[0]):
            for j in range(self.block_shape[1]):
                out[r[i]:r[i + 1], c[j]:c[j + 1]] = self.A[i, j].expand()
        return out
