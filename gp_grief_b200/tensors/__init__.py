from .kron_matrix import KronMatrix
from .selection_matrix import SelectionMatrix, SelectionMatrixSparse
from .block_matrix import BlockMatrix
from .tensors import TensorProduct, TensorSum, Array, expand_SKC
from .khatri_rao_matrix import KhatriRaoMatrix, RowColKhatriRaoMatrix, RowColKhatriRaoMatrixTransposed
