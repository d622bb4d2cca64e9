"""B200-native GP-GRIEF hot path behind the gp_grief Python API.

    import gp_grief_b200 as gp_grief
    kern = gp_grief.kern.GriefKernel(kern_list, grid, n_eigs=p)
    m = gp_grief.models.GPGriefModel(X, Y, kern, noise_var=0.1)
    lml, grad = m.log_likelihood(return_gradient=True)

Importing the package does not touch CUDA; the shared library (gp_grief_b200/_lib/libgrief_b200.so, built by
`__graft_entry__.build()`) is loaded on first use and there is no CPU fallback.
"""
import logging
import sys

from . import linalg
from . import grid
from . import tensors
from . import kern
from . import models
from . import synthetic

__version__ = '0.1'


def debug():
    """Route DEBUG logging of the package to stdout (same helper as the reference package)."""
    for handler in logging.root.handlers[:]:
        logging.root.removeHandler(handler)
    logging.basicConfig(stream=sys.stdout, level=logging.DEBUG,
                        format='%(asctime)s %(name)s %(levelname)s: %(message)s', datefmt='[ %H:%M:%S ]')
