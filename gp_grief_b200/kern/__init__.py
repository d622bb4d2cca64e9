from .basekernel import BaseKernel
from .gpy_kernel import GPyKernel
from .stationary import Stationary, RBF, Exponential, Matern32, Matern52
from .grid_kernel import GridKernel
from .grief_kernel import GriefKernel
from .web_kernel import WEBKernel
