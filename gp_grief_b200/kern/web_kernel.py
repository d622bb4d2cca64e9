"""Weight holder for models on a fixed, user-supplied basis (reference: gp_grief/kern/web_kernel.py)."""
import numpy as np


class WEBKernel(object):
    def __init__(self, initial_weights):
        assert isinstance(initial_weights, np.ndarray)
        assert np.ndim(initial_weights) == 1
        self.p = np.size(initial_weights)
        self.parameters = initial_weights
        self.constraints = ['+ve', ] * self.p
