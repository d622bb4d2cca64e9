"""GP with a weighted-basis (WEB) kernel on a fixed, user-supplied basis (reference: gp_grief/models/gp_web_model.py).

Works purely from the reduced statistics A = Phi^T Phi, r = Phi^T y, y^T y: every evaluation is the p x p stage of the
GP-GRIEF path (csrc/solve.cu) and is independent of n.  This is the consumer of the statistics in the Type-I workflow:
`GPwebModel.from_statistics` / `GPGriefModel.to_web_model()` hand over the device-resident (A | r | s) without a copy.
"""
from logging import getLogger

import numpy as np

from ..kern import WEBKernel
from .basemodel import BaseModel

logger = getLogger(__name__)


class GPwebModel(BaseModel):
    def __init__(self, Phi, y, noise_var=1.):
        """Phi (n, p) basis evaluated at the training inputs, y (n,) or (n, 1)."""
        import torch
        from .. import device
        device._torch()
        y = np.asarray(y, dtype=float).reshape((-1, 1))
        Phi = np.asarray(Phi, dtype=float)
        assert Phi.shape[0] == y.shape[0]
        Pd = torch.as_tensor(np.ascontiguousarray(Phi)).cuda()
        yd = torch.as_tensor(np.ascontiguousarray(y[:, 0])).cuda()
        PdT = Pd.T.contiguous()                           # (p, n): data rows contiguous, the K dimension of both products
        A = device.gemm_nt(PdT, PdT)                      # Phi^T Phi on the library's FP64 DMMA GEMM
        r = device.gemm_nt(PdT, yd.reshape(1, -1)).reshape(-1)
        s = (yd * yd).sum().reshape(1)
        self._init_from(A, r, s, y.shape[0], noise_var)

    @classmethod
    def from_statistics(cls, A, r, yty, n, noise_var=1.):
        """A (p, p), r (p), yty (1) as CUDA float64 tensors (or NumPy arrays); n = number of data rows."""
        import torch
        from .. import device
        device._torch()
        self = cls.__new__(cls)
        as_dev = lambda a: (a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))).cuda()
        self._init_from(as_dev(A), as_dev(r).reshape(-1), as_dev(yty).reshape(1), int(n), noise_var)
        return self

    def _init_from(self, A, r, s, n, noise_var):
        from .. import device
        super(GPwebModel, self).__init__()
        self.n = int(n)
        self.p = int(A.shape[0])
        self._A_dev, self._r_dev, self._s_dev = A, r, s
        self._device_mod = device
        self.noise_var = np.float64(noise_var)
        self.kern = WEBKernel(initial_weights=np.ones(self.p))
        self.grad_method = 'adjoint'
        self.dependent_attributes = np.unique(np.concatenate(
            (self.dependent_attributes, ['_P', '_Pchol', '_Pinv_r', '_alpha_p', '_solve'])))
        self._solve = None

    # host views of the statistics, same names as the reference
    @property
    def A(self):
        return self._A_dev.cpu().numpy()

    @property
    def r(self):
        return self._r_dev.cpu().numpy().reshape((-1, 1))

    @property
    def yTy(self):
        return self._s_dev.cpu().numpy().reshape((1, 1))

    def _run(self, want_grad):
        import torch
        have = self._solve
        if have is not None and (have['Pinv'] is not None or not want_grad):
            return have
        w = torch.as_tensor(np.ascontiguousarray(self.kern.parameters, dtype=np.float64)).cuda()
        self._solve = self._device_mod.shared_solver().solve(self._A_dev, self._r_dev, self._s_dev, w, float(self.noise_var), self.n,
                                         want_grad=want_grad, want_G2=False)
        return self._solve

    def _compute_log_likelihood(self, parameters):
        self.parameters = parameters
        return np.float64(self._run(False)['lml'])

    def _adjoint_gradient(self, parameters):
        assert isinstance(parameters, np.ndarray)
        self.parameters = parameters
        free = np.logical_not(self._fixed_indicies)
        out = self._run(True)
        gradient = np.zeros(parameters.shape) + np.nan
        gradient[1:] = out['grad_w'].cpu().numpy()
        gradient[0] = out['grad_noise']
        assert not np.any(np.isnan(gradient[free])), "gradient missed!"
        return np.float64(out['lml']), gradient

    def predict(self, Phi_new):
        """Posterior mean (M, 1) and covariance (M, M) at basis rows Phi_new (M, p)."""
        import torch
        logger.debug('Predicting model at new points.')
        assert Phi_new.ndim == 2
        assert Phi_new.shape[1] == self.p
        self.parameters
        out = self._run(True)
        Pn = torch.as_tensor(np.ascontiguousarray(Phi_new, dtype=np.float64)).cuda()
        nv = float(self.noise_var)
        from ..device import gemm_nt
        Yhat = gemm_nt(Pn, out['b'].reshape(1, -1)).cpu().numpy().reshape((-1, 1))   # alpha_p == b = P^-1 r
        Yhatvar = nv * torch.eye(Pn.shape[0], dtype=torch.float64, device="cuda")
        gemm_nt(gemm_nt(Pn, out['Pinv']), Pn, alpha=nv, beta=1.0, out=Yhatvar)       # P^-1 is symmetric
        return Yhat, Yhatvar.cpu().numpy()
