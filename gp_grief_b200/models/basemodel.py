"""Model base class: parameter vector <-> object state, cache invalidation, likelihood dispatch,
L-BFGS-B driver, positivity transforms, gradient check.

Host control flow with the semantics of gp_grief/models/basemodel.py (reference :34-59 log_likelihood,
:62-112 optimize, :115-148 checkgrad, :150-184 parameters, :210-246 _objective_grad, :264-325 transforms,
:328-361 _finite_diff_gradient).  No arithmetic beyond O(#parameters) happens here.
"""
import inspect
from logging import getLogger
from traceback import format_exc

import numpy as np
from numpy.linalg import LinAlgError
from numpy.testing import assert_array_almost_equal
from scipy.optimize import fmin_l_bfgs_b

from ..linalg import solver_counter, LogexpTransformation

logger = getLogger(__name__)


class BaseModel(object):
    param_shift = {'+ve': 1e-200, '-ve': -1e-200}
    _transformations = {'+ve': LogexpTransformation()}

    def __init__(self):
        logger.debug('Initializing %s model.' % self.__class__.__name__)
        self.dependent_attributes = ['_alpha', '_log_like', '_gradient', '_K', '_log_det']
        self._previous_parameters = None
        self.grad_method = None              # 'adjoint' or 'finite_difference'
        self.noise_var_constraint = '+ve'
        self._counter = None
        self._log_like = None
        self._gradient = None

    # ------------------------------------------------------------------ likelihood
    def log_likelihood(self, return_gradient=False):
        """Log marginal likelihood (and its gradient) at the current parameters, cached until they change."""
        p = self.parameters                  # must come first: invalidates stale caches
        if return_gradient and (self._gradient is None):
            if 'adjoint' in self.grad_method:
                self._log_like, self._gradient = self._adjoint_gradient(p)
            elif 'finite_difference' in self.grad_method:
                self._log_like, self._gradient = self._finite_diff_gradient(p)
            else:
                raise RuntimeError('unknown grad_method %s' % repr(self.grad_method))
        elif self._log_like is None:
            self._log_like = self._compute_log_likelihood(p)
        if return_gradient:
            return self._log_like, self._gradient
        return self._log_like

    def optimize(self, max_iters=1e3, messages=False, use_counter=False, factr=1e7, pgtol=1e-05):
        """Maximise the log likelihood over the free parameters with L-BFGS-B (in the transformed space)."""
        logger.debug('Beginning MLE to optimize hyperparams. grad_method=%s' % self.grad_method)
        try:
            x0 = self._transform_parameters(self.parameters)
            assert np.all(np.isfinite(x0))
        except Exception:
            logger.error('Transformation failed for initial values. '
                         'Ensure constraints are met or the value is not too small.')
            raise
        free = np.logical_not(self._fixed_indicies)
        x0 = x0[free]
        self._counter = solver_counter(disp=True) if use_counter else None
        kwargs = dict(func=self._objective_grad, x0=x0, factr=factr, pgtol=pgtol, maxiter=int(max_iters))
        if 'disp' in inspect.signature(fmin_l_bfgs_b).parameters and messages:   # removed in recent SciPy
            kwargs['disp'] = messages
        opt = None
        try:
            x_opt, f_opt, opt = fmin_l_bfgs_b(**kwargs)
        except (KeyboardInterrupt, IndexError):
            logger.info('Keyboard interrupt raised. Cleaning up...')
            if self._counter is not None and self._counter.backup is not None:
                self.parameters = self._counter.backup[1]
                logger.info('will return best parameter set with log-likelihood = %.4g' % self._counter.backup[0])
        else:
            logger.info('Function Evals: %d. Exit status: %s' % (opt['funcalls'], opt['warnflag']))
            transformed = self._transform_parameters(self._previous_parameters)
            transformed[free] = x_opt
            self.parameters = self._untransform_parameters(transformed)
        return opt

    def checkgrad(self, decimal=3, raise_if_fails=True):
        """Analytic gradient against forward differences, as a ratio (reference :115-148)."""
        grad_exact = self._finite_diff_gradient(self.parameters)[1]
        grad_exact[self._fixed_indicies] = 1
        self.parameters = self.parameters
        self._gradient = None
        grad_analytic = np.array(self.log_likelihood(return_gradient=True)[1], dtype=float)
        grad_analytic[self._fixed_indicies] = 1
        tiny = np.logical_and(np.abs(grad_exact) < 1e-8, np.abs(grad_analytic) < 1e-8)
        close = np.abs(grad_exact - grad_analytic) < 1e-5
        mask = np.logical_or(tiny, close)
        grad_exact[mask] = 1.
        grad_analytic[mask] = 1.
        try:
            assert_array_almost_equal(grad_exact / grad_analytic, np.ones(grad_exact.shape), decimal=decimal)
        except AssertionError:
            logger.info('Gradient check failed.')
            logger.debug('[[Finite-Diff Gradient], [Analytic Gradient]]:\n%s\n' % repr(np.asarray([grad_exact, grad_analytic])))
            if raise_if_fails:
                raise
            logger.info(format_exc())
            return False
        logger.info('Gradient check passed.')
        return True

    # ------------------------------------------------------------------ parameter vector [noise_var, kernel...]
    def _invalidate_if_changed(self, parameters):
        if not np.array_equal(parameters, self._previous_parameters):
            for attr in self.dependent_attributes:
                setattr(self, attr, None)
            self._previous_parameters = parameters.copy()

    @property
    def parameters(self):
        parameters = np.concatenate((np.ravel(self.noise_var), self.kern.parameters), axis=0)
        self._invalidate_if_changed(parameters)
        return parameters.copy()

    @parameters.setter
    def parameters(self, parameters):
        self.noise_var = parameters[0]
        self.kern.parameters = parameters[1:]
        self._invalidate_if_changed(parameters)

    @property
    def constraints(self):
        return np.concatenate((np.ravel(self.noise_var_constraint), self.kern.constraints), axis=0)

    def predict(self, Xnew, compute_var=None):
        raise NotImplementedError('')

    def fit(self):
        raise NotImplementedError('')

    # ------------------------------------------------------------------ optimiser objective
    def _objective_grad(self, transformed_free_parameters):
        """Negative log likelihood and its gradient w.r.t. the transformed free parameters."""
        free = np.logical_not(self._fixed_indicies)
        transformed = self._transform_parameters(self._previous_parameters)
        transformed[free] = transformed_free_parameters
        try:
            self.parameters = self._untransform_parameters(transformed)
            objective, gradient = self.log_likelihood(return_gradient=True)
            objective = -np.float64(np.asarray(objective).squeeze())
            gradient = -np.asarray(gradient, dtype=float)
            if not np.isfinite(objective):
                logger.debug('objective is not finite')
            if not np.all(np.isfinite(gradient[free])):
                logger.debug('some derivatives are non-finite')
            gradient = self._transform_gradient(self.parameters, gradient)
        except (LinAlgError, ZeroDivisionError, ValueError):
            logger.error('numerical issue computing log-likelihood or gradient')
            raise
        free_gradient = gradient[free]
        if self._counter is not None:
            msg = 'log-likelihood=%.4g, gradient_norm=%.2g' % (-objective, np.linalg.norm(free_gradient))
            if self._counter.backup is None or self._counter.backup[0] < -objective:
                self._counter(msg=msg, store=(-objective, self.parameters.copy()))
            else:
                self._counter(msg=msg)
        return objective, free_gradient

    @property
    def _fixed_indicies(self):
        return np.asarray(self.constraints) == 'fixed'

    @property
    def _free_indicies(self):
        return np.logical_not(self._fixed_indicies)

    def _is_plain(self, constraint):
        return constraint is None or constraint == 'fixed' or constraint == ''

    def _transform_parameters(self, parameters):
        constraints = self.constraints
        assert parameters.size == np.size(constraints)
        out = np.zeros(parameters.size)
        for i, (param, c) in enumerate(zip(parameters, constraints)):
            out[i] = param if self._is_plain(c) else self._transformations[c].transform(param - self.param_shift[c])
        if not np.all(np.isfinite(out)):
            logger.debug('transformation led to non-finite value')
        return out

    def _transform_gradient(self, parameters, gradients):
        constraints = self.constraints
        assert parameters.size == gradients.size == np.size(constraints)
        out = np.zeros(parameters.size)
        for i, (param, grad, c) in enumerate(zip(parameters, gradients, constraints)):
            if c is None or c == '':
                out[i] = grad
            elif c != 'fixed':
                out[i] = self._transformations[c].transform_grad(param - self.param_shift[c], grad)
        if not np.all(np.isfinite(out)):
            logger.debug('transformation led to non-finite value')
        return out

    def _untransform_parameters(self, transformed_parameters):
        constraints = self.constraints
        assert transformed_parameters.size == np.size(constraints)
        out = np.zeros(transformed_parameters.size)
        for i, (t, c) in enumerate(zip(transformed_parameters, constraints)):
            out[i] = t if self._is_plain(c) else self._transformations[c].inverse_transform(t) + self.param_shift[c]
        if not np.all(np.isfinite(out)):
            logger.debug('transformation led to non-finite value')
        return out

    def _finite_diff_gradient(self, parameters):
        """Forward differences, step 1e-6: one extra likelihood evaluation per free parameter."""
        assert isinstance(parameters, np.ndarray)
        free_inds = np.nonzero(np.logical_not(self._fixed_indicies))[0]
        step = 1e-6
        stepped = np.zeros(free_inds.size)
        for i, idx in enumerate(free_inds):
            p_fs = parameters.copy()
            p_fs[idx] += step
            stepped[i] = np.asarray(self._compute_log_likelihood(p_fs)).squeeze()
        log_like = self._compute_log_likelihood(parameters)
        gradient = np.zeros(parameters.shape)
        gradient[free_inds] = (stepped - np.asarray(log_like).squeeze()) / step
        return log_like, gradient

    def _compute_log_likelihood(self, parameters):
        raise NotImplementedError('')

    def _adjoint_gradient(self, parameters):
        raise NotImplementedError('')
