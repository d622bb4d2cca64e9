"""Inducing-point grids (reference: gp_grief/grid.py).  Host only: O(n d) once per data set."""
import logging

import numpy as np

logger = logging.getLogger(__name__)


def nd_grid(*xg):
    """MATLAB-style ndgrid: d arrays, each of the full grid shape (reference grid.py:8-26)."""
    axes = [np.asarray(g).reshape(-1) for g in xg]
    for g in xg:
        if np.ndim(g) > 1:
            assert np.shape(g)[1] == 1, "currently supports 1d grid dims"
    mesh = np.meshgrid(*axes, indexing="ij") if len(axes) > 1 else [axes[0]]
    out = np.empty(len(axes), dtype=object)
    for i, m in enumerate(mesh):
        out[i] = np.array(m)
    return out


def grid2mat(*xg):
    """All grid points as an (N, d) matrix, last dimension varying fastest (reference grid.py:29-45)."""
    mesh = nd_grid(*xg)
    return np.stack([m.reshape(-1, order='C') for m in mesh], axis=1).astype(float)


class InducingGrid(object):
    """Cartesian grid of inducing points, built from scattered data or given explicitly.

    Same constructor contract and attributes as the reference (grid.py:48-165):
    `xg` (object array of (m_i, 1) arrays), `grid_shape`, `grid_dim`, `grid_sub_dim`, `input_dim`,
    `num_data` (float: product of the grid shape), `eq`.
    """

    def __init__(self, x=None, mbar=10, eq=True, mbar_min=1, xg=None, beyond_domain=None):
        logger.debug('Initializing inducing grid.')
        if xg is None:
            xg = self._from_scattered(x, mbar, eq, mbar_min, beyond_domain)
            if xg is None:
                return
        self._from_user_grid(xg)

    # -- grid from scattered training inputs (reference grid.py:86-142) --
    def _from_scattered(self, x, k, eq, k_min, beyond_domain):
        assert isinstance(x, np.ndarray)
        assert x.ndim == 2
        self.eq = eq
        n_train, d = x.shape
        if not isinstance(k, (tuple, list, np.ndarray)):
            k = (k,) * d
        self.grid_dim = d
        self.grid_sub_dim = np.ones(d, dtype=int)
        self.input_dim = int(np.sum(self.grid_sub_dim))
        lo, hi = np.amin(x, axis=0), np.amax(x, axis=0)
        span = hi - lo
        uniq = [np.unique(x[:, i]) for i in range(d)]
        n_unq = np.array([u.size for u in uniq])
        if not np.all(n_unq >= 2):
            logger.debug('some dimension have < 2 unique points')
        shape = np.zeros(d, dtype=int)
        for i, ki in enumerate(k):
            if ki <= 1:      # a fraction of the unique values
                shape[i] = int(max(np.ceil(ki * n_unq[i]), k_min))
            else:
                assert np.mod(ki, 1) == 0, "if k > 1 then k must be integer"
                shape[i] = int(max(min(ki, n_unq[i]), k_min))
        self.grid_shape = shape
        self.num_data = np.prod(np.float64(shape))
        if beyond_domain is not None:
            assert np.all(shape >= 2), "need >=2 points per dim"
            inner = InducingGrid(x=x, mbar=tuple(int(s) for s in shape - 2), eq=eq, mbar_min=0).xg
            out = np.empty(d, dtype=object)
            for i in range(d):
                out[i] = np.vstack((lo[i] - beyond_domain * span[i], inner[i], hi[i] + beyond_domain * span[i]))
            return out       # handled as a user grid by the caller
        self.xg = np.empty(d, dtype=object)
        for i in range(d):
            if shape[i] == n_unq[i]:                 # grid on the unique values
                self.xg[i] = uniq[i].reshape((-1, 1))
            elif eq:
                self.xg[i] = np.linspace(lo[i], hi[i], num=shape[i]).reshape((-1, 1))
            elif shape[i] == 2:
                self.xg[i] = np.array([lo[i], hi[i]]).reshape((-1, 1))
            else:
                raise NotImplementedError
        return None

    # -- user supplied grid (reference grid.py:143-155) --
    def _from_user_grid(self, xg):
        grid = np.empty(len(xg), dtype=object)      # explicit object array: ragged grids are fine
        for i, g in enumerate(xg):
            g = np.asarray(g, dtype=float)
            assert g.ndim == 2, "each element in xg must be a 2d array"
            grid[i] = g
        self.xg = grid
        self.grid_dim = grid.shape[0]
        self.grid_shape = np.array([g.shape[0] for g in grid], dtype=int)
        self.grid_sub_dim = np.array([g.shape[1] for g in grid], dtype=int)
        self.input_dim = int(np.sum(self.grid_sub_dim))
        self.num_data = np.prod(np.float64(self.grid_shape))
        self.eq = None

    def __getitem__(self, key):
        return self.xg[key]

    def __setitem__(self, key, value):
        self.xg[key] = value
