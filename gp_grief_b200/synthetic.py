"""Synthetic GP-GRIEF workloads: the inputs of BASELINE.json's configs.

Pure NumPy, no device code.  Rows are generated in independent chunks so that any row range
(and therefore any row shard of a multi-GPU run) can be produced without generating the rest.
"""
import numpy as np

SEED = 20180710
CHUNK = 1 << 20


def synthetic_chunk(chunk_id, rows, d, noise=0.1):
    """Rows of chunk `chunk_id`: x ~ U[0,1)^d, y = sum_i sin(2 pi (1+i/d) x_i)/sqrt(d) + noise*N(0,1)."""
    rng = np.random.default_rng([SEED, int(chunk_id)])
    x = rng.random((rows, d))
    freq = 2.0 * np.pi * (1.0 + np.arange(d) / float(d))
    y = np.sin(x * freq[None, :]).sum(axis=1) / np.sqrt(d) + noise * rng.standard_normal(rows)
    return x, y.reshape(-1, 1)


def synthetic_xy(n, d, chunk=CHUNK, chunk_id0=0, row0=0):
    """First `n` rows starting at global row `row0` of the chunked stream (chunks of `chunk` rows)."""
    xs, ys = [], []
    first = row0 // chunk
    skip = row0 - first * chunk
    c = first
    need = n
    while need > 0:
        x, y = synthetic_chunk(chunk_id0 + c, chunk, d)
        x, y = x[skip:], y[skip:]
        skip = 0
        take = min(need, x.shape[0])
        xs.append(x[:take]); ys.append(y[:take])
        need -= take
        c += 1
    return np.ascontiguousarray(np.concatenate(xs, axis=0)), np.ascontiguousarray(np.concatenate(ys, axis=0))


def linspace_grid(d, m):
    """d one-dimensional grids of m equally spaced points on [0,1] (InducingGrid(xg=...) input)."""
    return [np.linspace(0.0, 1.0, m) for _ in range(d)]


def bench_lengthscales(d):
    """Distinct per-dimension lengthscales => tie-free top-p selection."""
    return [0.3 + 0.05 * i for i in range(d)]


CONFIGS = {
    # name: (n, d, m, p, type2)
    "C2": (1_000_000, 6, 10, 1024, False),
    "C3": (10_000_000, 10, 20, 4096, True),
    "C4": (5_000_000, 32, 8, 2048, False),
    "C5": (50_000_000, 8, 16, 8192, False),
}
