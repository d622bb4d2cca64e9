"""Re-export of the synthetic workload generator for the oracle scripts (test infrastructure)."""
from gp_grief_b200.synthetic import synthetic_xy, synthetic_chunk, linspace_grid, bench_lengthscales, CONFIGS  # noqa: F401
