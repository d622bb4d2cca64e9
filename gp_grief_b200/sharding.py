"""Row sharding of a data set over the ranks of one job, and the layout of the all-reduced statistics buffer.

Pure host logic (no CUDA): exercised on CPU with the gloo backend in tests/test_distributed_cpu.py.
"""


def row_shard(n_total, world, rank):
    """Contiguous row range [r0, r1) of `rank`: ceil(n/world) rows per rank, the tail ranks may be short or empty."""
    per = (n_total + world - 1) // world
    r0 = min(n_total, rank * per)
    r1 = min(n_total, (rank + 1) * per)
    return r0, r1


def stats_layout(p):
    """Offsets of A (p x p), r (p) and s (1) inside the packed buffer that is all-reduced once per evaluation."""
    return {"A": (0, p * p), "r": (p * p, p * p + p), "s": (p * p + p, p * p + p + 1), "size": p * p + p + 1}
