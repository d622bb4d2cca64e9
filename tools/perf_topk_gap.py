"""Diagnose host-side latency of the top-k call that follows a long GPU phase."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_grief_b200 import device as dev, _native as nat

rng = np.random.default_rng(0)
eigs = [np.sort(rng.random(20))[::-1].copy() + 1e-3 for _ in range(10)]
dev.topk_kron(eigs, 4096)
big = torch.randn((8192, 8192), dtype=torch.float64, device="cuda")
for rep in range(4):
    for _ in range(10):
        c = big @ big                      # ~0.3 s of GPU work
    torch.cuda.synchronize()
    t0 = time.perf_counter(); a = dev.topk_kron(eigs, 4096); t1 = time.perf_counter()
    b = dev.topk_kron(eigs, 4096); t2 = time.perf_counter()
    x = torch.empty(10, device="cuda"); y = x.cpu(); t3 = time.perf_counter()
    print(f"rep {rep}: first topk {1e3*(t1-t0):.2f} ms, second {1e3*(t2-t1):.2f} ms, tiny alloc+D2H {1e3*(t3-t2):.3f} ms")
