"""Time the p x p stage (grief_solve_lml) alone: python tools/perf_solve.py [p ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_grief_b200 import device as dev

for p in [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192]:
    g = torch.Generator(device="cuda").manual_seed(p)
    n = 2 * p
    Phi = torch.randn((n, p), dtype=torch.float64, device="cuda", generator=g) / n ** 0.5
    A = Phi.T @ Phi
    r = torch.randn(p, dtype=torch.float64, device="cuda", generator=g)
    w = torch.rand(p, dtype=torch.float64, device="cuda", generator=g) + 0.5
    s = torch.ones(1, dtype=torch.float64, device="cuda")
    sol = dev.DeviceSolver()
    for want in (False, True):
        for _ in range(2):
            sol.solve(A, r, s, w, 0.1, n, want_grad=want, want_G2=want)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            out = sol.solve(A, r, s, w, 0.1, n, want_grad=want, want_G2=want)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        fl = p ** 3 / 3 + (2 * p ** 3 / 3 + p ** 3 / 3 if want else 0)     # potrf (+ trtri-equivalent + lauum-equivalent)
        print(f"p={p} inverse={want}: {dt*1e3:.2f} ms  ({fl/dt/1e12:.2f} TFLOP/s nominal)")
    P = A + torch.diag(0.1 / w)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        Lt = torch.linalg.cholesky(P); Pi = torch.cholesky_inverse(Lt)
    torch.cuda.synchronize(); print(f"p={p} torch cholesky+cholesky_inverse (cuSOLVER): {(time.perf_counter()-t0)/3*1e3:.2f} ms")
    err = (out["Pinv"] @ P - torch.eye(p, dtype=torch.float64, device="cuda")).abs().max().item()
    print(f"p={p} |Pinv P - I|max = {err:.2e}")
