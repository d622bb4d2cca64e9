"""Quick device-side timing of the prepass + fused Gram at a benchmark shape (not the bench)."""
import sys, time, json
import numpy as np, scipy.linalg as la, torch
sys.path.insert(0, ".")
from gp_grief_b200 import device
from gp_grief_b200.synthetic import CONFIGS, linspace_grid, bench_lengthscales, synthetic_xy

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 0
_, d, m, p, _ = CONFIGS[cfg]
xg = linspace_grid(d, m); ls = bench_lengthscales(d)
Q, eig = [], []
for i in range(d):
    diff = xg[i][:, None] - xg[i][None, :]
    K = np.exp(-0.5 * diff ** 2 / ls[i] ** 2) + 1e-12 * np.eye(m)
    T_, Z = la.schur(np.asarray(K, order="F")); Q.append(Z); eig.append(np.diag(T_))
t0 = time.time()
loc_k, lam = device.topk_kron(eig[::-1], p)          # KronMatrix order = reversed dims
torch.cuda.synchronize(); t_topk = time.time() - t0
t0 = time.time()
loc_k, lam = device.topk_kron(eig[::-1], p)
torch.cuda.synchronize(); t_topk2 = time.time() - t0
loc = loc_k[:, ::-1]
t0 = time.time()
plan = device.DevicePlan(["RBF"] * d, [1.0] * d, ls, xg, Q, eig, loc, width_cap=cap)
t_plan = time.time() - t0
print("cfg", cfg, "n", n, "G", plan.n_groups, "width", plan.width, "stride", plan.stride, "topk_s", t_topk, t_topk2, "plan_s", t_plan)
X = torch.rand((n, d), dtype=torch.float64, device="cuda")
def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
T = plan.build_tables(X)
ms_tab = timed(lambda: plan.build_tables(X))
ws = torch.empty((plan.gram_workspace_bytes(n),), dtype=torch.uint8, device="cuda")
A = torch.empty((p, p), dtype=torch.float64, device="cuda")
ms_gram = timed(lambda: plan.gram(T, n, out=A, workspace=ws))
y = torch.rand((n,), dtype=torch.float64, device="cuda")
ms_r = timed(lambda: plan.phi_t_vec(T, n, y))
flops = float(n) * p * (p + 1)
tab_bytes = n * (d + plan.stride) * 8.0
print(json.dumps({"cfg": cfg, "n": n, "p": p, "G": plan.n_groups, "width": plan.width,
                  "tables_ms": ms_tab, "tables_GBs": tab_bytes / ms_tab * 1e-6,
                  "gram_ms": ms_gram, "gram_TFLOPs": flops / ms_gram * 1e-9, "phi_t_y_ms": ms_r}))
