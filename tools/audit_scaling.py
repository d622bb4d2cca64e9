"""How does the INT8-vs-FP64 difference of the gradient pass / the Gram scale with the number of rows?  (design of the a-posteriori audit)"""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import gp_grief_b200 as gp
from gp_grief_b200 import _native as nat
from gp_grief_b200.synthetic import CONFIGS, linspace_grid, bench_lengthscales, synthetic_xy

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
_, d, m, p, _ = CONFIGS[cfg]
x, y = synthetic_xy(n, d)
grid = gp.grid.InducingGrid(xg=[g.reshape(-1, 1) for g in linspace_grid(d, m)])
kern = gp.kern.GriefKernel([gp.kern.RBF(1, lengthscale=l) for l in bench_lengthscales(d)], grid, n_eigs=p, reweight_eig_funs=False, opt_kernel_params=True)
model = gp.models.GPGriefModel(x, y, kern, noise_var=0.1)
model.audit_rows = 0
ll, g_full = model.log_likelihood(return_gradient=True)
out = model._dev['solve']
plan = model._plan()
T, X, Y = model._dev['tables'], model._X_dev, model._y_dev
gmax = np.nanmax(np.abs(g_full))
for ns in (4096, 16384, 65536, 262144, n):
    row = {"rows": ns}
    for D in (4, 5, 7):
        plan.set_option(nat.OPT_GEMM_MODE, 1); plan.set_option(nat.OPT_DIGITS_Z, D)
        g8 = plan.grad_theta(T, X, Y, ns, out['Pinv'], out['b'], 0.1)
        plan.set_option(nat.OPT_GEMM_MODE, 0)
        g64 = plan.grad_theta(T, X, Y, ns, out['Pinv'], out['b'], 0.1)
        dabs = float((g8 - g64).abs().max())
        row["D%d" % D] = {"abs": dabs, "rel_to_sample_grad": dabs / float(g64.abs().max()), "linear_extrapolated_rel_to_full": dabs * n / ns / gmax}
    plan.set_option(nat.OPT_GEMM_MODE, 1); plan.set_option(nat.OPT_DIGITS_Z, 4)
    print(json.dumps(row))
