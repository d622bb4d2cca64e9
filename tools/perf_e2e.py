"""Wall-clock cost of a fresh model's first evaluation (what bench.py's e2e leg pays every step): upload, tables, audit, passes."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import gp_grief_b200 as gp
from gp_grief_b200.synthetic import CONFIGS, linspace_grid, bench_lengthscales, synthetic_xy

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
_, d, m, p, type2 = CONFIGS[cfg]
x, y = synthetic_xy(n, d)
xp, yp = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
grid = gp.grid.InducingGrid(xg=[g.reshape(-1, 1) for g in linspace_grid(d, m)])

def fresh(step, audit_rows):
    ls = np.array(bench_lengthscales(d)) * (1 + 1e-3 * step)
    kw = dict(reweight_eig_funs=False, opt_kernel_params=True) if type2 else {}
    kern = gp.kern.GriefKernel([gp.kern.RBF(1, lengthscale=l) for l in ls], grid, n_eigs=p, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model = gp.models.GPGriefModel(xp.numpy(), yp.numpy(), kern, noise_var=0.1)
    model.audit_rows = audit_rows
    torch.cuda.synchronize(); t1 = time.perf_counter()
    model.log_likelihood(return_gradient=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    prm = model.parameters
    if type2:
        prm[2] *= 1.0001
    else:
        prm[0] *= 1.0001
    model.parameters = prm
    if not type2:
        for name in ('_A', '_P', '_Pchol', '_alpha', '_alpha_p', '_log_like', '_gradient'):
            setattr(model, name, None)
    model.log_likelihood(return_gradient=True)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    del model
    return {"construct_ms": (t1 - t0) * 1e3, "first_eval_ms": (t2 - t1) * 1e3, "second_eval_ms": (t3 - t2) * 1e3}

fresh(0, 16384)
for rep in range(3):
    print(json.dumps({"audit_rows": 16384, **{k: round(v, 1) for k, v in fresh(1 + rep, 16384).items()}}))
    print(json.dumps({"audit_rows": 0, **{k: round(v, 1) for k, v in fresh(10 + rep, 0).items()}}))
