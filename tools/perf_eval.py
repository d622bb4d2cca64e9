"""Wall-clock breakdown of one Type-II evaluation through the public API (host + device, synchronised per phase)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import gp_grief_b200 as gp
from gp_grief_b200.synthetic import CONFIGS, linspace_grid, bench_lengthscales, synthetic_xy

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
_, d, m, p, _ = CONFIGS[cfg]
x, y = synthetic_xy(n, d)
grid = gp.grid.InducingGrid(xg=[g.reshape(-1, 1) for g in linspace_grid(d, m)])
ls = bench_lengthscales(d)
kern = gp.kern.GriefKernel([gp.kern.RBF(1, lengthscale=l) for l in ls], grid, n_eigs=p, reweight_eig_funs=False, opt_kernel_params=True)
model = gp.models.GPGriefModel(x, y, kern, noise_var=0.1)
model.log_likelihood(return_gradient=True)   # warm

def tic():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(2):
    prm = model.parameters
    prm[2:1 + 2 * d:2] = np.array(ls) * (1 + 1e-3 * (rep + 1))
    t = {}
    t0 = tic(); model.parameters = prm; t["set_params"] = tic() - t0
    # sub-phases of _setup_inducing_cov, timed on a throw-away copy of the same calls
    t0 = tic(); Kuu = kern.cov_grid(kern.grid.xg, dim_noise_var=kern.dim_noise_var); t["  cov_grid"] = tic() - t0
    t0 = tic(); Quu, Tm = Kuu.schur(); t["  schur"] = tic() - t0
    t0 = tic(); eg = Tm.diag(); pos = eg.find_extremum_eigs(n_eigs=p, mode='largest', log_expand=True)[0]; t["  topk"] = tic() - t0
    t0 = tic(); kern._setup_inducing_cov(); t["setup_inducing_cov(host schur + gpu topk)"] = tic() - t0
    t0 = tic(); plan = kern.device_plan(); t["plan_create"] = tic() - t0
    t0 = tic(); st = model._stats(); t["tables+gram+phity"] = tic() - t0
    t0 = tic(); out = model._cov_setup(want_grad=True, want_G2=True); t["solve"] = tic() - t0
    pmap = kern.base_parameter_map()
    free = np.nonzero(~(np.asarray(model.constraints) == 'fixed'))[0]
    theta_free = [i - 1 for i in free if 1 <= i <= 2 * d]
    active = [pmap[i] for i in theta_free]
    t0 = tic(); dqs = kern.scaled_eigvec_derivatives(active); t["eigvec_derivatives(host)"] = tic() - t0
    t0 = tic(); plan.grad_setup([a[0] for a in active], [0 if a[1] == 'variance' else 1 for a in active], dqs); t["grad_setup"] = tic() - t0
    t0 = tic(); g = plan.grad_theta(model._dev['tables'], model._X_dev, model._y_dev, n, out['Pinv'], out['b'], 0.1); t["grad_theta"] = tic() - t0
    t0 = tic(); ll, gr = model.log_likelihood(return_gradient=True); t["log_likelihood(cached parts + repeat grad)"] = tic() - t0
    print(json.dumps({k: round(v * 1e3, 2) for k, v in t.items()}))
