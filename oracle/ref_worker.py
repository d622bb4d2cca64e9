#!/usr/bin/env python
"""CPU arm of bench.py: the UNMODIFIED reference (scwolof/gp_grief) timed on the host cores, in its own process.

TEST / MEASUREMENT INFRASTRUCTURE -- NOT PRODUCT CODE.  Executed only by bench.py (`--impl reference` and the
`cpu_baseline` block); the product never imports it.

The reference package is looked for in
  1. <repo>/baseline/_ref      pip-installed copy (`__graft_entry__.build()` installs it when /root/reference exists;
                               git-ignored, travels to the GPU box with the snapshot),
  2. $GP_GRIEF_REFERENCE or /root/reference   (build container).
It imports GPy at kern/basekernel.py:3 and kern/gpy_kernel.py:3; GPy cannot be installed (no network), so a 4-file IMPORT
stub (no arithmetic) is written to a temporary directory; only the in-house kernels (kern/stationary.py) are used.
If neither location is importable the worker prints {"available": false} and bench.py falls back to the oracle port.

What is timed (all through the reference's own classes):
  setup    GriefKernel._setup_inducing_cov()                        kern/grief_kernel.py:168-190
  per row  kern.cov(X_chunk)[0]; Phi.T.dot(Phi); Phi.T.dot(y)       kern/grief_kernel.py:68-111, models/gp_grief_model.py:148-149,234
           on `chunks` chunks of `rows` rows -> median seconds per row (the path is exactly linear in n and the reference cannot
           hold n rows: it materialises ~4 p x n temporaries)
  p x p    cho_factor(P), cho_solve(P, r), log-det, and cho_solve(P, A) for the w / noise gradient
           models/gp_grief_model.py:152-153,171-191,238-245
Prints one JSON object on stdout.
"""
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _find_reference():
    cands = [os.path.join(ROOT, "baseline", "_ref"), os.environ.get("GP_GRIEF_REFERENCE", "/root/reference")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "gp_grief", "kern")) and os.path.isfile(os.path.join(c, "gp_grief", "models", "gp_grief_model.py")):
            return c
    return None


def _write_gpy_stub(dst):
    files = {
        "GPy/__init__.py": "from . import kern\n",
        "GPy/kern/__init__.py": "class Kern(object):\n    pass\nfrom . import src\n",
        "GPy/kern/src/__init__.py": "from . import stationary\n",
        "GPy/kern/src/stationary.py": "class Stationary(object):\n    pass\n",
    }
    for rel, body in files.items():
        path = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(body)


def main():
    cfg = json.loads(sys.argv[1])
    ref = _find_reference()
    if ref is None:
        print(json.dumps({"available": False, "why": "no importable copy of the reference (baseline/_ref, /root/reference)"}))
        return
    stub = tempfile.mkdtemp(prefix="gpy_stub_")
    _write_gpy_stub(stub)
    sys.path.insert(0, ref)
    sys.path.insert(0, stub)
    import warnings
    warnings.filterwarnings("ignore")
    import numpy as np
    from scipy.linalg import cho_factor, cho_solve
    try:
        import gp_grief                                           # the reference
        from gp_grief.grid import InducingGrid
        from gp_grief.kern import RBF, GriefKernel
    except Exception as e:                                        # pragma: no cover
        print(json.dumps({"available": False, "why": "import of the reference failed: %r" % (e,)}))
        return
    assert os.path.realpath(gp_grief.__file__).startswith(os.path.realpath(ref)), gp_grief.__file__
    sys.path.append(ROOT)
    from gp_grief_b200.synthetic import synthetic_xy, linspace_grid, bench_lengthscales   # input generator only (pure NumPy)

    d, m, p, rows, chunks, type2 = cfg["d"], cfg["m"], cfg["p"], cfg["rows"], cfg["chunks"], cfg["type2"]
    xg = linspace_grid(d, m)
    grid = InducingGrid(xg=[g.reshape(-1, 1) for g in xg])
    ls = bench_lengthscales(d)
    kw = dict(reweight_eig_funs=False, opt_kernel_params=True) if type2 else {}
    kern = GriefKernel([RBF(1, variance=1.0, lengthscale=l) for l in ls], grid, n_eigs=p, **kw)
    t0 = time.perf_counter()
    kern._setup_inducing_cov()
    t_setup = time.perf_counter() - t0
    per_row, A, r = [], None, None
    t_phi = t_gram = 0.0
    for c in range(chunks):
        x, y = synthetic_xy(rows, d, chunk=rows, chunk_id0=7_000_000 + c)
        t0 = time.perf_counter()
        Phi = kern.cov(x)[0]
        t1 = time.perf_counter()
        A = Phi.T.dot(Phi)
        r = Phi.T.dot(y)
        t2 = time.perf_counter()
        per_row.append((t2 - t0) / rows)
        t_phi += t1 - t0
        t_gram += t2 - t1
        del Phi
    w = np.ones(p)
    t0 = time.perf_counter()
    P = A + np.diag(0.1 / w)
    Pchol = cho_factor(P)
    b = cho_solve(Pchol, r)
    2.0 * np.sum(np.log(np.diag(Pchol[0])))
    t_pp_lml = time.perf_counter() - t0
    t0 = time.perf_counter()
    cho_solve(Pchol, A)                                           # adjoint gradient: models/gp_grief_model.py:175
    t_pp_grad = time.perf_counter() - t0
    print(json.dumps({"available": True, "reference_path": ref, "cores": os.cpu_count() or 1, "t_setup": t_setup,
                      "t_row": float(np.median(per_row)), "t_rows_all": per_row, "t_phi_share": t_phi / max(t_phi + t_gram, 1e-30),
                      "t_pp_lml": t_pp_lml, "t_pp_grad": t_pp_grad, "rows": rows, "chunks": chunks,
                      "check_b_norm": float(np.linalg.norm(b))}))


if __name__ == "__main__":
    main()
