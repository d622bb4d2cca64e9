#!/usr/bin/env python
"""Benchmark of the GP-GRIEF hot path: LML + gradient evaluations per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   the reference's CPU implementation on the host cores
    python bench.py --config C2|C3|C4|C5 ...                  the other BASELINE.json configs (C3 is the metric's config)

One "step" = one complete evaluation at NEW kernel hyper-parameters:
  Type-II (C3: n = 10M, d = 10, m = 20, p = 4096): host Schur of the d grid matrices -> GPU top-p selection -> table prepass ->
      Gram + Phi^T y (pass 1) -> [all-reduce] -> Cholesky / LML / d/dnoise -> Phi*P^-1 GEMM + contraction (pass 2) -> [all-reduce]
  Type-I (C2, C4, C5): the same without pass 2 (gradient w.r.t. the p weights and the noise from the p x p stage);
      `type1_reevals_per_s` is the O(p^3) re-evaluation on cached statistics that a Type-I optimiser actually iterates.
Rows are sharded over ranks (strong scaling: the total n is fixed).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gp_grief_b200.synthetic import CONFIGS, bench_lengthscales, linspace_grid, synthetic_xy  # noqa: E402

UNIT = "evals/s"
NOISE_VAR = 0.1


def metric_name(cfg):
    n, d, m, p, _ = CONFIGS[cfg]
    return "GRIEF LML+grad evals/sec (n=%s,d=%d,p=%d)" % ("%dM" % (n // 10 ** 6), d, p)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--rows", type=float, default=0, help="override n (testing only; reported in config)")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows per CPU-baseline sample chunk (0: 2^14 inline, 2^16 for --impl reference)")
    ap.add_argument("--cpu-chunks", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the parity blocks (full-n FP64 mode, oracle sample, N-vs-1 GPU sample)")
    ap.add_argument("--no-peaks", action="store_true", help="skip the in-run DGEMM / int8 GEMM peak measurement")
    ap.add_argument("--oracle-rows", type=int, default=1 << 15, help="rows of the oracle parity sample (<= 2e5)")
    ap.add_argument("--predict-rows", type=int, default=-1, help="rows of the prediction leg (-1: 100000 for C5, else 0)")
    ap.add_argument("--gemm", default="int8", choices=["int8", "int8x2", "int8x1", "fp64"],
                    help="arithmetic of the two O(n p^2) products: FP64 emulated on the INT8 tensor cores (int8 = int8x2: CTA pairs, the library "
                         "default; int8x1: one CTA per tile) or the FP64 DMMA GEMM")
    ap.add_argument("--digits", default="", help="'Dgram,Dz': int8 digits per operand of the two products (default: library defaults)")
    ap.add_argument("--slab-mb", type=int, default=0, help="HBM budget (MiB) of the Phi^T slab staged per pass-1 GEMM launch (0: library default 1280)")
    ap.add_argument("--power-trace", default="", help="write the clock / power samples of the timed region to this JSON file")
    return ap.parse_args()


def step_lengthscales(d, step):
    """New hyper-parameters every step (a Type-II optimiser never evaluates the same point twice)."""
    base = np.array(bench_lengthscales(d))
    return base * (1.0 + 1e-3 * ((step % 7) + 1))


# ------------------------------------------------------------------------------------------ clocks + power
class ClockSampler(object):
    """One long-running `nvidia-smi -lms 250` process (the recipe's clocks line + power.draw) sampled over a region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit")

    def __init__(self, index=0):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "250"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def stop(self, keep_trace=False):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=10)
                rows = [[t.strip() for t in ln.split(",")] for ln in out.splitlines() if ln.count(",") >= 7]
            except Exception:
                self.proc.kill()
        sm, mx, pw, lim, reasons, trace = [], [], [], [], set(), []
        for p in rows:
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
            try:
                pw.append(float(p[6])); lim.append(float(p[7]))
            except ValueError:
                pw.append(float("nan"))
            trace.append([sm[-1], pw[-1], 1 if p[5].lower().startswith("active") else 0])
        pw_ok = [v for v in pw if v == v]
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm),
               "power_w_median": float(np.median(pw_ok)) if pw_ok else None, "power_w_max": max(pw_ok) if pw_ok else None,
               "power_limit_w": max(lim) if lim else None}
        if keep_trace:
            out["trace_sm_mhz_power_w_powercap"] = trace
        return out


# ------------------------------------------------------------------------------------------ CPU baseline
def _reference_sample(cfg, rows, chunks):
    """The unmodified reference in a subprocess (oracle/ref_worker.py); None when no copy of it is importable."""
    _, d, m, p, type2 = CONFIGS[cfg]
    arg = json.dumps({"d": d, "m": m, "p": p, "rows": rows, "chunks": chunks, "type2": bool(type2)})
    env = dict(os.environ)
    env.setdefault("OPENBLAS_NUM_THREADS", str(os.cpu_count() or 1))
    try:
        res = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_worker.py"), arg], stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True, env=env, timeout=3600)
        out = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception:
        return None
    return out if out.get("available") else None


def _port_sample(cfg, rows, chunks):
    """The oracle port (oracle/grief_oracle.py, a NumPy restatement of the reference) on the same sample."""
    from oracle import grief_oracle as orc
    from scipy.linalg import cho_factor, cho_solve
    _, d, m, p, _ = CONFIGS[cfg]
    xg = linspace_grid(d, m)
    ls = bench_lengthscales(d)
    names, var = ["RBF"] * d, [1.0] * d
    t0 = time.perf_counter()
    basis = orc.setup_inducing_cov(names, var, ls, xg, p)
    t_setup = time.perf_counter() - t0
    per_row, A, r = [], None, None
    for c in range(chunks):
        x, y = synthetic_xy(rows, d, chunk=rows, chunk_id0=7_000_000 + c)
        t0 = time.perf_counter()
        Phi = orc.grief_phi(basis, names, var, ls, xg, x)
        A = Phi.T.dot(Phi)
        r = Phi.T.dot(y)
        per_row.append((time.perf_counter() - t0) / rows)
        del Phi
    t0 = time.perf_counter()
    Pchol = cho_factor(A + np.diag(NOISE_VAR / np.ones(p)))
    cho_solve(Pchol, r)
    t_pp_lml = time.perf_counter() - t0
    t0 = time.perf_counter()
    cho_solve(Pchol, A)
    t_pp_grad = time.perf_counter() - t0
    return {"t_setup": t_setup, "t_row": float(np.median(per_row)), "t_pp_lml": t_pp_lml, "t_pp_grad": t_pp_grad,
            "cores": os.cpu_count() or 1, "rows": rows, "chunks": chunks}


def cpu_baseline(cfg, n_total, rows, chunks):
    """Reference CPU path on a bounded row sample, extrapolated linearly in n.

    The reference materialises ~4 (p x n) float64 temporaries, so the full workload cannot run; its cost is exactly linear in n.
    seconds per LML = setup + n * (median per-row cost of kern.cov + Gram + Phi^T y) + p x p stage.  The reference's kernel-parameter
    gradient is forward finite differences (models/basemodel.py:328-361): a Type-II LML+gradient costs (d + 3) LML evaluations
    (d lengthscales + variance + noise + the base point); a Type-I one costs one LML + the adjoint p x p solve.
    """
    _, d, m, p, type2 = CONFIGS[cfg]
    s = _reference_sample(cfg, rows, chunks)
    kind = "reference"
    if s is None:
        s = _port_sample(cfg, rows, chunks)
        kind = "port"
    t_lml = s["t_setup"] + s["t_row"] * n_total + s["t_pp_lml"]
    t_analytic = t_lml + s["t_pp_grad"]                      # 1 LML + adjoint solve (what an analytic gradient would cost on the CPU)
    t_fd = (d + 3) * t_lml                                   # what the reference actually does for kernel parameters
    t_eval = t_fd if type2 else t_analytic
    what = ("unmodified reference (%s)" % s.get("reference_path", "?")) if kind == "reference" else "oracle port (NumPy restatement)"
    return {"value": 1.0 / t_eval, "unit": UNIT, "cores": s["cores"], "kind": kind,
            "sample": "%s, %d chunks x %d rows of the %s workload (p=%d, d=%d): median %.3e s/row -> %.0f s per LML at n=%d; p x p stage "
                      "%.2f s (+%.2f s adjoint solve); %s" % (what, s["chunks"], s["rows"], cfg, p, d, s["t_row"], t_lml, n_total, s["t_pp_lml"], s["t_pp_grad"],
                                                             "Type-II: x (d+3)=%d LML evaluations per LML+gradient (forward differences)" % (d + 3) if type2
                                                             else "Type-I: 1 LML + adjoint gradient"),
            "seconds_per_lml": t_lml, "seconds_per_eval_analytic_equivalent": t_analytic, "seconds_per_eval_finite_difference": t_fd,
            "seconds_per_eval": t_eval}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total, d, m, p, type2 = CONFIGS[args.config]
    if args.rows:
        n_total = int(args.rows)
    rows = args.cpu_rows or (1 << 16)
    t_start = time.perf_counter()
    if args.warmup > 0:
        cpu_baseline(args.config, n_total, min(rows, 4096), 1)                 # page in NumPy / BLAS threads
    best = cpu_baseline(args.config, n_total, rows, max(3, min(args.cpu_chunks, 8)))
    line = {"metric": metric_name(args.config), "value": best["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / best["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_string(args.config, n_total)},
            "cpu_baseline": best,
            "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_start}
    print(json.dumps(line))


def workload_string(cfg, n_total):
    _, d, m, p, type2 = CONFIGS[cfg]
    return "%s: %s GRIEF LML+gradient, n=%d, d=%d, m=%d grid pts/dim, p=%d, RBF kernels, new lengthscales every step" % (
        cfg, "Type-II" if type2 else "Type-I", n_total, d, m, p)


# ------------------------------------------------------------------------------------------ in-run peaks
def _time_loop(torch, fn, seconds):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, k = time.time(), 0
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(8):
            fn(); k += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return best, e0.elapsed_time(e1) / k


def fp64_peak(torch, seconds=1.5):
    """cuBLAS DGEMM 8192^3 on this GPU: burst (best of 8) and sustained (back to back for `seconds`), TFLOP/s."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    best, avg = _time_loop(torch, lambda: torch.matmul(a, b, out=c), seconds)
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best * 1e-9, 2.0 * n ** 3 / avg * 1e-9


def int8_peak(torch, local, seconds=2.0):
    """Dense int8 x int8 -> int32 GEMM 8192^3 (torch._int_mm = cuBLASLt) on this GPU, TOP/s: burst and sustained with the clock /
    power it ran at.  This is the denominator of the INT8 roofline: same tensor pipe, same power cap, measured in this process."""
    n = 8192
    try:
        a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda").t()      # K-major B (the "TN" layout)
        torch._int_mm(a, b)
        smp = ClockSampler(local).start()
        best, avg = _time_loop(torch, lambda: torch._int_mm(a, b), seconds)
        clk = smp.stop()
        del a, b
        torch.cuda.empty_cache()
        return {"burst_tops": 2.0 * n ** 3 / best * 1e-9, "sustained_tops": 2.0 * n ** 3 / avg * 1e-9, "seconds": seconds,
                "sm_mhz": clk["sm_mhz"], "power_w_median": clk["power_w_median"], "reasons": clk["reasons"],
                "how": "torch._int_mm 8192^3 int8 (cuBLASLt), best of 8 / back to back for %.1f s" % seconds}
    except Exception as e:                                                               # pragma: no cover
        return {"error": repr(e)}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import gp_grief_b200 as gp
    from gp_grief_b200 import _native as nat
    from gp_grief_b200.sharding import row_shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = args.config
    n_total, d, m, p, type2 = CONFIGS[cfg]
    if args.rows:
        n_total = int(args.rows)
    lib = nat.lib()
    mode_id = {"int8": 3, "int8x2": 3, "int8x1": 1, "fp64": 0}[args.gemm]
    nat.check(lib.grief_set_default_option(nat.OPT_GEMM_MODE, mode_id))
    if args.digits:
        dg, dz = [int(t) for t in args.digits.split(",")]
        nat.check(lib.grief_set_default_option(nat.OPT_DIGITS_GRAM, dg))
        nat.check(lib.grief_set_default_option(nat.OPT_DIGITS_Z, dz))
    if args.slab_mb:
        nat.check(lib.grief_set_default_option(nat.OPT_SLAB_BUDGET, args.slab_mb << 20))
    dg, dz = int(lib.grief_get_default_option(nat.OPT_DIGITS_GRAM)), int(lib.grief_get_default_option(nat.OPT_DIGITS_Z))
    pairs = lambda D, sym=False: D * (D + 1) // 2 + (1 if (sym and D % 2 == 0) else 0)     # int8 digit GEMMs per FP64 GEMM

    r0, r1 = row_shard(n_total, world, rank)
    n_local = r1 - r0
    x_np, y_np = synthetic_xy(n_local, d, row0=r0)
    x_pin = torch.from_numpy(x_np).pin_memory()               # pinned host copies: the e2e leg copies from these every step
    y_pin = torch.from_numpy(y_np).pin_memory()
    del x_np, y_np
    xg = linspace_grid(d, m)
    grid = gp.grid.InducingGrid(xg=[g.reshape(-1, 1) for g in xg])

    def make_kern(ls):
        kw = dict(reweight_eig_funs=False, opt_kernel_params=True) if type2 else {}
        return gp.kern.GriefKernel([gp.kern.RBF(1, variance=1.0, lengthscale=l) for l in ls], grid, n_eigs=p, **kw)

    def make_model(step, x=None, y=None, dist_=None):
        return gp.models.GPGriefModel(x_pin.numpy() if x is None else x, y_pin.numpy() if y is None else y,
                                      make_kern(step_lengthscales(d, step)), noise_var=NOISE_VAR,
                                      distributed=distributed if dist_ is None else dist_)

    def set_step(model, step):
        """New lengthscales for all d kernels.  A Type-I model keeps its Gram across parameter changes (kernel parameters are
        fixed there, like the reference): invalidate it the way the reference does, by assigning None."""
        if type2:
            prm = model.parameters
            prm[2:1 + 2 * d:2] = step_lengthscales(d, step)
            model.parameters = prm
        else:
            for k, l in zip(model.kern.kern_list, step_lengthscales(d, step)):
                k.lengthscale = l
            model.parameters = model.parameters
            for name in ('_A', '_P', '_Pchol', '_alpha', '_alpha_p', '_log_like', '_gradient'):
                setattr(model, name, None)

    def evaluate(model, step):
        set_step(model, step)
        return model.log_likelihood(return_gradient=True)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def as_f(l_):
        return float(np.asarray(l_).squeeze())

    def grad_diff(ga, gb):
        ok = ~np.isnan(ga) & ~np.isnan(gb)
        return float(np.abs(ga[ok] - gb[ok]).max() / max(np.abs(gb[ok]).max(), 1e-300))

    peaks = {}
    if rank == 0 and not args.no_peaks:
        pb, ps = fp64_peak(torch)
        peaks = {"cublas_dgemm_tflops_burst": pb, "cublas_dgemm_tflops_sustained": ps, "int8": int8_peak(torch, local)}

    # ---- device-resident leg: one model, data stays in HBM, new hyper-parameters every step ----
    model = make_model(0)
    for s in range(args.warmup):
        evaluate(model, s)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nat.profile_enable(True)
    nat.profile_read()
    lib.grief_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        lml, grad = evaluate(model, args.warmup + s)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = lib.grief_launch_count()
    prof = nat.profile_read()
    nat.profile_enable(False)
    clocks = sampler.stop(keep_trace=bool(args.power_trace)) if rank == 0 else None
    if rank == 0 and args.power_trace:
        with open(args.power_trace, "w") as f:
            json.dump({"config": cfg, "gemm": args.gemm, "digits": [dg, dz], "period_ms": 250,
                       "columns": ["sm_mhz", "power_w", "sw_power_cap"], "samples": clocks.pop("trace_sm_mhz_power_w_powercap")}, f)
    ms_total = float(ms.item())
    value = args.steps / (ms_total * 1e-3)
    lml_val = as_f(lml)
    grad = np.asarray(grad, dtype=float).copy()
    last_step = args.warmup + args.steps - 1

    # Type-I: what the optimiser iterates on -- new weights / noise on cached statistics (O(p^3), independent of n)
    reevals = None
    if not type2:
        prm = model.parameters
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 0
        while k < 3 or time.perf_counter() - t0 < 1.0:
            prm[0] = NOISE_VAR * (1.0 + 1e-3 * (k + 1))
            model.parameters = prm
            model.log_likelihood(return_gradient=True)
            k += 1
        reevals = k / (time.perf_counter() - t0)
        prm[0] = NOISE_VAR
        model.parameters = prm

    # ---- parity blocks (outside the timed region) ----
    check = {"lml": lml_val, "grad_finite": bool(np.all(np.isfinite(grad[~np.isnan(grad)]))),
             # a-posteriori audit of the digit counts, done by the model at its first evaluation (models/gp_grief_model.py)
             "arithmetic_audit": model.arithmetic_audit}
    if not args.no_check:
        # (a) the timed arithmetic against the FP64 DMMA mode at FULL n: same step, same rows, all ranks
        if args.gemm != "fp64":
            nat.check(lib.grief_set_default_option(nat.OPT_GEMM_MODE, 0))
            model.kern._plan = None                                 # next evaluation builds a plan with the FP64 default
            for name in ('_A', '_P', '_Pchol', '_log_like', '_gradient'):
                setattr(model, name, None)
            l64, g64 = model.log_likelihood(return_gradient=True)
            nat.check(lib.grief_set_default_option(nat.OPT_GEMM_MODE, mode_id))
            g64 = np.asarray(g64, dtype=float)
            check["int8_vs_fp64_full_n"] = {"rows": n_total, "step": last_step, "digits": [dg, dz], "lml_int8_tensor": lml_val, "lml_fp64_dmma": as_f(l64),
                                            "lml_rel_diff": abs(lml_val - as_f(l64)) / abs(as_f(l64)),
                                            "grad_max_abs_diff_over_max_abs": grad_diff(grad, g64)}
            if type2:      # the kernel-parameter block alone (the noise component is ~100x larger and comes from the p x p stage)
                check["int8_vs_fp64_full_n"]["grad_theta_max_abs_diff_over_max_abs_theta"] = grad_diff(grad[1:1 + 2 * d], g64[1:1 + 2 * d])
        del model
        torch.cuda.empty_cache()
        # (b) N GPUs against 1 GPU on a row sample (every rank evaluates its shard of the sample; rank 0 also evaluates all of it)
        if distributed:
            ns = min(n_total, 200_000)
            s0, s1 = row_shard(ns, world, rank)
            xs, ys = synthetic_xy(max(s1 - s0, 0), d, row0=s0) if s1 > s0 else (np.zeros((0, d)), np.zeros((0, 1)))
            md = make_model(0, xs, ys, True)
            ld, gd = md.log_likelihood(return_gradient=True)
            del md
            if rank == 0:
                xa, ya = synthetic_xy(ns, d)
                m1 = make_model(0, xa, ya, False)
                l1, g1 = m1.log_likelihood(return_gradient=True)
                del m1
                check["n_gpus_vs_1_gpu_on_sample"] = {"rows": ns, "n_gpus": world, "lml_rel_diff": abs(as_f(ld) - as_f(l1)) / abs(as_f(l1)),
                                                      "grad_max_abs_diff_over_max_abs": grad_diff(np.asarray(gd, dtype=float), np.asarray(g1, dtype=float))}
                assert check["n_gpus_vs_1_gpu_on_sample"]["lml_rel_diff"] < 1e-9, check
                assert check["n_gpus_vs_1_gpu_on_sample"]["grad_max_abs_diff_over_max_abs"] < 1e-9, check
            torch.cuda.empty_cache()
        # (c) the oracle (NumPy restatement of the reference, pinned by tests/golden) on a row sample, rank 0
        if rank == 0 and args.oracle_rows > 0:
            from oracle import grief_oracle as orc
            ns = int(min(n_total, args.oracle_rows, 200_000))
            xa, ya = synthetic_xy(ns, d)
            ls0 = list(step_lengthscales(d, 0))
            mo = make_model(0, xa, ya, False)
            lo, go = mo.log_likelihood(return_gradient=True)
            t0 = time.perf_counter()
            fit, basis = orc.lml_full(["RBF"] * d, [1.0] * d, ls0, xg, p, xa, ya, 1.0, NOISE_VAR)
            gs, gw = orc.adjoint_gradient(fit, reweight=not type2)
            blk = {"rows": ns, "oracle_seconds": time.perf_counter() - t0,
                   "eig_index_selection_identical": bool(np.array_equal(np.asarray(mo.kern._eig_pos), basis.eig_loc)),
                   "lml_rel_diff": abs(as_f(lo) - fit.lml) / abs(fit.lml),
                   "grad_noise_rel_diff": abs(float(go[0]) - float(gs)) / max(abs(float(gs)), 1e-300)}
            if not type2:
                gw = np.asarray(gw, dtype=float).reshape(-1)
                blk["grad_w_max_abs_diff_over_max_abs"] = float(np.abs(go[-p:] - gw).max() / np.abs(gw).max())
            xq, _ = synthetic_xy(256, d, chunk_id0=10 ** 6)
            yhat = mo.predict(xq, compute_var=None)
            Phi_q = orc.grief_phi(basis, ["RBF"] * d, [1.0] * d, ls0, xg, xq)
            yo = np.asarray(orc.predict(fit, Phi_q)[0]).reshape(-1)
            blk["predictive_mean_max_abs_diff_over_max_abs"] = float(np.abs(yhat.reshape(-1) - yo).max() / np.abs(yo).max())
            check["oracle_on_sample"] = blk
            del mo, fit, basis
            torch.cuda.empty_cache()
    else:
        del model
        torch.cuda.empty_cache()

    # ---- end-to-end leg: host buffers in, host results out, every step (the public API call a user makes) ----
    e2e = None
    if not args.no_e2e:
        steps_e2e = max(1, args.steps)
        m0 = make_model(0)                                    # warm: library handles and the caching allocator's pool (a process that
        m0.log_likelihood(return_gradient=True)               # evaluates repeatedly keeps its device buffers between calls)
        del m0
        barrier()
        trace = os.environ.get("GRIEF_BENCH_E2E_TRACE") == "1" and rank == 0
        t0 = time.perf_counter()
        for s in range(steps_e2e):
            ta = time.perf_counter()
            mm = make_model(100 + s)                           # H2D copy of X and y from pinned host memory
            tb = time.perf_counter()
            l_, g_ = mm.log_likelihood(return_gradient=True)  # D2H of LML and gradient
            as_f(l_); np.asarray(g_)
            tc = time.perf_counter()
            del mm
            if trace:
                print("e2e step %d: construct %.1f ms, evaluate %.1f ms, release %.1f ms" % (s, (tb - ta) * 1e3, (tc - tb) * 1e3, (time.perf_counter() - tc) * 1e3), file=sys.stderr)
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": steps_e2e / float(t_e2e.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(n_local * (d + 1) * 8), "d2h_bytes_per_step": int(grad.size * 8 + 8),
               "steps": steps_e2e}

    # ---- prediction leg (C5: mean + diagonal variance at M = 100 000 rows, sharded over ranks, no collective) ----
    predict = None
    M = args.predict_rows if args.predict_rows >= 0 else (100_000 if cfg == "C5" else 0)
    if M > 0:
        q0, q1 = row_shard(M, world, rank)
        xq, _ = synthetic_xy(q1 - q0, d, chunk_id0=10 ** 6, row0=q0)
        mp_ = make_model(0)
        mp_.fit()
        mp_.predict(xq[:1024], compute_var='diag')             # warm
        barrier()
        t0 = time.perf_counter()
        yhat, yvar = mp_.predict(xq, compute_var='diag')      # host rows in, host mean / variance out
        barrier()
        tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        predict = {"rows": M, "seconds": float(tp.item()), "rows_per_s": M / float(tp.item()),
                   "what": "predict(Xnew, compute_var='diag'): mean + marginal variance, models/gp_grief_model.py:99-125",
                   "mean_finite": bool(np.all(np.isfinite(yhat))), "var_min": float(yvar.min()), "var_above_noise": bool(np.all(yvar > NOISE_VAR * (1 - 1e-9)))}
        if args.gemm != "fp64" and not args.no_check:           # the INT8 quadratic form against the FP64 DMMA mode on the same rows
            plan = mp_.kern.device_plan()
            plan.set_option(nat.OPT_GEMM_MODE, 0)
            mp_._Phi_last_pred = None
            _, yvar64 = mp_.predict(xq[:8192], compute_var='diag')
            predict["var_int8_vs_fp64_max_rel_diff"] = float(np.abs(yvar[:8192] - yvar64).max() / np.abs(yvar64).max())
            plan.set_option(nat.OPT_GEMM_MODE, mode_id)
        del mp_
        torch.cuda.empty_cache()

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event timing of the profile slots) ----
    p_pad = (p + 127) // 128 * 128
    rows128 = (n_local + 127) // 128 * 128
    i8 = args.gemm != "fp64"
    kern_rows = []
    # work actually issued by the GEMM launches (padded rows / columns, full diagonal tiles)
    flops = {"k_zgemm": 2.0 * rows128 * p_pad * p_pad, "k_gram": float(rows128) * p_pad * (p_pad + 128)}
    digits_of = {"k_zgemm": dz, "k_gram": dg}
    gk = {"int8": "k_ozaki<2,D> (cta_group::2 pairs)", "int8x2": "k_ozaki<2,D> (cta_group::2 pairs)", "int8x1": "k_ozaki<1,D>", "fp64": "k_gemm_nt"}[args.gemm]
    label = {"k_zgemm": gk + " [Z = Phi*P^-1, pass 2]", "k_gram": gk + " [A = Phi^T Phi, lower tiles, split K]",
             "k_build_phi": "k_build_phi (+ row maxima, fused residual a = (y - Phi b) / noise) [Phi slab, pass 2]", "k_build_phi_t": "k_build_phi_t (+ slot maxima, fused Phi^T y) [Phi^T slab, pass 1]",
             "solve": "dense p x p stage (k_potf2_inv, k_gemm_nt, k_trsv_step, k_assemble, ...)",
             "k_contract": "k_contract_back [W = dL/dT from Z^T, lane = row trie sweep]",
             "k_dtables": "k_contract_tail [W -> V -> parameter sums]"}
    non_gemm = 0.0
    for name in ("k_zgemm", "k_gram", "k_contract", "k_build_phi", "k_build_phi_t", "solve", "k_dtables", "k_tables", "phi_t_y", "k_topk"):
        t_ms, cnt = prof.get(name, (0.0, 0))
        if cnt:
            row = {"kernel": label.get(name, name), "slot": name, "launches": cnt, "ms_total": t_ms, "share_of_step": t_ms / ms_total}
            if name in flops:
                row["issued_fp64_equiv_tflops"] = flops[name] * args.steps / (t_ms * 1e-3) * 1e-12
                if i8:
                    row["int8_digit_products"] = pairs(digits_of[name], name == "k_gram")
                    row["issued_int8_tops"] = row["int8_digit_products"] * row["issued_fp64_equiv_tflops"]
            elif name not in ("solve", "k_topk"):
                non_gemm += t_ms
            kern_rows.append(row)
    gemm_rows = [r for r in kern_rows if r["slot"] in flops]
    dom = max(gemm_rows, key=lambda r: r["ms_total"]) if gemm_rows else None
    traffic = None
    try:      # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of the same launch shape
        key = ("k_ozaki_" if i8 else "k_gemm_nt_") + ("zgemm" if dom["slot"] == "k_zgemm" else "gram")
        tr = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))[key]
        if tr["config"] == cfg:
            traffic = {"bytes_per_launch": tr["dram_bytes_read"] + tr["dram_bytes_write"], "rows_per_launch": tr["rows_per_launch"],
                       "algorithmic_bytes_per_launch": tr["algorithmic_bytes_per_launch"], "source": tr["source"]}
    except Exception:
        traffic = None
    roofline = None
    flop_factor = 3.0 if type2 else 1.0                        # algorithmic flops per evaluation: 3 n p^2 (Type-II), n p^2 (Type-I)
    if dom:
        algo_flops = (2.0 if dom["slot"] == "k_zgemm" else 1.0) * n_local * p * p        # algorithmic, per evaluation and GPU
        algo = algo_flops * args.steps / (dom["ms_total"] * 1e-3) * 1e-12
        common = {"traffic": traffic, "avg_launch_ms": dom["ms_total"] / dom["launches"], "kernels": kern_rows,
                  "non_gemm_row_kernels_share_of_step": non_gemm / ms_total,
                  "fp64_equivalent_tflops": algo, "cublas_dgemm_tflops_this_run": peaks.get("cublas_dgemm_tflops_sustained"),
                  "fp64_dmma_issue_peak_tflops": 37.2,
                  "whole_eval_fp64_equivalent_tflops_per_gpu": flop_factor * n_total * p * p * value / world * 1e-12}
        if i8:
            tops = dom["issued_int8_tops"]
            pk = peaks.get("int8", {})
            if "sustained_tops" in pk:
                peak_i8 = pk["sustained_tops"]
                src = ("dense int8 GEMM measured in this process: %s -> sustained %.0f / burst %.0f TOP/s at %s MHz, %s W (kernel timed inside a "
                       "long step: sustained)" % (pk["how"], pk["sustained_tops"], pk["burst_tops"], pk["sm_mhz"], pk["power_w_median"]))
            else:
                try:
                    mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                    peak_i8, src = 2.0 * float(mp["bf16_tflops_sustained"]), "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (no int8 GEMM could be timed in this run)"
                except Exception:
                    peak_i8, src = 4500.0, "nominal dense INT8 4.5 POP/s (MEASURED_PEAKS.json missing)"
            roofline = dict(common, bound="tensor",
                            kernel=gk + " (tcgen05 kind::i8, TMEM accumulators, TMA digit planes): %s as %d exact int8 x int8 -> int32 digit "
                                   "GEMMs (%d digits per operand); digit planes written straight from the tables by the slab builders"
                                   % ("Z = Phi*P^-1 for pass 2, one launch per slab of 37888 rows" if dom["slot"] == "k_zgemm" else
                                      "A = Phi^T Phi for pass 1, one launch per slab", dom["int8_digit_products"], digits_of[dom["slot"]]),
                            achieved=tops, peak=peak_i8, unit="TOP/s (int8)", frac=tops / peak_i8, peak_source=src,
                            int8_peak_measured=pk)
        else:
            ps = peaks.get("cublas_dgemm_tflops_sustained") or 35.4
            roofline = dict(common, bound="tensor", kernel="k_gemm_nt (TMA-fed FP64 DMMA GEMM), one launch per slab; Phi slab staged in HBM by the builders",
                            achieved=algo, peak=ps, unit="TFLOP/s", frac=algo / ps,
                            peak_source="cuBLAS DGEMM 8192^3 measured in this run (sustained); MEASURED_PEAKS.json has no FP64 row; FP64 DMMA issue-rate "
                                        "peak 37.2 TFLOP/s (profiles/r01_fp64_pipes_microbench.txt)")
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(cfg, n_total, args.cpu_rows or (1 << 14), max(2, args.cpu_chunks))
    dtype = "f64" if not i8 else ("f64 (O(n p^2) products as exact int8 digit GEMMs on the tensor cores: %d-bit operands for A = Phi^T Phi, %d-bit for "
                                  "Z = Phi P^-1, relative to the operand row maximum)" % (8 * dg - 2, 8 * dz - 2))
    line = {"metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload_string(cfg, n_total), "gemm_arithmetic": args.gemm, "int8_digits": [dg, dz] if i8 else None,
                       "rows_per_gpu": n_local,
                       "parallelism": "rows sharded over %d rank(s), NCCL all-reduce of (A|r|s)%s" % (world, " and of the theta-gradient" if type2 else ""),
                       "l2": "inputs (X %.1f GB, tables %.1f GB per GPU) exceed the 126 MB L2" % (n_local * d * 8e-9, rows128 * 105 * 8e-9)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "check": check,
            "tflops_whole_eval": flop_factor * n_total * p * p * value * 1e-12}
    if reevals is not None:
        line["type1_reevals_per_s"] = reevals
    if predict is not None:
        line["predict"] = predict
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
