#!/usr/bin/env python
"""Benchmark of the GP-GRIEF hot path: LML + gradient evaluations per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   the reference algorithm on the host cores (oracle port)

One "step" = one complete Type-II evaluation at NEW kernel hyper-parameters on the C3 workload
(n = 10M rows, d = 10, m = 20 grid points per dimension, p = 4096 basis functions):
host Schur of the d grid matrices -> GPU top-p selection -> table prepass -> fused Gram / Phi^T y (pass 1)
-> [all-reduce] -> Cholesky / LML / d/dnoise -> fused Phi*G2 GEMM + contraction (pass 2) -> [all-reduce].
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

METRIC = "GRIEF LML+grad evals/sec (n=10M,d=10,p=4096)"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--rows", type=float, default=0, help="override n (testing only; reported in config)")
    ap.add_argument("--cpu-rows", type=int, default=1 << 13, help="rows per CPU-baseline sample chunk")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gemm", default="int8", choices=["int8", "int8x2", "fp64"],
                    help="arithmetic of the two O(n p^2) products: FP64 emulated on the INT8 tensor cores (54-bit operands, "
                         "same parity tests) or the FP64 DMMA GEMM")
    return ap.parse_args()


def step_lengthscales(d, step):
    """New hyper-parameters every step (a Type-II optimiser never evaluates the same point twice)."""
    base = np.array(bench_lengthscales(d))
    return base * (1.0 + 1e-3 * ((step % 7) + 1))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """One long-running `nvidia-smi -lms 500` process (the recipe's clocks line) sampled over the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "500"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=10)
                rows = [[t.strip() for t in ln.split(",")] for ln in out.splitlines() if ln.count(",") >= 5]
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for p in rows:
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline(cfg, n_total, chunk_rows, chunks=2):
    """The reference algorithm (oracle/grief_oracle.py, a NumPy restatement) on the host cores, bounded sample.

    The reference materialises ~4 (p x n) float64 temporaries, so the full workload cannot run; its cost is
    exactly linear in n.  We time setup + Phi build + Gram + Phi^T y on `chunks` chunks of `chunk_rows` rows,
    take the median per-row cost, scale to n, add the p x p stage measured in full, and multiply the
    n-dependent part by (d + 3): the reference's only kernel-parameter gradient is forward finite
    differences over the d + 2 free parameters plus the base point (models/basemodel.py:328-361).
    """
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
        x, y = synthetic_xy(chunk_rows, d, chunk=chunk_rows, chunk_id0=7_000_000 + c)
        t0 = time.perf_counter()
        Phi = orc.grief_phi(basis, names, var, ls, xg, x)
        A = Phi.T.dot(Phi)
        r = Phi.T.dot(y)
        per_row.append((time.perf_counter() - t0) / chunk_rows)
        del Phi
    t_row = float(np.median(per_row))
    t0 = time.perf_counter()
    P = A + np.diag(0.1 / np.ones(p))
    Pchol = cho_factor(P)
    cho_solve(Pchol, r)
    t_pp = time.perf_counter() - t0
    t_lml = t_setup + t_row * n_total + t_pp
    t_eval = (d + 3) * t_lml
    cores = os.cpu_count() or 1
    return {"value": 1.0 / t_eval, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d chunks x %d rows of the %s workload (p=%d, d=%d): %.3e s/row -> %.0f s per LML at n=%d, "
                      "x (d+3)=%d LML evaluations per reference LML+gradient (forward differences); p x p stage %.2f s"
                      % (chunks, chunk_rows, cfg, p, d, t_row, t_lml, n_total, d + 3, t_pp),
            "seconds_per_lml": t_lml, "seconds_per_eval": t_eval}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total, d, m, p, _ = CONFIGS[args.config]
    if args.rows:
        n_total = int(args.rows)
    vals = []
    t_start = time.perf_counter()
    for _ in range(max(1, args.warmup > 0)):
        cpu_baseline(args.config, n_total, args.cpu_rows, chunks=1)
    for _ in range(max(1, min(args.steps, 3))):
        vals.append(cpu_baseline(args.config, n_total, args.cpu_rows, chunks=1))
    best = max(vals, key=lambda v: v["value"])
    line = {"metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / best["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "%s: Type-II GRIEF LML+gradient, n=%d, d=%d, m=%d, p=%d (extrapolated from a bounded "
                                   "row sample; the reference materialises Phi and cannot hold n rows)" % (args.config, n_total, d, m, p)},
            "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_start}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def fp64_peak(torch, seconds=2.0):
    """cuBLAS DGEMM 8192^3 on this GPU: burst (best of 5) and sustained (back to back for `seconds`)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    flops = 2.0 * n ** 3
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, k = time.time(), 0
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(4):
            torch.matmul(a, b, out=c); k += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sustained = flops * k / e0.elapsed_time(e1) * 1e-9
    del a, b, c
    torch.cuda.empty_cache()
    return flops / best * 1e-9, sustained


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gp_grief_b200 as gp
    from gp_grief_b200 import _native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_total, d, m, p, type2 = CONFIGS[args.config]
    if args.rows:
        n_total = int(args.rows)
    # contiguous row shard of this rank
    from gp_grief_b200.sharding import row_shard
    r0, r1 = row_shard(n_total, world, rank)
    n_local = r1 - r0
    x_np, y_np = synthetic_xy(n_local, d, row0=r0)
    # pinned host copies: the e2e leg copies from these every step
    x_pin = torch.from_numpy(x_np).pin_memory()
    y_pin = torch.from_numpy(y_np).pin_memory()
    del x_np, y_np
    xg = linspace_grid(d, m)
    grid = gp.grid.InducingGrid(xg=[g.reshape(-1, 1) for g in xg])

    def make_model(step):
        ls = step_lengthscales(d, step)
        kern = gp.kern.GriefKernel([gp.kern.RBF(1, variance=1.0, lengthscale=l) for l in ls], grid, n_eigs=p,
                                   reweight_eig_funs=False, opt_kernel_params=True)
        return gp.models.GPGriefModel(x_pin.numpy(), y_pin.numpy(), kern, noise_var=0.1, distributed=distributed)

    peak_burst = peak_sust = None
    arith = None
    if rank == 0:
        peak_burst, peak_sust = fp64_peak(torch)
        # same workload on a row sample in BOTH arithmetic modes (outside the timed region): the INT8 tensor-core path must reproduce
        # the FP64 DMMA path -- evidence inside the bench line that the timed arithmetic is double-precision class
        ns = min(n_local, 200000)
        res = {}
        for mode in (0, 1):
            nat.lib().grief_set_gemm_mode(mode)
            kern = gp.kern.GriefKernel([gp.kern.RBF(1, variance=1.0, lengthscale=l) for l in step_lengthscales(d, 0)], grid, n_eigs=p,
                                       reweight_eig_funs=False, opt_kernel_params=True)
            ms_ = gp.models.GPGriefModel(x_pin.numpy()[:ns], y_pin.numpy()[:ns], kern, noise_var=0.1)
            l_, g_ = ms_.log_likelihood(return_gradient=True)
            res[mode] = (float(np.asarray(l_).squeeze()), np.asarray(g_, dtype=float).copy())
            del ms_, kern
        torch.cuda.empty_cache()
        ok_ = ~np.isnan(res[0][1])
        arith = {"rows": ns, "lml_fp64_dmma": res[0][0], "lml_int8_tensor": res[1][0],
                 "lml_rel_diff": abs(res[1][0] - res[0][0]) / abs(res[0][0]),
                 "grad_max_abs_diff_over_max_abs": float(np.abs(res[1][1][ok_] - res[0][1][ok_]).max() / np.abs(res[0][1][ok_]).max())}
    nat.lib().grief_set_gemm_mode({"int8": 1, "int8x2": 3, "fp64": 0}[args.gemm])

    # ---- device-resident leg: one model, data stays in HBM, new hyper-parameters every step ----
    model = make_model(0)
    d_theta = 1 + 2 * d

    def evaluate(step):
        prm = model.parameters
        prm[2:d_theta:2] = step_lengthscales(d, step)         # lengthscales of all d kernels
        model.parameters = prm
        return model.log_likelihood(return_gradient=True)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        evaluate(s)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nat.profile_enable(True)
    nat.profile_read()
    nat.lib().grief_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        lml, grad = evaluate(args.warmup + s)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = nat.lib().grief_launch_count()
    prof = nat.profile_read()
    nat.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = float(ms.item())
    value = args.steps / (ms_total * 1e-3)
    lml_val = float(np.asarray(lml).squeeze())

    # ---- end-to-end leg: host buffers in, host results out, every step (public API call a user makes) ----
    e2e = None
    if not args.no_e2e:
        del model
        torch.cuda.empty_cache()
        steps_e2e = max(1, args.steps)
        m0 = make_model(0)                                    # warm: library handles and the caching allocator's pool (a process that
        m0.log_likelihood(return_gradient=True)               # evaluates repeatedly keeps its device buffers between calls)
        del m0
        barrier()
        t0 = time.perf_counter()
        for s in range(steps_e2e):
            mm = make_model(100 + s)                           # H2D copy of X and y from pinned host memory
            l_, g_ = mm.log_likelihood(return_gradient=True)  # D2H of LML and gradient
            float(np.asarray(l_).squeeze()); np.asarray(g_)
            del mm
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": steps_e2e / float(t_e2e.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(n_local * (d + 1) * 8), "d2h_bytes_per_step": int((d_theta + p) * 8 + 8),
               "steps": steps_e2e}

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel: k_gemm_nt computing Z = Phi*G2 (pass 2, 2/3 of the flops); live CUDA-event timing ----
    p_pad = (p + 127) // 128 * 128
    rows128 = (n_local + 127) // 128 * 128
    kern_rows = []
    # work actually issued by the GEMM launches (padded rows / columns, full diagonal tiles)
    flops = {"k_zgemm": 2.0 * rows128 * p_pad * p_pad, "k_gram": float(rows128) * p_pad * (p_pad + 128)}
    gk = {"int8": "k_ozaki<1>", "int8x2": "k_ozaki<2> (cta_group::2 pairs)", "fp64": "k_gemm_nt"}[args.gemm]
    sl = " (exponents + int8 digit planes)" if args.gemm != "fp64" else ""
    label = {"k_zgemm": gk + " [Z = Phi*G2, pass 2]", "k_gram": gk + " [A = Phi^T Phi, lower tiles, split K]",
             "k_build_phi": "k_build_phi%s [Phi slab, pass 2]" % sl, "k_build_phi_t": "k_build_phi_t%s [Phi^T slab, pass 1]" % sl,
             "solve": "dense p x p stage (k_potf2_inv, k_gemm_nt, k_trsv_step, k_assemble, ...)"}
    for name in ("k_zgemm", "k_gram", "k_contract", "k_build_phi", "k_build_phi_t", "solve", "k_dtables", "k_tables", "phi_t_y", "k_topk"):
        t_ms, cnt = prof.get(name, (0.0, 0))
        if cnt:
            row = {"kernel": label.get(name, name), "launches": cnt, "ms_total": t_ms, "share_of_step": t_ms / ms_total}
            if name in flops:
                row["issued_tflops"] = flops[name] * args.steps / (t_ms * 1e-3) * 1e-12
            kern_rows.append(row)
    dom = kern_rows[0] if kern_rows else None
    roofline = None
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))["k_ozaki_zgemm" if args.gemm != "fp64" else "k_gemm_nt_zgemm"]
        if tr["config"] == args.config:
            traffic = {"bytes_per_launch": tr["dram_bytes_read"] + tr["dram_bytes_write"], "rows_per_launch": tr["rows_per_launch"],
                       "algorithmic_bytes_per_launch": tr["algorithmic_bytes_per_launch"], "source": tr["source"]}
    except Exception:
        traffic = None
    if dom:
        algo = 2.0 * n_local * p * p * args.steps / (dom["ms_total"] * 1e-3) * 1e-12      # algorithmic: 2 n p^2 per evaluation
        common = {"traffic": traffic, "avg_launch_ms": dom["ms_total"] / dom["launches"], "kernels": kern_rows,
                  "fp64_equivalent_tflops": algo, "cublas_dgemm_tflops_this_run": peak_sust, "fp64_dmma_issue_peak_tflops": 37.2,
                  "whole_eval_fp64_equivalent_tflops": 3.0 * n_total * p * p * value / world * 1e-12}
        if args.gemm != "fp64":
            # 28 exact int8 digit GEMMs per FP64 GEMM; peak = dense INT8 = 2 x the dense bf16 rate measured on this pool
            tops = 28.0 * flops["k_zgemm"] * args.steps / (dom["ms_total"] * 1e-3) * 1e-12
            try:
                mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                peak_i8, src = 2.0 * float(mp["bf16_tflops_sustained"]), "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (kernel timed inside a long step)"
            except Exception:
                peak_i8, src = 4500.0, "nominal dense INT8 4.5 POP/s (MEASURED_PEAKS.json missing)"
            roofline = dict(common, bound="tensor", kernel="k_ozaki (tcgen05 kind::i8, TMEM accumulators, TMA digit planes): Z = Phi*G2 for pass 2 as 28 "
                            "exact int8 x int8 -> int32 digit GEMMs per slab of 37888 rows; digits cut by k_slice from the Phi slab of k_build_phi",
                            achieved=tops, peak=peak_i8, unit="TOP/s (int8)", frac=tops / peak_i8, peak_source=src)
        else:
            roofline = dict(common, bound="tensor", kernel="k_gemm_nt (TMA-fed FP64 DMMA GEMM) computing Z = Phi*G2 for pass 2, one launch per slab "
                            "of 37888 rows; Phi slab staged in HBM by k_build_phi",
                            achieved=algo, peak=peak_sust, unit="TFLOP/s", frac=algo / peak_sust,
                            peak_source="cuBLAS DGEMM 8192^3 measured in this run, sustained %.1f / burst %.1f TFLOP/s; MEASURED_PEAKS.json has "
                            "no FP64 row; FP64 DMMA issue-rate peak 37.2 TFLOP/s (profiles/r01_fp64_pipes_microbench.txt)" % (peak_sust, peak_burst))
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(args.config, n_total, args.cpu_rows)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.gemm == "fp64" else "f64 (O(n p^2) products as exact int8 digit GEMMs on the tensor cores, 54-bit operands)",
            "data": "synthetic",
            "config": {"workload": "%s: Type-II GRIEF LML+gradient, n=%d, d=%d, m=%d grid pts/dim, p=%d, RBF kernels, "
                                   "new lengthscales every step" % (args.config, n_total, d, m, p),
                       "gemm_arithmetic": args.gemm, "rows_per_gpu": n_local, "parallelism": "rows sharded over %d rank(s), NCCL all-reduce of (A|r|s) and of the "
                                                               "theta-gradient" % world,
                       "l2": "inputs (X %.1f GB, tables %.1f GB per GPU) exceed the 126 MB L2" % (n_local * d * 8e-9, rows128 * 105 * 8e-9)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "check": {"lml": lml_val, "grad_finite": bool(np.all(np.isfinite(grad[~np.isnan(grad)]))), "int8_vs_fp64_on_sample": arith},
            "tflops_whole_eval": 3.0 * n_total * p * p * value * 1e-12}
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
