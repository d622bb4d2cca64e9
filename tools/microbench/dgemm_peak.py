"""Measure the cuBLAS FP64 GEMM throughput of this B200 (roofline denominator for the fused SYRK).

Burst = best of 10 launches of an 8192^3 DGEMM; sustained = back-to-back launches for ~4 s,
with nvidia-smi clocks sampled during the sustained loop.  Writes one JSON line.
"""
import json, subprocess, sys, time, threading
import torch

def clocks_sampler(stop, out):
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                                "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=5)
            out.append(r.stdout.strip())
        except Exception as e:  # pragma: no cover
            out.append("err " + str(e))
        time.sleep(0.2)

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    dev = torch.device("cuda:0")
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    flops = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = flops / best * 1e-9
    # A^T A as cuBLAS would do it when Phi is materialised (gemm with transposed operand)
    bestT = 1e30
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a.t(), a, out=c); e1.record(); torch.cuda.synchronize()
        bestT = min(bestT, e0.elapsed_time(e1))
    stop = threading.Event(); samples = []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples)); th.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.time(); k = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(5):
            torch.matmul(a, b, out=c); k += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    sustained = flops * k / e0.elapsed_time(e1) * 1e-9
    print(json.dumps({"n": n, "fp64_tflops_burst": burst, "fp64_tflops_sustained": sustained,
                      "fp64_tflops_AtA_burst": flops / bestT * 1e-9, "launches": k, "clock_samples": samples[:40]}))

if __name__ == "__main__":
    main()
