"""Rate of the TMA-fed FP64 GEMM core (grief_gemm_nt) at the hot-path shapes: python tools/perf_gemm.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gp_grief_b200 import _native as nat

def run(M, N, K, tag):
    A = torch.randn((M, K), dtype=torch.float64, device="cuda")
    B = torch.randn((N, K), dtype=torch.float64, device="cuda")
    C = torch.empty((M, N), dtype=torch.float64, device="cuda")
    call = lambda: nat.check(nat.lib().grief_gemm_nt(nat.dev_ptr(A), K, nat.dev_ptr(B), K, nat.dev_ptr(C), N, M, N, K, 1.0, 0.0, nat.stream_ptr()))
    call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [call() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); [torch.matmul(A, B.T, out=C) for _ in range(3)]; t1.record(); torch.cuda.synchronize()
    ref = A[:256] @ B[:256].T
    call(); err = (C[:256, :256] - ref).abs().max().item()
    print(f"{tag}: M={M} N={N} K={K}: ours {ms:.2f} ms = {2*M*N*K/ms/1e9:.2f} TFLOP/s; cuBLAS {t0.elapsed_time(t1)/3:.2f} ms = {2*M*N*K/(t0.elapsed_time(t1)/3)/1e9:.2f} TFLOP/s; err {err:.1e}")

run(148 * 256, 4096, 4096, "zgemm-shape")
run(4096, 4096, 148 * 256, "gram-shape (full, not SYRK)")
run(8192, 8192, 8192, "square")
