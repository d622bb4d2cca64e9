"""Device time of the top-p selection kernel at the benchmark shapes (CUDA events via the library profiler)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gp_grief_b200 import device, _native as nat
import torch
for name in ["topk_c2_d6_m10_p1024", "topk_c3_d10_m20_p4096", "topk_c4_d32_m8_p2048", "topk_c5_d8_m16_p8192"]:
    z = np.load("tests/golden/%s.npz" % name)
    d, p = int(z["d"]), int(z["p"])
    eigs = [z["eigs_%d" % k] for k in range(d)]
    device.topk_kron(eigs, p)
    nat.profile_enable(True); nat.profile_read()
    t0 = time.perf_counter()
    for _ in range(5):
        device.topk_kron(eigs, p)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5
    ms, cnt = nat.profile_read()["k_topk"]
    nat.profile_enable(False)
    print("%-26s kernel %.3f ms  wall %.3f ms" % (name, ms / cnt, wall * 1e3))
