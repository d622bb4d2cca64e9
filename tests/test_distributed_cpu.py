"""World-size-2 gloo test of the multi-GPU host logic: row shards + one all-reduce of the packed statistics
reproduce the single-rank statistics (the arithmetic per shard is done by the CPU oracle here; on the GPU box
the same packing / reduction code path runs over NCCL, see tests/test_multi_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gp_grief_b200.sharding import row_shard, stats_layout
    from gp_grief_b200.synthetic import synthetic_xy, linspace_grid
    from oracle import grief_oracle as orc
    n, d, m, p = 1500, 3, 6, 40
    ls = [0.3, 0.4, 0.5]
    xg = linspace_grid(d, m)
    basis = orc.setup_inducing_cov(["RBF"] * d, [1.0] * d, ls, xg, p)
    r0, r1 = row_shard(n, world, rank)
    x, y = synthetic_xy(r1 - r0, d, chunk=256, row0=r0)
    Phi = orc.grief_phi(basis, ["RBF"] * d, [1.0] * d, ls, xg, x)
    lay = stats_layout(p)
    buf = torch.zeros(lay["size"], dtype=torch.float64)
    buf[lay["A"][0]:lay["A"][1]] = torch.from_numpy(Phi.T.dot(Phi).reshape(-1))
    buf[lay["r"][0]:lay["r"][1]] = torch.from_numpy(Phi.T.dot(y).reshape(-1))
    buf[lay["s"][0]:lay["s"][1]] = float((y ** 2).sum())
    cnt = torch.tensor([r1 - r0], dtype=torch.int64)
    dist.all_reduce(buf)
    dist.all_reduce(cnt)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.concatenate([buf.numpy(), [float(cnt.item())]]))
    dist.destroy_process_group()


def test_two_rank_statistics_match_single_rank(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from gp_grief_b200.sharding import stats_layout
    from gp_grief_b200.synthetic import synthetic_xy, linspace_grid
    from oracle import grief_oracle as orc
    n, d, m, p = 1500, 3, 6, 40
    ls = [0.3, 0.4, 0.5]
    xg = linspace_grid(d, m)
    basis = orc.setup_inducing_cov(["RBF"] * d, [1.0] * d, ls, xg, p)
    x, y = synthetic_xy(n, d, chunk=256)
    Phi = orc.grief_phi(basis, ["RBF"] * d, [1.0] * d, ls, xg, x)
    lay = stats_layout(p)
    r0 = np.load(os.path.join(str(tmp_path), "rank0.npy"))
    r1 = np.load(os.path.join(str(tmp_path), "rank1.npy"))
    np.testing.assert_array_equal(r0, r1)                       # every rank ends with identical statistics
    assert r0[-1] == n
    A = r0[lay["A"][0]:lay["A"][1]].reshape(p, p)
    np.testing.assert_allclose(A, Phi.T.dot(Phi), rtol=0, atol=1e-12 * np.abs(A).max())
    np.testing.assert_allclose(r0[lay["r"][0]:lay["r"][1]], Phi.T.dot(y).reshape(-1), rtol=0, atol=1e-12 * np.abs(A).max())
    np.testing.assert_allclose(r0[lay["s"][0]], float((y ** 2).sum()), rtol=1e-13)
    f1 = orc.fit_from_phi(Phi, y, np.ones(p), 0.1)
    lml2, _ = orc.gpweb_lml_grad(A, r0[lay["r"][0]:lay["r"][1]], r0[lay["s"][0]], n, np.ones(p), 0.1)
    np.testing.assert_allclose(lml2, f1.lml, rtol=1e-11)
