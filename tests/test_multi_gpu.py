"""Row-sharded (NCCL) evaluation equals the single-GPU evaluation.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(x, y, distributed):
    import gp_grief_b200 as gp
    from gp_grief_b200.synthetic import linspace_grid
    d, m, p = x.shape[1], 8, 96
    grid = gp.grid.InducingGrid(xg=[g.reshape(-1, 1) for g in linspace_grid(d, m)])
    kern = gp.kern.GriefKernel([gp.kern.RBF(1, lengthscale=0.3 + 0.05 * i) for i in range(d)], grid, n_eigs=p,
                               reweight_eig_funs=False, opt_kernel_params=True)
    return gp.models.GPGriefModel(x, y, kern, noise_var=0.1, distributed=distributed)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gp_grief_b200.sharding import row_shard
    from gp_grief_b200.synthetic import synthetic_xy
    n, d = 50_001, 4
    r0, r1 = row_shard(n, world, rank)
    x, y = synthetic_xy(r1 - r0, d, chunk=4096, row0=r0)
    m = _build(x, y, True)
    ll, g = m.log_likelihood(return_gradient=True)
    yhat, var = m.predict(x[:5], compute_var='diag')
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ll=np.asarray(ll), g=g, n=m.num_data, yhat=yhat)
    dist.destroy_process_group()


def test_two_gpus_match_one(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from gp_grief_b200.synthetic import synthetic_xy
    n, d = 50_001, 4
    x, y = synthetic_xy(n, d, chunk=4096)
    m = _build(x, y, False)
    ll, g = m.log_likelihood(return_gradient=True)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]
    assert int(r[0]["n"]) == n
    np.testing.assert_array_equal(r[0]["ll"], r[1]["ll"])            # identical on every rank
    np.testing.assert_array_equal(r[0]["g"], r[1]["g"])
    np.testing.assert_allclose(float(r[0]["ll"].squeeze()), float(np.asarray(ll).squeeze()), rtol=1e-12)
    free = ~np.isnan(g)
    np.testing.assert_allclose(r[0]["g"][free], g[free], rtol=1e-9, atol=1e-9 * np.abs(g[free]).max())


def _native_comm_worker(rank, world, out_dir):
    """The C ABI's own communicator (grief_comm_*): rank 0 publishes the NCCL id through a file, both ranks all-reduce the packed
    statistics of their row shards -- the path a host without torch.distributed uses."""
    sys.path.insert(0, ROOT)
    import time
    import torch
    torch.cuda.set_device(rank)
    from gp_grief_b200.device import NativeComm
    from gp_grief_b200.sharding import row_shard
    from gp_grief_b200.synthetic import synthetic_xy
    id_path = os.path.join(out_dir, "nccl_id.bin")
    if rank == 0:
        uid = NativeComm.unique_id()
        with open(id_path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(id_path + ".tmp", id_path)
    else:
        for _ in range(600):
            if os.path.exists(id_path):
                break
            time.sleep(0.05)
        uid = open(id_path, "rb").read()
    comm = NativeComm(uid, world, rank)
    n, d = 20_003, 4
    r0, r1 = row_shard(n, world, rank)
    x, y = synthetic_xy(r1 - r0, d, chunk=4096, row0=r0)
    m = _build(x, y, False)                        # local statistics of the shard
    st = m._stats()
    buf = st["buf"].clone()
    comm.all_reduce(buf)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, "native_rank%d.npy" % rank), buf.cpu().numpy())
    comm.close()


def test_native_comm_allreduce_matches_single_gpu_statistics(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_native_comm_worker, args=(world, str(tmp_path)), nprocs=world, join=True)
    from gp_grief_b200.synthetic import synthetic_xy
    n, d = 20_003, 4
    x, y = synthetic_xy(n, d, chunk=4096)
    ref = _build(x, y, False)._stats()["buf"].cpu().numpy()
    r = [np.load(os.path.join(str(tmp_path), "native_rank%d.npy" % k)) for k in range(world)]
    np.testing.assert_array_equal(r[0], r[1])
    np.testing.assert_allclose(r[0], ref, rtol=0, atol=1e-12 * np.abs(ref).max())
