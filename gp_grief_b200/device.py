"""Thin object layer over the C ABI: owns device buffers (torch tensors) and marshals pointers.

Nothing here computes; every method is one or two calls into libgrief_b200.so.
"""
import ctypes

import numpy as np

from . import _native as nat


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise nat.NativeLibraryError("gp_grief_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch


def topk_kron(eig_list, n_eigs):
    """Device top-p selection; same contract as KronMatrix.find_extremum_eigs(..., 'largest', log_expand=True).

    eig_list: 1-D arrays in KronMatrix.K order.  Returns (eig_loc (p,d) int64, log_lam (p,)) as NumPy.
    """
    torch = _torch()
    d = len(eig_list)
    eig_list = [np.ascontiguousarray(e, dtype=np.float64).reshape(-1) for e in eig_list]
    m = np.array([e.size for e in eig_list], dtype=np.int32)
    with np.errstate(divide="ignore", invalid="ignore"):
        logeig = np.concatenate([np.log(e) for e in eig_list])   # the reference's own call (linalg.py:86-89)
    raw0 = eig_list[0]
    p = int(n_eigs)
    idx = torch.empty((p, d), dtype=torch.int32, device="cuda")
    lam = torch.empty((p,), dtype=torch.float64, device="cuda")
    n_out = ctypes.c_int(0)
    nat.check(nat.lib().grief_topk_kron(d, nat.host_ptr(m), nat.host_ptr(raw0), nat.host_ptr(logeig), p,
                                        nat.dev_ptr(idx), nat.dev_ptr(lam), ctypes.byref(n_out), nat.stream_ptr()))
    k = n_out.value
    return idx[:k].cpu().numpy().astype(np.int64), lam[:k].cpu().numpy()


class DevicePlan(object):
    """One basis on the device (grief_plan). All list arguments are in INPUT-dimension order."""

    #: rows of K_xu evaluated on the host and uploaded per call of grief_build_tables_kxu (host-kernel dimensions only)
    host_chunk_rows = 1 << 16

    def __init__(self, kernel_names, variances, lengthscales, xg, Q, eig, eig_loc, width_cap=0, host_kernels=None):
        """kernel_names[i]: 'RBF' | 'Exponential' | 'Matern32' | 'Matern52' (evaluated on the device) or 'host' -- then
        host_kernels[i] is the BaseKernel whose `cov` (and `grad_x`, for d/dx tables) is called per row chunk."""
        _torch()
        d = len(xg)
        self.host_kernels = {int(i): k for i, k in (host_kernels or {}).items()}
        assert sorted(self.host_kernels) == [i for i, k in enumerate(kernel_names) if k == "host"], "host_kernels must match the 'host' entries"
        self._xg = [np.asarray(g, dtype=np.float64).reshape(-1, 1) for g in xg]
        self._grid_off = np.concatenate([[0], np.cumsum([g.size for g in self._xg])]).astype(int)
        eig_loc = np.asarray(eig_loc)
        p = eig_loc.shape[0]
        m = np.array([np.size(g) for g in xg], dtype=np.int32)
        kid = np.array([nat.KERNEL_IDS[k] for k in kernel_names], dtype=np.int32)
        var = np.ascontiguousarray(variances, dtype=np.float64)
        ls = np.ascontiguousarray(lengthscales, dtype=np.float64)
        grid = np.concatenate([np.asarray(g, dtype=np.float64).reshape(-1) for g in xg])
        u = np.zeros(d, dtype=np.int32)
        uinv = np.zeros((p, d), dtype=np.int32)
        qs = []
        self.unique = []
        for i in range(d):
            uniq, inv = np.unique(eig_loc[:, i], return_inverse=True)   # tensors/selection_matrix.py:78-79
            self.unique.append(uniq)
            u[i] = uniq.size
            uinv[:, i] = inv.reshape(-1)
            lam = np.asarray(eig[i], dtype=np.float64)[uniq]
            qs.append(np.ascontiguousarray(np.asarray(Q[i], dtype=np.float64)[:, uniq] / np.sqrt(lam)[None, :]).reshape(-1))
        qs = np.concatenate(qs)
        handle = ctypes.c_void_p()
        nat.check(nat.lib().grief_plan_create(ctypes.byref(handle), d, nat.host_ptr(m), nat.host_ptr(kid),
                                              nat.host_ptr(var), nat.host_ptr(ls), nat.host_ptr(grid), nat.host_ptr(u),
                                              nat.host_ptr(qs), p, nat.host_ptr(np.ascontiguousarray(uinv)),
                                              int(width_cap)))
        self._h = handle
        info = nat.lib().grief_plan_info
        self.n_groups, self.width, self.stride = info(handle, 0), info(handle, 1), info(handle, 2)
        self.p, self.p_pad, self.d = info(handle, 3), info(handle, 4), info(handle, 5)

    def close(self):
        if getattr(self, "_h", None):
            nat.lib().grief_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- row kernels ----
    def build_tables(self, X_dev, deriv_dim=None):
        """X_dev: (n, d) float64 CUDA tensor (row-major).  Returns the (n_pad, stride) table tensor.

        deriv_dim: input dimension whose kernel is replaced by its x-derivative (tables of d Phi / d x[:, deriv_dim])."""
        torch = _torch()
        assert X_dev.is_cuda and X_dev.dtype == torch.float64 and X_dev.dim() == 2 and X_dev.shape[1] == self.d
        assert X_dev.stride(1) == 1 or X_dev.shape[0] == 0
        n = X_dev.shape[0]
        rows = nat.lib().grief_table_rows(n)
        T = torch.empty((rows, self.stride), dtype=torch.float64, device=X_dev.device)
        ldx = X_dev.stride(0) if n > 1 else self.d
        if self.host_kernels:
            return self._build_tables_host_kxu(X_dev, ldx, n, T, deriv_dim)
        if deriv_dim is None:
            nat.check(nat.lib().grief_build_tables(self._h, nat.dev_ptr(X_dev), ldx, n, nat.dev_ptr(T), nat.stream_ptr()))
        else:
            nat.check(nat.lib().grief_build_tables_dx(self._h, nat.dev_ptr(X_dev), ldx, n, int(deriv_dim), nat.dev_ptr(T),
                                                      nat.stream_ptr()))
        return T

    def _build_tables_host_kxu(self, X_dev, ldx, n, T, deriv_dim):
        """Tables when some dimensions have host-only kernels (GridKernel.cov_kr, kern/grid_kernel.py:148-179, for those dimensions):
        per chunk of rows, kern_i.cov(x[:, i], U_i) is evaluated on the host into a pinned (rows, sum m) buffer, uploaded, and
        grief_build_tables_kxu projects / multiplies on the device.  Two pinned buffers alternate so that the host evaluation of
        chunk c + 1 overlaps the upload and the table kernel of chunk c."""
        torch = _torch()
        sum_m = int(self._grid_off[-1])
        chunk = int(self.host_chunk_rows) // 128 * 128
        assert chunk >= 128
        if n == 0:
            return T
        pins = [torch.empty((min(chunk, n), sum_m), dtype=torch.float64).pin_memory() for _ in range(2)]
        devs = [torch.empty((min(chunk, n), sum_m), dtype=torch.float64, device=X_dev.device) for _ in range(2)]
        done = [None, None]
        stream = torch.cuda.current_stream()
        dd = -1 if deriv_dim is None else int(deriv_dim)
        for c, r0 in enumerate(range(0, n, chunk)):
            rows = min(chunk, n - r0)
            b = c & 1
            if done[b] is not None:
                done[b].synchronize()                        # the copy out of this pinned buffer has finished
            xh = X_dev[r0:r0 + rows].cpu().numpy()
            kb = pins[b].numpy()
            for i, kern in self.host_kernels.items():
                xi = np.ascontiguousarray(xh[:, (i,)])
                Ki = kern.grad_x(xi, self._xg[i]) if i == dd else kern.cov(x=xi, z=self._xg[i])
                kb[:rows, self._grid_off[i]:self._grid_off[i + 1]] = Ki
            devs[b][:rows].copy_(pins[b][:rows], non_blocking=True)
            done[b] = torch.cuda.Event()
            done[b].record(stream)
            Xc = X_dev[r0:r0 + rows]
            nat.check(nat.lib().grief_build_tables_kxu(self._h, nat.dev_ptr(Xc), ldx, nat.dev_ptr(devs[b]), sum_m, rows, dd,
                                                       nat.dev_ptr(T[r0:]), nat.stream_ptr()))
        return T

    def phi_rows(self, T, n):
        torch = _torch()
        Phi = torch.empty((n, self.p), dtype=torch.float64, device=T.device)
        nat.check(nat.lib().grief_phi_rows(self._h, nat.dev_ptr(T), n, nat.dev_ptr(Phi), nat.stream_ptr()))
        return Phi

    def set_option(self, what, value):
        """Per-plan option (nat.OPT_GEMM_MODE / OPT_DIGITS_GRAM / OPT_DIGITS_Z / OPT_DIGITS_VAR / OPT_SLAB_BUDGET, include/grief_b200.h)."""
        nat.check(nat.lib().grief_plan_set_option(self._h, int(what), int(value)))

    def get_option(self, what):
        return int(nat.lib().grief_plan_get_option(self._h, int(what)))

    def gram(self, T, n, out=None, workspace=None, y=None, r_out=None, rowmax_out=None):
        """A = Phi^T Phi (p, p) from the tables of n rows; with y (n,) also r = Phi^T y from the same sweep (returned in r_out) and,
        with rowmax_out (int32, T.shape[0]), the row maxima of |Phi| that grad_theta can reuse for the same tables."""
        torch = _torch()
        A = out if out is not None else torch.empty((self.p, self.p), dtype=torch.float64, device=T.device)
        need = nat.lib().grief_gram_workspace_bytes(self._h, n)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty((need,), dtype=torch.uint8, device=T.device)
        if y is None:
            nat.check(nat.lib().grief_gram(self._h, nat.dev_ptr(T), n, nat.dev_ptr(A), A.stride(0), nat.dev_ptr(workspace),
                                           workspace.numel(), nat.stream_ptr()))
            return A
        assert r_out is not None and r_out.numel() == self.p and y.numel() == n
        if rowmax_out is not None:
            assert rowmax_out.dtype == torch.int32 and rowmax_out.numel() >= T.shape[0] and rowmax_out.is_contiguous()
        nat.check(nat.lib().grief_gram_ry(self._h, nat.dev_ptr(T), n, nat.dev_ptr(y), nat.dev_ptr(A), A.stride(0),
                                          nat.dev_ptr(r_out), nat.dev_ptr(rowmax_out), nat.dev_ptr(workspace), workspace.numel(),
                                          nat.stream_ptr()))
        return A

    def gram_workspace_bytes(self, n):
        return nat.lib().grief_gram_workspace_bytes(self._h, n)

    def phi_t_vec(self, T, n, v, out=None):
        torch = _torch()
        out = out if out is not None else torch.empty((self.p,), dtype=torch.float64, device=T.device)
        need = nat.lib().grief_phi_t_vec_workspace_bytes(self._h, n)
        ws = torch.empty((max(need, 8),), dtype=torch.uint8, device=T.device)
        nat.check(nat.lib().grief_phi_t_vec(self._h, nat.dev_ptr(T), n, nat.dev_ptr(v), nat.dev_ptr(out), nat.dev_ptr(ws),
                                            nat.stream_ptr()))
        return out

    def phi_vec(self, T, n, v):
        torch = _torch()
        out = torch.empty((n,), dtype=torch.float64, device=T.device)
        nat.check(nat.lib().grief_phi_vec(self._h, nat.dev_ptr(T), n, nat.dev_ptr(v), nat.dev_ptr(out), nat.stream_ptr()))
        return out


    # ---- pass 2: analytic hyper-parameter gradient ----
    def grad_setup(self, dims, kinds, dqs_list):
        """Declare the active parameters: dims[a], kinds[a] (0 variance, 1 lengthscale), dqs_list[a] (m_i, u_i)."""
        na = len(dims)
        dims_a = np.ascontiguousarray(dims, dtype=np.int32)
        kinds_a = np.ascontiguousarray(kinds, dtype=np.int32)
        dqs = np.concatenate([np.ascontiguousarray(q, dtype=np.float64).reshape(-1) for q in dqs_list]) if na else np.zeros(1)
        nat.check(nat.lib().grief_grad_setup(self._h, na, nat.host_ptr(dims_a), nat.host_ptr(kinds_a), nat.host_ptr(dqs)))
        self.n_active = na

    def grad_theta(self, T, X_dev, y_dev, n, Pinv, b, noise_var, rowmax=None):
        """d LML / d theta of the active parameters from P^-1 (p, p) and b = P^-1 r (device tensors); rowmax: the row maxima
        `gram(..., rowmax_out=)` recorded for the same tables (optional, saves a sweep over Phi)."""
        torch = _torch()
        g = torch.zeros((max(self.n_active, 1),), dtype=torch.float64, device=T.device)
        need = nat.lib().grief_grad_workspace_bytes(self._h, n)
        ws = torch.empty((max(need, 256),), dtype=torch.uint8, device=T.device)
        ldx = X_dev.stride(0) if n > 1 else self.d
        Pinv = _even_ld(Pinv)
        nat.check(nat.lib().grief_grad_theta(self._h, nat.dev_ptr(T), nat.dev_ptr(X_dev), ldx, nat.dev_ptr(y_dev), n,
                                             nat.dev_ptr(Pinv), Pinv.stride(0), nat.dev_ptr(b), float(noise_var), nat.dev_ptr(rowmax),
                                             nat.dev_ptr(g), nat.dev_ptr(ws), ws.numel(), nat.stream_ptr()))
        return g[:self.n_active]

    def quadform_rows(self, T, n, B):
        """q[i] = phi_i^T B phi_i for a symmetric (p, p) device matrix B."""
        torch = _torch()
        q = torch.empty((n,), dtype=torch.float64, device=T.device)
        need = nat.lib().grief_quadform_workspace_bytes(self._h, n)
        ws = torch.empty((max(need, 256),), dtype=torch.uint8, device=T.device)
        B = _even_ld(B)
        nat.check(nat.lib().grief_quadform_rows(self._h, nat.dev_ptr(T), n, nat.dev_ptr(B), B.stride(0), nat.dev_ptr(q),
                                                nat.dev_ptr(ws), ws.numel(), nat.stream_ptr()))
        return q


def gemm_nt(A, B, alpha=1.0, beta=0.0, out=None):
    """out = beta * out + alpha * A @ B.T on the library's FP64 DMMA GEMM (A: (M, K), B: (N, K) device tensors)."""
    torch = _torch()
    assert A.dim() == 2 and B.dim() == 2 and A.shape[1] == B.shape[1], "gemm_nt: A (M, K) and B (N, K)"
    A, B = _even_ld(A), _even_ld(B)
    M, K = A.shape
    N = B.shape[0]
    if out is None:
        assert beta == 0.0
        out = torch.empty((M, N), dtype=torch.float64, device=A.device)
    if M and N and K:
        nat.check(nat.lib().grief_gemm_nt(nat.dev_ptr(A), A.stride(0), nat.dev_ptr(B), B.stride(0), nat.dev_ptr(out), out.stride(0),
                                          M, N, K, float(alpha), float(beta), nat.stream_ptr()))
    elif beta == 0.0:
        out.zero_()
    return out


def kron_matvec(factors, x):
    """(K_1 kron ... kron K_d) x on the device: one GEMM per factor (reference tensors/kron_matrix.py:52-97, host loop there).

    factors: list of 2-D NumPy arrays, x: (N,) or (N, 1) NumPy array.  Returns a NumPy column vector.
    The running vector is kept as a (rest, cols_i) row-major matrix Yc (the memory image of the reference's F-ordered
    reshape); factor i is applied as (Yc @ K_i^T)^T through grief_gemm_nt_t, whose transposed store leaves the result in the
    layout the next factor needs -- each step reads and writes the running matrix once.
    """
    torch = _torch()
    y = torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))).cuda()
    for Ki in reversed(list(factors)):
        Kd = _even_ld(torch.as_tensor(np.ascontiguousarray(Ki, dtype=np.float64)).cuda())
        rows_i, cols_i = Kd.shape
        Yc = _even_ld(y.view(-1, cols_i))
        rest = Yc.shape[0]
        out_t = torch.empty((rows_i, rest), dtype=torch.float64, device=y.device)     # = (Yc @ K_i^T)^T
        if rest and rows_i:
            nat.check(nat.lib().grief_gemm_nt_t(nat.dev_ptr(Yc), Yc.stride(0), nat.dev_ptr(Kd), Kd.stride(0), nat.dev_ptr(out_t), rest,
                                                rest, rows_i, cols_i, 1.0, 0.0, nat.stream_ptr()))
        y = out_t.view(-1)
    return y.cpu().numpy().reshape((-1, 1))


def rowcol_kr_matvec(R, C, x):
    """y = (R K C) x for a row- and column-partitioned Khatri-Rao product on the device (grief_rowcol_kr_matvec).

    R: per factor a dense (rows, m_t) array, or a 1-D integer array of length rows (selection matrix: row i picks row R[t][i] of
    C[t]); C: per factor a dense (m_t, cols) array (already multiplied by the Kronecker factor, if any); x: (cols,) or (cols, 1).
    Returns a NumPy column vector (rows, 1)."""
    torch = _torch()
    d = len(C)
    assert len(R) == d and d >= 1
    keep = []                                              # device tensors must outlive the launch
    m = np.zeros(d, dtype=np.int32)
    Rp, Ip, Cp = (ctypes.c_void_p * d)(), (ctypes.c_void_p * d)(), (ctypes.c_void_p * d)()
    rows = cols = None
    for t in range(d):
        Ct = torch.as_tensor(np.ascontiguousarray(C[t], dtype=np.float64)).cuda()
        assert Ct.dim() == 2
        m[t] = Ct.shape[0]
        cols = Ct.shape[1] if cols is None else cols
        assert Ct.shape[1] == cols, "all C factors must have the same number of columns"
        Rt = np.asarray(R[t])
        if Rt.ndim == 1:
            assert np.issubdtype(Rt.dtype, np.integer) and (Rt.size == 0 or (Rt.min() >= 0 and Rt.max() < m[t]))
            Rd = torch.as_tensor(np.ascontiguousarray(Rt, dtype=np.int32)).cuda()
            Ip[t], Rp[t] = nat.dev_ptr(Rd), None
        else:
            assert Rt.ndim == 2 and Rt.shape[1] == m[t], "R[%d] must be (rows, %d)" % (t, m[t])
            Rd = torch.as_tensor(np.ascontiguousarray(Rt, dtype=np.float64)).cuda()
            Rp[t], Ip[t] = nat.dev_ptr(Rd), None
        rows = Rt.shape[0] if rows is None else rows
        assert Rt.shape[0] == rows, "all R factors must have the same number of rows"
        Cp[t] = nat.dev_ptr(Ct)
        keep += [Ct, Rd]
    xd = torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))).cuda()
    assert xd.numel() == cols
    y = torch.zeros((rows,), dtype=torch.float64, device="cuda")
    nat.check(nat.lib().grief_rowcol_kr_matvec(d, nat.host_ptr(m), Rp, Ip, Cp, rows, cols, nat.dev_ptr(xd), nat.dev_ptr(y), nat.stream_ptr()))
    return y.cpu().numpy().reshape((-1, 1))


def _even_ld(B):
    """TMA needs 16-byte aligned rows: unit column stride, even row stride, aligned base -- else a padded copy (small operands)."""
    if B.stride(1) == 1 and B.stride(0) % 2 == 0 and B.data_ptr() % 16 == 0 and B.stride(0) >= B.shape[1]:
        return B
    import torch
    pad = torch.zeros((B.shape[0], B.shape[1] + (B.shape[1] & 1)), dtype=B.dtype, device=B.device)
    pad[:, :B.shape[1]] = B
    return pad[:, :B.shape[1]]


def sumsq(y_dev):
    torch = _torch()
    out = torch.empty((1,), dtype=torch.float64, device=y_dev.device)
    ws = torch.empty((512,), dtype=torch.float64, device=y_dev.device)
    nat.check(nat.lib().grief_sumsq(nat.dev_ptr(y_dev), y_dev.numel(), nat.dev_ptr(out), nat.dev_ptr(ws), nat.stream_ptr()))
    return out


class NativeComm(object):
    """grief_comm: the library's own NCCL communicator (one rank per process / GPU), for hosts without torch.distributed.
    `NativeComm.unique_id()` on rank 0 gives the 128 bytes every rank passes to the constructor."""

    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(128)
        nat.check(nat.lib().grief_comm_unique_id(buf))
        return buf.raw

    def __init__(self, unique_id, world_size, rank):
        _torch()
        assert len(unique_id) == 128
        h = ctypes.c_void_p()
        nat.check(nat.lib().grief_comm_create(ctypes.byref(h), ctypes.c_char_p(unique_id), int(world_size), int(rank)))
        self._h, self.world_size, self.rank = h, int(world_size), int(rank)

    def all_reduce(self, t):
        """In-place sum over ranks of a contiguous float64 CUDA tensor, on the current stream."""
        torch = _torch()
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        nat.check(nat.lib().grief_comm_allreduce_sum(self._h, nat.dev_ptr(t), t.numel(), nat.stream_ptr()))
        return t

    def close(self):
        if getattr(self, "_h", None):
            nat.lib().grief_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_SHARED_SOLVERS = {}


def shared_solver():
    """The DeviceSolver of the calling thread's current device and stream.  A grief_ctx owns the dense-stage scratch (3 p_pad^2
    doubles: 0.4 GB at p = 4096, 1.6 GB at p = 8192); creating one per model costs a cudaMalloc of that size at its first solve and a
    cudaFree -- a device-wide synchronisation -- when the model is released (measured: 70-180 ms per released model in a process that
    builds a new model per evaluation).  Calls on one stream are ordered, so models of one thread can share the context."""
    import threading
    torch = _torch()
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream, threading.get_ident())
    solver = _SHARED_SOLVERS.get(key)
    if solver is None:
        solver = _SHARED_SOLVERS[key] = DeviceSolver()
    return solver


def release_shared_solvers():
    """Free the scratch of every shared solver (they are re-created on demand)."""
    for solver in list(_SHARED_SOLVERS.values()):
        solver.close()
    _SHARED_SOLVERS.clear()


class DeviceSolver(object):
    """grief_ctx: the p x p stage (Cholesky, solve, LML, w / noise gradients, pass-2 operand)."""

    def __init__(self):
        _torch()
        h = ctypes.c_void_p()
        nat.check(nat.lib().grief_ctx_create(ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            nat.lib().grief_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, A, r, yty, w, noise_var, n_rows, want_grad=True, want_G2=False):
        """Returns dict(lml, yt_alpha, logdet, grad_noise, L, b, Pinv, grad_w, G2) -- tensors stay on the device."""
        torch = _torch()
        p = A.shape[0]
        dev = A.device
        L = torch.empty((p, p), dtype=torch.float64, device=dev)
        b = torch.empty((p,), dtype=torch.float64, device=dev)
        Pinv = torch.empty((p, p), dtype=torch.float64, device=dev) if (want_grad or want_G2) else None
        grad_w = torch.empty((p,), dtype=torch.float64, device=dev) if want_grad else None
        G2 = torch.empty((p, p), dtype=torch.float64, device=dev) if want_G2 else None
        scal = np.zeros(nat.SC_COUNT, dtype=np.float64)
        info = ctypes.c_int(0)
        nat.check(nat.lib().grief_solve_lml(self._h, p, nat.dev_ptr(A), A.stride(0), nat.dev_ptr(r), nat.dev_ptr(yty),
                                            nat.dev_ptr(w), float(noise_var), int(n_rows), nat.dev_ptr(L), nat.dev_ptr(b),
                                            nat.dev_ptr(Pinv), nat.dev_ptr(grad_w), nat.dev_ptr(G2), nat.host_ptr(scal),
                                            ctypes.byref(info), nat.stream_ptr()))
        return dict(lml=float(scal[nat.SC_LML]), yt_alpha=float(scal[nat.SC_YT_ALPHA]), logdet=float(scal[nat.SC_LOGDET]),
                    grad_noise=float(scal[nat.SC_GRAD_NOISE]), L=L, b=b, Pinv=Pinv, grad_w=grad_w, G2=G2)
