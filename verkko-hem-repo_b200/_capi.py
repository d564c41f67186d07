"""ctypes binding of include/vh_femgl.h (lib/libvhfemgl.so).  No fallback: a missing library is an error."""
import ctypes
import os

import numpy as np

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "verkko-hem-repo_b200")
if not os.path.isdir(_PKG):
    _PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_PKG, "lib", "libvhfemgl.so")

_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p

VH_OK = 0
VH_ERR_NOT_CONVERGED = -4


class _MGParams(ctypes.Structure):
    _fields_ = [("pre", ctypes.c_int32), ("post", ctypes.c_int32), ("smoothing_range", ctypes.c_double),
                ("coarse_degree", ctypes.c_int32), ("coarse_range", ctypes.c_double), ("n_power", ctypes.c_int32),
                ("safety", ctypes.c_double)]


MG_DEFAULTS = dict(pre=1, post=1, smoothing_range=4.0, coarse_degree=8, coarse_range=30.0, n_power=8, safety=1.1)


class _Info(ctypes.Structure):
    _fields_ = [("n_owned_dofs", ctypes.c_int64), ("n_local_dofs", ctypes.c_int64), ("nnzb", ctypes.c_int64),
                ("n_fast_rows", ctypes.c_int64), ("n_slow_cells", ctypes.c_int64), ("device_bytes", ctypes.c_int64),
                ("n_packed_blocks", ctypes.c_int64), ("spmv_matrix_free", ctypes.c_int64)]


_lib = None


def have_cuda_lib():
    return os.path.exists(_LIB)


def cuda_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise RuntimeError("CUDA extension %s is missing: build it with __graft_entry__.build() "
                               "(there is no CPU fallback for the hot path)" % _LIB)
        L = ctypes.CDLL(_LIB)
        L.vh_last_error.restype = ctypes.c_char_p
        L.vh_last_error.argtypes = [_vp]
        L.vh_create.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(_vp)]
        L.vh_destroy.argtypes = [_vp]
        L.vh_nccl_unique_id.argtypes = [_vp]
        L.vh_comm_init.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _vp]
        L.vh_set_coefficients.argtypes = [_vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _dp,
                                          ctypes.c_double]
        for f in ("vh_set_solution", "vh_get_solution", "vh_get_newton_update", "vh_get_rhs", "vh_get_residual"):
            getattr(L, f).argtypes = [_vp, _dp]
        L.vh_transfer_solution.argtypes = [_vp, _vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _dp]
        L.vh_mg_attach.argtypes = [_vp, _vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _dp]
        L.vh_set_preconditioner.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(_MGParams)]
        L.vh_comm_share.argtypes = [_vp, _vp]
        L.vh_snapshot_begin.argtypes = [_vp]
        L.vh_snapshot_wait.argtypes = [_vp, ctypes.POINTER(_dp), ctypes.POINTER(_dp)]
        L.vh_mg_get_lambda.argtypes = [_vp, ctypes.c_int, _dp]
        L.vh_assemble.argtypes = [_vp, _dp]
        L.vh_solve.argtypes = [_vp, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), _dp]
        L.vh_line_search_trial.argtypes = [_vp, ctypes.c_double]
        L.vh_residual.argtypes = [_vp, _dp]
        L.vh_accept_trial.argtypes = [_vp]
        L.vh_energy.argtypes = [_vp, ctypes.c_int, _dp]
        L.vh_get_info.argtypes = [_vp, ctypes.POINTER(_Info)]
        L.vh_export_matrix_bsr.argtypes = [_vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _dp]
        L.vh_spmv.argtypes = [_vp, _dp, _dp]
        L.vh_precondition.argtypes = [_vp, _dp, _dp]
        L.vh_time_kernel.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
        L.vh_get_timers.argtypes = [_vp, _dp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
        L.vh_timer_start.argtypes = [_vp]
        L.vh_timer_stop.argtypes = [_vp, ctypes.POINTER(ctypes.c_float)]
        L.vh_measure_fp64_peak.argtypes = [_vp, _dp]
        L.vh_set_spmv_matrix_free.argtypes = [_vp, ctypes.c_int]
        L.vh_validate_mesh_desc.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int]
        _lib = L
    return _lib


def validate_tables(tables):
    """Host-only consistency check of a RankTables descriptor (vh_validate_mesh_desc; needs no GPU).  Returns "" or the finding."""
    msg = ctypes.create_string_buffer(512)
    rc = cuda_lib().vh_validate_mesh_desc(tables.desc_ptr(), msg, 512)
    return "" if rc == VH_OK else (msg.value.decode() or "invalid")


class Context:
    """One GPU context for one mesh on one rank (``vh_ctx``)."""

    def __init__(self, tables, device=0):
        from . import VhError
        self._E = VhError
        self.L = cuda_lib()
        self.tables = tables
        self._h = _vp()
        rc = self.L.vh_create(tables.desc_ptr(), device, ctypes.byref(self._h))
        if rc != VH_OK:
            raise VhError(rc, self.L.vh_last_error(None).decode())
        self.n_owned = 18 * tables.n_owned_nodes

    def _chk(self, rc):
        if rc != VH_OK:
            raise self._E(rc, self.L.vh_last_error(self._h).decode())

    def close(self):
        if self._h:
            self.L.vh_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- multi-GPU ---
    @staticmethod
    def nccl_unique_id():
        buf = ctypes.create_string_buffer(128)
        rc = cuda_lib().vh_nccl_unique_id(ctypes.cast(buf, _vp))
        if rc != VH_OK:
            raise RuntimeError("vh_nccl_unique_id failed: %s" % cuda_lib().vh_last_error(None).decode())
        return buf.raw

    def comm_init(self, rank, n_ranks, unique_id):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._chk(self.L.vh_comm_init(self._h, rank, n_ranks, ctypes.cast(buf, _vp)))

    def comm_share(self, donor):
        """Use the NCCL communicator of another context of this rank (vh_comm_share): no new ncclCommInitRank."""
        self._chk(self.L.vh_comm_share(self._h, donor._h))

    # --- coefficients / state ---
    def set_coefficients(self, K1, K2, K3, alpha, betas, bt):
        b = np.ascontiguousarray(betas, dtype=np.float64)
        self._chk(self.L.vh_set_coefficients(self._h, K1, K2, K3, alpha, b.ctypes.data_as(_dp), bt))

    def set_coef_vector(self, coef):
        """coef = [K1,K2,K3,alpha,beta1..5,bt] (the oracle's layout)."""
        self.set_coefficients(coef[0], coef[1], coef[2], coef[3], coef[4:9], coef[9])

    def set_solution(self, x_owned):
        x = np.ascontiguousarray(x_owned, dtype=np.float64)
        assert x.size == self.n_owned
        self._chk(self.L.vh_set_solution(self._h, x.ctypes.data_as(_dp)))

    def _get(self, fn, out=None):
        if out is None:
            out = np.zeros(self.n_owned)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == self.n_owned
        self._chk(fn(self._h, out.ctypes.data_as(_dp)))
        return out

    def get_solution(self, out=None):
        """Owned part of local_solution; pass a (pinned) float64 array as `out` to receive it without a staging copy."""
        return self._get(self.L.vh_get_solution, out)

    def get_newton_update(self):
        return self._get(self.L.vh_get_newton_update)

    # --- output path (io.cc:106-170): asynchronous snapshot of the two LOCAL vectors DataOut reads ---
    def snapshot_begin(self):
        self._chk(self.L.vh_snapshot_begin(self._h))

    def snapshot_wait(self):
        """(local_solution, Newton update), 18 * n_local doubles each, copied out of the library's pinned buffer."""
        s, u = _dp(), _dp()
        self._chk(self.L.vh_snapshot_wait(self._h, ctypes.byref(s), ctypes.byref(u)))
        n = 18 * self.tables.n_local_nodes
        return np.ctypeslib.as_array(s, shape=(n,)).copy(), np.ctypeslib.as_array(u, shape=(n,)).copy()

    def get_rhs(self):
        return self._get(self.L.vh_get_rhs)

    def get_residual(self):
        return self._get(self.L.vh_get_residual)

    def transfer_solution_from(self, src_ctx, ptr, src_node, weight):
        """Device-side SolutionTransfer::interpolate + constraints_solution.distribute (vh_transfer_solution)."""
        ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        src_node = np.ascontiguousarray(src_node, dtype=np.int32)
        weight = np.ascontiguousarray(weight, dtype=np.float64)
        i32 = ctypes.POINTER(ctypes.c_int32)
        self._chk(self.L.vh_transfer_solution(self._h, src_ctx._h, ptr.size - 1, ptr.ctypes.data_as(i32),
                                              src_node.ctypes.data_as(i32), weight.ctypes.data_as(_dp)))

    # --- preconditioner ---
    def mg_attach(self, coarse_ctx, ptr, coarse_node, weight):
        """Attach the next coarser multigrid level with its prolongation table (vh_mg_attach)."""
        ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        coarse_node = np.ascontiguousarray(coarse_node, dtype=np.int32)
        weight = np.ascontiguousarray(weight, dtype=np.float64)
        i32 = ctypes.POINTER(ctypes.c_int32)
        self._chk(self.L.vh_mg_attach(self._h, coarse_ctx._h, ptr.size - 1, ptr.ctypes.data_as(i32), coarse_node.ctypes.data_as(i32),
                                      weight.ctypes.data_as(_dp)))
        self._mg_coarse = coarse_ctx  # keep the level alive as long as this context

    def set_preconditioner(self, kind, **params):
        """kind: "block-jacobi" | "multigrid" (or 0 | 1); params: fields of vh_mg_params (MG_DEFAULTS).  "chebyshev" = kind 1 on a
        context without attached levels: the Chebyshev polynomial (degree coarse_degree) of block-Jacobi."""
        k = {"block-jacobi": 0, "bj": 0, "multigrid": 1, "mg": 1, "chebyshev": 1}.get(kind, kind)
        p = None
        if params:
            p = _MGParams(**{**MG_DEFAULTS, **params})
        self._chk(self.L.vh_set_preconditioner(self._h, int(k), ctypes.byref(p) if p is not None else None))

    def mg_lambda(self, level=0):
        v = ctypes.c_double()
        self._chk(self.L.vh_mg_get_lambda(self._h, level, ctypes.byref(v)))
        return v.value

    # --- hot path ---
    def assemble(self):
        v = ctypes.c_double()
        self._chk(self.L.vh_assemble(self._h, ctypes.byref(v)))
        return v.value

    def solve(self, tol_rel, max_it=10000, restart=30):
        its = ctypes.c_int()
        res = ctypes.c_double()
        self._chk(self.L.vh_solve(self._h, tol_rel, max_it, restart, ctypes.byref(its), ctypes.byref(res)))
        return its.value, res.value

    def line_search_trial(self, alpha):
        self._chk(self.L.vh_line_search_trial(self._h, alpha))

    def residual(self):
        v = ctypes.c_double()
        self._chk(self.L.vh_residual(self._h, ctypes.byref(v)))
        return v.value

    def accept_trial(self):
        self._chk(self.L.vh_accept_trial(self._h))

    def energy(self, which=0):
        v = ctypes.c_double()
        self._chk(self.L.vh_energy(self._h, which, ctypes.byref(v)))
        return v.value

    # --- introspection ---
    def info(self):
        i = _Info()
        self._chk(self.L.vh_get_info(self._h, ctypes.byref(i)))
        return {k: int(getattr(i, k)) for k, _ in _Info._fields_}

    def export_matrix_bsr(self):
        nnzb = self.info()["nnzb"]
        nb = self.tables.n_owned_nodes
        row_ptr = np.zeros(nb + 1, dtype=np.int32)
        col = np.zeros(nnzb, dtype=np.int32)
        vals = np.zeros((nnzb, 18, 18))
        self._chk(self.L.vh_export_matrix_bsr(self._h, row_ptr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                              col.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), vals.ctypes.data_as(_dp)))
        return row_ptr, col, vals

    def spmv(self, x_owned):
        x = np.ascontiguousarray(x_owned, dtype=np.float64)
        y = np.zeros(self.n_owned)
        self._chk(self.L.vh_spmv(self._h, x.ctypes.data_as(_dp), y.ctypes.data_as(_dp)))
        return y

    def precondition(self, x_owned):
        x = np.ascontiguousarray(x_owned, dtype=np.float64)
        y = np.zeros(self.n_owned)
        self._chk(self.L.vh_precondition(self._h, x.ctypes.data_as(_dp), y.ctypes.data_as(_dp)))
        return y

    def time_kernel(self, what, reps=20, flush_l2=True):
        ms = ctypes.c_float()
        self._chk(self.L.vh_time_kernel(self._h, what, reps, 1 if flush_l2 else 0, ctypes.byref(ms)))
        return ms.value

    def timer_start(self):
        self._chk(self.L.vh_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_float()
        self._chk(self.L.vh_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def set_spmv_matrix_free(self, on=True):
        """0 / False: packed SpMV; 1 / True: matrix-free from the H_q tables; 2: matrix-free and table-free."""
        self._chk(self.L.vh_set_spmv_matrix_free(self._h, int(on)))

    def measure_fp64_peak(self):
        v = ctypes.c_double()
        self._chk(self.L.vh_measure_fp64_peak(self._h, ctypes.byref(v)))
        return v.value

    def timers(self, reset=False):
        ms = (ctypes.c_double * 5)()
        n = ctypes.c_int64()
        self._chk(self.L.vh_get_timers(self._h, ms, ctypes.byref(n), 1 if reset else 0))
        return dict(assemble=ms[0], residual=ms[1], solve=ms[2], vector=ms[3], precond_setup=ms[4], launches=int(n.value))


# ------------------------------------------------------------------------------------------
# FemGL driver mirror (host/femgl.cc) through lib/libvhdriver.so
# ------------------------------------------------------------------------------------------
_drv = None


def driver_lib():
    global _drv
    if _drv is None:
        p = os.path.join(_PKG, "lib", "libvhdriver.so")
        if not os.path.exists(p):
            raise RuntimeError("%s is missing: build it with __graft_entry__.build()" % p)
        L = ctypes.CDLL(p)
        L.vhd_run.restype = _vp
        L.vhd_run.argtypes = [ctypes.c_char_p]
        L.vhd_error.restype = ctypes.c_char_p
        L.vhd_error.argtypes = [_vp]
        L.vhd_log.restype = ctypes.c_char_p
        L.vhd_log.argtypes = [_vp]
        L.vhd_n_steps.argtypes = [_vp]
        L.vhd_step.argtypes = [_vp, ctypes.c_int, _dp]
        L.vhd_solution_size.restype = ctypes.c_longlong
        L.vhd_solution_size.argtypes = [_vp]
        L.vhd_solution.argtypes = [_vp, _dp]
        L.vhd_free.argtypes = [_vp]
        _drv = L
    return _drv


def run_prm(prm_text):
    """FemGL<3>(degree, prm).run() of the C++ driver mirror on cuda:0, configured by .prm text.
    Returns dict(history=[{...}], log=str, solution=ndarray); raises RuntimeError with the driver's exception text."""
    L = driver_lib()
    h = L.vhd_run(prm_text.encode())
    try:
        err = L.vhd_error(h).decode()
        log = L.vhd_log(h).decode()
        if err:
            raise RuntimeError(err + "\n--- log ---\n" + log[-2000:])
        keys = ["cycle", "iteration", "rhs_norm", "linear_its", "residual", "alpha", "trials", "energy", "t_assemble_ms",
                "t_solve_ms", "t_newton_ms", "t_setup_ms"]
        hist = []
        buf = np.zeros(12)
        for i in range(L.vhd_n_steps(h)):
            L.vhd_step(h, i, buf.ctypes.data_as(_dp))
            d = dict(zip(keys, buf.tolist()))
            for k in ("cycle", "iteration", "linear_its", "trials"):
                d[k] = int(d[k])
            hist.append(d)
        sol = np.zeros(L.vhd_solution_size(h))
        if sol.size:
            L.vhd_solution(h, sol.ctypes.data_as(_dp))
        return dict(history=hist, log=log, solution=sol)
    finally:
        L.vhd_free(h)
