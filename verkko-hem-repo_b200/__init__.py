"""verkko-hem-repo_b200 — B200-native femgl Newton hot path (VerHem drop-in) behind a C ABI.

Python is only the test / benchmark harness here: ``ctypes`` bindings over

* ``lib/libvhfemgl.so``  — hand-written sm_100a CUDA (csrc/), C ABI ``include/vh_femgl.h``;
* ``lib/libvhhost.so``   — host-side tables and the ``FemGL`` driver mirror (host/), the stand-in for deal.II.

There is no CPU fallback: every compute entry point lives in ``libvhfemgl.so`` and raises if it is missing.
"""
import ctypes
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__)) if "__file__" in globals() else None
if _PKG is None or not os.path.isdir(os.path.join(_PKG, "csrc")):
    _PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "verkko-hem-repo_b200")
ROOT = os.path.dirname(_PKG)
LIBDIR = os.path.join(_PKG, "lib")

_dp = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


class VhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vh error %d: %s" % (code, msg))
        self.code = code


def build(cuda=True, host=True, verbose=False):
    """Compile the in-tree libraries (nvcc cross-compiles sm_100a without a GPU)."""
    targets = (["host"] if host else []) + (["cuda", "driver"] if cuda else [])
    cmd = ["make", "-C", _PKG, "-j8"] + targets
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)


# ------------------------------------------------------------------------------------------
# host tables (stand-in for deal.II): libvhhost.so
# ------------------------------------------------------------------------------------------
_host = None


def host_lib():
    global _host
    if _host is None:
        p = os.path.join(LIBDIR, "libvhhost.so")
        if not os.path.exists(p):
            build(cuda=False, host=True)
        L = ctypes.CDLL(p)
        L.vhh_last_error.restype = ctypes.c_char_p
        L.vhh_mesh_create.restype = ctypes.c_void_p
        L.vhh_mesh_create.argtypes = [ctypes.c_int, _dp, _dp, _i32p, _i32p, ctypes.c_int]
        L.vhh_mesh_free.argtypes = [ctypes.c_void_p]
        L.vhh_mesh_n_cells.restype = ctypes.c_int64
        L.vhh_mesh_n_cells.argtypes = [ctypes.c_void_p]
        L.vhh_mesh_cell_centers.argtypes = [ctypes.c_void_p, _dp]
        L.vhh_mesh_refine.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int64]
        L.vhh_mesh_finalize.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.vhh_mesh_global_sizes.argtypes = [ctypes.c_void_p, _i64p]
        L.vhh_mesh_rank_node_begin.argtypes = [ctypes.c_void_p, _i64p]
        L.vhh_tables_create.restype = ctypes.c_void_p
        L.vhh_tables_create.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.vhh_tables_free.argtypes = [ctypes.c_void_p]
        L.vhh_tables_desc.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.vhh_tables_array.restype = ctypes.c_void_p
        L.vhh_tables_array.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _i64p, ctypes.POINTER(ctypes.c_int)]
        L.vhh_tables_sizes.argtypes = [ctypes.c_void_p, _i64p]
        L.vhh_matep.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp]
        L.vhh_prm_dump.restype = ctypes.c_char_p
        L.vhh_prm_dump.argtypes = [ctypes.c_char_p]
        L.vhh_mesh_interpolate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, _dp, _dp]
        L.vhh_mesh_transfer_table.restype = ctypes.c_int64
        L.vhh_mesh_transfer_table.argtypes = [ctypes.c_void_p, ctypes.c_void_p, _i32p, _i32p, _dp]
        L.vhh_mesh_kelly.argtypes = [ctypes.c_void_p, _dp, _dp]
        L.vhh_mesh_clone.restype = ctypes.c_void_p
        L.vhh_mesh_clone.argtypes = [ctypes.c_void_p]
        L.vhh_mesh_node_xyz.argtypes = [ctypes.c_void_p, _dp]
        L.vhh_write_vtu.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, _dp, _dp]
        _host = L
    return _host


_FIELD_DTYPES = {
    "node_global": np.int64, "node_xyz": np.float64, "cell_nodes": np.int32, "cell_global": np.int64,
    "cell_origin": np.float64, "cell_h": np.float64, "cell_owned": np.uint8, "wall_face_cell": np.int32,
    "wall_face_no": np.int8, "wall_face_bid": np.int8, "c_dof": np.int32, "c_ptr": np.int32, "c_master": np.int32,
    "c_weight": np.float64, "peer_rank": np.int32, "send_ptr": np.int32, "send_nodes": np.int32, "recv_ptr": np.int32,
    "recv_nodes": np.int32,
}


class RankTables:
    """Flat per-rank tables (numpy views onto the C++ arrays) + the matching ``vh_mesh_desc`` blob."""

    def __init__(self, mesh, rank):
        L = host_lib()
        self._mesh = mesh  # keep alive
        self._h = L.vhh_tables_create(mesh._h, rank)
        if not self._h:
            raise RuntimeError(L.vhh_last_error().decode())
        sz = (ctypes.c_int64 * 4)()
        L.vhh_tables_sizes(self._h, sz)
        self.degree, self.n_owned_nodes, self.n_ghost_nodes, self.n_cells = (int(v) for v in sz)
        self.n_local_nodes = self.n_owned_nodes + self.n_ghost_nodes
        self.rank = rank
        n = 8 if self.degree == 1 else 27
        for name, dt in _FIELD_DTYPES.items():
            cnt = ctypes.c_int64()
            es = ctypes.c_int()
            ptr = L.vhh_tables_array(self._h, name.encode(), ctypes.byref(cnt), ctypes.byref(es))
            assert cnt.value >= 0 and es.value == np.dtype(dt).itemsize, name
            if cnt.value == 0:
                arr = np.zeros(0, dtype=dt)
            else:
                buf = (ctypes.c_char * (cnt.value * es.value)).from_address(ptr)
                arr = np.frombuffer(buf, dtype=dt)
            setattr(self, name, arr)
        self.cell_nodes = self.cell_nodes.reshape(-1, n)
        self.node_xyz = self.node_xyz.reshape(-1, 3)
        self.cell_origin = self.cell_origin.reshape(-1, 3)
        self.cell_h = self.cell_h.reshape(-1, 3)
        self._desc = ctypes.create_string_buffer(L.vhh_sizeof_mesh_desc())
        L.vhh_tables_desc(self._h, ctypes.cast(self._desc, ctypes.c_void_p))

    def desc_ptr(self):
        return ctypes.cast(self._desc, ctypes.c_void_p)

    def write_vtu(self, directory, counter, solution_local=None, update_local=None, n_ranks=1):
        """DataOut stand-in of the host mirror (host/vtu.cc; io.cc:106-170): <directory>/solution_<counter>.<rank>.vtu with the
        owned cells of this rank, + the .pvtu record on rank 0.  The fields hold 18 values per LOCAL node (owned, then ghosts)."""
        def arr(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            assert a.size == 18 * self.n_local_nodes
            return a
        s, u = arr(solution_local), arr(update_local)
        d = directory if directory.endswith("/") else directory + "/"
        if host_lib().vhh_write_vtu(self._h, self.rank, n_ranks, d.encode(), int(counter), s.ctypes.data_as(_dp) if s is not None else None,
                                    u.ctypes.data_as(_dp) if u is not None else None) != 0:
            raise RuntimeError(host_lib().vhh_last_error().decode())
        return "%ssolution_%02d.%d.vtu" % (d, counter, self.rank)

    def face_csr(self):
        """Wall faces regrouped per cell: (face_ptr[n_cells+1], face_no[], face_bid[]) as int32."""
        order = np.argsort(self.wall_face_cell, kind="stable")
        cnt = np.bincount(self.wall_face_cell, minlength=self.n_cells)
        ptr = np.zeros(self.n_cells + 1, dtype=np.int32)
        np.cumsum(cnt, out=ptr[1:])
        return ptr, self.wall_face_no[order].astype(np.int32), self.wall_face_bid[order].astype(np.int32)

    def __del__(self):
        try:
            if self._h:
                host_lib().vhh_tables_free(self._h)
                self._h = None
        except Exception:
            pass


class Mesh:
    """Box mesh of hexahedra (hyper_cube / hyper_rectangle + refine_global, optional local refinement)."""

    def __init__(self, degree, lo, hi, base=(1, 1, 1), face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=0):
        L = host_lib()
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        base = np.ascontiguousarray(base, dtype=np.int32)
        bid = np.ascontiguousarray(face_bid, dtype=np.int32)
        self.degree = degree
        self.lo, self.hi, self.base, self.n_global_refine, self.locally_refined = lo.copy(), hi.copy(), base.copy(), n_global_refine, False
        self._h = L.vhh_mesh_create(degree, lo.ctypes.data_as(_dp), hi.ctypes.data_as(_dp), base.ctypes.data_as(_i32p),
                                    bid.ctypes.data_as(_i32p), n_global_refine)
        if not self._h:
            raise RuntimeError(L.vhh_last_error().decode())
        self.n_ranks = 0

    @property
    def n_cells(self):
        return int(host_lib().vhh_mesh_n_cells(self._h))

    def cell_centers(self):
        out = np.zeros((self.n_cells, 3))
        host_lib().vhh_mesh_cell_centers(self._h, out.ctypes.data_as(_dp))
        return out

    def refine(self, flags):
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        if host_lib().vhh_mesh_refine(self._h, flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), flags.size) != 0:
            raise RuntimeError(host_lib().vhh_last_error().decode())
        self.locally_refined = True

    def finalize(self, n_ranks=1):
        if host_lib().vhh_mesh_finalize(self._h, n_ranks) != 0:
            raise RuntimeError(host_lib().vhh_last_error().decode())
        self.n_ranks = n_ranks
        sz = (ctypes.c_int64 * 5)()
        host_lib().vhh_mesh_global_sizes(self._h, sz)
        self.n_nodes, _, self.n_constraint_lines, self.n_hanging_nodes, self.n_periodic_nodes = (int(v) for v in sz)
        rb = (ctypes.c_int64 * (n_ranks + 1))()
        host_lib().vhh_mesh_rank_node_begin(self._h, rb)
        self.rank_node_begin = np.array(list(rb), dtype=np.int64)
        return self

    def tables(self, rank=0):
        return RankTables(self, rank)

    def clone(self):
        m = Mesh.__new__(Mesh)
        m.degree = self.degree
        m.lo, m.hi, m.base, m.n_global_refine, m.locally_refined = self.lo, self.hi, self.base, self.n_global_refine, self.locally_refined
        m._h = host_lib().vhh_mesh_clone(self._h)
        m.n_ranks = 0
        return m

    def node_xyz(self):
        out = np.zeros((self.n_nodes, 3))
        host_lib().vhh_mesh_node_xyz(self._h, out.ctypes.data_as(_dp))
        return out

    def interpolate_from(self, old_mesh, old_values):
        """SolutionTransfer stand-in: FE-interpolate old_values (old mesh, global node order) onto this mesh."""
        ov = np.ascontiguousarray(old_values, dtype=np.float64)
        nv = np.zeros(18 * self.n_nodes)
        if host_lib().vhh_mesh_interpolate(self._h, old_mesh._h, ov.ctypes.data_as(_dp), nv.ctypes.data_as(_dp)) != 0:
            raise RuntimeError(host_lib().vhh_last_error().decode())
        return nv

    def transfer_table(self, old_mesh):
        """The interpolation of interpolate_from as a CSR table (ptr, old node ids, weights) for vh_transfer_solution."""
        nn = 8 if self.degree == 1 else 27
        ptr = np.zeros(self.n_nodes + 1, dtype=np.int32)
        src = np.zeros(self.n_nodes * nn, dtype=np.int32)
        w = np.zeros(self.n_nodes * nn)
        nnz = host_lib().vhh_mesh_transfer_table(self._h, old_mesh._h, ptr.ctypes.data_as(_i32p), src.ctypes.data_as(_i32p),
                                                 w.ctypes.data_as(_dp))
        if nnz < 0:
            raise RuntimeError(host_lib().vhh_last_error().decode())
        return ptr, src[:nnz].copy(), w[:nnz].copy()

    def kelly_indicator(self, values):
        """Kelly-type face-jump indicator per cell (refine.cc:144-148 stand-in); values in global node order."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        eta = np.zeros(self.n_cells)
        if host_lib().vhh_mesh_kelly(self._h, v.ctypes.data_as(_dp), eta.ctypes.data_as(_dp)) != 0:
            raise RuntimeError(host_lib().vhh_last_error().decode())
        return eta

    def __del__(self):
        try:
            if self._h:
                host_lib().vhh_mesh_free(self._h)
                self._h = None
        except Exception:
            pass


def matep(p, t, scc):
    """Material coefficients from the host-side Matep restatement (host/matep.cc)."""
    out = np.zeros(12)
    host_lib().vhh_matep(float(p), float(t), int(bool(scc)), out.ctypes.data_as(_dp))
    keys = ["alpha", "beta1", "beta2", "beta3", "beta4", "beta5", "gapA", "gapB", "fA", "fB", "Tcp_mK", "tAB_RWS"]
    return dict(zip(keys, out.tolist()))


def mg_prolongation(mesh_f, Tf, mesh_c, Tc):
    """Prolongation table between two consecutive levels of a globally refined Q1 box mesh for ``vh_mg_attach``: row i =
    LOCAL node i of the fine level (owned and ghost), entries = LOCAL nodes of the coarse level with the trilinear weights
    (1, 1/2, 1/4, 1/8); parents that are not local on this rank's coarse level are dropped (they can only belong to ghost
    rows).  Stand-in for what deal.II's MGTransfer holds.  Returns (ptr, coarse_node, weight)."""
    if mesh_f.degree != 1 or mesh_c.degree != 1 or mesh_f.locally_refined or mesh_c.locally_refined:
        raise ValueError("mg_prolongation: globally refined Q1 meshes only")
    if mesh_f.n_global_refine != mesh_c.n_global_refine + 1 or Tf.c_master.size or Tc.c_master.size:
        raise ValueError("mg_prolongation: consecutive refinement levels of one box without periodic / hanging constraints")
    nf_side = mesh_f.base.astype(np.int64) << mesh_f.n_global_refine           # fine cells per direction
    hf = (mesh_f.hi - mesh_f.lo) / nf_side
    If = np.rint((Tf.node_xyz - mesh_f.lo[None, :]) / hf[None, :]).astype(np.int64)       # fine lattice coordinates
    Ic = np.rint((Tc.node_xyz - mesh_f.lo[None, :]) / (2.0 * hf[None, :])).astype(np.int64)
    nc1 = (nf_side // 2) + 1
    key_c = (Ic[:, 2] * nc1[1] + Ic[:, 1]) * nc1[0] + Ic[:, 0]
    order = np.argsort(key_c)
    skeys = key_c[order]
    n = Tf.n_local_nodes
    rows, cols, vals = [], [], []
    for corner in range(8):
        w = np.ones(n)
        J = np.zeros((n, 3), dtype=np.int64)
        ok = np.ones(n, dtype=bool)
        for d in range(3):
            odd = (If[:, d] & 1) == 1
            b = (corner >> d) & 1
            J[:, d] = (If[:, d] >> 1) + np.where(odd, b, 0)
            w *= np.where(odd, 0.5, 1.0)
            ok &= odd | (b == 0)          # an even coordinate has one parent in this direction
        key = (J[:, 2] * nc1[1] + J[:, 1]) * nc1[0] + J[:, 0]
        pos = np.minimum(np.searchsorted(skeys, key), skeys.size - 1)
        hit = ok & (skeys[pos] == key)
        rows.append(np.nonzero(hit)[0])
        cols.append(order[pos[hit]])
        vals.append(w[hit])
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    o = np.lexsort((cols, rows))
    rows, cols, vals = rows[o], cols[o], vals[o]
    ptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=n), out=ptr[1:])
    return ptr, cols.astype(np.int32), vals


def periodic_slab(degree, refine, half=(0.5, 0.5, 0.5), base=(1, 1, 1), n_ranks=1):
    """The reference's ACTIVE grid (femgl/CMakeLists.txt:51): hyper_rectangle(-half, half), x faces ids 5/6 and y faces 7/8
    periodic, z faces AdGR walls (id 4), refine_global(refine) (makegrid_retangle-z-AdGR_xy-periodic.cc:167-236)."""
    half = np.asarray(half, dtype=np.float64)
    return Mesh(degree, -half, half, base, (5, 6, 7, 8, 4, 4), refine).finalize(n_ranks)


def parse_prm(text=""):
    """Parse .prm text with the confreader mirror; returns {"subsection/key": value-string}."""
    r = host_lib().vhh_prm_dump(text.encode())
    if r is None:
        raise RuntimeError(host_lib().vhh_last_error().decode())
    return dict(line.split("=", 1) for line in r.decode().splitlines() if line)


def unit_cube(degree, refine, half=0.5, face_bid=(1, 1, 1, 1, 4, 4), n_ranks=1):
    """The BASELINE configs' cube: hyper_cube(-half, half) + refine_global(refine); z faces are AdGR walls (id 4)
    as in makegrid_cube-z-normal_AdGR.cc:164-195."""
    return Mesh(degree, [-half] * 3, [half] * 3, (1, 1, 1), face_bid, refine).finalize(n_ranks)


from ._capi import Context, cuda_lib, have_cuda_lib, run_prm, validate_tables  # noqa: E402,F401
