"""CPU tests of the host side: Matep / confreader mirrors, mesh + constraint + partition tables, C-ABI surface."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import femgl_oracle as O
import verkko_hem_repo_b200 as vh
from helpers import b_phase_state, coef_vector

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REFERENCE_KEYS = {  # SURVEY.md App. D == /root/reference/confreader/src/declare.cc:115-291 (misspellings included)
    "physical parameters": ["pressure in bar", "t_reduced", "AdGR diffuse length", "gaussian random mean value", "gaussian random STD",
                            "trun on Strong Coupling Correction", "amplitude u parameter"],
    "control parameters": ["cube half side length", "half x length of retangle", "half y length of retangle",
                           "half z length of retangle", "B-phase inner plate radius ratio", "B-phase ball radius ratio",
                           "A-phase block range ratio", "Number of refinements", "Number of interations",
                           "Cycle 0 refinement threshold", "Cycle 0 linear solver tol", "Cycle 1 refinement threshold",
                           "Cycle 1 do global refinement", "Cycle 1 linear solver tol", "Cycle 2 refinement threshold",
                           "Cycle 2 do global refinement", "Cycle 2 linear solver tol", "Cycle 3 refinement threshold",
                           "Cycle 3 do global refinement", "Cycle 3 linear solver tol", "Cycle 4 do global refinement",
                           "Cycle 4 linear solver tol", "converge accuracy", "adaptive refinment ratio", "adaptive coarsen ratio",
                           "Number of initial global refinments", "Number of n-cycle in AdditionalData",
                           "maximum linear iteration number", "Using dampped Newton iteration",
                           "primary step length of dampped newton iteration"],
}


def test_matep_restatement_is_bit_exact(golden_dir):
    with open(os.path.join(golden_dir, "matep.json")) as f:
        grid = json.load(f)
    for row in grid:
        got = vh.matep(row["p"], row["t"], row["scc"])
        for k, v in got.items():
            assert v == row[k], (row["p"], row["t"], row["scc"], k)


def test_confreader_declares_every_reference_key_with_its_default():
    d = vh.parse_prm("")
    n = 0
    for sec, keys in REFERENCE_KEYS.items():
        for k in keys:
            assert sec + "/" + k in d, k
            n += 1
    assert n == 37
    assert d["physical parameters/AdGR diffuse length"] == "1.0e10"
    assert d["control parameters/primary step length of dampped newton iteration"] == "0.83"
    assert d["control parameters/Cycle 0 linear solver tol"] == "1.0e-1"
    # additive GPU-side keys have defaults, so a reference configuration.prm parses unchanged
    assert d["control parameters/GMRES restart length"] == "30"


@pytest.mark.skipif(not os.path.isdir("/root/reference/confreader"), reason="reference tree not present")
def test_confreader_keys_match_reference_source():
    src = "\n".join(l for l in open("/root/reference/confreader/src/declare.cc").read().splitlines()
                    if not l.lstrip().startswith("//"))
    ref = re.findall(r'declare_entry\("([^"]+)",\s*"([^"]*)"', src)
    assert len(ref) == 37
    d = vh.parse_prm("")
    flat = {k.split("/", 1)[1]: v for k, v in d.items()}
    for key, default in ref:
        assert flat[key] == default, key


def test_prm_parsing_and_errors():
    d = vh.parse_prm("# comment\nsubsection physical parameters\n  set pressure in bar = 25.0 # trailing\n  set t_reduced=0.5\nend\n"
                     "subsection control parameters\n set Number of initial global refinments = 3\nend\n")
    assert d["physical parameters/pressure in bar"] == "25.0" and d["physical parameters/t_reduced"] == "0.5"
    assert d["control parameters/Number of initial global refinments"] == "3"
    with pytest.raises(RuntimeError):
        vh.parse_prm("subsection physical parameters\n set no such key = 1\nend\n")
    with pytest.raises(RuntimeError):
        vh.parse_prm("subsection physical parameters\n set t_reduced = 0.5\n")


@pytest.mark.parametrize("degree,refine", [(1, 3), (2, 2)])
def test_uniform_cube_tables(degree, refine):
    m = vh.unit_cube(degree, refine, half=0.5)
    T = m.tables(0)
    nc1 = 2 ** refine
    n1 = degree * nc1 + 1
    assert m.n_cells == nc1 ** 3 and m.n_nodes == n1 ** 3 and T.n_ghost_nodes == 0
    # every cell's nodes sit where the FE_Q support points of that box are
    _, _, _, xi = O.fe_tables(degree)
    want = T.cell_origin[:, None, :] + xi[None, :, :] * T.cell_h[:, None, :]
    got = T.node_xyz[T.cell_nodes]
    assert np.abs(want - got).max() < 1e-13
    # z faces carry boundary id 4 -> 6 masked components on the two wall planes (femgl.h:300-301)
    assert T.c_dof.size == 2 * n1 * n1 * 6 and T.c_master.size == 0
    assert set((T.c_dof % 18).tolist()) == {2, 5, 8, 11, 14, 17}
    wall = np.abs(np.abs(T.node_xyz[T.c_dof // 18, 2]) - 0.5) < 1e-13
    assert wall.all()
    assert T.wall_face_cell.size == 2 * nc1 * nc1 and set(T.wall_face_bid.tolist()) == {4} and set(T.wall_face_no.tolist()) == {4, 5}


def test_morton_order_and_partition_cover():
    m = vh.unit_cube(1, 3, n_ranks=4)
    owned = np.zeros(m.n_nodes, dtype=int)
    cells_owned = 0
    for r in range(4):
        T = m.tables(r)
        owned[T.node_global[:T.n_owned_nodes]] += 1
        cells_owned += int(T.cell_owned.sum())
        # ghosts are grouped by owner: receive lists are contiguous ascending runs
        assert (np.diff(T.recv_nodes) > 0).all() if T.recv_nodes.size else True
        # every visited cell touches an owned node
        assert (T.cell_nodes < T.n_owned_nodes).any(axis=1).all()
    assert (owned == 1).all() and cells_owned == m.n_cells
    # halo plans are mutually consistent: what r sends to p is what p expects from r, in the same order
    tabs = [m.tables(r) for r in range(4)]
    for r, T in enumerate(tabs):
        for k, p in enumerate(T.peer_rank):
            send = T.node_global[T.send_nodes[T.send_ptr[k]:T.send_ptr[k + 1]]]
            Tp = tabs[p]
            kk = list(Tp.peer_rank).index(r)
            recv = Tp.node_global[Tp.recv_nodes[Tp.recv_ptr[kk]:Tp.recv_ptr[kk + 1]]]
            assert np.array_equal(send, recv)


@pytest.mark.parametrize("degree", [1, 2])
def test_hanging_node_constraints_reproduce_polynomials(degree):
    m = vh.Mesh(degree, [-1, -1, -1], [1, 1, 1], n_global_refine=1 if degree == 2 else 2)
    c = m.cell_centers()
    m.refine((c[:, 0] < 0) & (c[:, 2] > 0))
    m.finalize(1)
    T = m.tables(0)
    assert m.n_hanging_nodes > 0
    cnt = np.diff(T.c_ptr)
    # hanging-node lines of components that are not Dirichlet-masked anywhere (c % 3 != 2 with z walls): no master is
    # ever dropped by closing the constraints, so the weights are the plain FE interpolation weights
    lines = np.nonzero((cnt > 0) & (T.c_dof % 3 != 2))[0]
    assert lines.size > 0
    x, y, z = T.node_xyz.T
    f = (1 + 2 * x - y + 0.5 * z) if degree == 1 else (1 + x * y - 2 * z * z + x * x + 0.3 * y)
    field = np.repeat(f[:, None], 18, axis=1).ravel()
    for l in lines:
        s = slice(T.c_ptr[l], T.c_ptr[l + 1])
        assert abs(T.c_weight[s].sum() - 1.0) < 1e-13                       # partition of unity
        assert (T.c_master[s] % 18 == T.c_dof[l] % 18).all()                # same component
        # a polynomial of the element degree, sampled at the nodes, already satisfies the constraint
        assert abs((T.c_weight[s] * field[T.c_master[s]]).sum() - field[T.c_dof[l]]) < 1e-12
    # masked components of hanging nodes: masters on the wall are Dirichlet DoFs and drop out when the object is closed
    masked = np.nonzero((T.c_dof % 3 == 2) & (np.abs(np.abs(T.node_xyz[T.c_dof // 18, 2]) - 1.0) < 1e-12))[0]
    assert (cnt[masked] == 0).all()


def test_solution_transfer_reproduces_trilinear_field():
    old = vh.Mesh(1, [-1, -1, -1], [1, 1, 1], n_global_refine=2).finalize(1)
    new = old.clone()
    c = new.cell_centers()
    new.refine(c[:, 1] > 0.2)
    new.finalize(1)
    f = lambda X: 0.5 + X[:, 0] - 2 * X[:, 1] * X[:, 2] + X[:, 0] * X[:, 1] * X[:, 2]  # noqa: E731
    ov = np.repeat(f(old.node_xyz())[:, None], 18, axis=1) * np.arange(1, 19)[None, :]
    nv = new.interpolate_from(old, ov.ravel()).reshape(-1, 18)
    want = f(new.node_xyz())[:, None] * np.arange(1, 19)[None, :]
    assert np.abs(nv - want).max() < 1e-13


def test_cabi_library_exports_every_declared_symbol():
    """include/vh_femgl.h is the drop-in boundary: the built library must export exactly those entry points."""
    hdr = open(os.path.join(ROOT, "include", "vh_femgl.h")).read()
    names = sorted(set(re.findall(r"\b(vh_[a-z0-9_]+)\s*\(", hdr)))
    assert "vh_assemble" in names and "vh_solve" in names and "vh_residual" in names and len(names) >= 25
    lib = os.path.join(ROOT, "verkko-hem-repo_b200", "lib", "libvhfemgl.so")
    if not os.path.exists(lib):
        vh.build(cuda=True, host=False)
    L = ctypes.CDLL(lib)  # loads without a GPU (no compute call is made)
    for n in names:
        assert hasattr(L, n), "missing export " + n


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    T = vh.unit_cube(1, 1).tables(0)
    with pytest.raises(vh.VhError) as e:
        vh.Context(T)
    assert "no CPU fallback" in str(e.value) or e.value.code in (-2, -5)


def test_oracle_partition_independence():
    """Rows assembled per rank from its owned + ghost-layer cells equal the rows of the 1-rank assembly (the reason the
    GPU path needs no compress(add) exchange)."""
    coef = coef_vector(bt=2.0)
    m1 = vh.unit_cube(1, 2, half=2.0)
    T1 = m1.tables(0)
    mP = vh.unit_cube(1, 2, half=2.0, n_ranks=3)
    # global numbering differs between the two partitions: compare through coordinates
    key1 = {tuple(np.round(p, 9)): i for i, p in enumerate(T1.node_xyz)}
    x1 = b_phase_state(T1, seed=5)
    A1, r1 = O.assemble_global(T1, x1, coef, True)
    A1 = A1.tocsr()
    for r in range(3):
        T = mP.tables(r)
        perm = np.array([key1[tuple(np.round(p, 9))] for p in T.node_xyz])
        x = x1.reshape(-1, 18)[perm].ravel()
        A, rhs = O.assemble_global(T, x, coef, True)
        dof_perm = (18 * perm[:, None] + np.arange(18)[None, :]).ravel()
        want = A1[dof_perm[:18 * T.n_owned_nodes]][:, dof_perm]
        assert abs(A - want).max() <= 1e-13 * abs(A1).max()
        assert np.abs(rhs - r1[dof_perm[:18 * T.n_owned_nodes]]).max() <= 1e-13 * np.abs(r1).max()


def _native_pointwise_lib():
    """Host build of the product's __host__ __device__ pointwise math (csrc/vh_pointwise.cuh) — test-only."""
    import subprocess
    src = os.path.join(ROOT, "tests", "native", "pointwise_host.cc")
    out = os.path.join(ROOT, "tests", "native", "_build", "libvhpw.so")
    hdr = os.path.join(ROOT, "verkko-hem-repo_b200", "csrc", "vh_pointwise.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O2", "-shared", "-fPIC", "-o", out, src])
    return ctypes.CDLL(out)


@pytest.mark.parametrize("entry", ["vht_pointwise", "vht_points_q1"])
def test_device_pointwise_math_matches_oracle_on_host(entry):
    """The closed forms the kernels evaluate (column form and the register-resident per-entry form of k_points_q1)
    agree with the oracle's restatement of cell_mat_lhs_*/cell_vec_rhs_* to rounding for random complex A."""
    L = _native_pointwise_lib()
    P = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(7)
    coef = coef_vector()
    for trial in range(20):
        A = rng.standard_normal(18) * (1.0 if trial else 0.0) + (0.3 if trial == 0 else 0.0)
        g = np.zeros(18); H = np.zeros((18, 18)); f = np.zeros(1)
        getattr(L, entry)(A.ctypes.data_as(P), coef.ctypes.data_as(P), g.ctypes.data_as(P), H.ctypes.data_as(P), f.ctypes.data_as(P))
        g0, H0, f0 = O.pointwise(A, coef)
        assert np.abs(g - g0).max() <= 1e-13 * max(1.0, np.abs(g0).max())
        assert np.abs(H - H0).max() <= 1e-13 * max(1.0, np.abs(H0).max())
        assert abs(f[0] - f0) <= 1e-13 * max(1.0, abs(f0))
        assert np.abs(H - H.T).max() <= 1e-13 * max(1.0, np.abs(H0).max())


def test_hq8_layout_is_a_bijection():
    L = _native_pointwise_lib()
    seen = {L.vht_hq8_index(q, e) for q in range(8) for e in range(180)}
    assert seen == set(range(1440))
    for p in range(90):  # the eight quadrature points of one pair share one 128-byte line
        assert {L.vht_hq8_index(q, 2 * p) // 16 for q in range(8)} == {p}


def test_packed_symmetric_matvec_of_the_matrix_free_apply():
    """vh_sym_matvec (the H_q z product of k_apply_cells) on both storage layouts of the packed H_q tables equals the dense
    symmetric product; the zero dummies of the odd rows never contribute."""
    L = _native_pointwise_lib()
    P = ctypes.POINTER(ctypes.c_double)
    L.vht_sym_matvec.argtypes = [P, ctypes.c_int, P, P]
    rng = np.random.default_rng(5)
    H = rng.standard_normal((8, 18, 18))
    H = H + H.transpose(0, 2, 1)
    cell = np.full(1440, np.nan)
    plain = np.full((8, 180), 7.5)  # dummies deliberately non-zero in the plain layout: they must be ignored
    for q in range(8):
        for c in range(18):
            for d in range(c, 18):
                e = L.vht_sym_index(c, d)
                cell[L.vht_hq8_index(q, e)] = H[q, c, d]
                plain[q, e] = H[q, c, d]
        for c in range(1, 18, 2):
            cell[L.vht_hq8_index(q, L.vht_sym_index(c, c) - 1)] = 7.5
    assert not np.isnan(cell).any()
    for q in range(8):
        z = rng.standard_normal(18)
        t = np.zeros(18)
        L.vht_sym_matvec(cell.ctypes.data_as(P), q, z.ctypes.data_as(P), t.ctypes.data_as(P))
        assert np.abs(t - H[q] @ z).max() <= 1e-13 * np.abs(H[q] @ z).max()
        row = np.ascontiguousarray(plain[q])
        L.vht_sym_matvec(row.ctypes.data_as(P), -1, z.ctypes.data_as(P), t.ctypes.data_as(P))
        assert np.abs(t - H[q] @ z).max() <= 1e-13 * np.abs(H[q] @ z).max()


GRID_VARIANTS = {  # stem of /root/reference/femgl/src/makegrid_<stem>.cc -> (n DoFs at 2 global refinements, has periodic pairs)
    "cube-z-normal_AdGR": (18 * 125, False), "cube-xyz-Homo-Neumann": (18 * 125, False), "cube-xyz-periodic": (18 * 125, True),
    "cube-z-normal_AdGR-xy-periodic": (18 * 125, True), "retangle-z-AdGR-xy-HomoNeumann": (18 * 125, False),
    "retangle-xyz-homogenous-Neumann": (18 * 125, False), "retangle-xyz-periodic": (18 * 125, True),
    "retangle-z-AdGR_x-periodic-y-HomoNeumann": (18 * 125, True), "retangle-z-AdGR_xy-periodic": (18 * 125, True),
    "xz-normal_AdGR": (18 * 125, False), "cube": (18 * 125, False), "retangle": (18 * 125, False),
    "retangle-xy-periodic": (18 * 125, True),
}


@pytest.mark.parametrize("name", sorted(GRID_VARIANTS))
def test_driver_mirror_selects_every_box_grid_variant(name):
    """FemGL::make_grid() + setup_system() of the C++ driver mirror accept the stem of every box makegrid_*.cc variant that
    femgl/CMakeLists.txt:42-53 lists: the run gets as far as printing the DoF count and then needs the GPU (vh_create) —
    on this CPU-only container that is a loud error, never a fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by tests/test_gpu_driver.py and tests/test_gpu_periodic.py")
    prm = """
subsection control parameters
  set geometry = %s
  set Number of initial global refinments = 2
  set Number of refinements = 0
  set Number of interations = 0
end
""" % name
    with pytest.raises(RuntimeError) as e:
        vh.run_prm(prm)
    msg = str(e.value)
    assert "unknown geometry" not in msg and "vh_create" in msg, msg
    assert "Number of degrees of freedom: %d" % GRID_VARIANTS[name][0] in msg, msg


def test_driver_mirror_rejects_unknown_names():
    with pytest.raises(RuntimeError) as e:
        vh.run_prm("subsection control parameters\n  set geometry = cube_cylider_hole\nend\n")
    assert "unknown geometry" in str(e.value) and "retangle-z-AdGR_xy-periodic" in str(e.value)


def _all_kinds_of_tables():
    yield vh.unit_cube(1, 2, half=2.0).tables(0)
    yield vh.unit_cube(2, 1, half=2.0).tables(0)
    for r in range(3):
        yield vh.unit_cube(1, 2, half=2.0, n_ranks=3).tables(r)
        yield vh.periodic_slab(1, 2, n_ranks=3).tables(r)
    m = vh.Mesh(1, [-2, -2, -2], [2, 2, 2], n_global_refine=2)
    c = m.cell_centers()
    m.refine((np.abs(c[:, 2]) < 1.1) & (c[:, 0] < 0.1))
    m.finalize(2)
    yield m.tables(0)
    yield m.tables(1)


def test_descriptor_validation_accepts_every_table_the_host_builds():
    """vh_validate_mesh_desc (host-only part of vh_create) on uniform / periodic / hanging-node tables, 1-3 ranks."""
    n = 0
    for T in _all_kinds_of_tables():
        assert vh.validate_tables(T) == ""
        n += 1
    assert n == 10


@pytest.mark.parametrize("field,index,value,expect", [
    ("cell_nodes", (0, 0), -1, "cell_nodes entry out of range"),
    ("cell_nodes", (3, 1), 10 ** 6, "cell_nodes entry out of range"),
    ("cell_nodes", (2, 5), "dup", "same node twice"),
    ("cell_h", (1, 2), 0.0, "cell_h must be positive"),
    ("cell_h", (1, 0), float("nan"), "cell_h must be positive"),
    ("wall_face_bid", 0, 7, "wall face table entry out of range"),
    ("wall_face_no", 1, 6, "wall face table entry out of range"),
    ("wall_face_cell", 0, 10 ** 6, "wall face table entry out of range"),
    ("c_dof", 1, "swap", "strictly ascending"),
    ("c_dof", 0, -5, "constrained DoF out of range"),
    ("c_master", 0, 10 ** 8, "master out of range"),
    ("c_master", 0, "constrained", "not closed"),
    ("c_weight", 0, float("inf"), "weight is not finite"),
    ("send_nodes", 0, "ghost", "send_nodes must be owned"),
    ("recv_nodes", 0, 0, "recv_nodes must be ghost"),
    ("recv_nodes", 1, "dup", "received twice"),
    ("peer_rank", 0, -1, "negative peer rank"),
])
def test_descriptor_validation_names_the_defect(field, index, value, expect):
    """Corrupt one entry of an otherwise valid two-rank table with hanging nodes and Dirichlet walls: the host-only
    validator (and therefore vh_create, before it touches CUDA) must name the defect."""
    m = vh.Mesh(1, [-2, -2, -2], [2, 2, 2], n_global_refine=2)
    c = m.cell_centers()
    m.refine((np.abs(c[:, 2]) < 1.1) & (c[:, 0] < 0.1))
    m.finalize(2)
    T = m.tables(0)
    assert vh.validate_tables(T) == ""
    arr = getattr(T, field)
    arr.flags.writeable = True
    old = arr[index]
    if value == "dup":
        value = arr[(index[0], index[1] - 1)] if isinstance(index, tuple) else arr[index - 1]
    elif value == "swap":
        value = arr[index - 1]
    elif value == "constrained":
        value = T.c_dof[-1]
    elif value == "ghost":
        value = T.n_owned_nodes
    arr[index] = value
    try:
        msg = vh.validate_tables(T)
    finally:
        arr[index] = old
    assert expect in msg, msg
    assert vh.validate_tables(T) == ""


def test_abi_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 with -pedantic and a C program links against libvhfemgl.so."""
    import subprocess
    libdir = os.path.join(ROOT, "verkko-hem-repo_b200", "lib")
    exe = str(tmp_path / "abi_from_c")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-o", exe, os.path.join(ROOT, "tests", "native", "abi_from_c.c"), "-L", libdir, "-lvhfemgl",
                           "-Wl,-rpath," + libdir])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr)


def test_example_prm_files_parse_with_the_confreader_mirror():
    """examples/*.prm (the reference ships no femgl .prm): every key is one the confreader mirror declares."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "examples", "*.prm")))
    assert len(files) >= 5
    for f in files:
        d = vh.parse_prm(open(f).read())
        assert d["control parameters/geometry"] in GRID_VARIANTS, f
        assert d["control parameters/initial condition"] in ("B-phase", "A-phase", "BnA"), f


def test_table_free_hessian_apply_equals_the_hessian_times_z():
    """vh_hessian_apply (directional derivative of the bulk residual density, eight 3x3 complex products) = H z with the
    oracle's H (restating cell_mat_lhs_alpha/beta1..5) for random states and directions, and is linear in z."""
    L = _native_pointwise_lib()
    P = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(23)
    for trial in range(25):
        coef = coef_vector()
        if trial % 5 == 4:  # generic coefficients: every beta_k different, nothing cancels
            coef = coef.copy()
            coef[3:9] = rng.uniform(-1, 1, 6)
        A = rng.standard_normal(18)
        z = rng.standard_normal(18)
        out = np.zeros(18)
        L.vht_hessian_apply(A.ctypes.data_as(P), coef.ctypes.data_as(P), z.ctypes.data_as(P), out.ctypes.data_as(P))
        _, H0, _ = O.pointwise(A, coef)
        want = H0 @ z
        assert np.abs(out - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), trial
        out2 = np.zeros(18)
        z2 = -2.5 * z
        L.vht_hessian_apply(A.ctypes.data_as(P), coef.ctypes.data_as(P), z2.ctypes.data_as(P), out2.ctypes.data_as(P))
        assert np.abs(out2 + 2.5 * out).max() <= 1e-12 * max(1.0, np.abs(out).max())
