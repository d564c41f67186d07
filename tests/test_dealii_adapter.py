"""The deal.II side of the drop-in boundary (integration/dealii/vh_dealii_adapter.h) compiled and RUN against
a functional mock of the deal.II calls it makes (tests/native/dealii_mock/: DoFHandler cell iterators, FESystem,
IndexSet, AffineConstraints, Utilities::MPI with ranks as threads), backed by the repository's own box mesh.  On every
rank the tables the adapter derives from those deal.II-shaped objects must equal the tables of the mini host — cells,
ghost sets, both constraint tables, wall faces and the halo plan.  deal.II itself is not installed in this image."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    src = [os.path.join(ROOT, "tests", "native", "dealii_adapter_host.cc"), os.path.join(ROOT, "verkko-hem-repo_b200", "host", "mesh.cc")]
    deps = src + [os.path.join(ROOT, "integration", "dealii", "vh_dealii_adapter.h"),
                  os.path.join(ROOT, "tests", "native", "dealii_mock", "dealii_mock.h"), os.path.join(ROOT, "include", "vh_femgl.h"),
                  os.path.join(ROOT, "verkko-hem-repo_b200", "host", "mesh.h")]
    out = os.path.join(ROOT, "tests", "native", "_build", "libvhadapter.so")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(p) for p in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-shared", "-fPIC", "-pthread",
                               "-I", os.path.join(ROOT, "tests", "native", "dealii_mock"), "-I", os.path.join(ROOT, "include"),
                               "-o", out] + src)
    L = ctypes.CDLL(out)
    ip = ctypes.POINTER(ctypes.c_int)
    L.vht_adapter_check.argtypes = [ctypes.c_int, ctypes.c_int, ip, ip, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    return L


def _check(degree, refine, base, bid, local_refine, n_ranks):
    L = _lib()
    b = (ctypes.c_int * 3)(*base)
    f = (ctypes.c_int * 6)(*bid)
    msg = ctypes.create_string_buffer(1024)
    rc = L.vht_adapter_check(degree, refine, b, f, int(local_refine), n_ranks, msg, 1024)
    return rc, msg.value.decode()


WALLS = (1, 1, 1, 1, 4, 4)          # makegrid_cube-z-normal_AdGR.cc
ACTIVE = (5, 6, 7, 8, 4, 4)         # makegrid_retangle-z-AdGR_xy-periodic.cc (the variant the reference compiles)
XZ_WALLS = (2, 2, 1, 1, 4, 4)       # makegrid_xz-normal_AdGR.cc
ALL_PERIODIC = (5, 6, 7, 8, 9, 10)  # makegrid_retangle-xyz-periodic.cc


@pytest.mark.parametrize("degree,refine", [(1, 2), (2, 1)])
@pytest.mark.parametrize("bid", [WALLS, ACTIVE, XZ_WALLS, ALL_PERIODIC])
@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4])
def test_adapter_tables_equal_mini_host_on_conforming_meshes(degree, refine, bid, n_ranks):
    rc, msg = _check(degree, refine, (2, 1, 1), bid, False, n_ranks)
    assert rc == 0, msg


@pytest.mark.parametrize("degree,refine", [(1, 2), (2, 1)])
def test_adapter_tables_with_hanging_nodes_single_rank(degree, refine):
    rc, msg = _check(degree, refine, (2, 1, 1), WALLS, True, 1)
    assert rc == 0, msg


@pytest.mark.parametrize("degree,refine", [(1, 2), (2, 1)])
@pytest.mark.parametrize("n_ranks", [2, 3, 4, 5, 8])
def test_adapter_ships_cells_from_beyond_the_ghost_layer(degree, refine, n_ranks):
    """With hanging nodes the master of a hanging node can be owned by a rank whose ghost layer does not contain the fine
    cell (the reference ships such contributions with compress(add), assemble.cc:369-370; here the cell's owner ships the
    cell record and the constraint lines of its DoFs once per mesh).  Every rank must end up with exactly the cells, ghost
    nodes, constraint lines and halo plan of the mini host, whose partition-independence is checked against the oracle in
    tests/test_periodic_host.py::test_hanging_node_partition_independence."""
    rc, msg = _check(degree, refine, (2, 1, 1), WALLS, True, n_ranks)
    assert rc == 0, msg


@pytest.mark.parametrize("degree,refine", [(1, 2), (2, 1)])
@pytest.mark.parametrize("n_ranks", [1, 2, 4])
def test_adapter_on_the_active_grid_with_differently_refined_periodic_faces(degree, refine, n_ranks):
    """Periodic pairs + local refinement reaching one periodic face only (hanging nodes across the seam)."""
    rc, msg = _check(degree, refine, (2, 1, 1), ACTIVE, True, n_ranks)
    assert rc == 0, msg


@pytest.mark.parametrize("degree,refine,seed", [(1, 1, 2), (1, 1, 3), (2, 1, 4)])
@pytest.mark.parametrize("bid", [WALLS, ACTIVE])
@pytest.mark.parametrize("n_ranks", [1, 3, 8])
def test_adapter_on_randomly_refined_multi_level_meshes(degree, refine, seed, bid, n_ranks):
    """Three rounds of random refinement (several levels, chains of hanging nodes, periodic seams refined differently)."""
    rc, msg = _check(degree, refine, (2, 1, 1), bid, seed, n_ranks)
    assert rc == 0, msg


def test_the_mock_ghost_layer_really_misses_cells():
    """Guard of the test above: switch the shipping off (VH_ADAPTER_TEST_NO_SHIPPING) and the tables must differ."""
    os.environ["VH_ADAPTER_TEST_NO_SHIPPING"] = "1"
    try:
        rc, msg = _check(1, 2, (2, 1, 1), WALLS, True, 2)
    finally:
        del os.environ["VH_ADAPTER_TEST_NO_SHIPPING"]
    assert rc == 1, (rc, msg)
