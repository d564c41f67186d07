"""Output path of the host mirror (host/vtu.cc; FemGL::output_results, io.cc:106-170): the .vtu piece of every rank and the
.pvtu record, checked on the CPU by reading the appended binary arrays back."""
import os
import re

import numpy as np

import verkko_hem_repo_b200 as vh

NAMES = [p + "_" + ij for p in ("du", "dv", "u", "v") for ij in ("11", "12", "13", "21", "22", "23", "31", "32", "33")]


def read_vtu(path):
    raw = open(path, "rb").read()
    head, tail = raw.split(b'<AppendedData encoding="raw">\n_', 1)
    head = head.decode()
    assert 'header_type="UInt64"' in head and 'type="UnstructuredGrid"' in head
    npts, ncells = (int(v) for v in re.search(r'NumberOfPoints="(\d+)" NumberOfCells="(\d+)"', head).groups())
    out = {"n_points": npts, "n_cells": ncells}
    dt = {"Float64": np.float64, "Int64": np.int64, "UInt8": np.uint8, "Float32": np.float32}
    for m in re.finditer(r'<DataArray type="(\w+)"(?: Name="([\w]+)")?(?: NumberOfComponents="3")? format="appended" offset="(\d+)"/>', head):
        typ, name, off = m.group(1), m.group(2) or "points", int(m.group(3))
        nbytes = int(np.frombuffer(tail[off:off + 8], dtype=np.uint64)[0])
        out[name] = np.frombuffer(tail[off + 8:off + 8 + nbytes], dtype=dt[typ])
    return out


def test_vtu_pieces_hold_mesh_and_fields(tmp_path):
    m = vh.Mesh(1, [-3, -2, -4], [3, 2, 4], n_global_refine=2)
    d = np.abs(m.cell_centers()[:, 2] - 0.4)
    m.refine(d <= np.sort(d)[int(0.3 * m.n_cells)])       # hanging nodes: cells of two sizes
    m.finalize(2)
    total_cells = 0
    for rank in range(2):
        T = m.tables(rank)
        xyz = T.node_xyz
        sol = (xyz[:, :1] + 2.0 * xyz[:, 1:2] - 0.5 * xyz[:, 2:3]) * (1.0 + np.arange(18))[None, :]     # linear in x: u_c = (c+1) * f
        upd = 0.01 * sol + 7.0
        path = T.write_vtu(str(tmp_path / "refine-cycle_0"), 3, sol.ravel(), upd.ravel(), n_ranks=2)
        assert path.endswith("solution_03.%d.vtu" % rank) and os.path.exists(path)
        v = read_vtu(path)
        assert v["n_points"] == T.n_local_nodes and v["n_cells"] == int(T.cell_owned.sum())
        total_cells += v["n_cells"]
        assert np.array_equal(v["points"].reshape(-1, 3), xyz)
        conn = v["connectivity"].reshape(-1, 8)
        owned = np.nonzero(T.cell_owned)[0]
        assert np.array_equal(conn, T.cell_nodes[owned][:, [0, 1, 3, 2, 4, 5, 7, 6]])
        assert np.array_equal(v["offsets"], 8 * (1 + np.arange(v["n_cells"]))) and np.all(v["types"] == 12)
        # VTK hexahedron ordering: vertices 0-3 run counter-clockwise in the bottom face, 4-7 above them
        p = xyz[conn]
        assert np.all(p[:, 1, 0] > p[:, 0, 0]) and np.all(p[:, 2, 1] > p[:, 1, 1]) and np.all(p[:, 3, 0] < p[:, 2, 0])
        assert np.allclose(p[:, 4:, :2], p[:, :4, :2]) and np.all(p[:, 4:, 2] > p[:, :4, 2])
        for c in range(18):
            assert np.array_equal(v[NAMES[18 + c]], sol[:, c]) and np.array_equal(v[NAMES[c]], upd[:, c])
        assert np.all(v["subdomain"] == rank)
    assert total_cells == m.n_cells
    pvtu = open(tmp_path / "refine-cycle_0" / "solution_03.pvtu").read()
    assert pvtu.count("<Piece Source=") == 2 and "solution_03.1.vtu" in pvtu and 'Name="dv_33"' in pvtu


def test_vtu_null_fields_and_q2_vertices(tmp_path):
    T = vh.unit_cube(2, 1, half=1.0).tables(0)
    v = read_vtu(T.write_vtu(str(tmp_path), 0))
    assert v["n_cells"] == 8 and np.all(v["u_11"] == 0.0) and np.all(v["du_11"] == 0.0)
    corners = T.node_xyz[v["connectivity"].reshape(-1, 8)]
    assert np.allclose(np.ptp(corners, axis=1), 1.0)       # the 8 VERTICES of every Q2 cell (DataOut with one subdivision)
