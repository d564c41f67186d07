"""Q2 (BASELINE configs[2] shape) timing on one GPU: kernels of the general-scatter path at Q2 r3/r4."""
import os
import sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + "/tests"]
import verkko_hem_repo_b200 as vh  # noqa: E402
from helpers import b_phase_state, coef_vector  # noqa: E402

for refine in ([int(a) for a in sys.argv[1:]] or [3, 4]):
    m = vh.unit_cube(2, refine, half=20.0)
    T = m.tables(0)
    ctx = vh.Context(T)
    ctx.set_coef_vector(coef_vector())
    ctx.set_solution(b_phase_state(T, noise=0.0))
    bn = ctx.assemble()
    info = ctx.info()
    t_asm = ctx.time_kernel(1, reps=3, flush_l2=True)
    t_pw = ctx.time_kernel(5, reps=3, flush_l2=True)
    t_spmv = ctx.time_kernel(0, reps=10, flush_l2=True)
    its, res = ctx.solve(1e-1)
    nnzb = info["nnzb"]
    bytes_spmv = 8 * 324 * nnzb + 4 * nnzb + 16 * 18 * T.n_owned_nodes
    npk = info["n_packed_blocks"]
    moved = 8 * (180 * npk + 324 * (nnzb - npk)) + 4 * nnzb + 16 * 18 * T.n_owned_nodes
    t_rows = ctx.time_kernel(6, reps=3, flush_l2=True)
    mf = ""
    if npk:  # the matrix-free operator apply (vh_set_spmv_matrix_free): streams the H_q tables instead of the blocks
        ctx.set_spmv_matrix_free(True)
        t_mf = ctx.time_kernel(0, reps=10, flush_l2=True)
        ctx.set_spmv_matrix_free(False)
        mf_moved = 8 * 180 * 27 * T.n_cells + 2 * 8 * 18 * 27 * T.n_cells + 16 * 18 * T.n_owned_nodes
        mf = " | matrix-free apply %.3f ms = %.0f GB/s moved" % (t_mf, mf_moved / t_mf / 1e6)
    print("Q2 r%d: dofs %d nnzb %d slow_cells %d | assembly %.2f ms (pointwise %.2f, rows %.2f = %.0f GB/s stored) | spmv %.3f ms = %.0f GB/s "
          "algorithmic, %.0f GB/s moved%s | gmres its %d | device %.2f GB"
          % (refine, 18 * m.n_nodes, nnzb, info["n_slow_cells"], t_asm, t_pw, t_rows, 8 * 180 * npk / t_rows / 1e6, t_spmv,
             bytes_spmv / t_spmv / 1e6, moved / t_spmv / 1e6, mf, its, info["device_bytes"] / 1e9))
    ctx.close()
