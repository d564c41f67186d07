"""Q2 (BASELINE configs[2] shape) timing on one GPU: kernels of the general-scatter path at Q2 r3/r4."""
import os
import sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + "/tests", R + "/oracle"]
import verkko_hem_repo_b200 as vh  # noqa: E402
from helpers import b_phase_state, coef_vector  # noqa: E402

for refine in (3, 4):
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
    print("Q2 r%d: dofs %d nnzb %d slow_cells %d | assembly %.2f ms (pointwise %.2f) | spmv %.3f ms = %.0f GB/s | gmres its %d"
          % (refine, 18 * m.n_nodes, nnzb, info["n_slow_cells"], t_asm, t_pw, t_spmv, bytes_spmv / t_spmv / 1e6, its))
    ctx.close()
