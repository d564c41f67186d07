"""Adaptive-mesh timing on one GPU (BASELINE configs[3] shape): Q1 cube refined globally, then a slab |z| < w refined once
more -> two planes of hanging nodes; prints the split between lattice rows and general-scatter cells and the kernel times."""
import os
import sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + "/tests"]
import numpy as np  # noqa: E402
import verkko_hem_repo_b200 as vh  # noqa: E402
from helpers import b_phase_state, coef_vector  # noqa: E402

degree = int(sys.argv[1]) if len(sys.argv) > 1 else 1
refine = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m = vh.Mesh(degree, [-20, -20, -20], [20, 20, 20], face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=refine)
c = m.cell_centers()
m.refine(np.abs(c[:, 2]) < 5.0)
m.finalize(1)
T = m.tables(0)
ctx = vh.Context(T)
ctx.set_coef_vector(coef_vector())
ctx.set_solution(b_phase_state(T, noise=0.0))
ctx.assemble()
info = ctx.info()
t_asm = ctx.time_kernel(1, reps=3, flush_l2=True)
t_pw = ctx.time_kernel(5, reps=3, flush_l2=True)
t_rows = ctx.time_kernel(6, reps=3, flush_l2=True)
t_res = ctx.time_kernel(2, reps=3, flush_l2=True)
t_spmv = ctx.time_kernel(0, reps=10, flush_l2=True)
its, res = ctx.solve(1e-1)
print("Q%d r%d+slab: cells %d dofs %d hanging %d | fast rows %d of %d, slow cells %d | assembly %.2f ms (pointwise %.2f, lattice rows %.2f, "
      "scatter %.2f) | residual %.2f ms | spmv %.3f ms | gmres its %d"
      % (degree, refine, m.n_cells, 18 * m.n_nodes, m.n_hanging_nodes, info["n_fast_rows"], T.n_owned_nodes, info["n_slow_cells"], t_asm,
         t_pw, t_rows, t_asm - t_pw - t_rows, t_res, t_spmv, its))
ctx.close()
