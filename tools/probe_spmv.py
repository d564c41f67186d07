"""SpMV timing probe at C2 (Q1 r5): ms per launch with and without L2 flush."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + "/tests"]
import verkko_hem_repo_b200 as vh
from helpers import b_phase_state, coef_vector
m = vh.unit_cube(1, 5, half=20.0); T = m.tables(0)
ctx = vh.Context(T); ctx.set_coef_vector(coef_vector()); ctx.set_solution(b_phase_state(T, noise=0.0)); ctx.assemble()
nnzb = ctx.info()["nnzb"]
for fl in (True, False):
    ms = ctx.time_kernel(0, reps=20, flush_l2=fl)
    print("variant", os.environ.get("VH_SPMV_VARIANT", "0"), "flush", fl, "ms %.4f" % ms, "moved GB/s %.0f" % (8 * 180 * nnzb / ms / 1e6),
          "full-equivalent GB/s %.0f" % (8 * 324 * nnzb / ms / 1e6))
