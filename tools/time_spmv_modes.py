"""Device-timed operator apply at C2 size (or --refine R): packed SpMV over the assembled blocks vs the matrix-free apply
from the H_q tables (vh_set_spmv_matrix_free).  CUDA events on the library's stream, L2 flushed between launches."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT,):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import verkko_hem_repo_b200 as vh  # noqa: E402

degree = int(sys.argv[1]) if len(sys.argv) > 1 else 1
refine = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
t0 = time.time()
T = vh.unit_cube(degree, refine, half=2.0).tables(0)
ctx = vh.Context(T)
ctx.set_coefficients(0.42072, 0.42072, 0.42072, -0.5, (-0.010850915879921348, 0.020598836429398658, 0.02117724292551364,
                                                     0.019780381922869742, -0.023091499302424053), 2.0)
x = np.zeros((T.n_local_nodes, 18))
x[:, [0, 4, 8]] = 3.99 * 0.577350269
x += 0.05 * np.random.default_rng(1).uniform(-1, 1, x.shape)
ctx.set_solution(x.ravel())
ctx.assemble()
info = ctx.info()
n = 8 if degree == 1 else 27
print("Q%d r%d: %d DoFs, %d blocks, setup %.1f s" % (degree, refine, info["n_owned_dofs"], info["nnzb"], time.time() - t0), flush=True)
modes = (0, 1, 0, 1, 2, 2, 3, 3)
for mode in modes:
    ctx.set_spmv_matrix_free(mode)
    ms = ctx.time_kernel(0, reps, True)
    byts = (8 * 180 * n * T.n_cells) if mode in (1, 3) else ((8 * 180 * info["n_packed_blocks"]) if mode == 0 else 2 * 8 * 18 * T.n_local_nodes)
    print("mode %s: %.4f ms per apply, streams %.3f GB -> %.0f GB/s" % (("packed-spmv", "matrix-free", "table-free", "matrix-free-v2")[mode], ms, byts / 1e9, byts / ms / 1e6),
          flush=True)
ctx.close()
