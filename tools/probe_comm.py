"""Communication latencies under torchrun (N ranks): ghost refresh and inner product + all-reduce, device time and wall time."""
import os
import sys
import time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + "/tests"]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import verkko_hem_repo_b200 as vh  # noqa: E402
from helpers import b_phase_state, coef_vector  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
refine = int(sys.argv[1]) if len(sys.argv) > 1 else 5
m = vh.Mesh(1, [-20, -20, -20], [20, 20, -20 + 40 * world], base=(1, 1, world), face_bid=(1, 1, 1, 1, 4, 4), n_global_refine=refine).finalize(world)
T = m.tables(rank)
ctx = vh.Context(T, device=int(os.environ["LOCAL_RANK"]))
uid = [vh.Context.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
ctx.set_coef_vector(coef_vector())
ctx.set_solution(b_phase_state(T, noise=0.0)[:18 * T.n_owned_nodes])
ctx.assemble()
for what, name in ((9, "halo exchange"), (10, "dot + all-reduce")):
    ctx.time_kernel(what, reps=2, flush_l2=False)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    ms = ctx.time_kernel(what, reps=10, flush_l2=False)
    wall = (time.perf_counter() - t0) / 10
    if rank == 0:
        print("%-18s device %.1f us each, wall %.1f us each (VH_P2P=%s, %d ranks)" % (name, ms / 20 * 1e3, wall / 20 * 1e6,
                                                                                   os.environ.get("VH_P2P", "1"), world))
ctx.close()
dist.destroy_process_group()
