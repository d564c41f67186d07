"""Host logic of bench.py that can run without a GPU: the Newton-run bookkeeping under the reference's stop rule
(run.cc:234-250), the compact summary keys, the config object shared by both arms, and the roofline arithmetic."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class FakeCtx:
    """A context whose residual falls by 10x per Newton step from 1.0 (converges below 5e-6 in 6 steps)."""

    def __init__(self):
        self.res, self.sets, self.ms = 1.0, 0, 0.0

    def set_solution(self, x):
        self.sets += 1
        self.res = 1.0

    def get_solution(self, out=None):
        return out

    def timer_start(self):
        pass

    def timer_stop(self):
        return 2.0

    def assemble(self):
        return self.res

    def solve(self, tol, max_it, restart):
        return 3, 0.0

    def line_search_trial(self, a):
        self.trial = self.res * 0.1

    def residual(self):
        return self.trial

    def accept_trial(self):
        self.res = self.trial


def test_run_steps_restarts_from_the_initial_condition_when_a_run_has_converged():
    ctx = FakeCtx()
    x0 = np.zeros(4)
    ms, runs = bench.run_steps(ctx, x0, 14)
    assert ms == 28.0
    assert [len(r) for r in runs] == [6, 6, 2]               # 1e-6 <= 5e-6 after six steps; the third run is cut by K
    assert ctx.sets == 3                                      # the IC and one reset per completed run
    s = bench.summarize_runs(runs)
    assert s["newton_steps"] == 14 and s["gmres_its_total"] == 42 and s["line_search_trials_total"] == 14
    assert s["runs_completed"] == 2 and s["steps_to_converge"] == 6 and len(s["first_run_residuals"]) == 6
    json.dumps(s)
    # e2e variant: the state goes through the host buffer every step
    buf = np.ones(4)
    ms2, runs2 = bench.run_steps(FakeCtx(), x0, 3, host_buffer=buf)
    assert ms2 == 6.0 and [len(r) for r in runs2] == [3] and not buf.any()


def test_config_object_is_the_same_in_both_arms_and_names_the_workload():
    n_dofs, n_cells = bench.problem_size(1, None, 7, 8)
    assert (n_dofs, n_cells) == (38640402, 2097152)
    a = bench.config_dict(1, None, 7, n_dofs, n_cells, 8)
    b = bench.config_dict(1, None, 7, n_dofs, n_cells, 8)
    assert a == b and "38640402 DoFs" in a["workload"] and "L2" in a["l2"] and "x8" in a["parallelism"] and "run.cc:234-250" in a["stop_rule"]
    assert bench.problem_size(1, 5, None, 1)[0] == 646866 and bench.problem_size(2, None, 5, 2)[0] == 4943250


def test_roofline_uses_the_ncu_traffic_of_the_workload_when_there_is_a_capture():
    class K:
        def time_kernel(self, what, reps, flush_l2):
            return {0: 5.5, 1: 21.5, 5: 8.7, 2: 4.4, 3: 0.96}[what]

        def measure_fp64_peak(self):
            return 36.8

    class T:
        n_owned_nodes, n_cells = 2146689, 2097152

    info = {"nnzb": 57066625, "n_packed_blocks": 57066625, "spmv_matrix_free": 1}
    roof, asm, kern = bench.kernel_numbers(K(), T, info, 1, 6491.2, "measured", "q1_g7_1gpu")
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["q1_g7_1gpu"]["k_points_apply"]
    assert roof["traffic"] == tj["read_bytes"] + tj["write_bytes"]
    assert abs(roof["achieved"] - roof["traffic"] / 5.5e-3 / 1e9) < 1e-6 and 0.8 < roof["frac"] < 0.9
    assert roof["frac_moved_model"] > roof["frac"] and roof["achieved_algorithmic"] > roof["peak"]
    assert abs(asm["frac_fp64_algorithmic"] - 2.0 * 512 * 336 * T.n_cells / 21.5e-3 / 1e12 / 36.8) < 1e-9
    assert kern["operator_apply_mode"] == "matrix-free"
    roof2, _, _ = bench.kernel_numbers(K(), T, info, 1, 6491.2, "measured", "no_such_workload")
    assert roof2["traffic"] is None and roof2["frac"] == roof2["frac_moved_model"]
