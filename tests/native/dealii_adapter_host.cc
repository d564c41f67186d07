// TEST INFRASTRUCTURE: compiles integration/dealii/vh_dealii_adapter.h against the functional deal.II mock and compares the
// tables it builds, rank by rank (ranks = threads), with the tables of the repository's mini host for the same mesh.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>

#include "../../integration/dealii/vh_dealii_adapter.h"

using namespace vhhost;

namespace
{
struct RankView
{
  dealii::DoFHandler<3>            dof_handler;
  dealii::IndexSet                 owned;
  dealii::AffineConstraints<double> c_newton, c_solution;
};

// owned / ghost / artificial as a p4est ghost layer sees it: ghost = cell of another rank whose closed box touches the
// box of an owned cell, directly or across a periodic face pair
void build_view(const Mesh &M, int rank, RankView &V)
{
  const int     n = M.degree == 1 ? 8 : 27;
  const int64_t nc = M.n_cells();
  double        lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  bool          periodic[3] = {false, false, false};
  std::vector<double> box((size_t)nc * 6);
  for (int64_t e = 0; e < nc; ++e)
    {
      double o[3], h[3];
      M.cell_box(e, o, h);
      for (int d = 0; d < 3; ++d)
        {
          box[6 * e + d]     = o[d];
          box[6 * e + 3 + d] = o[d] + h[d];
          lo[d]              = std::min(lo[d], o[d]);
          hi[d]              = std::max(hi[d], o[d] + h[d]);
          if (M.cell_face_bid[(size_t)e * 6 + 2 * d] == 5 + 2 * d)
            periodic[d] = true;
        }
    }
  std::vector<int> state(nc, 0);
  for (int64_t e = 0; e < nc; ++e)
    if (M.cell_rank[e] == rank)
      state[e] = 1;
  const double tol = 1e-9;
  for (int64_t e = 0; e < nc; ++e)
    {
      if (state[e] == 1)
        continue;
      for (int64_t o = 0; o < nc && state[e] == 0; ++o)
        {
          if (M.cell_rank[o] != rank)
            continue;
          for (int sx = -1; sx <= 1 && state[e] == 0; ++sx)
            for (int sy = -1; sy <= 1 && state[e] == 0; ++sy)
              for (int sz = -1; sz <= 1 && state[e] == 0; ++sz)
                {
                  const int s[3] = {sx, sy, sz};
                  bool      ok = true;
                  for (int d = 0; d < 3 && ok; ++d)
                    {
                      if (s[d] != 0 && !periodic[d])
                        ok = false;
                      const double sh = s[d] * (hi[d] - lo[d]);
                      if (box[6 * e + d] + sh > box[6 * o + 3 + d] + tol || box[6 * e + 3 + d] + sh < box[6 * o + d] - tol)
                        ok = false;
                    }
                  if (ok)
                    state[e] = 2;
                }
        }
    }
  for (int64_t e = 0; e < nc; ++e)
    V.dof_handler.cells.push_back({{&M, e, state[e]}});
  V.owned.add_range(18ull * M.rank_node_begin[rank], 18ull * M.rank_node_begin[rank + 1]);
  // locally relevant DoFs = DoFs of owned and ghost cells; the constraint objects know only those lines
  std::vector<uint8_t> relevant(M.n_nodes, 0);
  for (int64_t e = 0; e < nc; ++e)
    if (state[e])
      for (int a = 0; a < n; ++a)
        relevant[M.cell_nodes[(size_t)e * n + a]] = 1;
  for (size_t l = 0; l < M.c_dof.size(); ++l)
    if (relevant[M.c_dof[l] / 18])
      {
        dealii::AffineConstraints<double>::Entries ent;
        for (int64_t p = M.c_ptr[l]; p < M.c_ptr[l + 1]; ++p)
          ent.push_back({(unsigned long long)M.c_master[p], M.c_weight[p]});
        V.c_newton.lines[(unsigned long long)M.c_dof[l]]   = ent;
        V.c_solution.lines[(unsigned long long)M.c_dof[l]] = ent;
      }
}

template <class A, class B>
bool same(const A &a, const B &b, const char *what, int rank, std::string &msg)
{
  bool ok = a.size() == b.size();
  for (size_t i = 0; ok && i < a.size(); ++i)
    ok = (double)a[i] == (double)b[i];
  if (!ok && msg.empty())
    msg = std::string("rank ") + std::to_string(rank) + ": " + what + " differs (sizes " + std::to_string(a.size()) + " / " +
          std::to_string(b.size()) + ")";
  return ok;
}
} // namespace

// returns 0 if the adapter's tables equal the mini host's on every rank, 1 if they differ, 2 if the adapter refused
// (hanging-node rows outside a ghost layer), -1 on any other exception; msg receives the first finding
extern "C" int vht_adapter_check(int degree, int n_global_refine, const int *base, const int *face_bid, int local_refine, int n_ranks,
                                 char *msg_out, int msg_len)
{
  std::string msg;
  int         rc = 0;
  try
    {
      const double lo[3] = {-2.0, -1.5, -1.0}, hi[3] = {2.0, 1.5, 1.0};
      Mesh         M(degree, lo, hi, base, face_bid, n_global_refine);
      if (local_refine >= 2)
        { // randomised multi-level refinement: three rounds, ~15 % of the cells each (seed = local_refine)
          unsigned long long st = 0x9E3779B97F4A7C15ull * (unsigned long long)local_refine;
          for (int round = 0; round < 3; ++round)
            {
              std::vector<uint8_t> fl(M.n_cells(), 0);
              for (int64_t e = 0; e < M.n_cells(); ++e)
                {
                  st = st * 6364136223846793005ull + 1442695040888963407ull;
                  fl[e] = ((st >> 33) % 100) < 15 ? 1 : 0;
                }
              M.refine(fl);
            }
        }
      else if (local_refine)
        {
          std::vector<uint8_t> fl(M.n_cells(), 0);
          for (int64_t e = 0; e < M.n_cells(); ++e)
            {
              double c[3];
              M.cell_center(e, c);
              fl[e] = (std::fabs(c[2]) < 0.55 && c[0] < 0.1 && std::fabs(c[1]) < 0.8) ? 1 : 0;
            }
          M.refine(fl);
        }
      M.finalize(n_ranks);
      MockWorld world;
      world.n = (unsigned)n_ranks;
      std::vector<RankView>              views(n_ranks);
      std::vector<vh_dealii::HostTables> got(n_ranks);
      std::vector<std::string>           errs(n_ranks);
      std::vector<MockComm>              comms(n_ranks);
      for (int r = 0; r < n_ranks; ++r)
        {
          build_view(M, r, views[r]);
          comms[r] = MockComm{&world, (unsigned)r};
        }
      std::vector<std::thread> th;
      for (int r = 0; r < n_ranks; ++r)
        th.emplace_back([&, r] {
          try
            {
              dealii::FESystem<3> fe((unsigned)degree);
              got[r] = vh_dealii::build_tables<3>(views[r].dof_handler, fe, views[r].owned, views[r].c_newton, views[r].c_solution,
                                                  &comms[r]);
            }
          catch (const std::exception &e)
            {
              errs[r] = e.what();
            }
        });
      for (auto &t : th)
        t.join();
      for (int r = 0; r < n_ranks; ++r)
        if (!errs[r].empty())
          {
            msg = "rank " + std::to_string(r) + ": " + errs[r];
            rc  = errs[r].find("outside this rank's ghost layer") != std::string::npos ? 2 : -1;
          }
      for (int r = 0; r < n_ranks && rc == 0; ++r)
        {
          const RankTables              W = M.tables(r);
          const vh_dealii::HostTables &G = got[r];
          bool ok = G.degree == W.degree && G.n_owned_nodes == W.n_owned_nodes && G.n_ghost_nodes == W.n_ghost_nodes && G.n_cells == W.n_cells;
          if (!ok && msg.empty())
            msg = "rank " + std::to_string(r) + ": sizes differ (cells " + std::to_string(G.n_cells) + " / " + std::to_string(W.n_cells) +
                  ", ghosts " + std::to_string(G.n_ghost_nodes) + " / " + std::to_string(W.n_ghost_nodes) + ")";
          ok = ok && same(G.node_global, W.node_global, "node_global", r, msg) && same(G.cell_nodes, W.cell_nodes, "cell_nodes", r, msg) &&
               same(G.cell_origin, W.cell_origin, "cell_origin", r, msg) && same(G.cell_h, W.cell_h, "cell_h", r, msg) &&
               same(G.cell_owned, W.cell_owned, "cell_owned", r, msg) && same(G.wall_face_cell, W.wall_face_cell, "wall_face_cell", r, msg) &&
               same(G.wall_face_no, W.wall_face_no, "wall_face_no", r, msg) && same(G.wall_face_bid, W.wall_face_bid, "wall_face_bid", r, msg) &&
               same(G.newton_update.dof, W.c_dof, "constraint dof", r, msg) && same(G.newton_update.ptr, W.c_ptr, "constraint ptr", r, msg) &&
               same(G.newton_update.master, W.c_master, "constraint master", r, msg) &&
               same(G.newton_update.weight, W.c_weight, "constraint weight", r, msg) && same(G.solution.dof, W.c_dof, "constraint(solution) dof", r, msg) &&
               same(G.solution.master, W.c_master, "constraint(solution) master", r, msg) && same(G.peer_rank, W.peer_rank, "peer_rank", r, msg) &&
               same(G.send_ptr, W.send_ptr, "send_ptr", r, msg) && same(G.send_nodes, W.send_nodes, "send_nodes", r, msg) &&
               same(G.recv_ptr, W.recv_ptr, "recv_ptr", r, msg) && same(G.recv_nodes, W.recv_nodes, "recv_nodes", r, msg);
          if (!ok)
            rc = 1;
          // the descriptor view is consistent with the vectors it aliases
          const vh_mesh_desc d = G.desc();
          if (d.n_cells != G.n_cells || d.n_peers != (int32_t)G.peer_rank.size() || d.constraints_solution.n_lines != (int32_t)G.solution.dof.size())
            rc = 1;
        }
    }
  catch (const std::exception &e)
    {
      msg = e.what();
      rc  = -1;
    }
  if (msg_out && msg_len > 0)
    {
      std::strncpy(msg_out, msg.c_str(), (size_t)msg_len - 1);
      msg_out[msg_len - 1] = 0;
    }
  return rc;
}

// vh_last_error is referenced by vh_dealii::check(); the test library does not link the CUDA library
extern "C" const char *vh_last_error(const vh_ctx *) { return "test stub"; }
