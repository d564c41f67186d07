// vh_dealii_adapter.h — builds the flat tables of include/vh_femgl.h (vh_mesh_desc) from the deal.II objects VerHem's
// FemGL<dim> already owns (/root/reference/femgl/inc/femgl.h:274-308):
//     dof_handler, fe                      femgl.h:274,276      (FESystem<dim>(FE_Q<dim>(p), 18))
//     locally_owned_dofs                   femgl.h:304
//     constraints_newton_update / _solution femgl.h:307-308     (closed, reinit(locally_relevant_dofs), setup_*.cc:162-258)
//     mpi_communicator                     femgl.h:273
// This is the reference-side half of the drop-in boundary: header-only, needs deal.II >= 9.3 with MPI.
// deal.II is not installed in this repository's build image; the file is compiled and RUN there against a small functional
// mock of exactly the deal.II calls it makes (tests/native/dealii_mock/, backed by the repository's own box mesh, with
// ranks as threads), and its output is compared with the tables the mini host produces (tests/test_dealii_adapter.py).
//
// What it relies on [deal.II-internal, asserted at run time where possible]:
//  * FESystem(FE_Q(p),18) numbers the 18 components of one support point consecutively, locally and globally
//    (dof = 18*node + c); checked for every cell through fe.system_to_component_index().
//  * parallel::distributed::Triangulation gives every rank one contiguous range of global DoF indices, ascending with rank.
//  * cells are axis-aligned boxes (every makegrid_* variant femgl/CMakeLists.txt:42-53 lists); checked per cell.
#ifndef VH_DEALII_ADAPTER_H
#define VH_DEALII_ADAPTER_H

#include <deal.II/base/geometry_info.h>
#include <deal.II/base/index_set.h>
#include <deal.II/base/mpi.h>
#include <deal.II/dofs/dof_handler.h>
#include <deal.II/fe/fe_system.h>
#include <deal.II/lac/affine_constraints.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "vh_femgl.h"

namespace vh_dealii
{
struct ConstraintTable
{
  std::vector<int32_t> dof, ptr, master;
  std::vector<double>  weight;
  vh_constraints       view() const
  {
    vh_constraints c;
    c.n_lines = (int32_t)dof.size();
    c.dof     = dof.data();
    c.ptr     = ptr.data();
    c.master  = master.data();
    c.weight  = weight.data();
    return c;
  }
};

// Everything vh_create() wants, owned by the host for as long as desc() is in use.
struct HostTables
{
  int                  degree = 1;
  int32_t              n_owned_nodes = 0, n_ghost_nodes = 0, n_cells = 0;
  std::vector<int64_t> node_global; // [n_owned + n_ghost]; ghosts ascending (= grouped by owner rank)
  std::vector<int32_t> cell_nodes;
  std::vector<double>  cell_origin, cell_h;
  std::vector<uint8_t> cell_owned;
  std::vector<int32_t> wall_face_cell;
  std::vector<int8_t>  wall_face_no, wall_face_bid;
  ConstraintTable      newton_update, solution;
  std::vector<int32_t> peer_rank, send_ptr, send_nodes, recv_ptr, recv_nodes;

  vh_mesh_desc desc() const
  {
    vh_mesh_desc d{};
    d.degree                    = degree;
    d.n_owned_nodes             = n_owned_nodes;
    d.n_ghost_nodes             = n_ghost_nodes;
    d.node_global               = node_global.data();
    d.n_cells                   = n_cells;
    d.cell_nodes                = cell_nodes.data();
    d.cell_origin               = cell_origin.data();
    d.cell_h                    = cell_h.data();
    d.cell_owned                = cell_owned.data();
    d.n_wall_faces              = (int32_t)wall_face_cell.size();
    d.wall_face_cell            = wall_face_cell.data();
    d.wall_face_no              = wall_face_no.data();
    d.wall_face_bid             = wall_face_bid.data();
    d.constraints_newton_update = newton_update.view();
    d.constraints_solution      = solution.view();
    d.n_peers                   = (int32_t)peer_rank.size();
    d.peer_rank                 = peer_rank.data();
    d.send_ptr                  = send_ptr.data();
    d.send_nodes                = send_nodes.data();
    d.recv_ptr                  = recv_ptr.data();
    d.recv_nodes                = recv_nodes.data();
    return d;
  }
};

namespace internal
{
using gdi     = dealii::types::global_dof_index;
using Entries = std::vector<std::pair<gdi, double>>;

// The constraint lines this rank can look up: those of the AffineConstraints object (locally relevant DoFs) plus the
// lines that arrived with cells from beyond the ghost layer (see build_tables, step 2).
struct Lines
{
  const dealii::AffineConstraints<double> *object = nullptr;
  std::map<gdi, Entries>                   shipped;
  // nullptr: unconstrained (as far as this rank knows)
  const Entries *find(gdi dof, Entries &scratch) const
  {
    if (object->is_constrained(dof))
      {
        scratch.clear();
        if (const auto *e = object->get_constraint_entries(dof))
          for (const auto &kv : *e)
            scratch.push_back({(gdi)kv.first, (double)kv.second});
        return &scratch;
      }
    const auto it = shipped.find(dof);
    return it == shipped.end() ? nullptr : &it->second;
  }
};

// nodes a node depends on through its constraint lines (union over its 18 components)
inline void master_nodes_of(const Lines &c, gdi node, std::vector<gdi> &out)
{
  Entries scratch;
  for (unsigned int comp = 0; comp < 18; ++comp)
    if (const Entries *e = c.find(18 * node + comp, scratch))
      for (const auto &kv : *e)
        if (std::find(out.begin(), out.end(), kv.first / 18) == out.end())
          out.push_back(kv.first / 18);
}

struct CellRec
{
  gdi              id = 0; // global active cell index
  std::vector<gdi> nodes;  // global node ids in FE_Q local order
  double           origin[3], h[3];
  bool             owned = false;
  int8_t           face_bid[6]; // 2|3|4 on AdGR wall faces, else 0
};
} // namespace internal

// Call at the end of setup_system() (setup_*.cc:340), after both AffineConstraints objects are closed.
template <int dim>
HostTables build_tables(const dealii::DoFHandler<dim> &dof_handler, const dealii::FESystem<dim> &fe,
                        const dealii::IndexSet &locally_owned_dofs, const dealii::AffineConstraints<double> &constraints_newton_update,
                        const dealii::AffineConstraints<double> &constraints_solution, const MPI_Comm mpi_communicator)
{
  static_assert(dim == 3, "the femgl hot path is three-dimensional (sol/src/main.cc:116)");
  using internal::CellRec;
  using internal::Entries;
  using internal::gdi;
  HostTables T;
  T.degree = (int)fe.degree;
  if (T.degree != 1 && T.degree != 2)
    throw std::runtime_error("vh_dealii: FE_Q degree must be 1 or 2");
  const unsigned int dpc = fe.n_dofs_per_cell(), n = dpc / 18;
  if (dpc != 18 * n || n != (unsigned)((T.degree + 1) * (T.degree + 1) * (T.degree + 1)))
    throw std::runtime_error("vh_dealii: expected FESystem(FE_Q(p), 18)");
  if (!locally_owned_dofs.is_contiguous() || locally_owned_dofs.n_elements() % 18 != 0)
    throw std::runtime_error("vh_dealii: locally owned DoFs must be one contiguous range of whole nodes");
  const unsigned int my_rank = dealii::Utilities::MPI::this_mpi_process(mpi_communicator);
  const unsigned int n_ranks = dealii::Utilities::MPI::n_mpi_processes(mpi_communicator);
  const gdi          first_dof = locally_owned_dofs.n_elements() ? locally_owned_dofs.nth_index_in_set(0) : 0;
  if (first_dof % 18 != 0)
    throw std::runtime_error("vh_dealii: owned range does not start on a node boundary");
  const gdi node0 = first_dof / 18, node1 = node0 + locally_owned_dofs.n_elements() / 18;
  T.n_owned_nodes = (int32_t)(node1 - node0);
  // owner lookup: the owned node range of every rank
  const std::vector<gdi> all_first = dealii::Utilities::MPI::all_gather(mpi_communicator, node0);
  const std::vector<gdi> all_count = dealii::Utilities::MPI::all_gather(mpi_communicator, (gdi)T.n_owned_nodes);
  auto                   owner_of  = [&](gdi node) -> unsigned int {
    for (unsigned int r = 0; r < n_ranks; ++r)
      if (node >= all_first[r] && node < all_first[r] + all_count[r])
        return r;
    throw std::runtime_error("vh_dealii: node without an owner");
  };
  auto is_owned = [&](gdi node) { return node >= node0 && node < node1; };

  internal::Lines L[2];
  L[0].object = &constraints_newton_update;
  L[1].object = &constraints_solution;
  std::vector<gdi> masters;
  auto             masters_of = [&](gdi node) -> const std::vector<gdi> & {
    masters.clear();
    internal::master_nodes_of(L[0], node, masters);
    internal::master_nodes_of(L[1], node, masters);
    return masters;
  };

  // ---- step 1: owned and ghost cells -> records; which of them feed an owned row ----
  std::vector<CellRec>                        cells;
  std::vector<gdi>                            dofs(dpc);
  std::map<unsigned int, std::vector<double>> ship; // rank p -> serialised records of my cells that feed p's rows through constraints
  for (const auto &cell : dof_handler.active_cell_iterators())
    {
      if (!(cell->is_locally_owned() || cell->is_ghost()))
        continue;
      cell->get_dof_indices(dofs);
      CellRec rec;
      rec.id = (gdi)cell->global_active_cell_index();
      rec.nodes.assign(n, 0);
      for (unsigned int i = 0; i < dpc; ++i)
        {
          const auto ci = fe.system_to_component_index(i); // (component, shape function of FE_Q)
          if (ci.first == 0)
            {
              if (dofs[i] % 18 != 0)
                throw std::runtime_error("vh_dealii: component 0 of a support point is not on a multiple of 18");
              rec.nodes[ci.second] = dofs[i] / 18;
            }
        }
      for (unsigned int i = 0; i < dpc; ++i)
        {
          const auto ci = fe.system_to_component_index(i);
          if (dofs[i] != 18 * rec.nodes[ci.second] + ci.first)
            throw std::runtime_error("vh_dealii: the 18 components of a support point are not numbered consecutively");
        }
      // axis-aligned box: vertex 0 is the origin, the last vertex the far corner
      const auto v0 = cell->vertex(0), v7 = cell->vertex(dealii::GeometryInfo<dim>::vertices_per_cell - 1);
      for (unsigned int d = 0; d < 3; ++d)
        {
          rec.origin[d] = v0[d];
          rec.h[d]      = v7[d] - v0[d];
          if (!(rec.h[d] > 0.0))
            throw std::runtime_error("vh_dealii: cell is not a positively oriented box");
        }
      for (unsigned int v = 1; v + 1 < dealii::GeometryInfo<dim>::vertices_per_cell; ++v)
        {
          const auto p = cell->vertex(v);
          for (unsigned int d = 0; d < 3; ++d)
            {
              const double want = rec.origin[d] + (((v >> d) & 1u) ? rec.h[d] : 0.0);
              if (std::fabs(p[d] - want) > 1e-10 * rec.h[d])
                throw std::runtime_error("vh_dealii: only axis-aligned box cells are supported by this round's kernels");
            }
        }
      rec.owned = cell->is_locally_owned();
      for (unsigned int f = 0; f < dealii::GeometryInfo<dim>::faces_per_cell; ++f)
        {
          rec.face_bid[f] = 0;
          if (cell->face(f)->at_boundary())
            {
              const unsigned int b = cell->face(f)->boundary_id();
              if (b >= 2 && b <= 4) // AdGR walls with normal x|y|z (assemble.cc:288-292)
                rec.face_bid[f] = (int8_t)b;
            }
        }
      // does this cell feed an owned row (a node itself, or a node whose constraint names an owned master)?  And, for my own
      // cells: does it feed rows of OTHER ranks through a constraint?  Those ranks may not see the cell (step 2).
      bool                      need = false;
      std::vector<unsigned int> foreign;
      for (unsigned int a = 0; a < n; ++a)
        {
          if (is_owned(rec.nodes[a]))
            need = true;
          for (const gdi m : masters_of(rec.nodes[a]))
            {
              if (is_owned(m))
                need = true;
              else if (rec.owned)
                {
                  const unsigned int p = owner_of(m);
                  if (std::find(foreign.begin(), foreign.end(), p) == foreign.end())
                    foreign.push_back(p);
                }
            }
        }
      for (const unsigned int p : foreign)
        { // serialise: id, nodes, box, wall ids, then the lines of its 18 n DoFs in both constraint objects
          std::vector<double> &out = ship[p];
          out.push_back((double)rec.id);
          for (const gdi nd : rec.nodes)
            out.push_back((double)nd);
          out.insert(out.end(), rec.origin, rec.origin + 3);
          out.insert(out.end(), rec.h, rec.h + 3);
          for (unsigned int f = 0; f < 6; ++f)
            out.push_back((double)rec.face_bid[f]);
          Entries scratch;
          for (int o = 0; o < 2; ++o)
            for (const gdi nd : rec.nodes)
              for (unsigned int comp = 0; comp < 18; ++comp)
                {
                  const Entries *e = L[o].find(18 * nd + comp, scratch);
                  out.push_back(e ? (double)e->size() : -1.0);
                  if (e)
                    for (const auto &kv : *e)
                      {
                        out.push_back((double)kv.first);
                        out.push_back(kv.second);
                      }
                }
        }
      if (need)
        cells.push_back(std::move(rec));
    }

  // ---- step 2: owner-computes needs EVERY cell that feeds an owned row.  The reference sends such contributions with
  //      compress(add) (assemble.cc:369-370); here the owner visits the cell itself.  On conforming meshes and across
  //      periodic faces (add_periodicity, makegrid_*periodic.cc) those cells are in the ghost layer.  With hanging nodes the
  //      master of a hanging node can belong to a rank that touches the fine cell in no point: its owner ships the cell
  //      (nodes, box, wall ids) and the constraint lines of its DoFs, once per mesh. ----
  {
    std::vector<gdi> visible;
    for (const auto &cell : dof_handler.active_cell_iterators())
      if (cell->is_locally_owned() || cell->is_ghost())
        visible.push_back((gdi)cell->global_active_cell_index());
    std::sort(visible.begin(), visible.end());
    const std::map<unsigned int, std::vector<double>> got = dealii::Utilities::MPI::some_to_some(mpi_communicator, ship);
    for (const auto &kv : got)
      {
        const std::vector<double> &in = kv.second;
        size_t                     k  = 0;
        while (k < in.size())
          {
            CellRec rec;
            rec.id = (gdi)in[k++];
            rec.nodes.resize(n);
            for (unsigned int a = 0; a < n; ++a)
              rec.nodes[a] = (gdi)in[k++];
            for (unsigned int d = 0; d < 3; ++d)
              rec.origin[d] = in[k++];
            for (unsigned int d = 0; d < 3; ++d)
              rec.h[d] = in[k++];
            for (unsigned int f = 0; f < 6; ++f)
              rec.face_bid[f] = (int8_t)in[k++];
            rec.owned        = false;
            const bool known = std::binary_search(visible.begin(), visible.end(), rec.id);
            for (int o = 0; o < 2; ++o)
              for (unsigned int a = 0; a < n; ++a)
                for (unsigned int comp = 0; comp < 18; ++comp)
                  {
                    const double cnt = in[k++];
                    if (cnt < 0)
                      continue;
                    Entries e((size_t)cnt);
                    for (auto &kv2 : e)
                      {
                        kv2.first  = (gdi)in[k++];
                        kv2.second = in[k++];
                      }
                    const gdi dof = 18 * rec.nodes[a] + comp;
                    if (!known && !L[o].object->is_constrained(dof))
                      L[o].shipped[dof] = std::move(e);
                  }
            if (!known) // otherwise step 1 has it already (it is in the ghost layer)
              cells.push_back(std::move(rec));
          }
      }
    std::sort(cells.begin(), cells.end(), [](const CellRec &x, const CellRec &y) { return x.id < y.id; });
  }

  // ---- step 3: ghost set = nodes of the visited cells and the masters of their constraints ----
  std::vector<gdi> ghosts;
  for (const CellRec &rec : cells)
    for (unsigned int a = 0; a < n; ++a)
      {
        if (!is_owned(rec.nodes[a]))
          ghosts.push_back(rec.nodes[a]);
        for (const gdi m : masters_of(rec.nodes[a]))
          if (!is_owned(m))
            ghosts.push_back(m);
      }
  std::sort(ghosts.begin(), ghosts.end());
  ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
  T.n_ghost_nodes = (int32_t)ghosts.size();
  T.node_global.resize((size_t)T.n_owned_nodes + ghosts.size());
  for (int32_t i = 0; i < T.n_owned_nodes; ++i)
    T.node_global[i] = (int64_t)(node0 + i);
  for (size_t i = 0; i < ghosts.size(); ++i)
    T.node_global[T.n_owned_nodes + i] = (int64_t)ghosts[i];
  auto local_of = [&](gdi node) -> int32_t { // -1: not local
    if (is_owned(node))
      return (int32_t)(node - node0);
    const auto it = std::lower_bound(ghosts.begin(), ghosts.end(), node);
    return (it != ghosts.end() && *it == node) ? (int32_t)(T.n_owned_nodes + (it - ghosts.begin())) : -1;
  };

  // ---- step 4: cell tables ----
  T.n_cells = (int32_t)cells.size();
  for (size_t k = 0; k < cells.size(); ++k)
    {
      const CellRec &rec = cells[k];
      for (unsigned int a = 0; a < n; ++a)
        T.cell_nodes.push_back(local_of(rec.nodes[a]));
      T.cell_origin.insert(T.cell_origin.end(), rec.origin, rec.origin + 3);
      T.cell_h.insert(T.cell_h.end(), rec.h, rec.h + 3);
      T.cell_owned.push_back(rec.owned ? 1 : 0);
      for (unsigned int f = 0; f < 6; ++f)
        if (rec.face_bid[f])
          {
            T.wall_face_cell.push_back((int32_t)k);
            T.wall_face_no.push_back((int8_t)f);
            T.wall_face_bid.push_back(rec.face_bid[f]);
          }
    }

  // ---- step 6: halo plan: tell every owner which of its nodes we ghost (in our ghost order = its send order) ----
  std::map<unsigned int, std::vector<gdi>>     wanted; // owner -> global nodes
  std::map<unsigned int, std::vector<int32_t>> recv_local;
  for (size_t i = 0; i < ghosts.size(); ++i)
    {
      const unsigned int p = owner_of(ghosts[i]);
      wanted[p].push_back(ghosts[i]);
      recv_local[p].push_back((int32_t)(T.n_owned_nodes + i));
    }
  const std::map<unsigned int, std::vector<gdi>> requested = dealii::Utilities::MPI::some_to_some(mpi_communicator, wanted);
  T.send_ptr.push_back(0);
  T.recv_ptr.push_back(0);
  for (unsigned int p = 0; p < n_ranks; ++p)
    {
      if (p == my_rank)
        continue;
      const auto rq = requested.find(p);
      const auto rc = recv_local.find(p);
      const bool has_send = rq != requested.end() && !rq->second.empty(), has_recv = rc != recv_local.end() && !rc->second.empty();
      if (!has_send && !has_recv)
        continue;
      T.peer_rank.push_back((int32_t)p);
      if (has_send)
        for (const gdi node : rq->second)
          {
            if (!is_owned(node))
              throw std::runtime_error("vh_dealii: a peer asked for a node this rank does not own");
            T.send_nodes.push_back((int32_t)(node - node0));
          }
      if (has_recv)
        T.recv_nodes.insert(T.recv_nodes.end(), rc->second.begin(), rc->second.end());
      T.send_ptr.push_back((int32_t)T.send_nodes.size());
      T.recv_ptr.push_back((int32_t)T.recv_nodes.size());
    }

  // ---- step 6b: a ghost node that is local only as somebody's master (or sits in a shipped cell) may carry constraint
  //      lines this rank's AffineConstraints objects do not hold (they know the locally relevant lines only), e.g. the
  //      Dirichlet-masked components of a wall node.  Its owner knows them: it answers the halo request with the lines. ----
  {
    std::map<unsigned int, std::vector<double>> reply;
    Entries                                     scratch;
    for (const auto &kv : requested)
      {
        std::vector<double> &out = reply[kv.first];
        for (const gdi node : kv.second)
          for (int o = 0; o < 2; ++o)
            for (unsigned int comp = 0; comp < 18; ++comp)
              {
                const Entries *e = L[o].find(18 * node + comp, scratch);
                out.push_back(e ? (double)e->size() : -1.0);
                if (e)
                  for (const auto &en : *e)
                    {
                      out.push_back((double)en.first);
                      out.push_back(en.second);
                    }
              }
      }
    const std::map<unsigned int, std::vector<double>> answers = dealii::Utilities::MPI::some_to_some(mpi_communicator, reply);
    for (const auto &kv : answers)
      {
        const std::vector<double> &in    = kv.second;
        const std::vector<gdi>    &nodes = wanted[kv.first];
        size_t                     k     = 0;
        for (const gdi node : nodes)
          for (int o = 0; o < 2; ++o)
            for (unsigned int comp = 0; comp < 18; ++comp)
              {
                const double cnt = in.at(k++);
                if (cnt < 0)
                  continue;
                Entries e((size_t)cnt);
                for (auto &en : e)
                  {
                    en.first  = (gdi)in.at(k++);
                    en.second = in.at(k++);
                  }
                const gdi dof = 18 * node + comp;
                if (!L[o].object->is_constrained(dof) && !L[o].shipped.count(dof))
                  L[o].shipped[dof] = std::move(e);
              }
      }
  }

  // ---- step 7: constraint tables over local DoFs (lines whose masters are all local; the others belong to ghost DoFs
  //      whose value simply arrives through the halo) ----
  auto fill = [&](const internal::Lines &c, ConstraintTable &out) {
    out.ptr.push_back(0);
    const int32_t n_local = T.n_owned_nodes + T.n_ghost_nodes;
    Entries       scratch;
    for (int32_t ln = 0; ln < n_local; ++ln)
      for (unsigned int comp = 0; comp < 18; ++comp)
        {
          const Entries *entries = c.find(18 * (gdi)T.node_global[ln] + comp, scratch);
          if (!entries)
            continue;
          bool all_local = true;
          for (const auto &e : *entries)
            if (local_of(e.first / 18) < 0)
              all_local = false;
          if (!all_local)
            continue;
          out.dof.push_back(18 * ln + (int32_t)comp);
          for (const auto &e : *entries)
            {
              out.master.push_back(18 * local_of(e.first / 18) + (int32_t)(e.first % 18));
              out.weight.push_back(e.second);
            }
          out.ptr.push_back((int32_t)out.master.size());
        }
  };
  fill(L[0], T.newton_update);
  fill(L[1], T.solution);

  return T;
}

// owned part of a (ghosted or distributed) deal.II vector, in the order vh_set_solution / vh_get_solution use
template <class VectorType>
void owned_values(const VectorType &v, const dealii::IndexSet &locally_owned_dofs, std::vector<double> &out)
{
  out.resize(locally_owned_dofs.n_elements());
  size_t k = 0;
  for (const auto i : locally_owned_dofs)
    out[k++] = v[i];
}
// ... and back into a distributed (non-ghosted) vector; follow with the usual ghosted assignment (solve.cc:183)
template <class VectorType>
void set_owned_values(const std::vector<double> &in, const dealii::IndexSet &locally_owned_dofs, VectorType &distributed)
{
  size_t k = 0;
  for (const auto i : locally_owned_dofs)
    distributed[i] = in[k++];
}

inline void check(int rc, vh_ctx *ctx, const char *what)
{ // reaches the handler of sol/src/main.cc:120-145 like a deal.II exception
  if (rc != VH_OK)
    throw std::runtime_error(std::string(what) + ": " + vh_last_error(ctx));
}
} // namespace vh_dealii
#endif
