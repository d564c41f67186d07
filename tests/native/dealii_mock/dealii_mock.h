// TEST INFRASTRUCTURE.  A small FUNCTIONAL mock of exactly the deal.II calls integration/dealii/vh_dealii_adapter.h makes,
// backed by the repository's own box mesh (verkko-hem-repo_b200/host/mesh.h), with MPI ranks played by threads.  It lets
// the adapter be compiled and run in an image without deal.II; it does not pretend to be deal.II anywhere else.
// Semantics imitated: cells are owned / ghost (geometrically adjacent to an owned cell, across periodic faces too, as a
// p4est ghost layer with add_periodicity) / artificial; AffineConstraints only knows the lines of locally relevant DoFs.
#ifndef VH_DEALII_MOCK_H
#define VH_DEALII_MOCK_H

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include <map>
#include <mutex>
#include <stdexcept>
#include <utility>
#include <vector>

#include "../../../verkko-hem-repo_b200/host/mesh.h"

// ---------------- MPI: ranks are threads of one process ----------------
struct MockWorld
{
  unsigned int            n = 1;
  std::mutex              mu;
  std::condition_variable cv;
  unsigned int            arrived = 0, generation = 0;
  std::vector<std::vector<unsigned long long>>                                  gather_slots;
  void barrier()
  {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned int           gen = generation;
    if (++arrived == n)
      {
        arrived = 0;
        ++generation;
        cv.notify_all();
      }
    else
      cv.wait(lk, [&] { return generation != gen; });
  }
};
struct MockComm
{
  MockWorld   *world;
  unsigned int rank;
};
typedef const MockComm *MPI_Comm;

namespace dealii
{
namespace types
{
using global_dof_index  = unsigned long long;
using global_cell_index = unsigned long long;
using boundary_id       = unsigned int;
} // namespace types

template <int dim>
struct GeometryInfo
{
  static constexpr unsigned int faces_per_cell    = 2 * dim;
  static constexpr unsigned int vertices_per_cell = 1u << dim;
};

template <int dim>
struct Point
{
  double        x[dim];
  double        operator[](unsigned int d) const { return x[d]; }
  double        operator()(unsigned int d) const { return x[d]; }
};

class IndexSet
{
public:
  void add_range(types::global_dof_index b, types::global_dof_index e)
  {
    for (types::global_dof_index i = b; i < e; ++i)
      v.push_back(i);
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  }
  void                    add_index(types::global_dof_index i) { add_range(i, i + 1); }
  types::global_dof_index n_elements() const { return v.size(); }
  bool                    is_contiguous() const { return v.empty() || v.back() - v.front() + 1 == v.size(); }
  types::global_dof_index nth_index_in_set(types::global_dof_index k) const { return v.at(k); }
  bool                    is_element(types::global_dof_index i) const { return std::binary_search(v.begin(), v.end(), i); }
  std::vector<types::global_dof_index>::const_iterator begin() const { return v.begin(); }
  std::vector<types::global_dof_index>::const_iterator end() const { return v.end(); }
  std::vector<types::global_dof_index>                 v;
};

template <int dim>
class FESystem
{
public:
  explicit FESystem(unsigned int p) : degree(p) {}
  unsigned int                          degree;
  unsigned int                          n_dofs_per_cell() const { return 18 * (degree + 1) * (degree + 1) * (degree + 1); }
  std::pair<unsigned int, unsigned int> system_to_component_index(unsigned int i) const { return {i % 18, i / 18}; }
};

template <int dim>
class DoFHandler
{
public:
  struct FaceAcc
  {
    int                bid;
    bool               at_boundary() const { return bid != 0; }
    types::boundary_id boundary_id() const { return (types::boundary_id)bid; }
    const FaceAcc     *operator->() const { return this; }
  };
  struct CellAcc
  {
    const vhhost::Mesh *mesh;
    int64_t             e;
    int                 state; // 0 artificial, 1 owned, 2 ghost
    bool                is_locally_owned() const { return state == 1; }
    bool                is_ghost() const { return state == 2; }
    bool                is_artificial() const { return state == 0; }
    types::global_cell_index global_active_cell_index() const { return (types::global_cell_index)e; }
    void get_dof_indices(std::vector<types::global_dof_index> &dofs) const
    {
      if (state == 0)
        throw std::runtime_error("mock deal.II: get_dof_indices on an artificial cell");
      const int n = mesh->degree == 1 ? 8 : 27;
      for (int a = 0; a < n; ++a)
        for (int c = 0; c < 18; ++c)
          dofs[18 * a + c] = 18ull * (types::global_dof_index)mesh->cell_nodes[(size_t)e * n + a] + c;
    }
    Point<dim> vertex(unsigned int v) const
    {
      double o[3], h[3];
      mesh->cell_box(e, o, h);
      Point<dim> p;
      for (int d = 0; d < dim; ++d)
        p.x[d] = o[d] + (((v >> d) & 1u) ? h[d] : 0.0);
      return p;
    }
    FaceAcc face(unsigned int f) const { return FaceAcc{mesh->cell_face_bid[(size_t)e * 6 + f]}; }
  };
  struct CellIt
  {
    CellAcc        acc;
    const CellAcc *operator->() const { return &acc; }
  };
  const std::vector<CellIt> &active_cell_iterators() const { return cells; }
  std::vector<CellIt>        cells;
};

template <typename number>
class AffineConstraints
{
public:
  using size_type = types::global_dof_index;
  using Entries   = std::vector<std::pair<size_type, number>>;
  // the lines of the locally relevant DoFs only, as after reinit(locally_relevant_dofs) + close()
  std::map<size_type, Entries> lines;
  bool                         is_constrained(size_type i) const { return lines.count(i) != 0; }
  const Entries               *get_constraint_entries(size_type i) const
  {
    const auto it = lines.find(i);
    return it == lines.end() ? nullptr : &it->second;
  }
};

namespace Utilities
{
namespace MPI
{
inline unsigned int this_mpi_process(MPI_Comm c) { return c->rank; }
inline unsigned int n_mpi_processes(MPI_Comm c) { return c->world->n; }
template <typename T>
std::vector<T> all_gather(MPI_Comm c, const T &v)
{
  MockWorld &w = *c->world;
  {
    std::lock_guard<std::mutex> lk(w.mu);
    w.gather_slots.resize(w.n);
    w.gather_slots[c->rank] = {(unsigned long long)v};
  }
  w.barrier();
  std::vector<T> out;
  for (unsigned int r = 0; r < w.n; ++r)
    out.push_back((T)w.gather_slots[r][0]);
  w.barrier();
  return out;
}
template <typename T>
std::map<unsigned int, std::vector<T>> some_to_some(MPI_Comm c, const std::map<unsigned int, std::vector<T>> &to_send)
{
  MockWorld &w = *c->world;
  // one board per (world, payload type); guarded by the world's mutex
  static std::map<MockWorld *, std::vector<std::map<unsigned int, std::vector<T>>>> boards;
  std::vector<std::map<unsigned int, std::vector<T>>>                             *board;
  {
    std::lock_guard<std::mutex> lk(w.mu);
    board = &boards[&w];
    board->resize(w.n);
    (*board)[c->rank] = to_send;
  }
  w.barrier();
  std::map<unsigned int, std::vector<T>> out;
  // test hook: lose the cell records the adapter ships beyond the ghost layer (they are the only double payloads)
  const bool drop = std::is_same<T, double>::value && std::getenv("VH_ADAPTER_TEST_NO_SHIPPING") != nullptr;
  for (unsigned int r = 0; r < w.n && !drop; ++r)
    {
      const auto it = (*board)[r].find(c->rank);
      if (it != (*board)[r].end())
        out[r] = it->second;
    }
  w.barrier(); // (boards are never erased: a later call overwrites every slot before its first barrier)
  return out;
}
} // namespace MPI
} // namespace Utilities
} // namespace dealii
#endif
