// Host-side mesh / DoF / constraint tables for the femgl hot path.
//
// In production these tables come from deal.II (parallel::distributed::Triangulation + DoFHandler +
// AffineConstraints; /root/reference/femgl/inc/femgl.h:274-308).  deal.II is not available in this
// image, so this file provides the same *tables* for the reference's box geometries:
//   GridGenerator::hyper_cube / hyper_rectangle + refine_global   (makegrid_cube-z-normal_AdGR.cc:151-160,
//                                                                  makegrid_retangle-z-AdGR-xy-HomoNeumann.cc:159-192)
//   boundary ids 1 (natural) and 2|3|4 (AdGR walls with normal x|y|z) (makegrid_cube-z-normal_AdGR.cc:164-195)
//   boundary id pairs (5,6)|(7,8)|(9,10): periodic in x|y|z (makegrid_retangle-z-AdGR_xy-periodic.cc:167-219,
//                                         setup_weak-coupling-PDW-configuration.cc:128-206; the reference's active variant)
//   FESystem(FE_Q(p),18) DoF layout, hanging-node + component-masked Dirichlet constraints
//                                                                 (setup_uniform_B-phase.cc:135-186, femgl.h:294-301)
//   p4est-style partition: contiguous ranges of the Morton (z-order) cell sequence, DoFs on a
//   subdomain interface owned by the lower rank.
// It emits exactly the flat arrays of include/vh_femgl.h (vh_mesh_desc).
#ifndef VH_HOST_MESH_H
#define VH_HOST_MESH_H

#include <cstdint>
#include <string>
#include <vector>

namespace vhhost
{
struct Leaf
{
  int     level;
  int32_t g[3]; // global integer cell coordinates at `level`: g[d] in [0, base[d] << level)
};

// Flat per-rank tables; field names follow vh_mesh_desc.
struct RankTables
{
  int                  degree = 1;
  int32_t              n_owned_nodes = 0, n_ghost_nodes = 0, n_cells = 0;
  std::vector<int64_t> node_global;
  std::vector<double>  node_xyz; // [n_local][3] support point coordinates (initial conditions, output)
  std::vector<int32_t> cell_nodes;
  std::vector<int64_t> cell_global; // global (Morton) cell index
  std::vector<double>  cell_origin, cell_h;
  std::vector<uint8_t> cell_owned;
  std::vector<int32_t> wall_face_cell;
  std::vector<int8_t>  wall_face_no, wall_face_bid;
  // constraints over local DoFs (the two AffineConstraints objects have identical structure here)
  std::vector<int32_t> c_dof, c_ptr, c_master;
  std::vector<double>  c_weight;
  // halo plan
  std::vector<int32_t> peer_rank, send_ptr, send_nodes, recv_ptr, recv_nodes;
};

class Mesh
{
public:
  // Box [lo,hi] with base[0] x base[1] x base[2] root cells, each refined `n_global_refine` times.
  // face_bid[f]: boundary id of box face f (deal.II order x0,x1,y0,y1,z0,z1).
  Mesh(int degree, const double lo[3], const double hi[3], const int base[3], const int face_bid[6], int n_global_refine);

  int64_t n_cells() const { return (int64_t)leaves.size(); }
  // Refine flagged leaves (flags indexed by current Morton cell index), then restore 2:1 balance
  // over faces and edges by refining coarser neighbours.
  void refine(const std::vector<uint8_t> &flags);
  void refine_global(int times);
  // Centre of cell e (current leaf order) — used by refinement indicators.
  void cell_center(int64_t e, double c[3]) const;

  // Number nodes, build constraints, partition into n_ranks.  Must be called after the last refine().
  void finalize(int n_ranks);

  RankTables tables(int rank) const;

  // SolutionTransfer::interpolate stand-in (refine.cc:128-130,171-175): FE interpolation of a field given on the
  // nodes of `old_mesh` (global node order, 18 values per node) at this mesh's nodes.  Both meshes must be finalized
  // and this mesh must be a refinement of old_mesh.
  void interpolate_from(const Mesh &old_mesh, const std::vector<double> &old_values, std::vector<double> &new_values) const;
  // The same interpolation as a table, for the device-side transfer (vh_transfer_solution): row i = node i of this mesh,
  // entries (node of old_mesh, weight) = the shape functions of the old cell that contains the node.  Global node ids.
  void transfer_table(const Mesh &old_mesh, std::vector<int32_t> &ptr, std::vector<int32_t> &src, std::vector<double> &weight) const;
  // Kelly-type refinement indicator (refine.cc:144-148, KellyErrorEstimator with no Neumann data): per cell
  //   eta_K^2 = sum_{interior faces F of K} h_K/24 * |F| * sum_c [d_n u_c]^2,
  // the jump of the normal derivative taken between the cell-mean gradients of the two leaves that share the face
  // (quarter-face sampling, so finer and coarser neighbours across hanging-node faces are handled; periodic faces wrap).
  void kelly_indicator(const std::vector<double> &values, std::vector<double> &eta) const;

  // ---- global data, valid after finalize() ----
  int                  degree;
  int                  n_ranks = 0;
  int64_t              n_nodes = 0;
  std::vector<int64_t> cell_nodes;          // [n_cells][n]
  std::vector<int>     cell_rank;           // [n_cells]
  std::vector<int>     node_rank;           // [n_nodes] owner
  std::vector<int64_t> rank_node_begin;     // [n_ranks+1] owned node ranges (rank-major numbering)
  std::vector<double>  node_xyz;            // [n_nodes][3]
  std::vector<int8_t>  cell_face_bid;       // [n_cells][6]; 0 = interior face
  // closed constraints over global DoFs (18*node+c), sorted by dof
  std::vector<int64_t> c_dof, c_ptr, c_master;
  std::vector<double>  c_weight;
  int64_t              n_hanging_nodes = 0;
  int64_t              n_periodic_nodes = 0; // nodes of an upper periodic face identified with their image on the lower face
  std::vector<int64_t> node_lattice;        // [n_nodes][3] integer lattice coordinates, root cell side = lattice_U units
  int64_t              lattice_U = 0;

  void cell_box(int64_t e, double origin[3], double h[3]) const;

private:
  double            lo[3], hi[3];
  int               base[3];
  int               bid[6];
  std::vector<Leaf> leaves; // Morton order (roots lexicographic, z-order inside a root)
  int               Lmax = 0;
  void              sort_leaves();
  // per-rank ghost node lists (global ids, sorted by (owner, id)), built lazily by finalize()
  std::vector<std::vector<int64_t>> rank_ghosts;
  std::vector<std::vector<int64_t>> rank_cells;
  // node -> nodes it is constrained to (union over its 18 DoFs); empty for regular / Dirichlet nodes
  std::vector<int64_t> nm_ptr, nm_node;
};

} // namespace vhhost
#endif
