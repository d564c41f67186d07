// C entry points over the host-side mesh tables (mesh.h), for ctypes (tests, bench.py) and for C callers.
// Not part of the GPU drop-in boundary (that is include/vh_femgl.h); this is the stand-in for what deal.II
// hands to the adapter.
#include "confreader.h"
#include "matep.h"
#include "mesh.h"
#include "vtu.h"

#include "../../include/vh_femgl.h"

#include <cstring>
#include <string>

using vhhost::Mesh;
using vhhost::RankTables;

static thread_local std::string g_err;

extern "C" {

const char *vhh_last_error() { return g_err.c_str(); }

void *vhh_mesh_create(int degree, const double *lo, const double *hi, const int *base, const int *face_bid, int n_global_refine)
{
  try
    {
      return new Mesh(degree, lo, hi, base, face_bid, n_global_refine);
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return nullptr;
    }
}
void vhh_mesh_free(void *m) { delete static_cast<Mesh *>(m); }
int64_t vhh_mesh_n_cells(void *m) { return static_cast<Mesh *>(m)->n_cells(); }
void vhh_mesh_cell_centers(void *m, double *out)
{
  Mesh *M = static_cast<Mesh *>(m);
  for (int64_t e = 0; e < M->n_cells(); ++e)
    M->cell_center(e, out + 3 * e);
}
int vhh_mesh_refine(void *m, const uint8_t *flags, int64_t n)
{
  try
    {
      static_cast<Mesh *>(m)->refine(std::vector<uint8_t>(flags, flags + n));
      return 0;
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return -1;
    }
}
int vhh_mesh_finalize(void *m, int n_ranks)
{
  try
    {
      static_cast<Mesh *>(m)->finalize(n_ranks);
      return 0;
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return -1;
    }
}
// out[5]: n_nodes, n_cells, n_constraint_lines, n_hanging_nodes, n_periodic_nodes
void vhh_mesh_global_sizes(void *m, int64_t *out)
{
  Mesh *M = static_cast<Mesh *>(m);
  out[0]  = M->n_nodes;
  out[1]  = M->n_cells();
  out[2]  = (int64_t)M->c_dof.size();
  out[3]  = M->n_hanging_nodes;
  out[4]  = M->n_periodic_nodes;
}
void vhh_mesh_rank_node_begin(void *m, int64_t *out)
{
  Mesh *M = static_cast<Mesh *>(m);
  for (size_t i = 0; i < M->rank_node_begin.size(); ++i)
    out[i] = M->rank_node_begin[i];
}

void *vhh_tables_create(void *m, int rank)
{
  try
    {
      return new RankTables(static_cast<Mesh *>(m)->tables(rank));
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return nullptr;
    }
}
void vhh_tables_free(void *t) { delete static_cast<RankTables *>(t); }

// Fill a vh_mesh_desc whose pointers alias the RankTables arrays (valid until vhh_tables_free).
void vhh_tables_desc(void *t, vh_mesh_desc *d)
{
  RankTables *T = static_cast<RankTables *>(t);
  std::memset(d, 0, sizeof(*d));
  d->degree         = T->degree;
  d->n_owned_nodes  = T->n_owned_nodes;
  d->n_ghost_nodes  = T->n_ghost_nodes;
  d->node_global    = T->node_global.data();
  d->n_cells        = T->n_cells;
  d->cell_nodes     = T->cell_nodes.data();
  d->cell_origin    = T->cell_origin.data();
  d->cell_h         = T->cell_h.data();
  d->cell_owned     = T->cell_owned.data();
  d->n_wall_faces   = (int32_t)T->wall_face_cell.size();
  d->wall_face_cell = T->wall_face_cell.data();
  d->wall_face_no   = T->wall_face_no.data();
  d->wall_face_bid  = T->wall_face_bid.data();
  vh_constraints c;
  c.n_lines                    = (int32_t)T->c_dof.size();
  c.dof                        = T->c_dof.data();
  c.ptr                        = T->c_ptr.data();
  c.master                     = T->c_master.data();
  c.weight                     = T->c_weight.data();
  d->constraints_newton_update = c;
  d->constraints_solution      = c;
  d->n_peers                   = (int32_t)T->peer_rank.size();
  d->peer_rank                 = T->peer_rank.data();
  d->send_ptr                  = T->send_ptr.data();
  d->send_nodes                = T->send_nodes.data();
  d->recv_ptr                  = T->recv_ptr.data();
  d->recv_nodes                = T->recv_nodes.data();
}
int vhh_sizeof_mesh_desc() { return (int)sizeof(vh_mesh_desc); }

// Array access by name for numpy views: returns pointer, writes element count and element size.
const void *vhh_tables_array(void *t, const char *name, int64_t *count, int *elem_size)
{
  RankTables *T = static_cast<RankTables *>(t);
  std::string s(name);
#define ARR(field)                                   \
  if (s == #field)                                   \
    {                                                \
      *count     = (int64_t)T->field.size();         \
      *elem_size = (int)sizeof(T->field[0]);         \
      return T->field.data();                        \
    }
  ARR(node_global)
  ARR(node_xyz)
  ARR(cell_nodes)
  ARR(cell_global)
  ARR(cell_origin)
  ARR(cell_h)
  ARR(cell_owned)
  ARR(wall_face_cell)
  ARR(wall_face_no)
  ARR(wall_face_bid)
  ARR(c_dof)
  ARR(c_ptr)
  ARR(c_master)
  ARR(c_weight)
  ARR(peer_rank)
  ARR(send_ptr)
  ARR(send_nodes)
  ARR(recv_ptr)
  ARR(recv_nodes)
#undef ARR
  *count     = -1;
  *elem_size = 0;
  return nullptr;
}
// out: degree, n_owned_nodes, n_ghost_nodes, n_cells
void vhh_tables_sizes(void *t, int64_t *out)
{
  RankTables *T = static_cast<RankTables *>(t);
  out[0]        = T->degree;
  out[1]        = T->n_owned_nodes;
  out[2]        = T->n_ghost_nodes;
  out[3]        = T->n_cells;
}

// ---- Matep restatement: out12 = alpha, beta1..5, gapA, gapB, fA, fB, Tcp_mK, tAB_RWS ----
void vhh_matep(double p, double t, int scc, double *out12)
{
  vhhost::Matep mat;
  bool          key = scc != 0;
  mat.with_SCC(key);
  out12[0]  = mat.alpha_td(t);
  out12[1]  = mat.beta1_td(p, t);
  out12[2]  = mat.beta2_td(p, t);
  out12[3]  = mat.beta3_td(p, t);
  out12[4]  = mat.beta4_td(p, t);
  out12[5]  = mat.beta5_td(p, t);
  out12[6]  = mat.gap_A_td(p, t);
  out12[7]  = mat.gap_B_td(p, t);
  out12[8]  = mat.f_A_td(p, t);
  out12[9]  = mat.f_B_td(p, t);
  out12[10] = mat.Tcp_mK(p);
  out12[11] = mat.tAB_RWS(p);
}

// ---- confreader / ParameterHandler: parse prm text, return "subsection/key=value" lines (newline separated) ----
const char *vhh_prm_dump(const char *prm_text)
{
  static thread_local std::string out;
  out.clear();
  try
    {
      vhhost::ParameterHandler prm;
      vhhost::confreader       cr(prm);
      prm.parse_input_from_string(prm_text ? prm_text : "");
      for (const std::string &k : prm.declared_keys())
        {
          const size_t slash = k.find('/');
          prm.enter_subsection(k.substr(0, slash));
          out += k + "=" + prm.get(k.substr(slash + 1)) + "\n";
          prm.leave_subsection();
        }
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return nullptr;
    }
  return out.c_str();
}

// ---- solution transfer between two finalized meshes (single rank: global node order) ----
int vhh_mesh_interpolate(void *new_mesh, void *old_mesh, const double *old_values, double *new_values)
{
  try
    {
      Mesh               *N = static_cast<Mesh *>(new_mesh), *O = static_cast<Mesh *>(old_mesh);
      std::vector<double> ov(old_values, old_values + 18 * O->n_nodes), nv;
      N->interpolate_from(*O, ov, nv);
      std::memcpy(new_values, nv.data(), nv.size() * sizeof(double));
      return 0;
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return -1;
    }
}
// transfer table of Mesh::transfer_table (single rank: global == local node ids).  ptr[n_nodes+1]; src / weight must hold
// n_nodes * (8 | 27) entries; returns the number of entries or -1.
int64_t vhh_mesh_transfer_table(void *new_mesh, void *old_mesh, int32_t *ptr, int32_t *src, double *weight)
{
  try
    {
      Mesh                *N = static_cast<Mesh *>(new_mesh), *O = static_cast<Mesh *>(old_mesh);
      std::vector<int32_t> p, s;
      std::vector<double>  w;
      N->transfer_table(*O, p, s, w);
      std::memcpy(ptr, p.data(), p.size() * sizeof(int32_t));
      std::memcpy(src, s.data(), s.size() * sizeof(int32_t));
      std::memcpy(weight, w.data(), w.size() * sizeof(double));
      return (int64_t)s.size();
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return -1;
    }
}
// Kelly-type refinement indicator of Mesh::kelly_indicator; values in global node order, eta[n_cells]
int vhh_mesh_kelly(void *mesh, const double *values, double *eta)
{
  try
    {
      Mesh               *M = static_cast<Mesh *>(mesh);
      std::vector<double> v(values, values + 18 * M->n_nodes), e;
      M->kelly_indicator(v, e);
      std::memcpy(eta, e.data(), e.size() * sizeof(double));
      return 0;
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return -1;
    }
}
void *vhh_mesh_clone(void *m) { return new Mesh(*static_cast<Mesh *>(m)); }
void  vhh_mesh_node_xyz(void *m, double *out)
{
  Mesh *M = static_cast<Mesh *>(m);
  std::memcpy(out, M->node_xyz.data(), M->node_xyz.size() * sizeof(double));
}


// DataOut stand-in (host/vtu.cc): piece of `rank` (of n_ranks) from tables `t`; returns 0 / -1 (vhh_last_error)
int vhh_write_vtu(void *t, int rank, int n_ranks, const char *dir, int counter, const double *solution_local, const double *update_local)
{
  try
    {
      write_vtu_piece(*static_cast<RankTables *>(t), rank, n_ranks, dir, "solution", counter, solution_local, update_local);
      return 0;
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return -1;
    }
}

} // extern "C"
