// See mesh.h.  Stand-in for the deal.II objects that produce the hot path's tables.
#include "mesh.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>
#include <unordered_map>

namespace vhhost
{
namespace
{
// Local node a -> tensor index t in {0..degree}^3.  deal.II FE_Q order: Q1 = vertices, lexicographic with x
// fastest; Q2 = vertices(8), lines(12), quads(6), hex(1) in GeometryInfo<3> numbering.
const int Q2_T[27][3] = {{0, 0, 0}, {2, 0, 0}, {0, 2, 0}, {2, 2, 0}, {0, 0, 2}, {2, 0, 2}, {0, 2, 2}, {2, 2, 2},
                         {0, 1, 0}, {2, 1, 0}, {1, 0, 0}, {1, 2, 0}, {0, 1, 2}, {2, 1, 2}, {1, 0, 2}, {1, 2, 2},
                         {0, 0, 1}, {2, 0, 1}, {0, 2, 1}, {2, 2, 1}, {0, 1, 1}, {2, 1, 1}, {1, 0, 1}, {1, 2, 1},
                         {1, 1, 0}, {1, 1, 2}, {1, 1, 1}};

inline void node_t(int degree, int a, int t[3])
{
  if (degree == 1)
    {
      t[0] = a & 1;
      t[1] = (a >> 1) & 1;
      t[2] = (a >> 2) & 1;
    }
  else
    {
      t[0] = Q2_T[a][0];
      t[1] = Q2_T[a][1];
      t[2] = Q2_T[a][2];
    }
}

inline double lagrange(int degree, int t, double xi)
{
  if (degree == 1)
    return t ? xi : 1.0 - xi;
  if (t == 0)
    return (2 * xi - 1) * (xi - 1);
  if (t == 1)
    return 4 * xi * (1 - xi);
  return xi * (2 * xi - 1);
}

inline uint64_t spread3(uint32_t v)
{ // interleave 21 bits with two zero bits between each
  uint64_t x = v & 0x1fffff;
  x          = (x | x << 32) & 0x1f00000000ffffULL;
  x          = (x | x << 16) & 0x1f0000ff0000ffULL;
  x          = (x | x << 8) & 0x100f00f00f00f00fULL;
  x          = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x          = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}
} // namespace

Mesh::Mesh(int degree_, const double lo_[3], const double hi_[3], const int base_[3], const int face_bid[6], int n_global_refine)
  : degree(degree_)
{
  if (degree != 1 && degree != 2)
    throw std::invalid_argument("Mesh: degree must be 1 or 2");
  for (int d = 0; d < 3; ++d)
    {
      lo[d]   = lo_[d];
      hi[d]   = hi_[d];
      base[d] = base_[d];
      if (base[d] < 1 || !(hi[d] > lo[d]))
        throw std::invalid_argument("Mesh: bad box");
    }
  for (int f = 0; f < 6; ++f)
    bid[f] = face_bid[f];
  for (int z = 0; z < base[2]; ++z)
    for (int y = 0; y < base[1]; ++y)
      for (int x = 0; x < base[0]; ++x)
        leaves.push_back(Leaf{0, {x, y, z}});
  refine_global(n_global_refine);
}

void Mesh::sort_leaves()
{
  Lmax = 0;
  for (const Leaf &l : leaves)
    Lmax = std::max(Lmax, l.level);
  if (Lmax > 20)
    throw std::runtime_error("Mesh: refinement too deep");
  const int L = Lmax;
  auto      key = [&](const Leaf &l, uint64_t &root, uint64_t &mort) {
    const int sh  = L - l.level;
    uint32_t  r[3], m[3];
    for (int d = 0; d < 3; ++d)
      {
        r[d] = (uint32_t)(l.g[d] >> l.level);
        m[d] = (uint32_t)((l.g[d] - ((int32_t)r[d] << l.level)) << sh);
      }
    root = r[0] + (uint64_t)base[0] * (r[1] + (uint64_t)base[1] * r[2]);
    mort = spread3(m[0]) | (spread3(m[1]) << 1) | (spread3(m[2]) << 2);
  };
  std::vector<std::pair<std::pair<uint64_t, uint64_t>, Leaf>> tmp;
  tmp.reserve(leaves.size());
  for (const Leaf &l : leaves)
    {
      uint64_t r, m;
      key(l, r, m);
      tmp.push_back({{r, m}, l});
    }
  std::sort(tmp.begin(), tmp.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
  for (size_t i = 0; i < tmp.size(); ++i)
    leaves[i] = tmp[i].second;
}

void Mesh::refine_global(int times)
{
  for (int t = 0; t < times; ++t)
    {
      std::vector<Leaf> next;
      next.reserve(leaves.size() * 8);
      for (const Leaf &l : leaves)
        for (int c = 0; c < 8; ++c)
          next.push_back(Leaf{l.level + 1, {2 * l.g[0] + (c & 1), 2 * l.g[1] + ((c >> 1) & 1), 2 * l.g[2] + ((c >> 2) & 1)}});
      leaves.swap(next);
    }
  sort_leaves();
  n_ranks = 0;
}

void Mesh::cell_box(int64_t e, double origin[3], double h[3]) const
{
  const Leaf &l = leaves[e];
  for (int d = 0; d < 3; ++d)
    {
      h[d]      = (hi[d] - lo[d]) / ((double)base[d] * (double)(1 << l.level));
      origin[d] = lo[d] + h[d] * l.g[d];
    }
}

void Mesh::cell_center(int64_t e, double c[3]) const
{
  double o[3], h[3];
  cell_box(e, o, h);
  for (int d = 0; d < 3; ++d)
    c[d] = o[d] + 0.5 * h[d];
}

void Mesh::refine(const std::vector<uint8_t> &flags_in)
{
  if ((int64_t)flags_in.size() != n_cells())
    throw std::invalid_argument("Mesh::refine: flags size");
  std::vector<uint8_t> flags(flags_in);
  // 2:1 balance over faces and edges: a leaf that will be at level l+1 forces every face/edge neighbour to level >= l.
  for (;;)
    {
      // map (level, g) -> leaf index
      std::unordered_map<uint64_t, int64_t> where;
      where.reserve(leaves.size() * 2);
      auto pack = [](int level, const int32_t g[3]) {
        return ((uint64_t)level << 58) | ((uint64_t)(uint32_t)g[0] << 38) | ((uint64_t)(uint32_t)g[1] << 19) | (uint64_t)(uint32_t)g[2];
      };
      for (int64_t e = 0; e < n_cells(); ++e)
        where[pack(leaves[e].level, leaves[e].g)] = e;
      bool changed = false;
      for (int64_t e = 0; e < n_cells(); ++e)
        {
          if (!flags[e])
            continue;
          const Leaf &l = leaves[e];
          for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
              for (int dx = -1; dx <= 1; ++dx)
                {
                  const int nz = (dx != 0) + (dy != 0) + (dz != 0);
                  if (nz == 0 || nz == 3)
                    continue; // faces and edges only
                  int32_t g[3] = {l.g[0] + dx, l.g[1] + dy, l.g[2] + dz};
                  bool    out  = false;
                  for (int d = 0; d < 3; ++d)
                    if (g[d] < 0 || g[d] >= (base[d] << l.level))
                      {
                        // across a periodic pair the neighbour is the cell at the opposite face (p4est does the same
                        // once add_periodicity has been called, makegrid_retangle-z-AdGR_xy-periodic.cc:204-219)
                        if (bid[2 * d] == 5 + 2 * d && bid[2 * d + 1] == 6 + 2 * d)
                          g[d] = g[d] < 0 ? (base[d] << l.level) - 1 : 0;
                        else
                          out = true;
                      }
                  if (out)
                    continue;
                  // find the leaf covering that same-level position: it is at level <= l.level (or finer: then fine)
                  for (int lev = l.level; lev >= 0; --lev)
                    {
                      int32_t gg[3] = {g[0] >> (l.level - lev), g[1] >> (l.level - lev), g[2] >> (l.level - lev)};
                      auto    it    = where.find(pack(lev, gg));
                      if (it != where.end())
                        {
                          if (lev < l.level && !flags[it->second])
                            {
                              flags[it->second] = 1;
                              changed           = true;
                            }
                          break;
                        }
                    }
                }
        }
      if (!changed)
        break;
    }
  std::vector<Leaf> next;
  next.reserve(leaves.size() * 2);
  for (int64_t e = 0; e < n_cells(); ++e)
    {
      const Leaf &l = leaves[e];
      if (!flags[e])
        next.push_back(l);
      else
        for (int c = 0; c < 8; ++c)
          next.push_back(Leaf{l.level + 1, {2 * l.g[0] + (c & 1), 2 * l.g[1] + ((c >> 1) & 1), 2 * l.g[2] + ((c >> 2) & 1)}});
    }
  leaves.swap(next);
  sort_leaves();
  n_ranks = 0;
}

void Mesh::finalize(int n_ranks_)
{
  if (n_ranks_ < 1)
    throw std::invalid_argument("Mesh::finalize: n_ranks");
  n_ranks         = n_ranks_;
  const int     n = degree == 1 ? 8 : 27;
  const int64_t nc = n_cells();
  // lattice: one root cell side = U units; a level-l cell has node spacing 2^(Lmax-l)... times 1 for both degrees
  // because U = degree << Lmax.  We use an extra factor 2 so that half-spacings (hanging-node search) stay integral.
  const int64_t U = ((int64_t)degree << Lmax) * 2;
  const int64_t D[3] = {base[0] * U + 1, base[1] * U + 1, base[2] * U + 1};
  auto          keyof = [&](const int64_t X[3]) { return (uint64_t)(X[0] + D[0] * (X[1] + D[1] * X[2])); };

  // ---- partition of the Morton cell sequence (p4est: equal counts, contiguous) ----
  cell_rank.assign(nc, 0);
  for (int r = 0; r < n_ranks; ++r)
    for (int64_t e = nc * r / n_ranks; e < nc * (r + 1) / n_ranks; ++e)
      cell_rank[e] = r;

  // ---- first-touch node numbering in cell order (as DoFHandler::distribute_dofs walks cells) ----
  std::unordered_map<uint64_t, int64_t> node_of;
  node_of.reserve((size_t)nc * (degree == 1 ? 2 : 9));
  std::vector<int64_t> tmp_cell_nodes((size_t)nc * n);
  std::vector<int64_t> node_X; // lattice coords per node
  std::vector<int>     owner;
  for (int64_t e = 0; e < nc; ++e)
    {
      const Leaf   &l  = leaves[e];
      const int64_t cs = U >> l.level;  // cell side in lattice units
      const int64_t ns = cs / degree;   // node spacing
      for (int a = 0; a < n; ++a)
        {
          int t[3];
          node_t(degree, a, t);
          int64_t X[3];
          for (int d = 0; d < 3; ++d)
            X[d] = l.g[d] * cs + t[d] * ns;
          const uint64_t k  = keyof(X);
          auto           it = node_of.find(k);
          int64_t        id;
          if (it == node_of.end())
            {
              id         = (int64_t)owner.size();
              node_of[k] = id;
              owner.push_back(cell_rank[e]);
              node_X.insert(node_X.end(), X, X + 3);
            }
          else
            {
              id        = it->second;
              owner[id] = std::min(owner[id], cell_rank[e]); // interface DoFs belong to the lower subdomain id
            }
          tmp_cell_nodes[(size_t)e * n + a] = id;
        }
    }
  n_nodes = (int64_t)owner.size();

  // ---- rank-major renumbering, stable inside a rank ----
  std::vector<int64_t> perm(n_nodes), newid(n_nodes);
  for (int64_t i = 0; i < n_nodes; ++i)
    perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return owner[a] < owner[b]; });
  for (int64_t i = 0; i < n_nodes; ++i)
    newid[perm[i]] = i;
  node_rank.assign(n_nodes, 0);
  node_xyz.assign((size_t)n_nodes * 3, 0.0);
  std::vector<int64_t> nX((size_t)n_nodes * 3);
  for (int64_t i = 0; i < n_nodes; ++i)
    {
      const int64_t j = newid[i];
      node_rank[j]    = owner[i];
      for (int d = 0; d < 3; ++d)
        {
          nX[(size_t)j * 3 + d]       = node_X[(size_t)i * 3 + d];
          node_xyz[(size_t)j * 3 + d] = lo[d] + (hi[d] - lo[d]) * ((double)node_X[(size_t)i * 3 + d] / (double)(base[d] * U));
        }
    }
  node_lattice = nX;
  lattice_U    = U;
  for (auto &kv : node_of)
    kv.second = newid[kv.second];
  cell_nodes.resize(tmp_cell_nodes.size());
  for (size_t i = 0; i < tmp_cell_nodes.size(); ++i)
    cell_nodes[i] = newid[tmp_cell_nodes[i]];
  rank_node_begin.assign(n_ranks + 1, 0);
  for (int64_t i = 0; i < n_nodes; ++i)
    rank_node_begin[node_rank[i] + 1]++;
  for (int r = 0; r < n_ranks; ++r)
    rank_node_begin[r + 1] += rank_node_begin[r];

  // ---- boundary ids per cell face ----
  cell_face_bid.assign((size_t)nc * 6, 0);
  for (int64_t e = 0; e < nc; ++e)
    {
      const Leaf &l = leaves[e];
      for (int d = 0; d < 3; ++d)
        {
          if (l.g[d] == 0)
            cell_face_bid[(size_t)e * 6 + 2 * d] = (int8_t)bid[2 * d];
          if (l.g[d] == (base[d] << l.level) - 1)
            cell_face_bid[(size_t)e * 6 + 2 * d + 1] = (int8_t)bid[2 * d + 1];
        }
    }

  // ---- constraints: hanging nodes first, then masked Dirichlet, then close (setup_uniform_B-phase.cc:135-186) ----
  // node-level hanging constraints (identical for the 18 components of FESystem(FE_Q,18))
  std::map<int64_t, std::vector<std::pair<int64_t, double>>> hang;
  for (int64_t e = 0; e < nc; ++e)
    {
      const Leaf &l = leaves[e];
      if (l.level == Lmax)
        continue;
      const int64_t cs = U >> l.level, half = cs / (2 * degree);
      const int     m = 2 * degree; // sub-lattice index range 0..m per dimension
      for (int iz = 0; iz <= m; ++iz)
        for (int iy = 0; iy <= m; ++iy)
          for (int ix = 0; ix <= m; ++ix)
            {
              const int i3[3] = {ix, iy, iz};
              if (!(ix == 0 || ix == m || iy == 0 || iy == m || iz == 0 || iz == m))
                continue; // interior of the cell
              if (!((ix | iy | iz) & 1))
                continue; // one of this cell's own nodes
              int64_t X[3];
              for (int d = 0; d < 3; ++d)
                X[d] = l.g[d] * cs + i3[d] * half;
              auto it = node_of.find(keyof(X));
              if (it == node_of.end())
                continue;
              const int64_t hn = it->second;
              if (hang.count(hn))
                continue;
              // value = coarse cell's FE function at this point
              std::vector<std::pair<int64_t, double>> ent;
              for (int a = 0; a < n; ++a)
                {
                  int t[3];
                  node_t(degree, a, t);
                  double w = 1.0;
                  for (int d = 0; d < 3; ++d)
                    w *= lagrange(degree, t[d], (double)i3[d] / (double)m);
                  if (std::fabs(w) > 1e-14)
                    ent.push_back({cell_nodes[(size_t)e * n + a], w});
                }
              hang[hn] = ent;
            }
    }
  n_hanging_nodes = (int64_t)hang.size();
  // ---- periodic pairs (makegrid_retangle-z-AdGR_xy-periodic.cc:167-219, setup_weak-coupling-PDW-configuration.cc:128-206):
  // boundary ids (5,6) / (7,8) / (9,10) on the lower/upper x / y / z face declare that direction periodic.
  // DoFTools::make_periodicity_constraints(b_id1 = lower, b_id2 = upper) makes every DoF of the upper face an identity-
  // constrained DoF of its image on the lower face, all 18 components alike [deal.II-internal: which of the two faces
  // keeps its DoFs is restated from the deal.II 9.3 documentation; it changes the numbering, not the discrete problem].
  // DoFs that are constrained already (hanging nodes) are left alone, as in deal.II; chains (corner/edge nodes periodic in
  // two directions, hanging-node masters on the upper face) are resolved by the closing loop below.
  n_periodic_nodes = 0;
  for (int d = 0; d < 3; ++d)
    {
      if (!(bid[2 * d] == 5 + 2 * d && bid[2 * d + 1] == 6 + 2 * d))
        {
          if ((bid[2 * d] >= 5 && bid[2 * d] <= 10) || (bid[2 * d + 1] >= 5 && bid[2 * d + 1] <= 10))
            throw std::invalid_argument("Mesh: periodic boundary ids must come as the pairs (5,6) on x, (7,8) on y, (9,10) on z");
          continue;
        }
      int min_level = Lmax;
      for (const Leaf &l : leaves)
        min_level = std::min(min_level, l.level);
      if (((int64_t)base[d] << min_level) < 2)
        throw std::invalid_argument("Mesh: a periodic direction needs at least two cells across");
      // (a) the two faces are refined differently somewhere (2:1 after refine()): the nodes of the finer side that have no
      //     counterpart hang on the coarse cell across the seam, exactly like hanging nodes on an interior face
      //     (deal.II: make_periodicity_constraints recurses into the children of the refined face)
      for (int64_t e = 0; e < nc; ++e)
        {
          const Leaf &l = leaves[e];
          if (l.level == Lmax)
            continue;
          const int64_t cs = U >> l.level, half = cs / (2 * degree);
          const int     m = 2 * degree;
          for (int side = 0; side < 2; ++side)
            {
              if (l.g[d] != (side ? (base[d] << l.level) - 1 : 0))
                continue;
              const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
              for (int i2 = 0; i2 <= m; ++i2)
                for (int i1 = 0; i1 <= m; ++i1)
                  {
                    if (!((i1 | i2) & 1))
                      continue; // the image of one of this cell's own nodes: identity, see (b)
                    int i3[3];
                    i3[d]  = side ? m : 0;
                    i3[d1] = i1;
                    i3[d2] = i2;
                    int64_t X[3];
                    for (int k = 0; k < 3; ++k)
                      X[k] = l.g[k] * cs + i3[k] * half;
                    X[d]    = side ? 0 : D[d] - 1; // the same point seen from the other side of the seam
                    auto it = node_of.find(keyof(X));
                    if (it == node_of.end() || hang.count(it->second))
                      continue;
                    std::vector<std::pair<int64_t, double>> ent;
                    for (int a = 0; a < n; ++a)
                      {
                        int t[3];
                        node_t(degree, a, t);
                        double w = 1.0;
                        for (int k = 0; k < 3; ++k)
                          w *= lagrange(degree, t[k], (double)i3[k] / (double)m);
                        if (std::fabs(w) > 1e-14)
                          ent.push_back({cell_nodes[(size_t)e * n + a], w});
                      }
                    hang[it->second] = ent;
                    ++n_hanging_nodes;
                  }
            }
        }
      // (b) identities: a node of the upper face equals its image on the lower face
      for (int64_t nd = 0; nd < n_nodes; ++nd)
        {
          if (nX[(size_t)nd * 3 + d] != D[d] - 1 || hang.count(nd))
            continue;
          int64_t X[3] = {nX[(size_t)nd * 3], nX[(size_t)nd * 3 + 1], nX[(size_t)nd * 3 + 2]};
          X[d]         = 0;
          auto it      = node_of.find(keyof(X));
          if (it == node_of.end())
            throw std::runtime_error("Mesh: a node of a periodic face has neither an image nor a coarse cell across the seam "
                                     "(the periodic faces are not 2:1 balanced; refine through Mesh::refine)");
          hang[nd] = {{it->second, 1.0}};
          ++n_periodic_nodes;
        }
    }
  // resolve chains among constrained nodes (a master that is itself hanging or periodic)
  for (int guard = 0; guard < 8; ++guard)
    {
      bool again = false;
      for (auto &kv : hang)
        {
          std::map<int64_t, double> acc;
          bool                      sub = false;
          for (auto &mw : kv.second)
            {
              auto it = hang.find(mw.first);
              if (it == hang.end())
                acc[mw.first] += mw.second;
              else
                {
                  sub = true;
                  for (auto &mw2 : it->second)
                    acc[mw2.first] += mw.second * mw2.second;
                }
            }
          if (sub)
            {
              kv.second.assign(acc.begin(), acc.end());
              again = true;
            }
        }
      if (!again)
        break;
    }
  // Dirichlet masks per node: bit c set <=> DoF (node,c) is a homogeneous Dirichlet DoF.  Boundary id 2|3|4
  // constrains the components whose orbital index equals the wall normal (femgl.h:294-301).
  std::vector<uint32_t> dmask(n_nodes, 0u);
  for (int64_t e = 0; e < nc; ++e)
    for (int f = 0; f < 6; ++f)
      {
        const int b = cell_face_bid[(size_t)e * 6 + f];
        if (b < 2 || b > 4)
          continue;
        const int nd = f / 2, side = f % 2, normal = b - 2;
        uint32_t  m = 0;
        for (int c = 0; c < 18; ++c)
          if (c % 3 == normal)
            m |= 1u << c;
        for (int a = 0; a < n; ++a)
          {
            int t[3];
            node_t(degree, a, t);
            if (t[nd] == (side ? degree : 0))
              dmask[cell_nodes[(size_t)e * n + a]] |= m;
          }
      }
  // assemble closed DoF-level table
  c_dof.clear();
  c_ptr.assign(1, 0);
  c_master.clear();
  c_weight.clear();
  for (int64_t nd = 0; nd < n_nodes; ++nd)
    {
      auto hit = hang.find(nd);
      if (hit == hang.end() && dmask[nd] == 0)
        continue;
      for (int c = 0; c < 18; ++c)
        {
          if (hit != hang.end())
            { // hanging constraints take precedence (interpolate_boundary_values skips constrained DoFs)
              c_dof.push_back(18 * nd + c);
              for (auto &mw : hit->second)
                {
                  // closing: masters that are Dirichlet DoFs drop out (their value is 0)
                  if (hang.find(mw.first) == hang.end() && (dmask[mw.first] >> c) & 1u)
                    continue;
                  c_master.push_back(18 * mw.first + c);
                  c_weight.push_back(mw.second);
                }
              c_ptr.push_back((int64_t)c_master.size());
            }
          else if ((dmask[nd] >> c) & 1u)
            {
              c_dof.push_back(18 * nd + c);
              c_ptr.push_back((int64_t)c_master.size());
            }
        }
    }
  // node -> master nodes
  nm_ptr.assign(n_nodes + 1, 0);
  nm_node.clear();
  {
    std::vector<std::vector<int64_t>> tmp;
    for (int64_t nd = 0; nd < n_nodes; ++nd)
      {
        auto hit = hang.find(nd);
        if (hit != hang.end())
          for (auto &mw : hit->second)
            nm_node.push_back(mw.first);
        nm_ptr[nd + 1] = (int64_t)nm_node.size();
      }
  }

  // ---- per-rank cell and ghost lists ----
  rank_cells.assign(n_ranks, {});
  rank_ghosts.assign(n_ranks, {});
  {
    std::vector<int> seen(n_nodes, -1);
    for (int r = 0; r < n_ranks; ++r)
      {
        std::vector<int64_t> &cells = rank_cells[r];
        std::vector<int64_t> &gh    = rank_ghosts[r];
        for (int64_t e = 0; e < nc; ++e)
          {
            bool need = false;
            for (int a = 0; a < n && !need; ++a)
              {
                const int64_t nd = cell_nodes[(size_t)e * n + a];
                if (node_rank[nd] == r)
                  need = true;
                for (int64_t p = nm_ptr[nd]; p < nm_ptr[nd + 1] && !need; ++p)
                  if (node_rank[nm_node[p]] == r)
                    need = true;
              }
            if (!need)
              continue;
            cells.push_back(e);
            for (int a = 0; a < n; ++a)
              {
                const int64_t nd = cell_nodes[(size_t)e * n + a];
                if (node_rank[nd] != r && seen[nd] != r)
                  {
                    seen[nd] = r;
                    gh.push_back(nd);
                  }
                for (int64_t p = nm_ptr[nd]; p < nm_ptr[nd + 1]; ++p)
                  {
                    const int64_t m = nm_node[p];
                    if (node_rank[m] != r && seen[m] != r)
                      {
                        seen[m] = r;
                        gh.push_back(m);
                      }
                  }
              }
          }
        std::sort(gh.begin(), gh.end()); // rank-major global numbering => sorted by (owner, id)
      }
  }
}

namespace
{
inline uint64_t leaf_key(int level, const int64_t g[3])
{
  return ((uint64_t)level << 58) | ((uint64_t)g[0] << 38) | ((uint64_t)g[1] << 19) | (uint64_t)g[2];
}
} // namespace

void Mesh::transfer_table(const Mesh &old_mesh, std::vector<int32_t> &ptr, std::vector<int32_t> &src, std::vector<double> &weight) const
{
  if (n_ranks < 1 || old_mesh.n_ranks < 1)
    throw std::runtime_error("Mesh::transfer_table: both meshes must be finalized");
  if (old_mesh.degree != degree)
    throw std::invalid_argument("Mesh::transfer_table: incompatible meshes");
  const int n = degree == 1 ? 8 : 27;
  std::unordered_map<uint64_t, int64_t> where;
  where.reserve(old_mesh.leaves.size() * 2);
  for (int64_t e = 0; e < old_mesh.n_cells(); ++e)
    {
      const Leaf   &l    = old_mesh.leaves[e];
      const int64_t g[3] = {l.g[0], l.g[1], l.g[2]};
      where[leaf_key(l.level, g)] = e;
    }
  ptr.assign(1, 0);
  src.clear();
  weight.clear();
  for (int64_t nd = 0; nd < n_nodes; ++nd)
    {
      const int64_t *X = &node_lattice[(size_t)3 * nd];
      int64_t        cell = -1;
      double         xi[3] = {0, 0, 0};
      for (int lev = old_mesh.Lmax; lev >= 0 && cell < 0; --lev)
        {
          const int64_t cs = lattice_U >> lev;
          int64_t       g[3];
          for (int d = 0; d < 3; ++d)
            {
              g[d] = X[d] / cs;
              if (g[d] > ((int64_t)base[d] << lev) - 1)
                g[d] = ((int64_t)base[d] << lev) - 1;
              xi[d] = (double)(X[d] - g[d] * cs) / (double)cs;
            }
          auto it = where.find(leaf_key(lev, g));
          if (it != where.end())
            cell = it->second;
        }
      if (cell < 0)
        throw std::runtime_error("Mesh::transfer_table: node outside the old mesh");
      for (int a = 0; a < n; ++a)
        {
          int t[3];
          node_t(degree, a, t);
          const double w = lagrange(degree, t[0], xi[0]) * lagrange(degree, t[1], xi[1]) * lagrange(degree, t[2], xi[2]);
          if (w == 0.0)
            continue;
          src.push_back((int32_t)old_mesh.cell_nodes[(size_t)cell * n + a]);
          weight.push_back(w);
        }
      ptr.push_back((int32_t)src.size());
    }
}

void Mesh::interpolate_from(const Mesh &old_mesh, const std::vector<double> &old_values, std::vector<double> &new_values) const
{
  if ((int64_t)old_values.size() != 18 * old_mesh.n_nodes)
    throw std::invalid_argument("Mesh::interpolate_from: incompatible meshes / value array");
  std::vector<int32_t> ptr, src;
  std::vector<double>  w;
  transfer_table(old_mesh, ptr, src, w);
  new_values.assign((size_t)18 * n_nodes, 0.0);
  for (int64_t nd = 0; nd < n_nodes; ++nd)
    for (int k = ptr[nd]; k < ptr[nd + 1]; ++k)
      for (int c = 0; c < 18; ++c)
        new_values[(size_t)18 * nd + c] += w[k] * old_values[(size_t)18 * src[k] + c];
}

void Mesh::kelly_indicator(const std::vector<double> &values, std::vector<double> &eta) const
{
  if (n_ranks < 1 || (int64_t)values.size() != 18 * n_nodes)
    throw std::invalid_argument("Mesh::kelly_indicator: mesh not finalized or value array of the wrong size");
  const int     n  = degree == 1 ? 8 : 27;
  const int64_t nc = n_cells();
  std::unordered_map<uint64_t, int64_t> where;
  where.reserve(leaves.size() * 2);
  for (int64_t e = 0; e < nc; ++e)
    {
      const int64_t g[3] = {leaves[e].g[0], leaves[e].g[1], leaves[e].g[2]};
      where[leaf_key(leaves[e].level, g)] = e;
    }
  // cell-mean gradient from the 8 vertex values (the first 8 local nodes are the vertices for Q1 and Q2)
  std::vector<double> grad((size_t)nc * 54), hh((size_t)nc * 3);
  for (int64_t e = 0; e < nc; ++e)
    {
      double o[3];
      cell_box(e, o, &hh[3 * (size_t)e]);
      for (int c = 0; c < 18; ++c)
        for (int d = 0; d < 3; ++d)
          {
            double s = 0.0;
            for (int v = 0; v < 8; ++v)
              s += (((v >> d) & 1) ? 1.0 : -1.0) * values[(size_t)18 * cell_nodes[(size_t)e * n + v] + c];
            grad[(size_t)e * 54 + 3 * c + d] = 0.25 * s / hh[3 * (size_t)e + d];
          }
    }
  // integer coordinates in units of the finest half-cell: a level-l cell spans 2 << (Lmax - l) units per direction
  eta.assign((size_t)nc, 0.0);
  for (int64_t e = 0; e < nc; ++e)
    {
      const Leaf   &l  = leaves[e];
      const int64_t cs = (int64_t)2 << (Lmax - l.level);
      const double *h  = &hh[3 * (size_t)e];
      const double  hK = std::sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]);
      double        s  = 0.0;
      for (int f = 0; f < 6; ++f)
        {
          const int    d = f / 2, side = f % 2, d0 = (d + 1) % 3, d1 = (d + 2) % 3;
          const double area = h[d0] * h[d1];
          for (int q = 0; q < 4; ++q)
            { // a point just across the face, at the centre of quarter q of the face
              int64_t X[3];
              X[d]  = l.g[d] * cs + (side ? cs : -1);
              X[d0] = l.g[d0] * cs + ((q & 1) ? (3 * cs) / 4 : cs / 4);
              X[d1] = l.g[d1] * cs + ((q & 2) ? (3 * cs) / 4 : cs / 4);
              const int64_t ext = ((int64_t)base[d] << Lmax) * 2;
              if (X[d] < 0 || X[d] >= ext)
                {
                  const int b = bid[f];
                  if (b < 5 || b > 10)
                    break; // a physical boundary face: no contribution (no Neumann data, refine.cc:146)
                  X[d] = X[d] < 0 ? ext - 1 : 0; // periodic pair: continue on the opposite face
                }
              int64_t nb = -1;
              for (int lev = Lmax; lev >= 0 && nb < 0; --lev)
                {
                  const int64_t c2   = (int64_t)2 << (Lmax - lev);
                  const int64_t g[3] = {X[0] / c2, X[1] / c2, X[2] / c2};
                  auto          it   = where.find(leaf_key(lev, g));
                  if (it != where.end())
                    nb = it->second;
                }
              if (nb < 0)
                continue;
              double j2 = 0.0;
              for (int c = 0; c < 18; ++c)
                {
                  const double j = grad[(size_t)nb * 54 + 3 * c + d] - grad[(size_t)e * 54 + 3 * c + d];
                  j2 += j * j;
                }
              s += hK / 24.0 * 0.25 * area * j2;
            }
        }
      eta[e] = std::sqrt(s);
    }
}

RankTables Mesh::tables(int r) const
{
  if (n_ranks < 1)
    throw std::runtime_error("Mesh::tables before finalize");
  if (r < 0 || r >= n_ranks)
    throw std::invalid_argument("Mesh::tables: rank");
  const int  n = degree == 1 ? 8 : 27;
  RankTables T;
  T.degree              = degree;
  const int64_t n0      = rank_node_begin[r], n1 = rank_node_begin[r + 1];
  T.n_owned_nodes       = (int32_t)(n1 - n0);
  const auto &gh        = rank_ghosts[r];
  T.n_ghost_nodes       = (int32_t)gh.size();
  const int64_t n_local = T.n_owned_nodes + (int64_t)gh.size();
  T.node_global.resize(n_local);
  for (int64_t i = 0; i < T.n_owned_nodes; ++i)
    T.node_global[i] = n0 + i;
  for (size_t i = 0; i < gh.size(); ++i)
    T.node_global[T.n_owned_nodes + i] = gh[i];
  auto local_of = [&](int64_t g) -> int32_t {
    if (g >= n0 && g < n1)
      return (int32_t)(g - n0);
    auto it = std::lower_bound(gh.begin(), gh.end(), g);
    if (it == gh.end() || *it != g)
      throw std::runtime_error("Mesh::tables: node not local");
    return (int32_t)(T.n_owned_nodes + (it - gh.begin()));
  };
  T.node_xyz.resize((size_t)n_local * 3);
  for (int64_t i = 0; i < n_local; ++i)
    for (int d = 0; d < 3; ++d)
      T.node_xyz[(size_t)i * 3 + d] = node_xyz[(size_t)T.node_global[i] * 3 + d];

  const auto &cells = rank_cells[r];
  T.n_cells         = (int32_t)cells.size();
  T.cell_nodes.resize(cells.size() * n);
  T.cell_global.assign(cells.begin(), cells.end());
  T.cell_origin.resize(cells.size() * 3);
  T.cell_h.resize(cells.size() * 3);
  T.cell_owned.resize(cells.size());
  for (size_t k = 0; k < cells.size(); ++k)
    {
      const int64_t e = cells[k];
      for (int a = 0; a < n; ++a)
        T.cell_nodes[k * n + a] = local_of(cell_nodes[(size_t)e * n + a]);
      cell_box(e, &T.cell_origin[3 * k], &T.cell_h[3 * k]);
      T.cell_owned[k] = cell_rank[e] == r;
      for (int f = 0; f < 6; ++f)
        {
          const int b = cell_face_bid[(size_t)e * 6 + f];
          if (b >= 2 && b <= 4)
            {
              T.wall_face_cell.push_back((int32_t)k);
              T.wall_face_no.push_back((int8_t)f);
              T.wall_face_bid.push_back((int8_t)b);
            }
        }
    }
  // constraints on local DoFs (owned and ghost), masters are local by construction of the ghost list
  T.c_ptr.push_back(0);
  {
    std::vector<std::pair<int32_t, int64_t>> lines; // (local dof, global line index)
    for (size_t k = 0; k < c_dof.size(); ++k)
      {
        const int64_t nd = c_dof[k] / 18;
        int32_t       ln;
        if (nd >= n0 && nd < n1)
          ln = (int32_t)(nd - n0);
        else
          {
            auto it = std::lower_bound(gh.begin(), gh.end(), nd);
            if (it == gh.end() || *it != nd)
              continue;
            ln = (int32_t)(T.n_owned_nodes + (it - gh.begin()));
          }
        lines.push_back({(int32_t)(18 * ln + c_dof[k] % 18), (int64_t)k});
      }
    std::sort(lines.begin(), lines.end());
    for (auto &ln : lines)
      {
        const int64_t k       = ln.second;
        bool          all_loc = true;
        for (int64_t p = c_ptr[k]; p < c_ptr[k + 1]; ++p)
          {
            const int64_t m = c_master[p] / 18;
            if (!(m >= n0 && m < n1) && !std::binary_search(gh.begin(), gh.end(), m))
              all_loc = false;
          }
        if (!all_loc)
          continue; // a ghost DoF whose masters are not visible here: its value arrives through the halo
        T.c_dof.push_back(ln.first);
        for (int64_t p = c_ptr[k]; p < c_ptr[k + 1]; ++p)
          {
            T.c_master.push_back(18 * local_of(c_master[p] / 18) + (int32_t)(c_master[p] % 18));
            T.c_weight.push_back(c_weight[p]);
          }
        T.c_ptr.push_back((int32_t)T.c_master.size());
      }
  }
  // halo plan: ghosts are sorted by owner, so each peer's receive list is a contiguous run
  T.send_ptr.push_back(0);
  T.recv_ptr.push_back(0);
  for (int p = 0; p < n_ranks; ++p)
    {
      if (p == r)
        continue;
      std::vector<int32_t> recv, send;
      for (size_t i = 0; i < gh.size(); ++i)
        if (node_rank[gh[i]] == p)
          recv.push_back((int32_t)(T.n_owned_nodes + i));
      for (int64_t g : rank_ghosts[p])
        if (g >= n0 && g < n1)
          send.push_back((int32_t)(g - n0));
      if (recv.empty() && send.empty())
        continue;
      T.peer_rank.push_back(p);
      T.send_nodes.insert(T.send_nodes.end(), send.begin(), send.end());
      T.recv_nodes.insert(T.recv_nodes.end(), recv.begin(), recv.end());
      T.send_ptr.push_back((int32_t)T.send_nodes.size());
      T.recv_ptr.push_back((int32_t)T.recv_nodes.size());
    }
  return T;
}

} // namespace vhhost
