// TEST INFRASTRUCTURE: the product's pointwise kernel family (csrc/vh_points_kernel.cuh) compiled by g++ and executed
// lane by lane under tests/native/cuda_emu.h, so that kernel logic written without GPU access (the matrix-free and
// table-free operator apply) is checked against the oracle in the CPU-only container.  Never part of the product.
#include <vector>
#include <cuda_runtime.h> // vector types (double2) and the __global__ / __launch_bounds__ macros for a host compiler

#include <algorithm>
#include <cmath>
using std::min;

#include "../../verkko-hem-repo_b200/csrc/vh_internal.h"

#include "cuda_emu.h"

#include "../../verkko-hem-repo_b200/csrc/vh_points_kernel.cuh"
#include "../../verkko-hem-repo_b200/csrc/vh_diag_kernel.cuh"
#include "../../verkko-hem-repo_b200/csrc/vh_apply_v2.cuh"
#include "../../verkko-hem-repo_b200/csrc/vh_gather_kernels.cuh"
#include "../../verkko-hem-repo_b200/csrc/vh_block_invert.cuh"

// mode 0: assembly (WANT_H: writes Hq, Rc = -cell residual, Dc, avgD)      x = Newton state
// mode 1: residual only (Rc)                                                x = trial state
// mode 2: operator apply from the H_q tables (Rc = K_cell z_cell)           x = z (masked Krylov vector), Hq from mode 0
// mode 3: table-free operator apply                                         x = z, x_state = Newton state
// mode 4: energy (Ec)
extern "C" int vht_points_emulated(int degree, int mode, int n_cells, const int32_t *cell_nodes, const double *cell_h4,
                                   const uint32_t *cell_faces, const uint8_t *cell_owned, const double *x, const double *x_state,
                                   const double *N, const double *dN, const double *wq, const double *Gref, const double *Mf,
                                   const double *coef10, double *Hq, double *Rc, double *Dc, double *avgD, double *Ec)
{
  VhTables tab{};
  tab.degree = degree;
  tab.nn = tab.nq = degree == 1 ? 8 : 27;
  tab.N    = const_cast<double *>(N);
  tab.dN   = const_cast<double *>(dN);
  tab.wq   = const_cast<double *>(wq);
  tab.Gref = const_cast<double *>(Gref);
  tab.Mf   = const_cast<double *>(Mf);
  VhCoef cf{};
  cf.K1    = coef10[0];
  cf.K23   = coef10[1] + coef10[2];
  cf.alpha = coef10[3];
  for (int k = 0; k < 5; ++k)
    cf.beta[k] = coef10[4 + k];
  cf.bt = coef10[9];
  const vh_hweights hw = vh_make_hweights(cf.alpha, cf.beta);
  try
    {
      if (degree == 1)
        {
          const unsigned grid = (unsigned)((n_cells + 4 * VH_PT_WARPS - 1) / (4 * VH_PT_WARPS));
          const size_t   sm = mode == 3 ? VhPt<8>::SMEM_TFREE : VhPt<8>::SMEM;
          auto           go = [&](auto kernel) { emu::launch(grid, VH_PT_WARPS * 32, sm, kernel); };
          if (mode == 0)
            go([&] { k_points<8, true, false>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec); });
          else if (mode == 1)
            go([&] { k_points<8, false, false>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec); });
          else if (mode == 2)
            go([&] { k_points<8, false, false, true>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec); });
          else if (mode == 3)
            go([&] { k_points<8, false, false, true, true>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec, x_state); });
          else
            go([&] { k_points<8, false, true>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec); });
        }
      else
        {
          const unsigned grid = (unsigned)((n_cells + VH_PT_WARPS - 1) / VH_PT_WARPS);
          const size_t   sm = mode == 3 ? VhPt<27>::SMEM_TFREE : VhPt<27>::SMEM;
          auto           go = [&](auto kernel) { emu::launch(grid, VH_PT_WARPS * 32, sm, kernel); };
          if (mode == 0)
            go([&] { k_points<27, true, false>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec); });
          else if (mode == 1)
            go([&] { k_points<27, false, false>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec); });
          else if (mode == 2)
            go([&] { k_points<27, false, false, true>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec); });
          else if (mode == 3)
            go([&] { k_points<27, false, false, true, true>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec, x_state); });
          else
            go([&] { k_points<27, false, true>(n_cells, cell_nodes, cell_h4, cell_faces, cell_owned, x, tab, cf, hw, Hq, Rc, Dc, avgD, Ec); });
        }
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "vht_points_emulated: %s\n", e.what());
      return -1;
    }
  return 0;
}

// diagonal packed blocks of the lattice rows from the H_q tables: k_diag_cells + k_diag_gather (vh_diag_kernel.cuh)
extern "C" int vht_diag_emulated(int degree, int n_cells, const double *N, const double *wq, const double *Hq, int n_fast,
                                 const int32_t *fast_rows, const int32_t *fast_cells, const int8_t *fast_a, const int32_t *diag_pos,
                                 double *Dblk, double *pvals)
{
  try
    {
      if (degree == 1)
        emu::launch((unsigned)n_cells, 192, 0, [&] { k_diag_cells<8>(n_cells, N, wq, Hq, Dblk); });
      else
        emu::launch((unsigned)n_cells, 192, 0, [&] { k_diag_cells<27>(n_cells, N, wq, Hq, Dblk); });
      std::vector<double> dpack((size_t)n_fast * VH_SYMP, 0.0); // the kernel writes [fast row][180]; the test reads block positions
      double *dp = dpack.data();
      emu::launch((unsigned)n_fast, 192, 0,
                  [&] { k_diag_gather(n_fast, degree == 1 ? 8 : 27, fast_rows, fast_cells, fast_a, Dblk, dp); });
      for (int r = 0; r < n_fast; ++r)
        for (int e = 0; e < VH_SYMP; ++e)
          pvals[(size_t)diag_pos[fast_rows[r]] * VH_SYMP + e] = dpack[(size_t)r * VH_SYMP + e];
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "vht_diag_emulated: %s\n", e.what());
      return -1;
    }
  return 0;
}

// second formulation of the Q1 matrix-free apply (vh_apply_v2.cuh)
extern "C" int vht_apply_v2_emulated(int n_cells, const int32_t *cell_nodes, const double *cell_h4, const uint32_t *cell_faces,
                                     const double *z, const double *N, const double *dN, const double *wq, const double *Gref,
                                     const double *Mf, const double *coef10, const double *Hq, double *Yc)
{
  VhTables tab{};
  tab.degree = 1;
  tab.nn = tab.nq = 8;
  tab.N    = const_cast<double *>(N);
  tab.dN   = const_cast<double *>(dN);
  tab.wq   = const_cast<double *>(wq);
  tab.Gref = const_cast<double *>(Gref);
  tab.Mf   = const_cast<double *>(Mf);
  VhCoef cf{};
  cf.K1    = coef10[0];
  cf.K23   = coef10[1] + coef10[2];
  cf.alpha = coef10[3];
  for (int k = 0; k < 5; ++k)
    cf.beta[k] = coef10[4 + k];
  cf.bt = coef10[9];
  try
    {
      emu::launch((unsigned)((n_cells + 4 * VH_V2_WARPS - 1) / (4 * VH_V2_WARPS)), VH_V2_WARPS * 32, 0,
                  [&] { k_apply_q1_v2(n_cells, cell_nodes, cell_h4, cell_faces, z, tab, cf, Hq, Yc); });
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "vht_apply_v2_emulated: %s\n", e.what());
      return -1;
    }
  return 0;
}

// the two row gathers (vh_gather_kernels.cuh): which = 0: k_rhs_fast (rhs + cdiag from Rc, Dc, avgD), 1: k_gather_apply
extern "C" int vht_gather_emulated(int which, int n_fast, int dpc, const int32_t *fast_rows, const int32_t *fast_cells, const int8_t *fast_a,
                                   const uint32_t *dirmask, const double *cellvec, const double *Dc, const double *avgD, double *cdiag,
                                   const double *x_orig, double *out)
{
  try
    {
      const unsigned grid = (unsigned)(((int64_t)n_fast * 18 + 255) / 256);
      if (which == 0)
        emu::launch(grid, 256, 0, [&] { k_rhs_fast(n_fast, dpc, fast_rows, fast_cells, fast_a, dirmask, cellvec, out, Dc, avgD, cdiag); });
      else
        emu::launch(grid, 256, 0, [&] { k_gather_apply(n_fast, dpc, fast_rows, fast_cells, fast_a, dirmask, cellvec, cdiag, x_orig, out); });
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "vht_gather_emulated: %s\n", e.what());
      return -1;
    }
  return 0;
}

// block-Jacobi setup (vh_block_invert.cuh).  packed = 1: row i is lattice row i (fast_index = identity, one geometry class with
// the 3x3 block M9 in slot 0), its diagonal block Sym(dpack[i]) + kron(I_6, M9) with the Dirichlet rule; packed = 0: full blocks.
extern "C" int vht_block_invert_emulated(int n_rows, int packed, const double *blocks, const double *M9, const uint32_t *dirmask,
                                         const double *cdiag, double *minv, int *n_singular)
{
  std::vector<int32_t> ident(n_rows), zeros(n_rows, 0);
  for (int i = 0; i < n_rows; ++i)
    ident[i] = i;
  double classM[10] = {0};
  for (int i = 0; i < 9; ++i)
    classM[i] = M9[i];
  *n_singular = 0;
  try
    {
      emu::launch((unsigned)((n_rows + VH_INV_WARPS - 1) / VH_INV_WARPS), VH_INV_WARPS * 32, 0, [&] {
        k_block_invert(n_rows, ident.data(), packed ? nullptr : blocks, minv, n_singular, nullptr, 10, 0, ident.data(), zeros.data(), classM,
                       dirmask, cdiag, packed ? blocks : nullptr, packed);
      });
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "vht_block_invert_emulated: %s\n", e.what());
      return -1;
    }
  return 0;
}
