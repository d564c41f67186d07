// Second formulation of the matrix-free operator apply at Q1 (VH_SPMV_MF=3; k_points<APPLY> is mode 1): same result,
//   Yc[cell] = K_cell z_cell,   K_cell = sum_q w_q (N_a N_b) (vol H_q) + geometry-only gradient / Robin forms,
// but built for occupancy and for less shared-memory traffic (static budget of the apply kernel):
//  * the bulk part keeps "lane = quadrature point" (z_q = sum_a N_a(q) z_a, t_q = w_q H_q z_q from the packed table,
//    back-projection y_a = sum_q N_a(q) t_q through a per-warp buffer);
//  * the gradient forms are NOT interpolated: with lane = node a they are applied as the per-cell 8x8 table
//      y_a[c] += D_ab[x(c)] z_b[c] + sum_j M_ab[x(c)][j] z_b[3*(c/3)+j],
//      M_ab[i][j] = vol K23 G_ab[i][j] / (h_i h_j),  D_ab[x] = vol K1 sum_i G_ab[i][i]/h_i^2 (+ Robin face mass, x != normal)
//    from the reference-cell table G, which removes the 54-value gradient staging of mode 1;
//  * 128 registers (4 CTAs of 4 warps per SM instead of 2), 19 KB of shared memory per CTA.
// Written after the round-1 GPU budget was spent; its logic is checked lane by lane under tests/native/cuda_emu.h
// (tests/test_kernel_emulation.py).  Same header for nvcc and for the emulation.
#ifndef VH_APPLY_V2_CUH
#define VH_APPLY_V2_CUH

#define VH_V2_WARPS 4
#define VH_V2_ZS 146 /* 8 nodes x 18 + 2: the four cells of a warp start in different banks */
#define VH_V2_GS 73  /* padded row of the geometry table: lanes a = 0..7 hit different banks */

__global__ void __launch_bounds__(VH_V2_WARPS * 32, 4)
  k_apply_q1_v2(int n_cells, const int32_t *__restrict__ cell_nodes, const double *__restrict__ cell_h,
                const uint32_t *__restrict__ cell_faces, const double *__restrict__ z, VhTables tab, VhCoef cf,
                const double *__restrict__ Hq, double *__restrict__ Yc)
{
  __shared__ __align__(16) double sNT[64];                          // [q][a]
  __shared__ __align__(16) double sG[8 * VH_V2_GS];                 // [a][b*9 + 3i + j]
  __shared__ __align__(16) double sZ[VH_V2_WARPS * 4 * VH_V2_ZS];   // [warp][cell of the warp][a*18 + c]
  __shared__ __align__(16) double sT[VH_V2_WARPS * 4 * VH_V2_ZS];   // [warp][cell of the warp][q*18 + c]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 3, i8 = lane & 7;
  for (int i = t; i < 64; i += VH_V2_WARPS * 32)
    sNT[(i & 7) * 8 + (i >> 3)] = tab.N[i]; // tab.N is [a][q]
  for (int i = t; i < 8 * 72; i += VH_V2_WARPS * 32)
    sG[(i / 72) * VH_V2_GS + i % 72] = tab.Gref[i];
  double   *wZ = sZ + warp * 4 * VH_V2_ZS, *wT = sT + warp * 4 * VH_V2_ZS;
  const int cell0 = (blockIdx.x * VH_V2_WARPS + warp) * 4;
#pragma unroll
  for (int kk = 0; kk < 9; ++kk)
    { // coalesced gather of the warp's 4 x 8 node rows in 16-byte pieces
      const int i  = lane + 32 * kk; // < 288 = 4 cells x 8 nodes x 9 pieces
      const int gg = i / 72, r = i - 72 * gg, a = r / 9, pp = r - 9 * a;
      const int e  = min(cell0 + gg, n_cells - 1);
      *reinterpret_cast<double2 *>(wZ + gg * VH_V2_ZS + a * 18 + 2 * pp) =
        *reinterpret_cast<const double2 *>(z + 18 * (int64_t)cell_nodes[(int64_t)e * 8 + a] + 2 * pp);
    }
  __syncthreads(); // tables of the CTA and the gathers of the warp

  const bool    live = cell0 + g < n_cells;
  const int64_t cell = min(cell0 + g, n_cells - 1);
  const double *zc   = wZ + g * VH_V2_ZS;
  double       *tc   = wT + g * VH_V2_ZS;
  // ---- bulk part, lane = quadrature point q ----
  {
    const int q = i8;
    double    zq[18], tq[18];
#pragma unroll
    for (int c = 0; c < 18; ++c)
      zq[c] = 0.0, tq[c] = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a)
      {
        const double n = sNT[q * 8 + a];
#pragma unroll
        for (int cp = 0; cp < 9; ++cp)
          {
            const double2 u = *reinterpret_cast<const double2 *>(zc + a * 18 + 2 * cp);
            zq[2 * cp]      = fma(n, u.x, zq[2 * cp]);
            zq[2 * cp + 1]  = fma(n, u.y, zq[2 * cp + 1]);
          }
      }
    const double *hbase = Hq + cell * (int64_t)(8 * VH_SYMP);
    vh_sym_matvec(
      [&](int pp, double &v0, double &v1) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(hbase + ((pp << 3) + (q ^ (pp & 7))) * 2));
        v0 = v.x;
        v1 = v.y;
      },
      zq, tq);
    const double w = tab.wq[q]; // the tables carry the cell volume already
#pragma unroll
    for (int cp = 0; cp < 9; ++cp)
      *reinterpret_cast<double2 *>(tc + q * 18 + 2 * cp) = make_double2(w * tq[2 * cp], w * tq[2 * cp + 1]);
  }
  __syncwarp();
  // ---- lane = node a: back-projection of the bulk part, then the geometry-only forms ----
  const int a = i8;
  double    y[18];
#pragma unroll
  for (int c = 0; c < 18; ++c)
    y[c] = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    {
      const double n = sNT[q * 8 + a];
#pragma unroll
      for (int cp = 0; cp < 9; ++cp)
        {
          const double2 v = *reinterpret_cast<const double2 *>(tc + q * 18 + 2 * cp);
          y[2 * cp]       = fma(n, v.x, y[2 * cp]);
          y[2 * cp + 1]   = fma(n, v.y, y[2 * cp + 1]);
        }
    }
  const double2 h01 = *reinterpret_cast<const double2 *>(cell_h + 4 * cell), h23 = *reinterpret_cast<const double2 *>(cell_h + 4 * cell + 2);
  const double  vol = h23.y, ih[3] = {1.0 / h01.x, 1.0 / h01.y, 1.0 / h23.x};
  double        cm[3][3], ck[3]; // vol K23 /(h_i h_j), vol K1 / h_i^2
#pragma unroll
  for (int i = 0; i < 3; ++i)
    {
      ck[i] = vol * cf.K1 * ih[i] * ih[i];
#pragma unroll
      for (int j = 0; j < 3; ++j)
        cm[i][j] = vol * cf.K23 * ih[i] * ih[j];
    }
  const uint32_t faces = cell_faces[cell];
  const bool     robin = (cf.bt < 1e10) && faces != 0u;
#pragma unroll 2
  for (int b = 0; b < 8; ++b)
    {
      const double *G = sG + a * VH_V2_GS + b * 9;
      double        M[3][3], D[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          M[i][j] = cm[i][j] * G[3 * i + j];
      const double d0 = ck[0] * G[0] + ck[1] * G[4] + ck[2] * G[8];
      D[0] = D[1] = D[2] = d0;
      if (robin)
        for (int f = 0; f < 6; ++f)
          { // AdGR diffuse wall faces: K1/bt * unit-face mass on the components whose orbital index is not the wall normal
            const int bid = (faces >> (4 * f)) & 15u;
            if (bid < 2 || bid > 4)
              continue;
            const double s = cf.K1 / cf.bt * (vol * (f / 2 == 0 ? ih[0] : (f / 2 == 1 ? ih[1] : ih[2]))) * tab.Mf[(size_t)(f * 8 + a) * 8 + b];
#pragma unroll
            for (int x = 0; x < 3; ++x)
              if (x != bid - 2)
                D[x] += s;
          }
#pragma unroll
      for (int tr = 0; tr < 6; ++tr)
        { // the divergence term couples the three components of one row of u or v
          double zb[3];
#pragma unroll
          for (int j = 0; j < 3; ++j)
            zb[j] = zc[b * 18 + 3 * tr + j];
#pragma unroll
          for (int i = 0; i < 3; ++i)
            y[3 * tr + i] += D[i] * zb[i] + M[i][0] * zb[0] + M[i][1] * zb[1] + M[i][2] * zb[2];
        }
    }
  if (live)
    {
      double *dst = Yc + cell * 144 + a * 18;
#pragma unroll
      for (int cp = 0; cp < 9; ++cp)
        *reinterpret_cast<double2 *>(dst + 2 * cp) = make_double2(y[2 * cp], y[2 * cp + 1]);
    }
}

#endif
