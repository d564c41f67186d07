// The pointwise kernel family k_points (one THREAD per quadrature point): assembly tables + cell rhs, residual, energy, and
// the matrix-free / table-free operator apply.  Kept in its own header so that the same source is compiled by nvcc into
// vh_assemble.cu (inside its anonymous namespace) and, for the CPU-only test container, by g++ against the lane-by-lane
// CUDA emulation of tests/native/cuda_emu.h (tests/test_kernel_emulation.py).  Needs vh_internal.h (VhTables, VhCoef).
#ifndef VH_POINTS_KERNEL_CUH
#define VH_POINTS_KERNEL_CUH

#ifndef VH_DYNAMIC_SMEM
#define VH_DYNAMIC_SMEM(name) extern __shared__ __align__(16) double name[]
#endif

// ------------------------------------------------------------------------------------------------
// 1b. pointwise kernel, one THREAD per quadrature point (Q1: 8 points per cell, Q2: 27)
// ------------------------------------------------------------------------------------------------
// Q1: a warp owns 4 cells x 8 quadrature points (lane = 8*g + q); Q2: a warp owns one cell (lanes 0..26 = q, 5 idle).
// The thread interpolates A and grad A at its point, keeps A and the 42 unique product entries in registers and
// evaluates g (18) and the packed H_q (171 entries) with fully unrolled, compile-time specialised formulas (vh_h_entry):
// ~12 FP64 instructions per entry and no index arithmetic, against ~75 instructions per entry of the table-driven
// k_pointwise.  Q1: H_q leaves the registers with 16-byte stores in the [pair][q XOR pair] layout (vh_hq8_index): the 8
// lanes of a cell fill one 128-byte line per store instruction.  Q2: [q][180], each lane streams its own 1440-byte row.
// The q-sums of the cell vectors (rhs, cell diagonal) go through a per-warp shared buffer: the lane then plays node
// a = lane % G and accumulates its 18 components over the NQ points; only __syncwarp() is needed.
#define VH_PT_WARPS 4
template <int NN>
struct VhPt
{
  static constexpr int G       = NN == 8 ? 8 : 32;          // lanes per cell
  static constexpr int CPW     = 32 / G;                    // cells per warp
  static constexpr int USTRIDE = NN * 18 + 2;               // +2: the cells of a warp start in different banks
  static constexpr int BSTRIDE = NN * 54 + 4;               // per cell: NQ points x 54 doubles (+4: bank offset)
  static constexpr int TAB     = 4 * NN * NN;               // sNT [q][a] | sdNT [q][x][a]
  static constexpr size_t SMEM = (size_t)(TAB + VH_PT_WARPS * CPW * (USTRIDE + BSTRIDE)) * sizeof(double);
  // table-free operator apply: a second gather buffer per cell (the Newton state next to the Krylov vector)
  static constexpr size_t SMEM_TFREE = (size_t)(TAB + VH_PT_WARPS * CPW * (2 * USTRIDE + BSTRIDE)) * sizeof(double);
};

//
// APPLY = true turns the kernel into the MATRIX-FREE OPERATOR APPLY of the lattice rows (VH_SPMV_MF=1, vhk_apply_fast):
// x is the (Dirichlet-masked) Krylov vector z, the bulk density g(A) is replaced by the linearisation H_q z_q read back
// from the packed tables this kernel wrote at assembly time (Hq is an input then), and Rc receives +K_cell z_cell.  The
// gradient and Robin forms are linear in the field, so that code is shared verbatim with the residual.  Per apply the
// kernel streams 8*180*NQ bytes per cell (the H_q tables) instead of the 8*180 bytes per matrix block of the packed SpMV:
// 3.5x fewer bytes at Q1 (27 blocks per row vs 8 tables per cell) and 19x fewer at Q2 (C3: 1.27 GB of tables vs 24.4 GB).
//
// APPLY + TFREE = the TABLE-FREE apply (VH_SPMV_MF=2, unverified on hardware in round 1): no H_q table is read; the thread
// also interpolates the Newton state A_q from x_state and evaluates H(A_q) z_q as the directional derivative
// vh_hessian_apply (eight 3x3 complex products).  18 doubles of state per node instead of 180 per point.
template <int NN, bool WANT_H, bool WANT_E, bool APPLY = false, bool TFREE = false>
__global__ void __launch_bounds__(VH_PT_WARPS * 32, 2)
  k_points(int n_cells, const int32_t *__restrict__ cell_nodes, const double *__restrict__ cell_h,
           const uint32_t *__restrict__ cell_faces, const uint8_t *__restrict__ cell_owned, const double *__restrict__ x,
           VhTables tab, VhCoef cf, vh_hweights hw, double *__restrict__ Hq, double *__restrict__ Rc, double *__restrict__ Dc,
           double *__restrict__ avgD, double *__restrict__ Ec, const double *__restrict__ x_state = nullptr)
{
  static_assert(!APPLY || (!WANT_H && !WANT_E), "the operator apply neither writes H_q nor integrates the energy");
  static_assert(!TFREE || APPLY, "table-free is a mode of the operator apply");
  using P = VhPt<NN>;
  constexpr int NQ = NN, G = P::G, CPW = P::CPW, DPC = 18 * NN;
  VH_DYNAMIC_SMEM(sm);
  double   *sNT  = sm;            // [q][a]
  double   *sdNT = sm + NN * NN;  // [q][x][a]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane / G, ql = lane % G;
  const bool pt = ql < NQ;                 // Q2: lanes 27..31 carry no point / node
  const int  q  = pt ? ql : NQ - 1;
  double    *sU = sm + P::TAB + warp * CPW * ((TFREE ? 2 : 1) * P::USTRIDE + P::BSTRIDE); // [CPW][USTRIDE]
  double    *sB = sU + CPW * P::USTRIDE;                                                  // [CPW][BSTRIDE]
  double    *sX = sB + CPW * P::BSTRIDE;                                                  // [CPW][USTRIDE] (TFREE only)
  for (int i = t; i < NN * NQ; i += VH_PT_WARPS * 32)
    sNT[(i % NQ) * NN + i / NQ] = tab.N[i];
  for (int i = t; i < NN * NQ * 3; i += VH_PT_WARPS * 32)
    {
      const int a = i / (3 * NQ), r = i - 3 * NQ * a, qq = r / 3, xx = r - 3 * qq;
      sdNT[(qq * 3 + xx) * NN + a] = tab.dN[i];
    }
  const int cell0 = (blockIdx.x * VH_PT_WARPS + warp) * CPW;
#pragma unroll
  for (int kk = 0; kk < (CPW * NN * 9 + 31) / 32; ++kk)
    { // coalesced gather of the warp's DoF values (16-byte pieces of the node rows), all loads in flight at once
      const int i = lane + 32 * kk;
      if (i >= CPW * NN * 9)
        break;
      const int gg = i / (NN * 9), r = i - NN * 9 * gg, a = r / 9, pp = r - 9 * a;
      const int e  = min(cell0 + gg, n_cells - 1);
      const double2 v = *reinterpret_cast<const double2 *>(x + 18 * (int64_t)cell_nodes[(int64_t)e * NN + a] + 2 * pp);
      *reinterpret_cast<double2 *>(sU + gg * P::USTRIDE + a * 18 + 2 * pp) = v;
      if constexpr (TFREE)
        *reinterpret_cast<double2 *>(sX + gg * P::USTRIDE + a * 18 + 2 * pp) =
          *reinterpret_cast<const double2 *>(x_state + 18 * (int64_t)cell_nodes[(int64_t)e * NN + a] + 2 * pp);
    }
  __syncthreads();

  const bool    live = pt && cell0 + g < n_cells;
  const int64_t cell = min(cell0 + g, n_cells - 1);
  const double2 h01 = *reinterpret_cast<const double2 *>(cell_h + 4 * cell), h23 = *reinterpret_cast<const double2 *>(cell_h + 4 * cell + 2);
  const double  vol = h23.y;
  const double  hh[3] = {h01.x, h01.y, h23.x};
  const double  ih[3] = {1.0 / h01.x, 1.0 / h01.y, 1.0 / h23.x};
  const double  JxW = tab.wq[q] * vol;
  const double *sUg = sU + g * P::USTRIDE;
  double       *gB  = sB + g * P::BSTRIDE;

  // ---- FE interpolation of the state and its gradient at this thread's point (s_vector2matrix.cc:154-162, 203-213) ----
  double a18[18];
  double eg = 0.0; // gradient energy density
  {
    double gt[54];
    if constexpr (NN == 8)
      {
        double Nq[8], dNq[3][8];
#pragma unroll
        for (int a = 0; a < 8; ++a)
          {
            Nq[a] = sNT[q * 8 + a];
#pragma unroll
            for (int xx = 0; xx < 3; ++xx)
              dNq[xx][a] = sdNT[(q * 3 + xx) * 8 + a] * ih[xx];
          }
#pragma unroll
        for (int cp = 0; cp < 9; ++cp)
          {
            double A0 = 0, A1 = 0, d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0};
#pragma unroll
            for (int a = 0; a < 8; ++a)
              {
                const double2 u = *reinterpret_cast<const double2 *>(sUg + a * 18 + 2 * cp);
                A0 = fma(Nq[a], u.x, A0);
                A1 = fma(Nq[a], u.y, A1);
#pragma unroll
                for (int xx = 0; xx < 3; ++xx)
                  {
                    d0[xx] = fma(dNq[xx][a], u.x, d0[xx]);
                    d1[xx] = fma(dNq[xx][a], u.y, d1[xx]);
                  }
              }
            a18[2 * cp]     = A0;
            a18[2 * cp + 1] = A1;
#pragma unroll
            for (int xx = 0; xx < 3; ++xx)
              {
                gt[6 * cp + xx]     = d0[xx];
                gt[6 * cp + 3 + xx] = d1[xx];
              }
          }
      }
    else
      { // Q2: the 27 x 4 shape values of this point stay in shared memory ([q][a]: conflict-free for lane = q)
#pragma unroll
        for (int c = 0; c < 18; ++c)
          a18[c] = 0.0;
#pragma unroll
        for (int i = 0; i < 54; ++i)
          gt[i] = 0.0;
#pragma unroll 1
        for (int a = 0; a < NN; ++a)
          {
            const double n = sNT[q * NN + a], n0 = sdNT[(q * 3 + 0) * NN + a] * ih[0], n1 = sdNT[(q * 3 + 1) * NN + a] * ih[1],
                         n2 = sdNT[(q * 3 + 2) * NN + a] * ih[2];
#pragma unroll
            for (int cp = 0; cp < 9; ++cp)
              {
                const double2 u = *reinterpret_cast<const double2 *>(sUg + a * 18 + 2 * cp);
                a18[2 * cp]     = fma(n, u.x, a18[2 * cp]);
                a18[2 * cp + 1] = fma(n, u.y, a18[2 * cp + 1]);
                gt[6 * cp + 0]  = fma(n0, u.x, gt[6 * cp + 0]);
                gt[6 * cp + 1]  = fma(n1, u.x, gt[6 * cp + 1]);
                gt[6 * cp + 2]  = fma(n2, u.x, gt[6 * cp + 2]);
                gt[6 * cp + 3]  = fma(n0, u.y, gt[6 * cp + 3]);
                gt[6 * cp + 4]  = fma(n1, u.y, gt[6 * cp + 4]);
                gt[6 * cp + 5]  = fma(n2, u.y, gt[6 * cp + 5]);
              }
          }
      }
    // gt[3c+x] = d_x A_c -> Gt[c][x] = JxW (K1 dA[c][x] + delta_{x,xc} K23 div): what the test gradient of
    // node a is contracted with
#pragma unroll
    for (int pm = 0; pm < 6; ++pm)
      {
        double dv[3][3];
#pragma unroll
        for (int y = 0; y < 3; ++y)
#pragma unroll
          for (int xx = 0; xx < 3; ++xx)
            {
              dv[y][xx] = gt[3 * (3 * pm + y) + xx];
              if (WANT_E)
                eg = fma(cf.K1 * dv[y][xx], dv[y][xx], eg);
            }
        const double div = dv[0][0] + dv[1][1] + dv[2][2]; // the divergence couples the three components of a row of A
        if (WANT_E)
          eg = fma(cf.K23 * div, div, eg);
#pragma unroll
        for (int y = 0; y < 3; ++y)
#pragma unroll
          for (int xx = 0; xx < 3; ++xx)
            gt[3 * (3 * pm + y) + xx] = JxW * (xx == y ? fma(cf.K23, div, cf.K1 * dv[y][xx]) : cf.K1 * dv[y][xx]);
      }
    double *myB = gB + q * 54;
    if (pt)
#pragma unroll
      for (int i = 0; i < 27; ++i)
        *reinterpret_cast<double2 *>(myB + 2 * i) = make_double2(gt[2 * i], gt[2 * i + 1]);
  }
  __syncwarp();
  // ---- round 1: lane = node a of its cell;  rc[c] = sum_q grad N_a(q) . Gt_q[c]  (assemble.cc:257-276) ----
  const int a_node = q;
  double    rc[18];
#pragma unroll
  for (int c = 0; c < 18; ++c)
    rc[c] = 0.0;
#pragma unroll(NN == 8 ? 8 : 1)
  for (int qq = 0; qq < NQ; ++qq)
    {
      double wx[3];
#pragma unroll
      for (int xx = 0; xx < 3; ++xx)
        wx[xx] = sdNT[(qq * 3 + xx) * NN + a_node] * ih[xx];
#pragma unroll
      for (int i = 0; i < 27; ++i)
        {
          const double2 v = *reinterpret_cast<const double2 *>(gB + qq * 54 + 2 * i);
          rc[(2 * i) / 3]     = fma(wx[(2 * i) % 3], v.x, rc[(2 * i) / 3]);
          rc[(2 * i + 1) / 3] = fma(wx[(2 * i + 1) % 3], v.y, rc[(2 * i + 1) / 3]);
        }
    }
  __syncwarp();

  // ---- bulk terms at this thread's point ----
  vh_prods pr;
  if constexpr (APPLY && TFREE)
    { // H(A_q) z_q without a table: A_q interpolated from the Newton state, then the directional derivative of g
      double Aq[18];
#pragma unroll
      for (int c = 0; c < 18; ++c)
        Aq[c] = 0.0;
      const double *sXg = sX + g * P::USTRIDE;
#pragma unroll(NN == 8 ? 8 : 1)
      for (int a = 0; a < NN; ++a)
        {
          const double n = sNT[q * NN + a];
#pragma unroll
          for (int cp = 0; cp < 9; ++cp)
            {
              const double2 u = *reinterpret_cast<const double2 *>(sXg + a * 18 + 2 * cp);
              Aq[2 * cp]      = fma(n, u.x, Aq[2 * cp]);
              Aq[2 * cp + 1]  = fma(n, u.y, Aq[2 * cp + 1]);
            }
        }
      double prod[72];
#pragma unroll
      for (int e = 0; e < 36; ++e)
        vh_product_entry(Aq, e, prod + 2 * e);
      double gv[18];
      vh_hessian_apply(Aq, prod, a18, cf.alpha, cf.beta, gv);
      double *myB = gB + q * 18;
      if (pt)
#pragma unroll
        for (int i = 0; i < 9; ++i)
          *reinterpret_cast<double2 *>(myB + 2 * i) = make_double2(JxW * gv[2 * i], JxW * gv[2 * i + 1]);
    }
  else if constexpr (APPLY)
    { // (vol H_q) z_q from the stored packed table of this (cell, point); the tables carry the cell volume already
      double        gv[18];
#pragma unroll
      for (int c = 0; c < 18; ++c)
        gv[c] = 0.0;
      const double *hbase = Hq + cell * (int64_t)(NQ * VH_SYMP) + (NN == 8 ? 0 : q * VH_SYMP);
      vh_sym_matvec(
        [&](int pp, double &v0, double &v1) {
          const double2 v = NN == 8 ? __ldg(reinterpret_cast<const double2 *>(hbase + ((pp << 3) + (q ^ (pp & 7))) * 2))
                                    : __ldg(reinterpret_cast<const double2 *>(hbase + 2 * pp));
          v0 = v.x;
          v1 = v.y;
        },
        a18, gv);
      const double w   = tab.wq[q];
      double      *myB = gB + q * 18;
      if (pt)
#pragma unroll
        for (int i = 0; i < 9; ++i)
          *reinterpret_cast<double2 *>(myB + 2 * i) = make_double2(w * gv[2 * i], w * gv[2 * i + 1]);
    }
  else
    {
      vh_prods_compute(a18, pr);
      double gv[18];
      vh_g_all(a18, pr, hw, gv);
      double *myB = gB + q * 18;
      if (pt)
#pragma unroll
        for (int i = 0; i < 9; ++i)
          *reinterpret_cast<double2 *>(myB + 2 * i) = make_double2(JxW * gv[2 * i], JxW * gv[2 * i + 1]);
    }
  __syncwarp();
  // ---- round 2: rc[c] += sum_q N_a(q) JxW g_q[c] ----
#pragma unroll(NN == 8 ? 8 : 1)
  for (int qq = 0; qq < NQ; ++qq)
    {
      const double n = sNT[qq * NN + a_node];
#pragma unroll
      for (int i = 0; i < 9; ++i)
        {
          const double2 v = *reinterpret_cast<const double2 *>(gB + qq * 18 + 2 * i);
          rc[2 * i]       = fma(n, v.x, rc[2 * i]);
          rc[2 * i + 1]   = fma(n, v.y, rc[2 * i + 1]);
        }
    }
  __syncwarp();
  const uint32_t faces = cell_faces[cell];
  const bool     robin = (cf.bt < 1e10) && faces != 0u;
  if (robin)
    for (int f = 0; f < 6; ++f)
      { // Robin (AdGR diffuse) wall faces: K1/bt * unit-face mass, components whose row index is the wall normal are skipped
        const int bid = (faces >> (4 * f)) & 15u;
        if (bid < 2 || bid > 4)
          continue;
        const double  s = cf.K1 / cf.bt * (vol / (f / 2 == 0 ? hh[0] : (f / 2 == 1 ? hh[1] : hh[2])));
        const double *M = tab.Mf + (size_t)(f * NN + a_node) * NN;
        for (int b = 0; b < NN; ++b)
          {
            const double m = s * M[b];
#pragma unroll
            for (int c = 0; c < 18; ++c)
              if (c % 3 != bid - 2)
                rc[c] = fma(m, sUg[b * 18 + c], rc[c]);
          }
      }
  if (live)
    {
      double      *dst = Rc + cell * DPC + a_node * 18;
      const double sg  = APPLY ? 1.0 : -1.0; // residual: system_rhs = -R (assemble.cc:257); apply: +K_cell z_cell
#pragma unroll
      for (int i = 0; i < 9; ++i)
        *reinterpret_cast<double2 *>(dst + 2 * i) = make_double2(sg * rc[2 * i], sg * rc[2 * i + 1]);
    }

  if (WANT_H)
    { // ---- the packed H_q, pre-multiplied by the cell volume, and the cell-matrix diagonal ----
      const vh_hdiag hd    = vh_make_hdiag(pr, hw);
      double        *hbase = Hq + cell * (int64_t)(NQ * VH_SYMP) + (NN == 8 ? 0 : q * VH_SYMP);
      double        *myB   = gB + q * 18;
#pragma unroll
      for (int cc = 0; cc < 18; ++cc)
        { // rows in the order 0, 9, 1, 10, ...: Re and Im rows of one matrix position share their complex products
          const int c = (cc >> 1) + 9 * (cc & 1);
#pragma unroll
          for (int d = 2 * (c >> 1); d < 18; d += 2)
            {
              const int    pp = vh_sym_index(c, d) >> 1;
              const double v0 = d >= c ? vh_h_entry(a18, pr, hw, hd, c, d) : 0.0; // (c, c-1) is the zero dummy of odd rows
              const double v1 = vh_h_entry(a18, pr, hw, hd, c, d + 1);
              if (d == c && pt)
                myB[c] = JxW * v0;
              if (d + 1 == c && pt)
                myB[c] = JxW * v1;
              if (live)
                {
                  if constexpr (NN == 8)
                    *reinterpret_cast<double2 *>(hbase + ((pp << 3) + (q ^ (pp & 7))) * 2) = make_double2(v0 * vol, v1 * vol);
                  else
                    *reinterpret_cast<double2 *>(hbase + 2 * pp) = make_double2(v0 * vol, v1 * vol);
                }
            }
        }
      __syncwarp();
      // round 3: diagonal of the cell matrix, dg[c] = sum_q N_a(q)^2 JxW H_q[c][c] + geometry (+ Robin)
      double dg[18];
      {
        const double *Gd = tab.Gref + (size_t)(a_node * NN + a_node) * 9;
        const double  g0 = Gd[0] * ih[0] * ih[0], g1 = Gd[4] * ih[1] * ih[1], g2 = Gd[8] * ih[2] * ih[2];
        const double  k1 = vol * cf.K1 * (g0 + g1 + g2);
#pragma unroll
        for (int c = 0; c < 18; ++c)
          dg[c] = k1 + vol * cf.K23 * (c % 3 == 0 ? g0 : (c % 3 == 1 ? g1 : g2));
      }
#pragma unroll(NN == 8 ? 8 : 1)
      for (int qq = 0; qq < NQ; ++qq)
        {
          const double n = sNT[qq * NN + a_node], n2 = n * n;
#pragma unroll
          for (int i = 0; i < 9; ++i)
            {
              const double2 v = *reinterpret_cast<const double2 *>(gB + qq * 18 + 2 * i);
              dg[2 * i]       = fma(n2, v.x, dg[2 * i]);
              dg[2 * i + 1]   = fma(n2, v.y, dg[2 * i + 1]);
            }
        }
      if (robin)
        for (int f = 0; f < 6; ++f)
          {
            const int bid = (faces >> (4 * f)) & 15u;
            if (bid < 2 || bid > 4)
              continue;
            const double s = cf.K1 / cf.bt * (vol / (f / 2 == 0 ? hh[0] : (f / 2 == 1 ? hh[1] : hh[2])));
            const double m = s * tab.Mf[(size_t)(f * NN + a_node) * NN + a_node];
#pragma unroll
            for (int c = 0; c < 18; ++c)
              if (c % 3 != bid - 2)
                dg[c] += m;
          }
      double absd = 0.0;
#pragma unroll
      for (int c = 0; c < 18; ++c)
        absd += fabs(dg[c]);
      if (!pt)
        absd = 0.0;
#pragma unroll
      for (int o = 1; o < G; o <<= 1)
        absd += __shfl_xor_sync(0xffffffffu, absd, o);
      if (live)
        {
          double *dst = Dc + cell * DPC + a_node * 18;
#pragma unroll
          for (int i = 0; i < 9; ++i)
            *reinterpret_cast<double2 *>(dst + 2 * i) = make_double2(dg[2 * i], dg[2 * i + 1]);
          if (a_node == 0)
            avgD[cell] = absd / (double)DPC;
        }
    }
  if (WANT_E)
    { // GL functional, SURVEY.md A.1 (only locally owned cells count; ghost cells are assembled redundantly)
      double e = JxW * (eg + vh_bulk_energy_u(pr, cf.alpha, cf.beta));
      if (robin)
        for (int f = 0; f < 6; ++f)
          { // lane = node a: its share  s * sum_c U[a][c] sum_b M[a][b] U[b][c]
            const int bid = (faces >> (4 * f)) & 15u;
            if (bid < 2 || bid > 4)
              continue;
            const double  s = cf.K1 / cf.bt * (vol / (f / 2 == 0 ? hh[0] : (f / 2 == 1 ? hh[1] : hh[2])));
            const double *M = tab.Mf + (size_t)(f * NN + a_node) * NN;
            double        ef = 0.0;
            for (int c = 0; c < 18; ++c)
              {
                if (c % 3 == bid - 2)
                  continue;
                double m = 0.0;
                for (int b = 0; b < NN; ++b)
                  m += M[b] * sUg[b * 18 + c];
                ef += sUg[a_node * 18 + c] * m;
              }
            e += s * ef;
          }
      if (!pt)
        e = 0.0;
#pragma unroll
      for (int o = 1; o < G; o <<= 1)
        e += __shfl_xor_sync(0xffffffffu, e, o);
      if (live && a_node == 0)
        Ec[cell] = cell_owned[cell] ? e : 0.0;
    }
}


#endif
