// Derived from VerHem (verkko-Hem-repo), Copyright (C) 2023-present by Kuang. Zhang (author: Quang. Zhang, timohyva@github,
// Helsinki Institute of Physics, University of Helsinki), GNU LGPL version 2.1 or later; original version:
// https://github.com/VerHem/verkko-Hem-repo.  THIS FILE IS MODIFIED: the loop nests of femgl/src/assemble.cc and residual.cc as a test driver for the reference's own term files.  See NOTICE and LICENSE.
// TEST INFRASTRUCTURE — not product code.  Never linked into the CUDA library.
//
// "O1": drives the reference's OWN pointwise term functions
//   /root/reference/femgl/src/cell_mat_vec/*.cc  (compiled verbatim, in place,
//   against oracle/shim/dealii_shim.h by oracle/Makefile; nothing is copied)
// with a literal transcription of the loop nests of
//   FemGL::assemble_system   /root/reference/femgl/src/assemble.cc:177-361
//   FemGL::compute_residual  /root/reference/femgl/src/residual.cc:166-289
// The deal.II objects those loops query (FEValues / FEFaceValues) are replaced
// by plain tables passed in by the caller:
//   N [a*n_q + q]            = scalar FE_Q shape value of node a at q          (fe_values[comp].value)
//   dN[(a*n_q + q)*3 + k]    = real-space gradient component k                 (fe_values[comp].gradient)
//   JxW[q]
// Local DoF i of FESystem(FE_Q(p),18) is (node a = i/18, component c = i%18)
// (SURVEY.md Appendix A.3), so  fe_values[comp].value(i,q) = (comp==c) ? N[a][q] : 0.
//
// The term functions touch no class members, so they are invoked on an
// uninitialised, suitably sized buffer reinterpreted as FemGL<3>.
#include <dealii_shim.h>

#define private public
#include "femgl.h"
#undef private

#include <cstring>
#include <new>

using namespace dealii;
typedef FemGL_mpi::FemGL<3> Ref;

static Ref *dummy()
{
  static void *buf = ::operator new(sizeof(Ref) + 64);
  return reinterpret_cast<Ref *>(buf);
}

static void fill3x3(FullMatrix<double> &M, const double *a)
{
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      M.set(r, c, a[3 * r + c]);
}

// --- transcriptions of the generator helpers -------------------------------------------------
// phi_matrix_generator (phi_vector2matrix.cc:116-129)
static void phi_gen(const double *N, int n_q, unsigned x, unsigned q, FullMatrix<double> &pu, FullMatrix<double> &pv)
{
  const unsigned a = x / 18, c = x % 18;
  for (unsigned comp = 0; comp <= 8; ++comp)
    {
      pu.set(comp / 3u, comp % 3u, (c == comp) ? N[a * n_q + q] : 0.0);
      pv.set(comp / 3u, comp % 3u, (c == comp + 9) ? N[a * n_q + q] : 0.0);
    }
}
// grad_phi_matrix_container_generator (phi_vector2matrix.cc:131-151)
static void grad_phi_gen(const double *dN, int n_q, unsigned x, unsigned q, std::vector<FullMatrix<double>> &gu,
                         std::vector<FullMatrix<double>> &gv)
{
  const unsigned a = x / 18, c = x % 18;
  for (unsigned comp = 0; comp <= 8; ++comp)
    for (unsigned k = 0; k < 3; ++k)
      {
        gu[k].set(comp / 3u, comp % 3u, (c == comp) ? dN[(a * n_q + q) * 3 + k] : 0.0);
        gv[k].set(comp / 3u, comp % 3u, (c == comp + 9) ? dN[(a * n_q + q) * 3 + k] : 0.0);
      }
}
// phi_matrix_face_generator (phi_vector2matrix.cc:153-195): wall-normal column zeroed
static void phi_face_gen(const double *Nf, int n_qf, unsigned x, unsigned q, FullMatrix<double> &pu, FullMatrix<double> &pv,
                         unsigned b_id)
{
  const unsigned a = x / 18, c = x % 18;
  for (unsigned comp = 0; comp <= 8; ++comp)
    {
      const bool zero = ((comp % 3 == 0) && b_id == 2) || ((comp % 3 == 1) && b_id == 3) || ((comp % 3 == 2) && b_id == 4);
      if (zero)
        {
          pu.set(comp / 3u, comp % 3u, 0.);
          pv.set(comp / 3u, comp % 3u, 0.);
        }
      else
        {
          pu.set(comp / 3u, comp % 3u, (c == comp) ? Nf[a * n_qf + q] : 0.0);
          pv.set(comp / 3u, comp % 3u, (c == comp + 9) ? Nf[a * n_qf + q] : 0.0);
        }
    }
}
// vector_matrix_generator (s_vector2matrix.cc:116-166): get_function_values = sum_i U_i phi_i
static void vec_gen(const double *N, int n_nodes, int n_q, const double *U, unsigned q, FullMatrix<double> &u, FullMatrix<double> &v)
{
  for (unsigned comp = 0; comp <= 8; ++comp)
    {
      double su = 0, sv = 0;
      for (int a = 0; a < n_nodes; ++a)
        {
          su += U[18 * a + comp] * N[a * n_q + q];
          sv += U[18 * a + comp + 9] * N[a * n_q + q];
        }
      u.set(comp / 3u, comp % 3u, su);
      v.set(comp / 3u, comp % 3u, sv);
    }
}
// grad_vector_matrix_generator (s_vector2matrix.cc:168-217)
static void grad_vec_gen(const double *dN, int n_nodes, int n_q, const double *U, unsigned q, std::vector<FullMatrix<double>> &gu,
                         std::vector<FullMatrix<double>> &gv)
{
  for (unsigned comp = 0; comp <= 8; ++comp)
    for (unsigned k = 0; k < 3; ++k)
      {
        double su = 0, sv = 0;
        for (int a = 0; a < n_nodes; ++a)
          {
            su += U[18 * a + comp] * dN[(a * n_q + q) * 3 + k];
            sv += U[18 * a + comp + 9] * dN[(a * n_q + q) * 3 + k];
          }
        gu[k].set(comp / 3u, comp % 3u, su);
        gv[k].set(comp / 3u, comp % 3u, sv);
      }
}
// vector_face_matrix_generator (s_vector2matrix.cc:219-299)
static void vec_face_gen(const double *Nf, int n_nodes, int n_qf, const double *U, unsigned q, FullMatrix<double> &u,
                         FullMatrix<double> &v, unsigned b_id)
{
  for (unsigned comp = 0; comp <= 8; ++comp)
    {
      const bool zero = ((comp % 3 == 0) && b_id == 2) || ((comp % 3 == 1) && b_id == 3) || ((comp % 3 == 2) && b_id == 4);
      double     su = 0, sv = 0;
      if (!zero)
        for (int a = 0; a < n_nodes; ++a)
          {
            su += U[18 * a + comp] * Nf[a * n_qf + q];
            sv += U[18 * a + comp + 9] * Nf[a * n_qf + q];
          }
      u.set(comp / 3u, comp % 3u, su);
      v.set(comp / 3u, comp % 3u, sv);
    }
}

extern "C" {

// The six pointwise RHS forms at one (q,i): out = {alpha, beta1..beta5}
void vhref_rhs_terms(const double *u9, const double *v9, const double *phiu9, const double *phiv9, double *out6)
{
  Ref               *F = dummy();
  FullMatrix<double> u(3, 3), v(3, 3), pu(3, 3), pv(3, 3);
  fill3x3(u, u9);
  fill3x3(v, v9);
  fill3x3(pu, phiu9);
  fill3x3(pv, phiv9);
  out6[0] = F->vec_rhs_alpha(pu, pv, u, v);
  out6[1] = F->vec_rhs_beta1(u, v, pu, pv);
  out6[2] = F->vec_rhs_beta2(u, v, pu, pv);
  out6[3] = F->vec_rhs_beta3(u, v, pu, pv);
  out6[4] = F->vec_rhs_beta4(u, v, pu, pv);
  out6[5] = F->vec_rhs_beta5(u, v, pu, pv);
}

// The six pointwise LHS forms at one (q,i,j)
void vhref_lhs_terms(const double *u9, const double *v9, const double *pui9, const double *pvi9, const double *puj9,
                     const double *pvj9, double *out6)
{
  Ref               *F = dummy();
  FullMatrix<double> u(3, 3), v(3, 3), pui(3, 3), pvi(3, 3), puj(3, 3), pvj(3, 3);
  fill3x3(u, u9);
  fill3x3(v, v9);
  fill3x3(pui, pui9);
  fill3x3(pvi, pvi9);
  fill3x3(puj, puj9);
  fill3x3(pvj, pvj9);
  out6[0] = F->mat_lhs_alpha(pui, puj, pvi, pvj);
  out6[1] = F->mat_lhs_beta1(u, v, pui, puj, pvi, pvj);
  out6[2] = F->mat_lhs_beta2(u, v, pui, puj, pvi, pvj);
  out6[3] = F->mat_lhs_beta3(u, v, pui, puj, pvi, pvj);
  out6[4] = F->mat_lhs_beta4(u, v, pui, puj, pvi, pvj);
  out6[5] = F->mat_lhs_beta5(u, v, pui, puj, pvi, pvj);
}

// Gradient forms: g*27 = [k][3x3] ; out = {lhs_K1, lhs_K2K3}
void vhref_lhs_grad_terms(const double *gui27, const double *gvi27, const double *guj27, const double *gvj27, double *out2)
{
  Ref                            *F = dummy();
  const FullMatrix<double>        id(IdentityMatrix(3));
  std::vector<FullMatrix<double>> gui(3, id), gvi(3, id), guj(3, id), gvj(3, id);
  for (int k = 0; k < 3; ++k)
    {
      fill3x3(gui[k], gui27 + 9 * k);
      fill3x3(gvi[k], gvi27 + 9 * k);
      fill3x3(guj[k], guj27 + 9 * k);
      fill3x3(gvj[k], gvj27 + 9 * k);
    }
  out2[0] = F->mat_lhs_K1(gui, gvi, guj, gvj);
  out2[1] = F->mat_lhs_K2K3(gui, gvi, guj, gvj);
}

// Literal cell loops.  coef = {K1,K2,K3,alpha,beta1..beta5,bt}.
// Faces: n_faces wall faces of this cell, face f has boundary id face_bid[f] (2|3|4),
// tables Nf[f][a*n_qf+q], JxWf[f][q].  want_matrix=0 gives the residual.cc variant (rhs only).
void vhref_cell(int n_nodes, int n_q, const double *N, const double *dN, const double *JxW, const double *U, const double *coef,
                int n_faces, const int *face_bid, int n_qf, const double *Nf, const double *JxWf, int want_matrix,
                double *cell_matrix_out, double *cell_rhs_out)
{
  Ref         *F  = dummy();
  const double K1 = coef[0], K2 = coef[1], K3 = coef[2], alpha = coef[3], beta1 = coef[4], beta2 = coef[5], beta3 = coef[6],
               beta4 = coef[7], beta5 = coef[8], bt = coef[9];
  const unsigned int dofs_per_cell = 18 * n_nodes;
  const unsigned int n_q_points = n_q, n_face_q_points = n_qf;

  std::vector<double> cell_matrix(want_matrix ? (size_t)dofs_per_cell * dofs_per_cell : 0, 0.0);
  std::vector<double> cell_rhs(dofs_per_cell, 0.0);

  FullMatrix<double>              old_solution_u(3, 3), old_f_solution_u(3, 3);
  FullMatrix<double>              old_solution_v(3, 3), old_f_solution_v(3, 3);
  const FullMatrix<double>        identity(IdentityMatrix(3));
  std::vector<FullMatrix<double>> grad_old_u_q(3, identity), grad_old_v_q(3, identity);
  FullMatrix<double>              phi_u_i_q(3, 3), phi_uf_i_q(3, 3), phi_u_j_q(3, 3), phi_uf_j_q(3, 3);
  FullMatrix<double>              phi_v_i_q(3, 3), phi_vf_i_q(3, 3), phi_v_j_q(3, 3), phi_vf_j_q(3, 3);
  std::vector<FullMatrix<double>> grad_phi_u_i_q(3, identity), grad_phi_v_i_q(3, identity), grad_phi_u_j_q(3, identity),
    grad_phi_v_j_q(3, identity);

  for (unsigned int q = 0; q < n_q_points; ++q) // assemble.cc:188
    {
      old_solution_u = 0.0;
      old_solution_v = 0.0;
      for (auto &m : grad_old_u_q)
        m = 0.0;
      for (auto &m : grad_old_v_q)
        m = 0.0;
      vec_gen(N, n_nodes, n_q, U, q, old_solution_u, old_solution_v);         // assemble.cc:200
      grad_vec_gen(dN, n_nodes, n_q, U, q, grad_old_u_q, grad_old_v_q);       // assemble.cc:201

      for (unsigned int i = 0; i < dofs_per_cell; ++i) // assemble.cc:205
        {
          if (want_matrix)
            for (unsigned int j = 0; j < dofs_per_cell; ++j) // assemble.cc:207
              {
                phi_u_i_q = 0.0;
                phi_u_j_q = 0.0;
                phi_v_i_q = 0.0;
                phi_v_j_q = 0.0;
                for (int k = 0; k < 3; ++k)
                  {
                    grad_phi_u_i_q[k] = 0.0;
                    grad_phi_v_i_q[k] = 0.0;
                    grad_phi_u_j_q[k] = 0.0;
                    grad_phi_v_j_q[k] = 0.0;
                  }
                phi_gen(N, n_q, i, q, phi_u_i_q, phi_v_i_q);
                phi_gen(N, n_q, j, q, phi_u_j_q, phi_v_j_q);
                grad_phi_gen(dN, n_q, i, q, grad_phi_u_i_q, grad_phi_v_i_q);
                grad_phi_gen(dN, n_q, j, q, grad_phi_u_j_q, grad_phi_v_j_q);

                cell_matrix[(size_t)i * dofs_per_cell + j] += // assemble.cc:229-252
                  (((K1 * F->mat_lhs_K1(grad_phi_u_i_q, grad_phi_v_i_q, grad_phi_u_j_q, grad_phi_v_j_q)) +
                    ((K2 + K3) * F->mat_lhs_K2K3(grad_phi_u_i_q, grad_phi_v_i_q, grad_phi_u_j_q, grad_phi_v_j_q)) +
                    (alpha * F->mat_lhs_alpha(phi_u_i_q, phi_u_j_q, phi_v_i_q, phi_v_j_q)) +
                    2.0 * ((beta1 * F->mat_lhs_beta1(old_solution_u, old_solution_v, phi_u_i_q, phi_u_j_q, phi_v_i_q, phi_v_j_q)) +
                           (beta2 * F->mat_lhs_beta2(old_solution_u, old_solution_v, phi_u_i_q, phi_u_j_q, phi_v_i_q, phi_v_j_q)) +
                           (beta3 * F->mat_lhs_beta3(old_solution_u, old_solution_v, phi_u_i_q, phi_u_j_q, phi_v_i_q, phi_v_j_q)) +
                           (beta4 * F->mat_lhs_beta4(old_solution_u, old_solution_v, phi_u_i_q, phi_u_j_q, phi_v_i_q, phi_v_j_q)) +
                           (beta5 * F->mat_lhs_beta5(old_solution_u, old_solution_v, phi_u_i_q, phi_u_j_q, phi_v_i_q, phi_v_j_q)))) *
                   JxW[q]);
              }
          if (!want_matrix)
            { // residual.cc builds phi_i once per (q,i)
              phi_u_i_q = 0.0;
              phi_v_i_q = 0.0;
              for (int k = 0; k < 3; ++k)
                {
                  grad_phi_u_i_q[k] = 0.0;
                  grad_phi_v_i_q[k] = 0.0;
                }
              phi_gen(N, n_q, i, q, phi_u_i_q, phi_v_i_q);
              grad_phi_gen(dN, n_q, i, q, grad_phi_u_i_q, grad_phi_v_i_q);
            }
          cell_rhs[i] -= // assemble.cc:257-276 / residual.cc:206-226
            (((K1 * F->vec_rhs_K1(grad_old_u_q, grad_old_v_q, grad_phi_u_i_q, grad_phi_v_i_q)) +
              ((K2 + K3) * F->vec_rhs_K2K3(grad_phi_u_i_q, grad_phi_v_i_q, grad_old_u_q, grad_old_v_q)) +
              (alpha * F->vec_rhs_alpha(phi_u_i_q, phi_v_i_q, old_solution_u, old_solution_v)) +
              2.0 * ((beta1 * F->vec_rhs_beta1(old_solution_u, old_solution_v, phi_u_i_q, phi_v_i_q)) +
                     (beta2 * F->vec_rhs_beta2(old_solution_u, old_solution_v, phi_u_i_q, phi_v_i_q)) +
                     (beta3 * F->vec_rhs_beta3(old_solution_u, old_solution_v, phi_u_i_q, phi_v_i_q)) +
                     (beta4 * F->vec_rhs_beta4(old_solution_u, old_solution_v, phi_u_i_q, phi_v_i_q)) +
                     (beta5 * F->vec_rhs_beta5(old_solution_u, old_solution_v, phi_u_i_q, phi_v_i_q)))) *
             JxW[q]);
        }
    }

  // Robin ("AdGR diffuse") wall faces: assemble.cc:286-348 / residual.cc:237-281
  for (int f = 0; f < n_faces; ++f)
    {
      const unsigned b_id = face_bid[f];
      if ((b_id == 2 || b_id == 3 || b_id == 4) && (bt < 1e10))
        {
          const double *Nff   = Nf + (size_t)f * n_nodes * n_qf;
          const double *JxWff = JxWf + (size_t)f * n_qf;
          for (unsigned int q_face = 0; q_face < n_face_q_points; ++q_face)
            {
              old_f_solution_u = 0.0;
              old_f_solution_v = 0.0;
              vec_face_gen(Nff, n_nodes, n_qf, U, q_face, old_f_solution_u, old_f_solution_v, b_id);
              for (unsigned int i = 0; i < dofs_per_cell; ++i)
                {
                  if (want_matrix)
                    for (unsigned int j = 0; j < dofs_per_cell; ++j)
                      {
                        phi_uf_i_q = 0.0;
                        phi_vf_i_q = 0.0;
                        phi_uf_j_q = 0.0;
                        phi_vf_j_q = 0.0;
                        phi_face_gen(Nff, n_qf, i, q_face, phi_uf_i_q, phi_vf_i_q, b_id);
                        phi_face_gen(Nff, n_qf, j, q_face, phi_uf_j_q, phi_vf_j_q, b_id);
                        cell_matrix[(size_t)i * dofs_per_cell + j] -= // assemble.cc:322-326
                          ((K1 * (-1.0 / bt) * F->mat_face_lhs_K1(phi_uf_i_q, phi_uf_j_q, phi_vf_i_q, phi_vf_j_q)) *
                           JxWff[q_face]);
                      }
                  else
                    {
                      phi_uf_i_q = 0.0;
                      phi_vf_i_q = 0.0;
                      phi_face_gen(Nff, n_qf, i, q_face, phi_uf_i_q, phi_vf_i_q, b_id);
                    }
                  cell_rhs[i] -= // assemble.cc:333-337
                    (((-K1) * (-1.0 / bt) * F->vec_face_rhs_K1(phi_uf_i_q, phi_vf_i_q, old_f_solution_u, old_f_solution_v)) *
                     JxWff[q_face]);
                }
            }
        }
    }

  if (want_matrix && cell_matrix_out)
    std::memcpy(cell_matrix_out, cell_matrix.data(), cell_matrix.size() * sizeof(double));
  std::memcpy(cell_rhs_out, cell_rhs.data(), cell_rhs.size() * sizeof(double));
}

// ---- Matep (reference matep.cc compiled verbatim) ----
// out = {alpha, beta1..5, gapA, gapB, fA, fB, Tcp_mK, tAB_RWS}
void vhref_matep(double p, double t, int scc, double *out12)
{
  FemGL_mpi::Matep mat;
  bool             key = scc != 0;
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

} // extern "C"
