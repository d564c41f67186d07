// Derived from VerHem (verkko-Hem-repo), Copyright (C) 2023-present by Kuang. Zhang (author: Quang. Zhang, timohyva@github,
// Helsinki Institute of Physics, University of Helsinki), GNU LGPL version 2.1 or later; original version:
// https://github.com/VerHem/verkko-Hem-repo.  THIS FILE IS MODIFIED: compact rewrite of matep/src/matep.cc with the same coefficient tables.  See NOTICE and LICENSE.
#include "matep.h"

#include <cmath>
#include <iostream>

namespace vhhost
{
namespace
{
// The reference stores these in double variables initialised from float literals (matep.h:102, matep.cc:46-52).
const real_t PI      = 3.14159265358979323846264338328f;
const real_t ZETA3   = 1.2020569031595942f;
const real_t C_BETAI = (7.0f * ZETA3) / (80.0f * PI * PI);
const real_t U_AMU   = 1.66053906660f * (1.0e-27) * 1.0f;
const real_t M3      = 3.016293f * U_AMU;
const real_t NM      = (1.0e-9) * 1.0f;
const real_t HBAR    = 1.054571817f * (1.0e-34) * 1.0f * 1.0f;

// Strong-coupling-correction sheets (JWS2019) and Fermi-liquid data at p = 0, 2, ..., 34 bar (matep.cc:54-72).
const real_t SCC[5][18] = {
  {-0.0098, -0.0127, -0.0155, -0.0181, -0.0207, -0.0231, -0.0254, -0.0275, -0.0295, -0.0314, -0.0330, -0.0345, -0.0358, -0.0370,
   -0.0381, -0.0391, -0.0402, -0.0413},
  {-0.0419, -0.0490, -0.0562, -0.0636, -0.0711, -0.0786, -0.0861, -0.0936, -0.1011, -0.1086, -0.1160, -0.1233, -0.1306, -0.1378,
   -0.1448, -0.1517, -0.1583, -0.1645},
  {-0.0132, -0.0161, -0.0184, -0.0202, -0.0216, -0.0226, -0.0233, -0.0239, -0.0243, -0.0247, -0.0249, -0.0252, -0.0255, -0.0258,
   -0.0262, -0.0265, -0.0267, -0.0268},
  {-0.0047, -0.0276, -0.0514, -0.0760, -0.1010, -0.1260, -0.1508, -0.1751, -0.1985, -0.2208, -0.2419, -0.2614, -0.2795, -0.2961,
   -0.3114, -0.3255, -0.3388, -0.3518},
  {-0.0899, -0.1277, -0.1602, -0.1880, -0.2119, -0.2324, -0.2503, -0.2660, -0.2801, -0.2930, -0.3051, -0.3167, -0.3280, -0.3392,
   -0.3502, -0.3611, -0.3717, -0.3815}};
const real_t WEAK[5]    = {-1.0f, 2.0f, 2.0f, 2.0f, -2.0f}; // weak-coupling beta_k / c_betai
const real_t TC_MK[18]  = {0.929, 1.181, 1.388, 1.560, 1.705, 1.828, 1.934, 2.026, 2.106, 2.177, 2.239, 2.293, 2.339, 2.378, 2.411,
                           2.438, 2.463, 2.486};
const real_t MSTAR[18]  = {2.80, 3.05, 3.27, 3.48, 3.68, 3.86, 4.03, 4.20, 4.37, 4.53, 4.70, 4.86, 5.02, 5.18, 5.34, 5.50, 5.66, 5.82};
const real_t VFERMI[18] = {59.03, 55.41, 52.36, 49.77, 47.56, 45.66, 44.00, 42.51, 41.17, 39.92, 38.74, 37.61, 36.53, 35.50, 34.53,
                           33.63, 32.85, 32.23};
const real_t XI0_NM[18] = {77.21, 57.04, 45.85, 38.77, 33.91, 30.37, 27.66, 25.51, 23.76, 22.29, 21.03, 19.94, 18.99, 18.15, 17.41,
                           16.77, 16.22, 15.76};
} // namespace

// Piecewise-linear in p on the 2-bar grid; the result is rounded through float exactly like matep.cc:367-404.
// (The reference leaves the interval index uninitialised for p < 0 or p >= 34 — undefined behaviour there;
// here the interval is clamped to the table.)
real_t Matep::lininterp(const real_t *tab, real_t p)
{
  int k = (int)std::floor(p / 2.0);
  if (k < 0)
    k = 0;
  if (k > 16)
    k = 16;
  const float pk = (float)(2 * k);
  const float fp = (float)(((tab[k + 1] - tab[k]) / 2.0) * (p - pk) + tab[k]);
  return fp;
}

real_t Matep::Tcp(real_t p) { return lininterp(TC_MK, p) * (1.0e-3); }
real_t Matep::Tcp_mK(real_t p) { return lininterp(TC_MK, p); }
real_t Matep::mEffp(real_t p) { return lininterp(MSTAR, p) * M3; }
real_t Matep::vFp(real_t p) { return lininterp(VFERMI, p); }
real_t Matep::xi0p(real_t p) { return lininterp(XI0_NM, p) * NM; }
double Matep::N0p(real_t p) { return (std::pow(mEffp(p), 2) * vFp(p)) / ((2.0f * PI * PI) * (HBAR * HBAR * HBAR)); }

real_t Matep::alpha_td(real_t t) { return 1.f * (t - 1); }

real_t Matep::beta_k(int k, real_t p, real_t t)
{
  if (scc_on)
    return C_BETAI * (WEAK[k] + (t)*lininterp(SCC[k], p));
  return C_BETAI * (WEAK[k]);
}
real_t Matep::beta1_td(real_t p, real_t t) { return beta_k(0, p, t); }
real_t Matep::beta2_td(real_t p, real_t t) { return beta_k(1, p, t); }
real_t Matep::beta3_td(real_t p, real_t t) { return beta_k(2, p, t); }
real_t Matep::beta4_td(real_t p, real_t t) { return beta_k(3, p, t); }
real_t Matep::beta5_td(real_t p, real_t t) { return beta_k(4, p, t); }

real_t Matep::beta_A_td(real_t p, real_t t) { return beta2_td(p, t) + beta4_td(p, t) + beta5_td(p, t); }
real_t Matep::beta_B_td(real_t p, real_t t)
{
  return beta1_td(p, t) + beta2_td(p, t) + (1.f / 3.f) * (beta3_td(p, t) + beta4_td(p, t) + beta5_td(p, t));
}
real_t Matep::gap_A_td(real_t p, real_t t)
{
  if (t <= 1.0)
    return std::sqrt(-alpha_td(t) / (2.f * beta_A_td(p, t)));
  return 0.;
}
real_t Matep::gap_B_td(real_t p, real_t t)
{
  if (t <= 1.0)
    return std::sqrt(-alpha_td(t) / (2.f * beta_B_td(p, t)));
  return 0.;
}
real_t Matep::f_A_td(real_t p, real_t t)
{
  if (t <= 1.0)
    return (-1.f / 4.f) * (std::pow(alpha_td(t), 2.f)) / beta_A_td(p, t);
  return 0.;
}
real_t Matep::f_B_td(real_t p, real_t t)
{
  if (t <= 1.0)
    return (-1.f / 4.f) * (std::pow(alpha_td(t), 2.f)) / beta_B_td(p, t);
  return 0.;
}
real_t Matep::gap_td(real_t p, real_t t)
{ // same branches and messages as matep.cc:258-288 (including its A/B labelling)
  const real_t fa = f_A_td(p, t), fb = f_B_td(p, t);
  if (fa > fb)
    {
      std::cout << " \nnow p, T are: " << p << ", " << t << ", equlibrum bulk phase is B phase. " << std::endl;
      return gap_A_td(p, t);
    }
  if (fa < fb)
    {
      std::cout << " \nnow p, T are: " << p << ", " << t << ", equlibrum bulk phase is A phase. " << std::endl;
      return gap_B_td(p, t);
    }
  if (t < 1.0)
    {
      std::cout << " \nnow p, t are: " << p << ", " << t << ", and A and B degenerate, return as -1. " << std::endl;
      return -1.f;
    }
  std::cout << " \nnow p, t are: " << p << ", " << t << ", system is in normal phase. " << std::endl;
  return 0.f;
}
real_t Matep::tAB_RWS(real_t p)
{
  return 1.f / (3.f * lininterp(SCC[0], p) + lininterp(SCC[2], p) - 2.f * lininterp(SCC[3], p) - 2.f * lininterp(SCC[4], p));
}
real_t Matep::epsilon(int al, int be, int ga)
{ // Levi-Civita symbol
  if (al == be || al == ga || be == ga)
    return 0.0;
  return ((be - al) * (ga - al) * (ga - be)) > 0 ? 1.0 : -1.0;
}
} // namespace vhhost
