// Matep — 3He material parameters with the reference's interface (/root/reference/matep/inc/matep.h:43-127).
// In a production build the reference's own matep library is linked unchanged (the GPU only receives its six
// doubles); this restatement exists so that the stand-alone host of this repository has no dependency on
// /root/reference.  It reproduces the reference's values BIT-EXACTLY, including its float-literal quirks
// (pi and zeta(3) rounded to float, lininterp rounding through float: matep.h:102-103, matep.cc:51-52,367-404);
// tests/test_host_logic.py checks it against tests/golden/matep.json, recorded from the reference's matep.cc.
#ifndef VH_HOST_MATEP_H
#define VH_HOST_MATEP_H

namespace vhhost
{
using real_t = double;

class Matep
{
public:
  Matep() = default;
  void with_SCC(const bool &key) { scc_on = key; }

  real_t Tcp(real_t p);    // K
  real_t Tcp_mK(real_t p); // mK
  real_t mEffp(real_t p);
  real_t vFp(real_t p);
  real_t xi0p(real_t p);
  double N0p(real_t p);

  real_t alpha_td(real_t t);
  real_t beta1_td(real_t p, real_t t);
  real_t beta2_td(real_t p, real_t t);
  real_t beta3_td(real_t p, real_t t);
  real_t beta4_td(real_t p, real_t t);
  real_t beta5_td(real_t p, real_t t);
  real_t beta_A_td(real_t p, real_t t);
  real_t beta_B_td(real_t p, real_t t);
  real_t gap_A_td(real_t p, real_t t);
  real_t gap_B_td(real_t p, real_t t);
  real_t gap_td(real_t p, real_t t);
  real_t tAB_RWS(real_t p);
  real_t f_A_td(real_t p, real_t t);
  real_t f_B_td(real_t p, real_t t);
  real_t epsilon(int al, int be, int ga);

private:
  bool   scc_on = false;
  real_t beta_k(int k, real_t p, real_t t);
  real_t lininterp(const real_t *table, real_t p);
};
} // namespace vhhost
#endif
