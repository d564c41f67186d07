// C entry points over the FemGL driver mirror, for ctypes (tests, bench.py): run a .prm given as text and read back the
// per-Newton-step history.  Lives in libvhdriver.so, which links the CUDA library; libvhhost.so stays CUDA-free.
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

#include "confreader.h"
#include "femgl.h"

using namespace vhhost;

namespace
{
struct Run
{
  ParameterHandler           prm;
  std::unique_ptr<FemGL<3>>  femgl;
  std::ostringstream         log;
  std::string                log_copy, err;
};
} // namespace

extern "C" {

// Returns a handle (never null); check vhd_error().  prm_text overrides defaults exactly like configuration.prm.
void *vhd_run(const char *prm_text)
{
  Run *r = new Run();
  try
    {
      confreader cr(r->prm);
      r->prm.parse_input_from_string(prm_text ? prm_text : "");
      r->prm.enter_subsection("control parameters");
      const unsigned int degree = (unsigned int)r->prm.get_integer("polynomial degree");
      r->prm.leave_subsection();
      r->femgl.reset(new FemGL<3>(degree, r->prm, &r->log));
      r->femgl->run();
    }
  catch (const std::exception &e)
    {
      r->err = e.what();
    }
  return r;
}
const char *vhd_error(void *h) { return static_cast<Run *>(h)->err.c_str(); }
const char *vhd_log(void *h)
{
  Run *r      = static_cast<Run *>(h);
  r->log_copy = r->log.str();
  return r->log_copy.c_str();
}
int vhd_n_steps(void *h)
{
  Run *r = static_cast<Run *>(h);
  return r->femgl ? (int)r->femgl->history().size() : 0;
}
// out[12] = cycle, iteration, rhs_norm, linear_its, residual, alpha, trials, energy, t_assemble_ms, t_solve_ms, t_newton_ms, t_setup_ms
void vhd_step(void *h, int i, double *out)
{
  const auto &s = static_cast<Run *>(h)->femgl->history()[i];
  out[0] = s.cycle, out[1] = s.iteration, out[2] = s.rhs_norm, out[3] = s.linear_its, out[4] = s.residual, out[5] = s.alpha,
  out[6] = s.trials, out[7] = s.energy, out[8] = s.t_assemble_ms, out[9] = s.t_solve_ms, out[10] = s.t_newton_ms, out[11] = s.t_setup_ms;
}
long long vhd_solution_size(void *h)
{
  Run *r = static_cast<Run *>(h);
  return r->femgl ? (long long)r->femgl->solution().size() : 0;
}
void vhd_solution(void *h, double *out)
{
  const auto &v = static_cast<Run *>(h)->femgl->solution();
  std::memcpy(out, v.data(), v.size() * sizeof(double));
}
void vhd_free(void *h) { delete static_cast<Run *>(h); }

} // extern "C"
