// FemGL — host-side mirror of the reference's solver class (/root/reference/femgl/inc/femgl.h:121-349) whose three hot
// members are thin adapters over the C ABI of include/vh_femgl.h.  Same public surface
//     FemGL(unsigned int Q_degree, ParameterHandler &);   void run();
// same private member names, same .prm keys, same pcout lines.  What deal.II does in the reference (mesh, DoFs,
// constraints, refinement) is done by vhhost::Mesh here (deal.II is not available in this image); in a deal.II build the
// bodies of assemble_system / compute_residual / solve / newton_iteration below are what replaces the reference's
// (INTEGRATION.md).  One process drives one GPU; multi-GPU runs use one process per rank (bench.py).
#ifndef VH_HOST_FEMGL_H
#define VH_HOST_FEMGL_H

#include <memory>
#include <ostream>
#include <string>
#include <thread>
#include <vector>

#include "matep.h"
#include "mesh.h"
#include "param_handler.h"

struct vh_ctx;

namespace vhhost
{
template <int dim>
class FemGL
{
public:
  // `log` is where the reference's pcout lines go (default std::cout, as in the reference)
  FemGL(unsigned int Q_degree, ParameterHandler &, std::ostream *log = nullptr);
  ~FemGL();

  void run();

  // machine-readable per-Newton-step record (the printed lines are the reference's only observable API)
  struct StepRecord
  {
    unsigned int cycle, iteration;
    double       rhs_norm;     // system_rhs.l2_norm()            (solve.cc:158)
    int          linear_its;   // solver_control.last_step()      (solve.cc:176)
    double       residual;     // residual_vector.l2_norm()       (run.cc:234)
    double       alpha;        // accepted line-search step       (iteration.cc:172)
    int          trials;       // compute_residual() calls        (iteration.cc:187)
    double       energy;       // GL functional of the accepted state (not evaluated by the reference)
    double       t_assemble_ms, t_solve_ms, t_newton_ms;
    double       t_setup_ms;   // tables + vh_create (+ vh_transfer_solution after refine_grid) charged to the cycle's first step
  };
  const std::vector<StepRecord> &history() const { return records; }
  const std::vector<double>     &solution() const { return host_solution; }
  void                           set_output_stream(std::ostream *os) { out = os; }

private:
  void make_grid();
  void setup_system();
  void assemble_system();
  void compute_residual();
  void solve(const double &);
  void newton_iteration();
  void refine_grid(std::string &);
  void output_results(const std::string &dirc) const;

  void check(int rc, const char *what) const;

  unsigned int      degree, cycle, iteration_loop;
  ParameterHandler &conf;
  std::unique_ptr<Mesh>       triangulation; // + dof_handler + constraints, see mesh.h
  std::unique_ptr<RankTables> tables;
  vh_ctx                     *gpu = nullptr;
  std::vector<double>         host_solution; // local_solution (owned part), kept for output/refinement

  const double K1 = 0.42072; // femgl.h:320-322
  const double K2 = 0.42072;
  const double K3 = 0.42072;
  Matep        mat;
  double       alpha, beta1, beta2, beta3, beta4, beta5;
  double       bt;
  double       reduced_t, p;
  bool         SCC_key;

  // quantities the reference keeps in vectors and queries with l2_norm()
  double system_rhs_l2 = 0, residual_l2 = 0, last_alpha = 0;
  int    last_linear_its = 0, last_trials = 0;
  double last_setup_ms = 0;

  std::ostream           *out;
  std::vector<StepRecord> records;

  // output path: the snapshot leaves the device on a second stream and a writer thread produces the file while the next
  // Newton step runs (at most one file in flight; joined before the mesh or the context changes)
  bool                write_vtu = false;
  mutable std::thread writer;
  mutable std::string writer_error;
  mutable double      output_main_thread_ms = 0.0; // time output_results() kept the Newton loop waiting, summed
  void                finish_output() const;
};
} // namespace vhhost
#endif
