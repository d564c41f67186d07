#include "confreader.h"

namespace vhhost
{
confreader::confreader(ParameterHandler &p) : prm(p) { declare_parameters(); } // confreader.cc:111-114

void confreader::read_parameters(const std::string &file) { prm.parse_input(file); } // confreader.cc:116-121

void confreader::declare_parameters()
{
  // (key, default) pairs in the reference's order, declare.cc:115-291.  Spelling errors are part of the keys.
  prm.enter_subsection("physical parameters");
  {
    static const char *kv[][2] = {{"pressure in bar", "0.0"},
                                  {"t_reduced", "0.0"},
                                  {"AdGR diffuse length", "1.0e10"},
                                  {"gaussian random mean value", "2.0"},
                                  {"gaussian random STD", "0.1"},
                                  {"trun on Strong Coupling Correction", "true"},
                                  {"amplitude u parameter", "1.0"}};
    for (auto &e : kv)
      prm.declare_entry(e[0], e[1]);
  }
  prm.leave_subsection();
  prm.enter_subsection("control parameters");
  {
    static const char *kv[][2] = {{"cube half side length", "20"},
                                  {"half x length of retangle", "20"},
                                  {"half y length of retangle", "20"},
                                  {"half z length of retangle", "20"},
                                  {"B-phase inner plate radius ratio", "0.4666"},
                                  {"B-phase ball radius ratio", "0.5"},
                                  {"A-phase block range ratio", "0.0"},
                                  {"Number of refinements", "4"},
                                  {"Number of interations", "40"},
                                  {"Cycle 0 refinement threshold", "1.0e0"},
                                  {"Cycle 0 linear solver tol", "1.0e-1"},
                                  {"Cycle 1 refinement threshold", "1.0e0"},
                                  {"Cycle 1 do global refinement", "false"},
                                  {"Cycle 1 linear solver tol", "1.0e-1"},
                                  {"Cycle 2 refinement threshold", "1.0e0"},
                                  {"Cycle 2 do global refinement", "false"},
                                  {"Cycle 2 linear solver tol", "1.0e-1"},
                                  {"Cycle 3 refinement threshold", "1.0e0"},
                                  {"Cycle 3 do global refinement", "false"},
                                  {"Cycle 3 linear solver tol", "1.0e-1"},
                                  {"Cycle 4 do global refinement", "false"},
                                  {"Cycle 4 linear solver tol", "1.0e-1"},
                                  {"converge accuracy", "5.0e-6"},
                                  {"adaptive refinment ratio", "0.3"},
                                  {"adaptive coarsen ratio", "0.0"},
                                  {"Number of initial global refinments", "4"},
                                  {"Number of n-cycle in AdditionalData", "4"}, // AMG-only: parsed, ignored by the GPU solver
                                  {"maximum linear iteration number", "10000"},
                                  {"Using dampped Newton iteration", "true"},
                                  {"primary step length of dampped newton iteration", "0.83"}};
    for (auto &e : kv)
      prm.declare_entry(e[0], e[1]);
    // ---- additive entries of the GPU path (not in the reference; defaults reproduce its behaviour) ----
    prm.declare_entry("GMRES restart length", "30");      // SolverFGMRES default max_basis_size
    prm.declare_entry("polynomial degree", "1");          // main.cc:116 hard-codes FemGL<3>(1, prm)
    // the reference picks geometry / initial condition by (un)commenting sources in femgl/CMakeLists.txt:42-64
    prm.declare_entry("geometry", "cube");                // stem of a reference makegrid_<stem>.cc box variant (list: host/femgl.cc grid_variants); aliases cube, retangle, retangle-xy-periodic
    // the reference writes a .vtu per rank + a .pvtu record after every Newton step (run.cc:221-227); the mirror does so on request
    prm.declare_entry("write vtu output", "false");
    prm.declare_entry("initial condition", "B-phase");    // B-phase: setup_uniform_B-phase.cc | A-phase: setup_uniform_A-phase.cc | BnA: setup_uniform_BnA-flatwall-configuration.cc
  }
  prm.leave_subsection();
}
} // namespace vhhost
