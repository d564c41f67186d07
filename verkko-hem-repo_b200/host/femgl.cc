// Derived from VerHem (verkko-Hem-repo), Copyright (C) 2023-present by Kuang. Zhang (author: Quang. Zhang, timohyva@github,
// Helsinki Institute of Physics, University of Helsinki), GNU LGPL version 2.1 or later; original version:
// https://github.com/VerHem/verkko-Hem-repo.  THIS FILE IS MODIFIED: the control flow of femgl/src/{run,solve,iteration,refine}.cc around calls into the CUDA library.  See NOTICE and LICENSE.
// See femgl.h.  Control flow and printed lines follow the reference:
//   ctor                 /root/reference/femgl/src/femgl.cc:107-203
//   run()                /root/reference/femgl/src/run.cc:108-260
//   assemble_system()    /root/reference/femgl/src/assemble.cc:108-372      -> vh_assemble
//   solve(tol)           /root/reference/femgl/src/solve.cc:108-187         -> vh_solve
//   newton_iteration()   /root/reference/femgl/src/iteration.cc:109-215     -> vh_line_search_trial / vh_residual / vh_accept_trial
//   compute_residual()   /root/reference/femgl/src/residual.cc:109-297      -> vh_residual
//   make_grid()          /root/reference/femgl/src/makegrid_cube-z-normal_AdGR.cc:140-197,
//                        makegrid_retangle-z-AdGR-xy-HomoNeumann.cc:159-192
//   setup_system()       /root/reference/femgl/src/setup_uniform_B-phase.cc:109-341,
//                        setup_uniform_BnA-flatwall-configuration.cc:201-245
//   refine_grid()        /root/reference/femgl/src/refine.cc:109-181
#include "femgl.h"
#include "vtu.h"

#include "../../include/vh_femgl.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <iostream>
#include <stdexcept>

namespace vhhost
{
namespace
{
double now_ms()
{
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
} // namespace

template <int dim>
void FemGL<dim>::check(int rc, const char *what) const
{
  if (rc != VH_OK)
    throw std::runtime_error(std::string(what) + ": " + vh_last_error(gpu)); // reaches main()'s handler, main.cc:120-145
}

template <int dim>
FemGL<dim>::FemGL(unsigned int Q_degree, ParameterHandler &prmHandler, std::ostream *log)
  : degree(Q_degree), cycle(0), iteration_loop(0), conf(prmHandler), out(log ? log : &std::cout)
{
  static_assert(dim == 3, "femgl is three-dimensional");
  conf.enter_subsection("physical parameters");
  p         = conf.get_double("pressure in bar");
  reduced_t = conf.get_double("t_reduced");
  bt        = conf.get_double("AdGR diffuse length");
  SCC_key   = conf.get_bool("trun on Strong Coupling Correction");
  conf.leave_subsection();

  mat.with_SCC(SCC_key);
  alpha = mat.alpha_td(reduced_t);
  beta1 = mat.beta1_td(p, reduced_t);
  beta2 = mat.beta2_td(p, reduced_t);
  beta3 = mat.beta3_td(p, reduced_t);
  beta4 = mat.beta4_td(p, reduced_t);
  beta5 = mat.beta5_td(p, reduced_t);

  std::ostream &pcout = *out;
  pcout << "------------------------------------------------------" << "\n"
        << ">>>>>>>>>>  Physical Parameters in this run  <<<<<<<<<" << "\n"
        << "------------------------------------------------------" << "\n"
        << " Number of MPI processes is " << 1 << "\n"
        << " p is " << p << ", t is " << reduced_t << ", T is " << (reduced_t * mat.Tcp_mK(p)) << "\n"
        << " AdGR expolation lenghtn bt is " << bt << ", SCC_key is " << SCC_key << "\n"
        << " gapA is " << mat.gap_A_td(p, reduced_t) << ", gapB is " << mat.gap_B_td(p, reduced_t) << "\n"
        << " f_A is " << mat.f_A_td(p, reduced_t) << ", f_B is " << mat.f_B_td(p, reduced_t) << "\n"
        << " alpha is " << alpha << ", beta1 is " << beta1 << ", beta2 is " << beta2 << "\n"
        << " beta3 is " << beta3 << ", beta4 is " << beta4 << ", beta5 is " << beta5 << "\n"
        << "------------------------------------------------------" << "\n"
        << ">>>>>>>>>>  Physical Parameters in this run  <<<<<<<<<" << "\n"
        << "------------------------------------------------------" << "\n"
        << std::endl;
}

template <int dim>
FemGL<dim>::~FemGL()
{
  if (writer.joinable())
    writer.join();
  if (gpu)
    vh_destroy(gpu);
}

// The reference selects geometry and initial condition at compile time by (un)commenting one makegrid_*.cc and one
// setup_*.cc in femgl/CMakeLists.txt:42-64.  Here the additive key "geometry" selects among the box variants that list
// names, by the stem of the reference file (short aliases: cube, retangle, retangle-xy-periodic).
namespace
{
struct GridVariant
{
  const char *name;
  int         box;         // 0: hyper_cube(-half, half)   1: hyper_rectangle(-h, +h) from the .prm   2: hard-wired box of the file
  int         face_bid[6]; // x0 x1 y0 y1 z0 z1; 1 natural, 2|3|4 AdGR wall with normal x|y|z, (5,6)|(7,8)|(9,10) periodic pairs
  double      lo[3], hi[3];
};
const GridVariant grid_variants[] = {
  // makegrid_cube-z-normal_AdGR.cc:157-195
  {"cube", 0, {1, 1, 1, 1, 4, 4}, {}, {}},
  {"cube-z-normal_AdGR", 0, {1, 1, 1, 1, 4, 4}, {}, {}},
  // makegrid_cube-xyz-Homo-Neumann.cc:158-196
  {"cube-xyz-Homo-Neumann", 0, {1, 1, 1, 1, 1, 1}, {}, {}},
  // makegrid_cube-xyz-periodic.cc:159-226
  {"cube-xyz-periodic", 0, {5, 6, 7, 8, 9, 10}, {}, {}},
  // makegrid_cube-z-normal_AdGR-xy-periodic.cc:157-214
  {"cube-z-normal_AdGR-xy-periodic", 0, {5, 6, 7, 8, 4, 4}, {}, {}},
  // makegrid_retangle-z-AdGR-xy-HomoNeumann.cc:159-192
  {"retangle", 1, {1, 1, 1, 1, 4, 4}, {}, {}},
  {"retangle-z-AdGR-xy-HomoNeumann", 1, {1, 1, 1, 1, 4, 4}, {}, {}},
  // makegrid_retangle-xyz-homogenous-Neumann.cc:163-200
  {"retangle-xyz-homogenous-Neumann", 1, {1, 1, 1, 1, 1, 1}, {}, {}},
  // makegrid_retangle-xyz-periodic.cc:164-230
  {"retangle-xyz-periodic", 1, {5, 6, 7, 8, 9, 10}, {}, {}},
  // makegrid_retangle-z-AdGR_x-periodic-y-HomoNeumann.cc:164-209
  {"retangle-z-AdGR_x-periodic-y-HomoNeumann", 1, {5, 6, 1, 1, 4, 4}, {}, {}},
  // makegrid_retangle-z-AdGR_xy-periodic.cc:164-219 (the variant femgl/CMakeLists.txt:51 compiles)
  {"retangle-xy-periodic", 1, {5, 6, 7, 8, 4, 4}, {}, {}},
  {"retangle-z-AdGR_xy-periodic", 1, {5, 6, 7, 8, 4, 4}, {}, {}},
  // makegrid_xz-normal_AdGR.cc:156-197: box (-l,0,0)-(l,Ex,D) with l = 15, Ex = 8, D = 6; x and z faces are AdGR walls
  {"xz-normal_AdGR", 2, {2, 2, 1, 1, 4, 4}, {-15.0, 0.0, 0.0}, {15.0, 8.0, 6.0}},
};
} // namespace

template <int dim>
void FemGL<dim>::make_grid()
{
  conf.enter_subsection("control parameters");
  const int         number_global_refine = (int)conf.get_integer("Number of initial global refinments");
  const double      half_length          = conf.get_double("cube half side length");
  const double      hx = conf.get_double("half x length of retangle"), hy = conf.get_double("half y length of retangle"),
               hz        = conf.get_double("half z length of retangle");
  const std::string geom = conf.get("geometry");
  conf.leave_subsection();
  const int base[3] = {1, 1, 1};
  for (const GridVariant &v : grid_variants)
    if (geom == v.name)
      {
        double lo[3], hi[3];
        for (int d = 0; d < 3; ++d)
          {
            const double h = d == 0 ? hx : (d == 1 ? hy : hz);
            lo[d]          = v.box == 0 ? -half_length : (v.box == 1 ? -h : v.lo[d]);
            hi[d]          = v.box == 0 ? half_length : (v.box == 1 ? h : v.hi[d]);
          }
        triangulation.reset(new Mesh((int)degree, lo, hi, base, v.face_bid, number_global_refine));
        return;
      }
  std::string known;
  for (const GridVariant &v : grid_variants)
    known += std::string(known.empty() ? "" : " | ") + v.name;
  throw std::runtime_error("FemGL::make_grid: unknown geometry \"" + geom + "\" (" + known + ")");
}

template <int dim>
void FemGL<dim>::setup_system()
{
  std::ostream &pcout = *out;
  triangulation->finalize(1);
  tables.reset(new RankTables(triangulation->tables(0)));
  pcout << "   Number of degrees of freedom: " << 18 * triangulation->n_nodes << std::endl;

  if (gpu)
    {
      vh_destroy(gpu);
      gpu = nullptr;
    }
  const double t_create0 = now_ms();
  vh_mesh_desc d;
  std::memset(&d, 0, sizeof(d));
  RankTables &T    = *tables;
  d.degree         = T.degree;
  d.n_owned_nodes  = T.n_owned_nodes;
  d.n_ghost_nodes  = T.n_ghost_nodes;
  d.node_global    = T.node_global.data();
  d.n_cells        = T.n_cells;
  d.cell_nodes     = T.cell_nodes.data();
  d.cell_origin    = T.cell_origin.data();
  d.cell_h         = T.cell_h.data();
  d.cell_owned     = T.cell_owned.data();
  d.n_wall_faces   = (int32_t)T.wall_face_cell.size();
  d.wall_face_cell = T.wall_face_cell.data();
  d.wall_face_no   = T.wall_face_no.data();
  d.wall_face_bid  = T.wall_face_bid.data();
  vh_constraints c;
  c.n_lines                   = (int32_t)T.c_dof.size();
  c.dof                       = T.c_dof.data();
  c.ptr                       = T.c_ptr.data();
  c.master                    = T.c_master.data();
  c.weight                    = T.c_weight.data();
  d.constraints_newton_update = c; // identical structure: both use zero Dirichlet values (dirichlet.h:110-169)
  d.constraints_solution      = c;
  int rc                      = vh_create(&d, 0, &gpu);
  if (rc != VH_OK)
    throw std::runtime_error(std::string("vh_create: ") + vh_last_error(nullptr));
  const double beta[5] = {beta1, beta2, beta3, beta4, beta5};
  check(vh_set_coefficients(gpu, K1, K2, K3, alpha, beta, bt), "vh_set_coefficients");

  if (cycle == 0)
    { // initial condition
      conf.enter_subsection("control parameters");
      const std::string ic           = conf.get("initial condition");
      const double      rangeA_ratio = conf.get_double("A-phase block range ratio");
      const double      z_half       = conf.get_double("half z length of retangle");
      conf.leave_subsection();
      host_solution.assign((size_t)18 * T.n_owned_nodes, 0.0);
      if (ic == "BnA")
        { // flat A/B wall at z = ratio*Lz: BnA.h:130-163, setup_uniform_BnA-flatwall-configuration.cc:228-235
          const double z_ofA     = rangeA_ratio * z_half;
          const double matelem_A = mat.gap_A_td(p, reduced_t) * 0.7071067811865475f;
          const double matelem_B = mat.gap_B_td(p, reduced_t) * 0.5773502691896258f;
          for (int n = 0; n < T.n_owned_nodes; ++n)
            {
              double *v = &host_solution[(size_t)18 * n];
              if (T.node_xyz[(size_t)3 * n + 2] >= z_ofA)
                v[0] = v[4] = v[8] = matelem_B;
              else
                v[0] = v[10] = matelem_A;
            }
        }
      else if (ic == "A-phase")
        { // uniform A phase: u11 = v12 = gap_A * 0.707107f, setup_uniform_A-phase.cc:235-251
          const double amp = mat.gap_A_td(p, reduced_t) * 0.707107f;
          for (int n = 0; n < T.n_owned_nodes; ++n)
            {
              double *v = &host_solution[(size_t)18 * n];
              v[0] = v[10] = amp;
            }
        }
      else if (ic != "B-phase")
        throw std::runtime_error("FemGL::setup_system: unknown initial condition \"" + ic + "\" (B-phase | A-phase | BnA)");
      else
        { // uniform B phase, setup_uniform_B-phase.cc:245-259
          const double amp = mat.gap_B_td(p, reduced_t) * 0.577350269f;
          for (int n = 0; n < T.n_owned_nodes; ++n)
            {
              double *v = &host_solution[(size_t)18 * n];
              v[0] = v[4] = v[8] = amp;
            }
        }
      // constraints_solution.distribute (setup_uniform_B-phase.cc:262): zero the Dirichlet DoFs, interpolate hanging ones
      for (size_t l = 0; l < T.c_dof.size(); ++l)
        {
          double s = 0.0;
          for (int q = T.c_ptr[l]; q < T.c_ptr[l + 1]; ++q)
            s += T.c_weight[q] * host_solution[T.c_master[q]];
          host_solution[T.c_dof[l]] = s;
        }
      check(vh_set_solution(gpu, host_solution.data()), "vh_set_solution");
      last_setup_ms = now_ms() - t_create0;
    } // later cycles: refine_grid() transfers the state on the device (vh_transfer_solution)
}

template <int dim>
void FemGL<dim>::assemble_system()
{
  check(vh_assemble(gpu, &system_rhs_l2), "assemble_system");
}

template <int dim>
void FemGL<dim>::compute_residual()
{
  check(vh_residual(gpu, &residual_l2), "compute_residual");
}

template <int dim>
void FemGL<dim>::solve(const double &tol)
{
  std::ostream &pcout = *out;
  conf.enter_subsection("control parameters");
  const unsigned int no_n_cycles = (unsigned int)conf.get_integer("Number of n-cycle in AdditionalData"); // AMG only
  const unsigned int max_linear_solver_iterations = (unsigned int)conf.get_integer("maximum linear iteration number");
  const int          restart                      = (int)conf.get_integer("GMRES restart length");
  conf.leave_subsection();
  (void)no_n_cycles;
  pcout << " start to build block-Jacobi preconditioner." << std::endl;
  pcout << " system_rhs.l2_norm() is " << system_rhs_l2 << std::endl;
  pcout << " Starting linear solving." << std::endl;
  double final_res = 0;
  check(vh_solve(gpu, tol, (int)max_linear_solver_iterations, restart, &last_linear_its, &final_res), "solve");
  pcout << "   Solved in " << last_linear_its << " iterations." << std::endl;
}

template <int dim>
void FemGL<dim>::newton_iteration()
{
  std::ostream &pcout = *out;
  conf.enter_subsection("control parameters");
  const double line_search_step = conf.get_double("primary step length of dampped newton iteration");
  const bool   dampped_newton   = conf.get_bool("Using dampped Newton iteration");
  conf.leave_subsection();

  const double previous_residual = system_rhs_l2; // iteration.cc:130
  last_trials                    = 0;
  if (dampped_newton == false)
    {
      check(vh_line_search_trial(gpu, 1.0), "newton_iteration");
      compute_residual();
      last_trials = 1;
      last_alpha  = 1.0;
      pcout << " we are in full newton-iteration " << ", residual is: " << residual_l2 << ", previous_residual is: " << previous_residual
            << std::endl;
      if (residual_l2 < previous_residual)
        pcout << " ohh! current_residual < previous_residual, we get better solution ! " << std::endl;
      else
        pcout << " Humm ! current_residual >= previous_residual, maybe you need a better guess ! This is full-newton" << std::endl;
    }
  else
    {
      for (unsigned int i = 0; i < 100; ++i)
        {
          const double a = std::pow(line_search_step, static_cast<double>(i));
          check(vh_line_search_trial(gpu, a), "newton_iteration");
          compute_residual();
          ++last_trials;
          last_alpha = a;
          pcout << " step length alpha is: " << a << ", residual is: " << residual_l2 << ", previous_residual is: " << previous_residual
                << std::endl;
          if (residual_l2 < previous_residual)
            {
              pcout << " ohh! current_residual < previous_residual, we get better solution ! " << std::endl;
              break;
            }
          else
            pcout << " haa! current_residual >= previous_residual, more line search ! " << std::endl;
        }
    }
  check(vh_accept_trial(gpu), "newton_iteration"); // local_solution = distributed_solution (iteration.cc:210)
}

template <int dim>
void FemGL<dim>::refine_grid(std::string &refinement_strategy)
{
  std::ostream &pcout = *out;
  finish_output(); // the writer thread reads the tables and the context of the mesh that is about to be replaced
  const double  t_begin = now_ms();
  // the old mesh and its GPU context stay alive until the solution has been transferred on the device
  std::unique_ptr<Mesh> old_mesh(new Mesh(*triangulation));
  vh_ctx               *old_gpu = gpu;
  gpu                           = nullptr;

  std::vector<uint8_t> flags((size_t)triangulation->n_cells(), 0);
  if (refinement_strategy == "global")
    std::fill(flags.begin(), flags.end(), 1);
  else
    { // KellyErrorEstimator + refine_and_coarsen_fixed_number (refine.cc:144-153): the top `refine_ratio` fraction of the
      // cells by the face-jump indicator (Mesh::kelly_indicator) is refined.  Coarsening is not available in the mini host.
      conf.enter_subsection("control parameters");
      const double refine_ratio  = conf.get_double("adaptive refinment ratio");
      const double coarsen_ratio = conf.get_double("adaptive coarsen ratio");
      conf.leave_subsection();
      if (coarsen_ratio != 0.0)
        {
          gpu = old_gpu;
          throw std::runtime_error("FemGL::refine_grid: \"adaptive coarsen ratio\" != 0 is not supported by this host (no coarsening); "
                                   "the reference's default is 0.0 (declare.cc:258)");
        }
      check(vh_get_solution(old_gpu, host_solution.data()), "refine_grid"); // the estimator runs on the host mesh (refine.cc:144)
      std::vector<double> eta;
      triangulation->kelly_indicator(host_solution, eta);
      const int64_t        nc = triangulation->n_cells();
      std::vector<int64_t> order(nc);
      for (int64_t e = 0; e < nc; ++e)
        order[e] = e;
      std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return eta[x] > eta[y]; });
      const int64_t n_ref = (int64_t)std::floor(refine_ratio * (double)nc); // exactly this many cells (fixed NUMBER)
      for (int64_t k = 0; k < n_ref; ++k)
        flags[order[k]] = 1;
    }
  triangulation->refine(flags);
  const double t_mesh = now_ms();
  setup_system(); // tables + GPU context of the new mesh (no state yet)
  const double t_setup = now_ms();
  pcout << (refinement_strategy == "global" ? "setup_system() call is done ! this is global refinment" : "setup_system() call is done !")
        << std::endl;
  // SolutionTransfer::interpolate + constraints_solution.distribute + ghosted copy (refine.cc:128-130, 171-175), on the device
  std::vector<int32_t> tptr, tsrc;
  std::vector<double>  tw;
  triangulation->transfer_table(*old_mesh, tptr, tsrc, tw);
  check(vh_transfer_solution(gpu, old_gpu, (int32_t)tptr.size() - 1, tptr.data(), tsrc.data(), tw.data()), "refine_grid");
  vh_destroy(old_gpu);
  host_solution.assign((size_t)18 * tables->n_owned_nodes, 0.0);
  last_setup_ms = now_ms() - t_begin;
  pcout << " refine_grid timings [ms]: estimate + refine " << t_mesh - t_begin << ", tables + GPU context " << t_setup - t_mesh
        << ", solution transfer " << now_ms() - t_setup << std::endl;
  if (refinement_strategy != "global")
    pcout << "adaptive_refine_grid() call is done !" << std::endl;
}

// io.cc:106-170 + run.cc:221-227: the reference writes both vectors after EVERY Newton step.  At GPU speeds a synchronous
// download + file write would cost more than the step, so the snapshot is taken in stream order (vh_snapshot_begin returns at
// once), leaves the device on a second stream, and a writer thread turns it into the .vtu / .pvtu pair (host/vtu.cc) while
// the next Newton step runs.  Enabled by the additive key "write vtu output".
template <int dim>
void FemGL<dim>::finish_output() const
{
  if (writer.joinable())
    writer.join();
  if (!writer_error.empty())
    {
      const std::string e = writer_error;
      writer_error.clear();
      throw std::runtime_error("output_results: " + e);
    }
}

template <int dim>
void FemGL<dim>::output_results(const std::string &dirc) const
{
  if (!write_vtu || !gpu)
    return;
  const double t0 = now_ms();
  finish_output(); // the previous file still reads the pinned snapshot buffer
  check(vh_snapshot_begin(gpu), "output_results");
  vh_ctx            *ctx = gpu;
  const RankTables  *T = tables.get();
  const int          counter = (int)iteration_loop;
  writer = std::thread([this, ctx, T, dirc, counter] {
    try
      {
        const double *sol = nullptr, *upd = nullptr;
        if (vh_snapshot_wait(ctx, &sol, &upd) != VH_OK)
          throw std::runtime_error(vh_last_error(ctx));
        write_vtu_piece(*T, 0, 1, dirc, "solution", counter, sol, upd);
      }
    catch (const std::exception &e)
      {
        writer_error = e.what();
      }
  });
  output_main_thread_ms += now_ms() - t0;
}

template <int dim>
void FemGL<dim>::run()
{
  std::ostream &pcout = *out;
  pcout << "Running using the B200 CUDA hot path (vh_femgl)." << std::endl;
  conf.enter_subsection("control parameters");
  const unsigned int n_cycles    = (unsigned int)conf.get_integer("Number of refinements");
  const unsigned int n_iteration = (unsigned int)conf.get_integer("Number of interations");
  std::vector<double> cycleX_refine_threshold = {conf.get_double("Cycle 0 refinement threshold"), conf.get_double("Cycle 1 refinement threshold"),
                                                 conf.get_double("Cycle 2 refinement threshold"), conf.get_double("Cycle 3 refinement threshold")};
  std::vector<bool>   cycleX_refine_strategy  = {conf.get_bool("Cycle 1 do global refinement"), conf.get_bool("Cycle 2 do global refinement"),
                                                 conf.get_bool("Cycle 3 do global refinement"), conf.get_bool("Cycle 4 do global refinement")};
  std::vector<double> cycleX_solve_tol        = {conf.get_double("Cycle 0 linear solver tol"), conf.get_double("Cycle 1 linear solver tol"),
                                                 conf.get_double("Cycle 2 linear solver tol"), conf.get_double("Cycle 3 linear solver tol"),
                                                 conf.get_double("Cycle 4 linear solver tol")};
  const double        converge_acc            = conf.get_double("converge accuracy");
  write_vtu                                   = conf.get_bool("write vtu output");
  conf.leave_subsection();
  if (n_cycles > 4)
    throw std::runtime_error("Number of refinements > 4: the reference declares tolerances for cycles 0..4 only");

  std::string ref_str;
  for (cycle = 0; cycle <= n_cycles; ++cycle)
    {
      pcout << "\n" << "Refinement Cycle is " << cycle << "\n"
            << "------------------------------------------------------" << "\n"
            << "------------------------------------------------------" << "\n" << std::endl;
      if (cycle == 0)
        {
          pcout << " cycle 0 will be globally refined in make_grid() call " << "\n" << std::endl;
          make_grid();
          setup_system();
          output_results("./setup_config/");
        }
      else
        {
          pcout << " 0th rank has active cells : " << triangulation->n_cells() << " cycleX_refine_strategy[cycle-1] is "
                << cycleX_refine_strategy[cycle - 1] << "\n" << std::endl;
          ref_str = cycleX_refine_strategy[cycle - 1] ? "global" : "adaptive";
          refine_grid(ref_str);
        }
      double residual_last_iter = 0.0;
      for (iteration_loop = 0; iteration_loop <= n_iteration; ++iteration_loop)
        {
          pcout << "Refinement cycle : " << cycle << ", " << "iteration_loop: " << iteration_loop << std::endl;
          StepRecord rec{};
          rec.cycle     = cycle;
          rec.iteration = iteration_loop;
          double t0     = now_ms();
          assemble_system();
          rec.t_assemble_ms = now_ms() - t0;
          pcout << " assembly is done !" << std::endl;
          t0 = now_ms();
          solve(cycleX_solve_tol[cycle]);
          rec.t_solve_ms = now_ms() - t0;
          pcout << " block-Jacobi preconditioned solving is done ! With solver_tol " << cycleX_solve_tol[cycle] << std::endl;
          t0 = now_ms();
          newton_iteration();
          rec.t_newton_ms = now_ms() - t0;
          pcout << " newton iteration is done !" << std::endl;
          output_results("./refine-cycle_" + std::to_string(cycle) + "/");
          rec.t_setup_ms = last_setup_ms; // context (re)build + state upload / transfer before the first step of a cycle
          last_setup_ms  = 0.0;
          rec.rhs_norm   = system_rhs_l2;
          rec.linear_its = last_linear_its;
          rec.residual   = residual_l2;
          rec.alpha      = last_alpha;
          rec.trials     = last_trials;
          check(vh_energy(gpu, 0, &rec.energy), "energy");
          records.push_back(rec);
          pcout << " timings [ms]: assembly " << rec.t_assemble_ms << ", solve " << rec.t_solve_ms << ", newton_iteration "
                << rec.t_newton_ms << ", free energy " << rec.energy << "\n" << std::endl;

          const double residual_l2_norm = residual_l2; // run.cc:234-250
          if ((std::fabs(residual_l2_norm - residual_last_iter) < cycleX_refine_threshold[cycle < 4 ? cycle : 3]) &&
              (residual_l2_norm > converge_acc) && (cycle < n_cycles))
            break;
          else if (residual_l2_norm <= converge_acc)
            break;
          else
            residual_last_iter = residual_l2_norm;
        }
      if (residual_l2 <= converge_acc)
        break;
    }
  finish_output();
  if (write_vtu)
    pcout << " output: the Newton loop waited " << output_main_thread_ms << " ms in total for output_results (snapshot + file asynchronous)"
          << std::endl;
  if (gpu)
    check(vh_get_solution(gpu, host_solution.data()), "get_solution");
}

template class FemGL<3>;
} // namespace vhhost
