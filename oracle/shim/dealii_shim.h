// TEST INFRASTRUCTURE — not product code.
//
// Minimal stand-in for the deal.II types that the reference's femgl.h and its
// 18 femgl/src/cell_mat_vec/*.cc term files mention, so that those files
// compile VERBATIM from /root/reference (they are never copied into this
// repository).  Only FullMatrix<double> / Vector<double> carry arithmetic; all
// other types are inert placeholders that let `class FemGL` be declared.
//
// Arithmetic semantics follow deal.II's documented FullMatrix interface:
//   A.mmult (C,B,adding)  : C (+)= A  * B
//   A.mTmult(C,B,adding)  : C (+)= A  * B^T
//   A.Tmmult(C,B,adding)  : C (+)= A^T* B
//   A.add(a,B)            : A += a*B ;  A.add(a,B,b,C) : A += a*B + b*C
#ifndef VH_ORACLE_DEALII_SHIM_H
#define VH_ORACLE_DEALII_SHIM_H

#define DEAL_II_WITH_TRILINOS 1

#include <cmath>
#include <cstddef>
#include <iostream>
#include <string>
#include <vector>

typedef int MPI_Comm;
#ifndef MPI_COMM_WORLD
#define MPI_COMM_WORLD 0
#endif

#define Assert(cond, exc) ((void)0)
#define AssertThrow(cond, exc) ((void)0)
#define ExcDimensionMismatch(a, b) 0
#define ExcInternalError() 0
#define ExcNotImplemented() 0

namespace dealii
{
namespace types
{
typedef unsigned long long global_dof_index;
typedef unsigned int       boundary_id;
} // namespace types

struct IdentityMatrix
{
  explicit IdentityMatrix(unsigned int n) : n(n) {}
  unsigned int n;
};

template <typename T>
class Vector
{
public:
  Vector() {}
  explicit Vector(unsigned int n) : v(n, T(0)) {}
  unsigned int size() const { return (unsigned int)v.size(); }
  T &operator()(unsigned int i) { return v[i]; }
  const T &operator()(unsigned int i) const { return v[i]; }
  T &operator[](unsigned int i) { return v[i]; }
  const T &operator[](unsigned int i) const { return v[i]; }
  Vector &operator=(T s)
  {
    for (auto &x : v)
      x = s;
    return *this;
  }
  Vector &operator+=(const Vector &o)
  {
    for (unsigned int i = 0; i < v.size(); ++i)
      v[i] += o.v[i];
    return *this;
  }
  T operator*(const Vector &o) const
  {
    T s = T(0);
    for (unsigned int i = 0; i < v.size(); ++i)
      s += v[i] * o.v[i];
    return s;
  }
  typename std::vector<T>::iterator begin() { return v.begin(); }
  typename std::vector<T>::iterator end() { return v.end(); }
  typename std::vector<T>::const_iterator begin() const { return v.begin(); }
  typename std::vector<T>::const_iterator end() const { return v.end(); }

private:
  std::vector<T> v;
};

template <typename T>
class FullMatrix
{
public:
  // deal.II: explicit FullMatrix(const size_type n = 0) -> square n x n.  The reference relies on it with a
  // double literal, `phi_ut_i_q_v0(3.3)` (cell_vec_rhs_beta3.cc:125), which converts to n = 3.
  explicit FullMatrix(unsigned int n = 0) : r(n), c(n), a(n * n, T(0)) {}
  FullMatrix(unsigned int rows, unsigned int cols) : r(rows), c(cols), a(rows * cols, T(0)) {}
  FullMatrix(const IdentityMatrix &id) : r(id.n), c(id.n), a(id.n * id.n, T(0))
  {
    for (unsigned int i = 0; i < r; ++i)
      a[i * c + i] = T(1);
  }
  unsigned int m() const { return r; }
  unsigned int n() const { return c; }
  T &operator()(unsigned int i, unsigned int j) { return a[i * c + j]; }
  const T &operator()(unsigned int i, unsigned int j) const { return a[i * c + j]; }
  void set(unsigned int i, unsigned int j, T v) { a[i * c + j] = v; }
  FullMatrix &operator=(T s)
  {
    for (auto &x : a)
      x = s;
    return *this;
  }
  T trace() const
  {
    T t = T(0);
    for (unsigned int i = 0; i < r; ++i)
      t += a[i * c + i];
    return t;
  }
  void add(T s, const FullMatrix &A)
  {
    for (unsigned int i = 0; i < a.size(); ++i)
      a[i] += s * A.a[i];
  }
  void add(T s, const FullMatrix &A, T t, const FullMatrix &B)
  {
    for (unsigned int i = 0; i < a.size(); ++i)
      a[i] += s * A.a[i] + t * B.a[i];
  }
  // C = this * B
  void mmult(FullMatrix &C, const FullMatrix &B, bool adding = false) const
  {
    for (unsigned int i = 0; i < r; ++i)
      for (unsigned int j = 0; j < B.c; ++j)
        {
          T s = adding ? C(i, j) : T(0);
          for (unsigned int k = 0; k < c; ++k)
            s += (*this)(i, k) * B(k, j);
          C(i, j) = s;
        }
  }
  // C = this * B^T
  void mTmult(FullMatrix &C, const FullMatrix &B, bool adding = false) const
  {
    for (unsigned int i = 0; i < r; ++i)
      for (unsigned int j = 0; j < B.r; ++j)
        {
          T s = adding ? C(i, j) : T(0);
          for (unsigned int k = 0; k < c; ++k)
            s += (*this)(i, k) * B(j, k);
          C(i, j) = s;
        }
  }
  // C = this^T * B
  void Tmmult(FullMatrix &C, const FullMatrix &B, bool adding = false) const
  {
    for (unsigned int i = 0; i < c; ++i)
      for (unsigned int j = 0; j < B.c; ++j)
        {
          T s = adding ? C(i, j) : T(0);
          for (unsigned int k = 0; k < r; ++k)
            s += (*this)(k, i) * B(k, j);
          C(i, j) = s;
        }
  }

private:
  unsigned int   r, c;
  std::vector<T> a;
};

// ---- inert placeholders (declaration of class FemGL only) ----
struct Subscriptor
{
};
struct ParameterHandler
{
};
template <int dim>
struct Point
{
  double x[dim];
  double operator()(unsigned int i) const { return x[i]; }
  double &operator()(unsigned int i) { return x[i]; }
  double operator[](unsigned int i) const { return x[i]; }
};
template <int dim>
class Function
{
public:
  explicit Function(unsigned int n_components = 1) : n_components(n_components) {}
  virtual ~Function() {}
  virtual void vector_value(const Point<dim> &, Vector<double> &) const {}
  virtual void vector_value_list(const std::vector<Point<dim>> &, std::vector<Vector<double>> &) const {}
  const unsigned int n_components;
};
struct ComponentMask
{
  ComponentMask() {}
  ComponentMask(const std::vector<bool> &) {}
};
namespace FEValuesExtractors
{
struct Scalar
{
  Scalar() : component(0) {}
  explicit Scalar(unsigned int c) : component(c) {}
  unsigned int component;
};
} // namespace FEValuesExtractors
struct IndexSet
{
};
template <typename T>
struct AffineConstraints
{
};
template <int dim>
struct FE_Q
{
  explicit FE_Q(unsigned int) {}
};
template <int dim>
struct FESystem
{
  FESystem(const FE_Q<dim> &, unsigned int) : degree(0) {}
  unsigned int degree;
  unsigned int n_dofs_per_cell() const { return 0; }
};
template <int dim>
struct Triangulation
{
  enum MeshSmoothing
  {
    none                    = 0,
    smoothing_on_refinement = 1,
    smoothing_on_coarsening = 2
  };
};
namespace parallel
{
namespace distributed
{
template <int dim>
struct Triangulation
{
  Triangulation(MPI_Comm, typename dealii::Triangulation<dim>::MeshSmoothing) {}
};
} // namespace distributed
} // namespace parallel
template <int dim>
struct DoFHandler
{
  DoFHandler() {}
  template <typename Tria>
  explicit DoFHandler(const Tria &)
  {}
};
template <int dim>
struct FEValues
{
};
template <int dim>
struct FEFaceValues
{
};
struct ConditionalOStream
{
  ConditionalOStream(std::ostream &, bool) {}
};
struct TimerOutput
{
  enum OutputFrequency
  {
    summary
  };
  enum OutputType
  {
    wall_times
  };
  TimerOutput(MPI_Comm, ConditionalOStream &, OutputFrequency, OutputType) {}
};
namespace LinearAlgebraTrilinos
{
namespace MPI
{
struct SparseMatrix
{
};
struct Vector
{
};
struct PreconditionAMG
{
};
} // namespace MPI
} // namespace LinearAlgebraTrilinos
} // namespace dealii

#endif
