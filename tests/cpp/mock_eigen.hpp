// Minimal dense matrix / vector stand-in with the slice of Eigen's interface the overlay uses
// (rows, cols, operator(), resize).  Test infrastructure: Eigen is not installed in this image.
#pragma once
#include <vector>
namespace mock {
template<typename S> struct Matrix
{
  long r{0}, c{0};
  std::vector<S> d;
  Matrix() = default;
  Matrix(long rr, long cc) : r(rr), c(cc), d(rr * cc, S(0)) {}
  long rows() const { return r; }
  long cols() const { return c; }
  void resize(long rr, long cc) { r = rr; c = cc; d.assign(rr * cc, S(0)); }
  S & operator()(long i, long j) { return d[i + r * j]; }
  const S & operator()(long i, long j) const { return d[i + r * j]; }
};
template<typename S> struct Vector
{
  std::vector<S> d;
  Vector() = default;
  explicit Vector(long n) : d(n, S(0)) {}
  long size() const { return static_cast<long>(d.size()); }
  long rows() const { return size(); }
  void resize(long n) { d.resize(n, S(0)); }
  S & operator()(long i) { return d[i]; }
  const S & operator()(long i) const { return d[i]; }
};
// compressed sparse matrix with Eigen::SparseMatrix's raw-pointer interface (outer = columns for P, rows for A)
template<typename S> struct SparseMatrix
{
  long r{0}, c{0};
  std::vector<int> outer, inner;
  std::vector<S> vals;
  long rows() const { return r; }
  long cols() const { return c; }
  long nonZeros() const { return static_cast<long>(vals.size()); }
  const int * outerIndexPtr() const { return outer.data(); }
  const int * innerIndexPtr() const { return inner.data(); }
  const S * valuePtr() const { return vals.data(); }
  bool isCompressed() const { return true; }
};
template<typename S = double> struct QuadraticProgramSparse  // qp.hpp:60-79: P column-major, A row-major
{
  SparseMatrix<S> P; Vector<S> q; SparseMatrix<S> A; Vector<S> l, u;
};
template<typename S = double> struct QuadraticProgram  // qp.hpp:31-45
{
  Matrix<S> P; Vector<S> q; Matrix<S> A; Vector<S> l, u;
};
}  // namespace mock
