// TEST-ONLY type-level stand-in for <RcppEigen.h>: just enough of Rcpp's and Eigen's interfaces for
// `g++ -fsyntax-only` to parse and TYPE-CHECK flgp_b200/r_shim/flgp_shim.cpp together with the reference's own
// headers (tests/test_abi_cpu.py::test_r_shim_typechecks_against_the_reference_headers).  The image has no R, Rcpp or
// Eigen; nothing here computes anything and nothing is ever linked.  Member sets are the ones the shim and the
// reference headers use; semantics (shapes, aliasing) are not modelled.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <limits>
#include <numeric>
#include <string>
#include <type_traits>
#include <vector>

#define NA_REAL (0.0 / 0.0)

namespace Eigen {
typedef std::ptrdiff_t Index;
enum { ColMajor = 0, RowMajor = 1, RowMajorBit = 1, Dynamic = -1 };

template <class Derived>
struct MatrixBase {
  const Derived& derived() const { return static_cast<const Derived&>(*this); }
};
template <class Func, class MatrixType>
struct CwiseNullaryOp {};

template <class S, int R, int C, int Opt = 0, int MR = R, int MC = C>
class Matrix : public MatrixBase<Matrix<S, R, C, Opt, MR, MC>> {
 public:
  typedef S Scalar;
  enum { SizeAtCompileTime = -1, MaxSizeAtCompileTime = -1, Flags = Opt };
  Matrix() {}
  explicit Matrix(Index) {}
  Matrix(Index, Index) {}
  template <class F, class M>
  Matrix(const CwiseNullaryOp<F, M>&) {}
  S* data() { return nullptr; }
  const S* data() const { return nullptr; }
  Index rows() const { return 0; }
  Index cols() const { return 0; }
  Index size() const { return 0; }
  void resize(Index) {}
  void resize(Index, Index) {}
  S& operator()(Index, Index = 0) { return *data(); }
  const S& operator()(Index, Index = 0) const { return *data(); }
  S& operator[](Index) { return *data(); }
  const S& operator[](Index) const { return *data(); }
  Matrix<S, Dynamic, Dynamic> topRows(Index) const { return {}; }
  Matrix<S, Dynamic, Dynamic> bottomRows(Index) const { return {}; }
  Matrix<S, Dynamic, 1> head(Index) const { return {}; }
  Matrix<S, Dynamic, 1> tail(Index) const { return {}; }
  Matrix<S, Dynamic, 1> col(Index) const { return {}; }
  static Matrix Constant(Index, S) { return {}; }
  static Matrix Constant(Index, Index, S) { return {}; }
  static Matrix Zero(Index, Index = 1) { return {}; }
  static Matrix LinSpaced(Index, S, S) { return {}; }
  template <class F>
  static CwiseNullaryOp<F, Matrix> NullaryExpr(Index, Index, const F&) { return {}; }
  // any dense matrix converts to any other (Eigen checks shapes at run time)
  template <class S2, int R2, int C2, int O2, int MR2, int MC2>
  Matrix(const Matrix<S2, R2, C2, O2, MR2, MC2>&) {}
};
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 1, Dynamic> RowVectorXd;
typedef Matrix<int, Dynamic, Dynamic> MatrixXi;
typedef Matrix<int, Dynamic, 1> VectorXi;

template <class M>
class Map : public M {
 public:
  Map() {}
  Map(typename M::Scalar*, Index, Index = 1) {}
};

template <class S, int Opt = 0>
class SparseMatrix {
 public:
  struct Ref {
    Ref& operator=(S) { return *this; }
  };
  SparseMatrix() {}
  SparseMatrix(Index, Index) {}
  void reserve(const VectorXi&) {}
  Ref insert(Index, Index) { return {}; }
  void makeCompressed() {}
  Index rows() const { return 0; }
  Index cols() const { return 0; }
  Index nonZeros() const { return 0; }
  int* innerIndexPtr() { return nullptr; }
  const int* innerIndexPtr() const { return nullptr; }
  S* valuePtr() { return nullptr; }
  const S* valuePtr() const { return nullptr; }
};
}  // namespace Eigen

namespace Rcpp {
[[noreturn]] inline void stop(const char*) { throw 0; }
[[noreturn]] inline void stop(const std::string&) { throw 0; }
static std::ostream& Rcout = std::cout;

class String {
 public:
  String(const char*) {}
  String(const std::string&) {}
  operator std::string() const { return {}; }
};

struct NamedValue {
  template <class T>
  NamedValue(const char*, const T&) {}
};
struct Named {
  const char* name;
  explicit Named(const char* n) : name(n) {}
  template <class T>
  NamedValue operator=(const T& v) const { return NamedValue(name, v); }
};

class List {
 public:
  struct Proxy {  // element access: assignable from anything wrap() accepts, readable through Rcpp::as<T>
    template <class T>
    Proxy& operator=(const T&) { return *this; }
    template <class T>
    operator T() const { return T(); }
  };
  List() {}
  template <class... A>
  static List create(const A&...) {
    static_assert((std::is_same<A, NamedValue>::value && ...), "List::create takes Named(...) = value arguments");
    return {};
  }
  Proxy operator[](const char*) const { return {}; }
  Proxy operator[](const std::string&) const { return {}; }
};

class NumericMatrix {
 public:
  int nrow() const { return 0; }
  int ncol() const { return 0; }
};
class NumericVector {
 public:
  Eigen::Index size() const { return 0; }
};
class IntegerVector {
 public:
  const int* begin() const { return nullptr; }
  const int* end() const { return nullptr; }
};
inline IntegerVector sample(int, int) { return {}; }

class Function {
 public:
  Function() {}
  template <class... A>
  List::Proxy operator()(const A&...) const { return {}; }
};
class Environment {
 public:
  static Environment namespace_env(const std::string&) { return {}; }
  Function operator[](const char*) const { return {}; }
};

template <class T>
T as(const NumericMatrix&) { return T(); }
template <class T>
T as(const NumericVector&) { return T(); }
template <class T>
T as(const List::Proxy&) { return T(); }
template <class T>
T as(const List&) { return T(); }
}  // namespace Rcpp
