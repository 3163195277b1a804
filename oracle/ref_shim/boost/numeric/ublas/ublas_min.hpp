// Minimal stand-in for the Boost.uBLAS headers the reference's libBoostMath / libMultiArray include -- TEST
// INFRASTRUCTURE, written for this repo (Boost is not installed here; see boost/multi_array.hpp next to this file).
// Dense row-major matrix / vector, eager evaluation.  prod() accumulates `t = 0; t += a(i,k) * b(k,j)` for ascending k,
// which is what uBLAS's dense matrix_matrix_binary / matrix_vector_binary functors do.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <vector>

namespace boost {
namespace numeric {
namespace ublas {

template <class T>
class vector {
 public:
  typedef T value_type;
  typedef typename std::vector<T>::iterator iterator;
  typedef typename std::vector<T>::const_iterator const_iterator;
  vector() {}
  explicit vector(std::size_t n) : v_(n) {}
  vector(std::size_t n, const T &x) : v_(n, x) {}
  std::size_t size() const { return v_.size(); }
  void resize(std::size_t n, bool preserve = true) {
    if (preserve) v_.resize(n); else v_.assign(n, T());
  }
  T &operator()(std::size_t i) { assert(i < v_.size()); return v_[i]; }
  const T &operator()(std::size_t i) const { assert(i < v_.size()); return v_[i]; }
  T &operator[](std::size_t i) { assert(i < v_.size()); return v_[i]; }
  const T &operator[](std::size_t i) const { assert(i < v_.size()); return v_[i]; }
  iterator begin() { return v_.begin(); }
  iterator end() { return v_.end(); }
  const_iterator begin() const { return v_.begin(); }
  const_iterator end() const { return v_.end(); }
  vector &operator*=(const T &s) { for (T &x : v_) x *= s; return *this; }
  vector &operator/=(const T &s) { for (T &x : v_) x /= s; return *this; }
  vector &operator+=(const vector &o) { assert(o.size() == size()); for (std::size_t i = 0; i < size(); ++i) v_[i] += o.v_[i]; return *this; }
  vector &operator-=(const vector &o) { assert(o.size() == size()); for (std::size_t i = 0; i < size(); ++i) v_[i] -= o.v_[i]; return *this; }

 private:
  std::vector<T> v_;
};
template <class T>
struct zero_vector : vector<T> {
  zero_vector() {}
  explicit zero_vector(std::size_t n) : vector<T>(n, T(0)) {}
};
template <class T>
struct scalar_vector : vector<T> {
  scalar_vector() {}
  scalar_vector(std::size_t n, const T &x) : vector<T>(n, x) {}
};

template <class T>
class matrix {
 public:
  typedef T value_type;
  matrix() : r_(0), c_(0) {}
  matrix(std::size_t r, std::size_t c) : r_(r), c_(c), v_(r * c) {}
  matrix(std::size_t r, std::size_t c, const T &x) : r_(r), c_(c), v_(r * c, x) {}
  std::size_t size1() const { return r_; }
  std::size_t size2() const { return c_; }
  void resize(std::size_t r, std::size_t c, bool preserve = true) {
    std::vector<T> n(r * c, T());
    if (preserve)
      for (std::size_t i = 0; i < std::min(r, r_); ++i)
        for (std::size_t j = 0; j < std::min(c, c_); ++j) n[i * c + j] = v_[i * c_ + j];
    v_.swap(n);
    r_ = r;
    c_ = c;
  }
  T &operator()(std::size_t i, std::size_t j) { assert(i < r_ && j < c_); return v_[i * c_ + j]; }
  const T &operator()(std::size_t i, std::size_t j) const { assert(i < r_ && j < c_); return v_[i * c_ + j]; }
  matrix &operator*=(const T &s) { for (T &x : v_) x *= s; return *this; }
  matrix &operator/=(const T &s) { for (T &x : v_) x /= s; return *this; }
  matrix &operator+=(const matrix &o) { assert(o.r_ == r_ && o.c_ == c_); for (std::size_t i = 0; i < v_.size(); ++i) v_[i] += o.v_[i]; return *this; }
  matrix &assign(const matrix &o) { *this = o; return *this; }

 private:
  std::size_t r_, c_;
  std::vector<T> v_;
};
template <class T>
struct zero_matrix : matrix<T> {
  zero_matrix() {}
  zero_matrix(std::size_t r, std::size_t c) : matrix<T>(r, c, T(0)) {}
};
template <class T>
struct scalar_matrix : matrix<T> {
  scalar_matrix() {}
  scalar_matrix(std::size_t r, std::size_t c, const T &x) : matrix<T>(r, c, x) {}
};
template <class T>
struct identity_matrix : matrix<T> {
  identity_matrix() {}
  explicit identity_matrix(std::size_t n) : matrix<T>(n, n, T(0)) { for (std::size_t i = 0; i < n; ++i) (*this)(i, i) = T(1); }
  identity_matrix(std::size_t r, std::size_t c) : matrix<T>(r, c, T(0)) { for (std::size_t i = 0; i < std::min(r, c); ++i) (*this)(i, i) = T(1); }
};

// ---- proxies: subrange / row / column (assignable, readable) -------------------------------------------------------
template <class T>
class matrix_range {
 public:
  typedef T value_type;
  matrix_range(matrix<T> &m, std::size_t r0, std::size_t r1, std::size_t c0, std::size_t c1) : m_(m), r0_(r0), r1_(r1), c0_(c0), c1_(c1) {
    assert(r0 <= r1 && r1 <= m.size1() && c0 <= c1 && c1 <= m.size2());
  }
  std::size_t size1() const { return r1_ - r0_; }
  std::size_t size2() const { return c1_ - c0_; }
  T &operator()(std::size_t i, std::size_t j) const { return m_(r0_ + i, c0_ + j); }
  const matrix_range &operator=(const matrix<T> &o) const {
    assert(o.size1() == size1() && o.size2() == size2());
    for (std::size_t i = 0; i < size1(); ++i)
      for (std::size_t j = 0; j < size2(); ++j) (*this)(i, j) = o(i, j);
    return *this;
  }
  operator matrix<T>() const {
    matrix<T> r(size1(), size2());
    for (std::size_t i = 0; i < size1(); ++i)
      for (std::size_t j = 0; j < size2(); ++j) r(i, j) = (*this)(i, j);
    return r;
  }

 private:
  matrix<T> &m_;
  std::size_t r0_, r1_, c0_, c1_;
};
template <class T>
matrix_range<T> subrange(matrix<T> &m, std::size_t r0, std::size_t r1, std::size_t c0, std::size_t c1) {
  return matrix_range<T>(m, r0, r1, c0, c1);
}
template <class T>
matrix_range<T> subrange(const matrix<T> &m, std::size_t r0, std::size_t r1, std::size_t c0, std::size_t c1) {
  return matrix_range<T>(const_cast<matrix<T> &>(m), r0, r1, c0, c1);  // read-only use by the callers
}

template <class T>
class matrix_row {
 public:
  typedef T value_type;
  matrix_row(matrix<T> &m, std::size_t i) : m_(m), i_(i) {}
  std::size_t size() const { return m_.size2(); }
  T &operator()(std::size_t j) const { return m_(i_, j); }
  const matrix_row &operator=(const vector<T> &v) const {
    assert(v.size() == size());
    for (std::size_t j = 0; j < size(); ++j) m_(i_, j) = v(j);
    return *this;
  }

 private:
  matrix<T> &m_;
  std::size_t i_;
};
template <class T>
matrix_row<T> row(matrix<T> &m, std::size_t i) { return matrix_row<T>(m, i); }

template <class T>
class matrix_column {
 public:
  typedef T value_type;
  matrix_column(const matrix<T> &m, std::size_t j) : v_(m.size1()) {
    for (std::size_t i = 0; i < m.size1(); ++i) v_[i] = m(i, j);
  }
  std::size_t size() const { return v_.size(); }
  typename std::vector<T>::const_iterator begin() const { return v_.begin(); }
  typename std::vector<T>::const_iterator end() const { return v_.end(); }
  const T &operator()(std::size_t i) const { return v_[i]; }

 private:
  std::vector<T> v_;  // a copy: the callers only read
};
template <class T>
matrix_column<T> column(const matrix<T> &m, std::size_t j) { return matrix_column<T>(m, j); }

// ---- operations -----------------------------------------------------------------------------------------------------
template <class A, class B>
matrix<typename A::value_type> prod_mm(const A &a, const B &b) {
  typedef typename A::value_type T;
  assert(a.size2() == b.size1());
  matrix<T> r(a.size1(), b.size2());
  for (std::size_t i = 0; i < a.size1(); ++i)
    for (std::size_t j = 0; j < b.size2(); ++j) {
      T t = T(0);
      for (std::size_t k = 0; k < a.size2(); ++k) t += a(i, k) * b(k, j);
      r(i, j) = t;
    }
  return r;
}
template <class T> matrix<T> prod(const matrix<T> &a, const matrix<T> &b) { return prod_mm(a, b); }
template <class T> matrix<T> prod(const matrix_range<T> &a, const matrix_range<T> &b) { return prod_mm(a, b); }
template <class T> matrix<T> prod(const matrix<T> &a, const matrix_range<T> &b) { return prod_mm(a, b); }
template <class T> matrix<T> prod(const matrix_range<T> &a, const matrix<T> &b) { return prod_mm(a, b); }
template <class T>
vector<T> prod(const matrix<T> &a, const vector<T> &x) {
  assert(a.size2() == x.size());
  vector<T> r(a.size1());
  for (std::size_t i = 0; i < a.size1(); ++i) {
    T t = T(0);
    for (std::size_t k = 0; k < a.size2(); ++k) t += a(i, k) * x(k);
    r(i) = t;
  }
  return r;
}
template <class T>
matrix<T> trans(const matrix<T> &a) {
  matrix<T> r(a.size2(), a.size1());
  for (std::size_t i = 0; i < a.size1(); ++i)
    for (std::size_t j = 0; j < a.size2(); ++j) r(j, i) = a(i, j);
  return r;
}
template <class T> matrix<T> operator-(const matrix<T> &a) { matrix<T> r(a.size1(), a.size2()); for (std::size_t i = 0; i < r.size1(); ++i) for (std::size_t j = 0; j < r.size2(); ++j) r(i, j) = -a(i, j); return r; }
template <class T> matrix<T> operator+(const matrix<T> &a, const matrix<T> &b) { matrix<T> r(a); r += b; return r; }
template <class T> matrix<T> operator*(const matrix<T> &a, const T &s) { matrix<T> r(a); r *= s; return r; }
template <class T> matrix<T> operator*(const T &s, const matrix<T> &a) { matrix<T> r(a); for (std::size_t i = 0; i < r.size1(); ++i) for (std::size_t j = 0; j < r.size2(); ++j) r(i, j) = s * a(i, j); return r; }
template <class T> vector<T> operator+(const vector<T> &a, const vector<T> &b) { vector<T> r(a); r += b; return r; }
template <class T> vector<T> operator-(const vector<T> &a, const vector<T> &b) { vector<T> r(a); r -= b; return r; }
template <class T> vector<T> operator-(const vector<T> &a) { vector<T> r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r(i) = -a(i); return r; }
template <class T> vector<T> operator*(const vector<T> &a, const T &s) { vector<T> r(a); r *= s; return r; }
template <class T> vector<T> operator*(const T &s, const vector<T> &a) { vector<T> r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r(i) = s * a(i); return r; }
// uBLAS promotes an int scalar against a double vector (partdef.cpp:381 multiplies by an int axis length)
inline vector<double> operator*(int s, const vector<double> &a) { return (double)s * a; }
inline vector<double> operator*(const vector<double> &a, int s) { return a * (double)s; }
template <class T> vector<T> operator/(const vector<T> &a, const T &s) { vector<T> r(a); r /= s; return r; }
template <class T>
T inner_prod(const vector<T> &a, const vector<T> &b) {
  assert(a.size() == b.size());
  T t = T(0);
  for (std::size_t i = 0; i < a.size(); ++i) t += a(i) * b(i);
  return t;
}
template <class T>
T norm_1(const vector<T> &a) {
  T t = T(0);
  for (std::size_t i = 0; i < a.size(); ++i) t += std::abs(a(i));
  return t;
}
template <class T>
T norm_2(const vector<T> &a) {
  T t = T(0);
  for (std::size_t i = 0; i < a.size(); ++i) t += a(i) * a(i);
  return std::sqrt(t);
}

// names that appear in never-instantiated templates of boost_math.hpp (det / inv for n > 2)
template <class T>
class permutation_matrix : public vector<T> {
 public:
  explicit permutation_matrix(std::size_t n) : vector<T>(n) { for (std::size_t i = 0; i < n; ++i) (*this)(i) = (T)i; }
};

// LU for n > 2 is reached by nothing on the pictorial-structures path (2 x 2 covariances take the closed forms)
template <class M, class P> int lu_factorize(M &, P &) { std::abort(); }
template <class M, class P, class X> void lu_substitute(const M &, const P &, X &) { std::abort(); }

}  // namespace ublas
}  // namespace numeric
}  // namespace boost
