// Minimal stand-in for <boost/multi_array.hpp> -- TEST INFRASTRUCTURE, written for this repo.
// Boost is not installed in this image.  This header implements just enough of the Boost.MultiArray interface
// (C-order dense arrays, extents / indices generators, strided views, element-wise assignment between array-likes) for
// the reference's own libMultiArray / libBoostMath sources to compile UNMODIFIED from /root/reference into
// oracle/_ref/ (see oracle/Makefile, target `ref`).  Only containers live here: every arithmetic rule that the
// oracle restates comes from the reference's files.
#pragma once
#include <algorithm>  // the real header pulls these in; the reference relies on it
#include <cassert>
#include <cstddef>
#include <functional>
#include <iostream>
#include <numeric>
#include <vector>

namespace boost {

struct c_storage_order {
  bool operator==(const c_storage_order &) const { return true; }
  bool operator!=(const c_storage_order &) const { return false; }
};

namespace multi_array_types {
typedef std::ptrdiff_t index;
typedef std::size_t size_type;
struct index_range {
  index start_, finish_;
  bool all_;
  index_range() : start_(0), finish_(0), all_(true) {}
  index_range(index s, index f) : start_(s), finish_(f), all_(false) {}
};
}  // namespace multi_array_types

namespace detail {
namespace multi_array {
using boost::multi_array_types::index;
using boost::multi_array_types::index_range;

template <int N>
struct extent_gen {
  std::size_t e[N ? N : 1];
  extent_gen<N + 1> operator[](std::size_t n) const {
    extent_gen<N + 1> r;
    for (int i = 0; i < N; ++i) r.e[i] = e[i];
    r.e[N] = n;
    return r;
  }
};

// N subscripts given so far, ND of them ranges (= dimensionality of the resulting view)
template <int N, int ND>
struct index_gen {
  index_range r[N ? N : 1];
  bool is_index[N ? N : 1];
  index_gen<N + 1, ND + 1> operator[](const index_range &x) const {
    index_gen<N + 1, ND + 1> g;
    for (int i = 0; i < N; ++i) { g.r[i] = r[i]; g.is_index[i] = is_index[i]; }
    g.r[N] = x;
    g.is_index[N] = false;
    return g;
  }
  index_gen<N + 1, ND> operator[](index i0) const {
    index_gen<N + 1, ND> g;
    for (int i = 0; i < N; ++i) { g.r[i] = r[i]; g.is_index[i] = is_index[i]; }
    g.r[N] = index_range(i0, i0 + 1);
    g.is_index[N] = true;
    return g;
  }
};

// element-wise copy between two array-likes of the same dimensionality, through chained operator[]
template <int N>
struct copier {
  template <class D, class S>
  static void run(D dst, const S &src, const std::size_t *shape) {
    for (std::size_t i = 0; i < shape[0]; ++i) copier<N - 1>::run(dst[(index)i], src[(index)i], shape + 1);
  }
};
template <>
struct copier<1> {
  template <class D, class S>
  static void run(D dst, const S &src, const std::size_t *shape) {
    for (std::size_t i = 0; i < shape[0]; ++i) dst[(index)i] = src[(index)i];
  }
};

// pointer + shape + strides: sub-arrays (a[i]) and views (a[indices[...]]) are both this
template <class T, int N>
class strided_ref {
 public:
  typedef T element;
  static const int dimensionality = N;
  T *base_;
  std::size_t shape_[N];
  index stride_[N];
  const std::size_t *shape() const { return shape_; }
  std::size_t num_elements() const {
    std::size_t n = 1;
    for (int i = 0; i < N; ++i) n *= shape_[i];
    return n;
  }
  strided_ref<T, N - 1> operator[](index i) const {
    assert(i >= 0 && (std::size_t)i < shape_[0]);
    strided_ref<T, N - 1> r;
    r.base_ = base_ + i * stride_[0];
    for (int k = 1; k < N; ++k) { r.shape_[k - 1] = shape_[k]; r.stride_[k - 1] = stride_[k]; }
    return r;
  }
  template <class S>
  const strided_ref &operator=(const S &src) const {
    for (int k = 0; k < N; ++k) assert(src.shape()[k] == shape_[k]);
    copier<N>::run(*this, src, shape_);
    return *this;
  }
  const strided_ref &operator=(const strided_ref &src) const {
    copier<N>::run(*this, src, shape_);
    return *this;
  }
};
template <class T>
class strided_ref<T, 1> {
 public:
  typedef T element;
  static const int dimensionality = 1;
  T *base_;
  std::size_t shape_[1];
  index stride_[1];
  const std::size_t *shape() const { return shape_; }
  std::size_t num_elements() const { return shape_[0]; }
  T &operator[](index i) const {
    assert(i >= 0 && (std::size_t)i < shape_[0]);
    return base_[i * stride_[0]];
  }
  template <class S>
  const strided_ref &operator=(const S &src) const {
    assert(src.shape()[0] == shape_[0]);
    for (std::size_t i = 0; i < shape_[0]; ++i) (*this)[(index)i] = src[(index)i];
    return *this;
  }
  const strided_ref &operator=(const strided_ref &src) const {
    for (std::size_t i = 0; i < shape_[0]; ++i) (*this)[(index)i] = src[(index)i];
    return *this;
  }
};

template <class T, int N, int M, int ND>
strided_ref<T, ND> make_view(T *base, const std::size_t *shape, const index *stride, const index_gen<M, ND> &g) {
  static_assert(M == N, "one subscript per dimension");
  strided_ref<T, ND> v;
  v.base_ = base;
  int d = 0;
  for (int k = 0; k < N; ++k) {
    const index s = g.r[k].all_ ? 0 : g.r[k].start_;
    const index f = g.r[k].all_ ? (index)shape[k] : g.r[k].finish_;
    assert(s >= 0 && f <= (index)shape[k] && s <= f);
    v.base_ += s * stride[k];
    if (!g.is_index[k]) {
      v.shape_[d] = (std::size_t)(f - s);
      v.stride_[d] = stride[k];
      ++d;
    }
  }
  return v;
}
}  // namespace multi_array
}  // namespace detail

static const detail::multi_array::extent_gen<0> extents = detail::multi_array::extent_gen<0>();
static const detail::multi_array::index_gen<0, 0> indices = detail::multi_array::index_gen<0, 0>();

template <class T, std::size_t NN>
class multi_array {
  enum { N = (int)NN };

 public:
  typedef T element;
  typedef multi_array_types::index index;
  typedef multi_array_types::size_type size_type;
  static const int dimensionality = (int)NN;

  multi_array() { set_shape(nullptr); }
  explicit multi_array(const detail::multi_array::extent_gen<(int)NN> &e) { set_shape(e.e); v_.assign(count(), T()); }
  multi_array(const multi_array &o) : v_(o.v_) { set_shape(o.shape_); }
  template <class S>
  multi_array(const S &src) {  // from a view / sub-array / other array of the same dimensionality
    static_assert(S::dimensionality == (int)NN, "dimensionality");
    set_shape(src.shape());
    v_.assign(count(), T());
    detail::multi_array::copier<(int)NN>::run(ref(), src, shape_);
  }
  multi_array &operator=(const multi_array &o) {
    // Boost asserts equal shapes here; the reference only assigns equal shapes or freshly default-constructed targets
    if (count() == 0) set_shape(o.shape_);
    for (int k = 0; k < N; ++k) assert(shape_[k] == o.shape_[k]);
    v_ = o.v_;
    return *this;
  }
  template <class S>
  multi_array &operator=(const S &src) {
    static_assert(S::dimensionality == (int)NN, "dimensionality");
    if (count() == 0) { set_shape(src.shape()); v_.assign(count(), T()); }
    for (int k = 0; k < N; ++k) assert(shape_[k] == src.shape()[k]);
    detail::multi_array::copier<(int)NN>::run(ref(), src, shape_);
    return *this;
  }
  void resize(const detail::multi_array::extent_gen<(int)NN> &e) {  // contents are not preserved (the callers do not rely on it)
    set_shape(e.e);
    v_.assign(count(), T());
  }
  const size_type *shape() const { return shape_; }
  c_storage_order storage_order() const { return c_storage_order(); }
  size_type num_elements() const { return v_.size(); }
  T *data() { return v_.data(); }
  const T *data() const { return v_.data(); }
  T *origin() { return v_.data(); }
  const T *origin() const { return v_.data(); }

  decltype(auto) operator[](index i) { return ref()[i]; }
  decltype(auto) operator[](index i) const { return cref()[i]; }
  template <int M, int ND>
  detail::multi_array::strided_ref<T, ND> operator[](const detail::multi_array::index_gen<M, ND> &g) {
    return detail::multi_array::make_view<T, (int)NN>(v_.data(), shape_, stride_, g);
  }
  template <int M, int ND>
  detail::multi_array::strided_ref<const T, ND> operator[](const detail::multi_array::index_gen<M, ND> &g) const {
    return detail::multi_array::make_view<const T, (int)NN>(v_.data(), shape_, stride_, g);
  }

 private:
  std::vector<T> v_;
  size_type shape_[NN];
  index stride_[NN];
  size_type count() const {
    size_type n = 1;
    for (int k = 0; k < N; ++k) n *= shape_[k];
    return n;
  }
  void set_shape(const size_type *s) {
    for (int k = 0; k < N; ++k) shape_[k] = s ? s[k] : 0;
    index st = 1;
    for (int k = N - 1; k >= 0; --k) { stride_[k] = st; st *= (index)shape_[k]; }
  }
  detail::multi_array::strided_ref<T, (int)NN> ref() {
    detail::multi_array::strided_ref<T, (int)NN> r;
    r.base_ = v_.data();
    for (int k = 0; k < N; ++k) { r.shape_[k] = shape_[k]; r.stride_[k] = stride_[k]; }
    return r;
  }
  detail::multi_array::strided_ref<const T, (int)NN> cref() const {
    detail::multi_array::strided_ref<const T, (int)NN> r;
    r.base_ = v_.data();
    for (int k = 0; k < N; ++k) { r.shape_[k] = shape_[k]; r.stride_[k] = stride_[k]; }
    return r;
  }
};

template <class Array, int ND>
struct array_view_gen {
  typedef detail::multi_array::strided_ref<typename Array::element, ND> type;
};
template <class Array, int ND>
struct const_array_view_gen {
  typedef detail::multi_array::strided_ref<const typename Array::element, ND> type;
};

}  // namespace boost
