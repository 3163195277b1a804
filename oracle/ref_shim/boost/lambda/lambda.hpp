// Stand-in for <boost/lambda/lambda.hpp> / <boost/lambda/bind.hpp> -- TEST INFRASTRUCTURE (see boost/multi_array.hpp in
// this tree).  The subset the reference's libMultiArray / libPictStruct sources spell out:
//   bind(&T::member_or_method, _1) > bind(&T::member_or_method, _2)     a comparator for std::sort (objectdetect_aux.cpp:249)
//   std::cout << _1 << " "                                              a printer for std::for_each
//   _1 = _1 / x                                                         inside a template that is never instantiated
#pragma once
#include <iostream>
#include <type_traits>
namespace boost {
namespace lambda {

template <int I>
struct placeholder {
  template <class A, class B>
  decltype(auto) pick(A &a, B &b) const {
    if constexpr (I == 1) return (a);
    else return (b);
  }
  template <class T> placeholder operator/(const T &) const { return *this; }
  template <class T> const placeholder &operator=(const T &) const { return *this; }
};
static const placeholder<1> _1 = placeholder<1>();
static const placeholder<2> _2 = placeholder<2>();

// bind(member pointer, placeholder): data members and nullary const member functions
template <class M, class T, int I>
struct bound_data {
  M T::*p;
  template <class A, class B> decltype(auto) operator()(const A &a, const B &b) const { return placeholder<I>().pick(a, b).*p; }
};
template <class Rr, class T, int I>
struct bound_method {
  Rr (T::*p)() const;
  template <class A, class B> decltype(auto) operator()(const A &a, const B &b) const { return (placeholder<I>().pick(a, b).*p)(); }
};
template <class M, class T, int I>
typename std::enable_if<!std::is_function<M>::value, bound_data<M, T, I> >::type bind(M T::*p, const placeholder<I> &) {
  return bound_data<M, T, I>{p};
}
template <class Rr, class T, int I>
bound_method<Rr, T, I> bind(Rr (T::*p)() const, const placeholder<I> &) {
  return bound_method<Rr, T, I>{p};
}
template <class L, class Rh>
struct greater_expr {
  L l;
  Rh r;
  template <class A, class B> bool operator()(const A &a, const B &b) const { return l(a, b) > r(a, b); }
};
template <class M1, class T1, int I1, class M2, class T2, int I2>
greater_expr<bound_data<M1, T1, I1>, bound_data<M2, T2, I2> > operator>(const bound_data<M1, T1, I1> &l, const bound_data<M2, T2, I2> &r) {
  return greater_expr<bound_data<M1, T1, I1>, bound_data<M2, T2, I2> >{l, r};
}
template <class M1, class T1, int I1, class M2, class T2, int I2>
greater_expr<bound_method<M1, T1, I1>, bound_method<M2, T2, I2> > operator>(const bound_method<M1, T1, I1> &l, const bound_method<M2, T2, I2> &r) {
  return greater_expr<bound_method<M1, T1, I1>, bound_method<M2, T2, I2> >{l, r};
}

// std::cout << _1 << " "
struct printer {
  std::ostream *os;
  const char *tail;
  template <class T> void operator()(const T &x) const { (*os) << x << (tail ? tail : ""); }
};
inline printer operator<<(std::ostream &os, const placeholder<1> &) { return printer{&os, 0}; }
inline printer operator<<(const printer &p, const char *tail) { return printer{p.os, tail}; }

}  // namespace lambda
}  // namespace boost
