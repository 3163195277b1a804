// Stand-in for <boost/lambda/lambda.hpp> -- TEST INFRASTRUCTURE (see boost/multi_array.hpp next to this tree).
// multi_array_op.hpp names the placeholder `_1` inside a template (normalize) that the oracle build never instantiates.
#pragma once
namespace boost {
namespace lambda {
struct placeholder1 {
  template <class T> placeholder1 operator/(const T &) const { return *this; }
  template <class T> const placeholder1 &operator=(const T &) const { return *this; }
  template <class T> void operator()(T &) const {}
};
static const placeholder1 _1 = placeholder1();
}  // namespace lambda
}  // namespace boost
