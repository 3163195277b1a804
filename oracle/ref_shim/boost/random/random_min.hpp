// Stand-in for the Boost.Random headers libPictStruct/objectdetect_aux.cpp includes -- TEST INFRASTRUCTURE.
// Only names: the sampling helpers that use them are outside what oracle/_ref pins and are never called.
#pragma once
#include <cstdlib>
namespace boost {
class mt19937 {
 public:
  typedef unsigned result_type;
  explicit mt19937(unsigned = 0) {}
  unsigned operator()() { abort(); }
};
template <class T = double>
class uniform_real {
 public:
  uniform_real(T = 0, T = 1) {}
};
template <class Engine, class Dist>
class variate_generator {
 public:
  variate_generator(Engine, Dist) {}
  double operator()() { abort(); }
};
}  // namespace boost
