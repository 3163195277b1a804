// Stand-in for libDiscPS/disc_ps.h -- TEST INFRASTRUCTURE: the one function objectdetect_findrot.cpp calls, declared
// as in the reference (disc_sample.cpp:123-138) and defined in oracle/ref_drivers.cpp.
#pragma once
#include <cstdlib>
namespace disc_ps {
void index_from_flat3(int shape0, int shape1, int shape2, int flat_idx, int &idx1, int &idx2, int &idx3);
template <class... A> double eval_joint_factor(const A &...) { abort(); }  // libDiscPS/factors.cpp is not part of oracle/_ref
}
