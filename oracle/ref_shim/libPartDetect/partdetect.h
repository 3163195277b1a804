// Stand-in for libPartDetect/partdetect.h -- TEST INFRASTRUCTURE.
#pragma once
#include <cstdlib>
#include <libPartDetect/partdef.h>
namespace part_detect {
const float NO_CLASS_VALUE = 0;
template <class... A> void partdetect(const A &...) { abort(); }
template <class... A> int getNumPartTypes(const A &...) { return 1; }
template <class... A> int getPartById(const A &...) { abort(); }
template <class... A> void runMatlabCode(const A &...) { abort(); }  // the detector is not part of oracle/_ref
}
