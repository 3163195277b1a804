// Stand-in for libPartDetect/partdetect.h -- TEST INFRASTRUCTURE.
#pragma once
#include <libPartDetect/partdef.h>
namespace part_detect {
const float NO_CLASS_VALUE = 0;
}
