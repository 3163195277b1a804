// Stand-in -- TEST INFRASTRUCTURE, see random_min.hpp.
#pragma once
#include "random_min.hpp"
