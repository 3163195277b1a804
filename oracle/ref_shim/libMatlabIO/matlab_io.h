// Stand-in -- TEST INFRASTRUCTURE.
#pragma once
#include <libMatlabIO/matlab_io.hpp>
