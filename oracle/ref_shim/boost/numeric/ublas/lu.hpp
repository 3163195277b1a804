// stand-in, see ublas_min.hpp
#pragma once
#include "ublas_min.hpp"
