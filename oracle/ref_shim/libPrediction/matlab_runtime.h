// Stand-in for libPrediction/matlab_runtime.h -- TEST INFRASTRUCTURE (the MATLAB runtime is not installed).
#pragma once
#define MATLAB_RUNTIME ""
