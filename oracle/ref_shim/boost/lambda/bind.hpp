// Stand-in -- TEST INFRASTRUCTURE.
#pragma once
#include <boost/lambda/lambda.hpp>
