// Stand-in -- TEST INFRASTRUCTURE: the filesys:: names live in the libPartApp/partapp.h stand-in.
#pragma once
#include <libPartApp/partapp.h>
