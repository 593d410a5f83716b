// oracle/shim/math_functions.h — TEST INFRASTRUCTURE: see optix.h in this directory.
#include "optix.h"
