// oracle/shim/curand.h — TEST INFRASTRUCTURE: see optix.h in this directory.
#include "optix.h"
