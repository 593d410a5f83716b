// oracle/shim/curand_kernel.h — TEST INFRASTRUCTURE: see optix.h in this directory.
#include "optix.h"
