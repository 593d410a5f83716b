// oracle/shim/cuda_runtime.h — TEST INFRASTRUCTURE: see optix.h in this directory.
#include "optix.h"
