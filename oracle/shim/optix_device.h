// oracle/shim/optix_device.h — TEST INFRASTRUCTURE: see optix.h in this directory.
#include "optix.h"
