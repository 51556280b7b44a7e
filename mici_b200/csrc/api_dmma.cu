// libmici_b200.so -- C-ABI entry points (include/mici_b200.h): dispatch of the FP64 tensor-core leapfrog kernel K1.
// Host-side argument checking and kernel dispatch only; all arithmetic is in the .cuh kernels.
#include "api_common.cuh"
#include "leapfrog_dmma.cuh"

// leapfrog_dmma.cuh defines mb200::leapfrog_dmma_dispatch (declared in api_common.cuh)
