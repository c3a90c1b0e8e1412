/* instantiation unit: solve kernels <LANES, TPL, MINB, BLOCK> = <4, 1, 2, 128> (see se3mpc_kernel.cuh) */
#include "se3mpc_kernel.cuh"

namespace dartb200 {
KernelSet kernel_set_l4() { return make_kernel_set<4, 1, 2, 128>(); }
}
