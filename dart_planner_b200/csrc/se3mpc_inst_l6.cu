/* instantiation unit: solve kernels <LANES, TPL, MINB, BLOCK> = <6, 1, 2, 128> (see se3mpc_kernel.cuh) */
#include "se3mpc_kernel.cuh"

namespace dartb200 {
KernelSet kernel_set_l6() { return make_kernel_set<6, 1, 2, 128>(); }
}
