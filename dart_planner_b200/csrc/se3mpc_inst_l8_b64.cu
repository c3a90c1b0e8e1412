/* instantiation unit: solve kernels <LANES, TPL, MINB, BLOCK> = <8, 1, 6, 64> (see se3mpc_kernel.cuh) */
#include "se3mpc_kernel.cuh"

namespace dartb200 {
KernelSet kernel_set_l8_b64() { return make_kernel_set<8, 1, 6, 64>(); }
}
