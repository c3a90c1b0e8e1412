/* instantiation unit: solve kernels <LANES, TPL, MINB, BLOCK> = <16, 1, 3, 128> (see se3mpc_kernel.cuh) */
#include "se3mpc_kernel.cuh"

namespace dartb200 {
KernelSet kernel_set_l16_occ3() { return make_kernel_set<16, 1, 3, 128>(); }
}
