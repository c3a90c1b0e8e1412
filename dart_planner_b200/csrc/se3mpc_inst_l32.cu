/* instantiation unit: solve kernels <LANES, TPL, MINB, BLOCK> = <32, 1, 2, 128> (see se3mpc_kernel.cuh) */
#include "se3mpc_kernel.cuh"

namespace dartb200 {
KernelSet kernel_set_l32() { return make_kernel_set<32, 1, 2, 128>(); }
}
