/* instantiation unit: solve kernels <LANES, TPL, MINB, BLOCK> = <4, 1, 2, 64> (see se3mpc_kernel.cuh).
 * 64-thread blocks: a block holds 16 problems (63 KB of shared memory), so three blocks stay
 * resident per SM; with 128 threads (32 problems, 127 KB) only one would. */
#include "se3mpc_kernel.cuh"

namespace dartb200 {
KernelSet kernel_set_l4()
{
    KernelSet k = make_kernel_set<4, 1, 2, 64>();
    k.resident = 3;
    return k;
}
}
