/* instantiation unit: solve kernels <LANES, TPL, MINB, BLOCK> = <8, 1, 2, 128> (see se3mpc_kernel.cuh) */
#include "se3mpc_kernel.cuh"

namespace dartb200 {
KernelSet kernel_set_l8() { return make_kernel_set<8, 1, 2, 128>(); }
}

#if defined(DART_PHASE_TIMING)
/* diagnostic builds only: copy out and reset the phase log (tools/phase_timing.py) */
extern "C" int dart_phase_log_read(long long *out, int cap)
{
    int n = 0;
    cudaMemcpyFromSymbol(&n, dartb200::g_phase_n, sizeof(int));
    if (n > cap) n = cap;
    cudaMemcpyFromSymbol(out, dartb200::g_phase_log, sizeof(long long) * 2 * n);
    int zero = 0;
    cudaMemcpyToSymbol(dartb200::g_phase_n, &zero, sizeof(int));
    return n;
}
#endif
