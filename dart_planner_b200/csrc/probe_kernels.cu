/*
 * probe_kernels.cu -- FP64 FMA throughput probe used by bench.py for the compute roofline
 * denominator (MEASURED_PEAKS.json has no FP64 vector peak).  8 independent DFMA chains per
 * thread, 148*8 blocks of 256 threads.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dart_se3mpc.h"

extern "C" void dart_count_launch_(void);

namespace {
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double seed, double *out)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.9999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s; /* never true: keeps the chains alive */
}
}

extern "C" int dart_fp64_probe(int32_t iters, int32_t *threads_out, double *scratch, void *stream)
{
    if (iters <= 0 || !scratch) return DART_E_BADARG;
    const int blocks = 148 * 8;
    dfma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, scratch);
    if (cudaGetLastError() != cudaSuccess) return DART_E_CUDA;
    if (threads_out) *threads_out = blocks * 256;
    dart_count_launch_();
    return DART_OK;
}

/* ---- self-test of the solver's division sequence (se3mpc_core.cuh ddiv) against the
 * compiler's IEEE division: n random operand pairs with magnitudes 2^[-emax, emax] and random
 * signs, plus zero dividends; counts the pairs whose bits differ. */
#include "se3mpc_core.cuh"

namespace {
__device__ __forceinline__ unsigned long long splitmix(unsigned long long &x)
{
    unsigned long long z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double rnd_operand(unsigned long long &st, int emax)
{
    const unsigned long long m = splitmix(st);
    const double frac = 1.0 + (double)(m >> 12) * (1.0 / 4503599627370496.0); /* [1, 2) */
    const int e = (int)(splitmix(st) % (unsigned)(2 * emax + 1)) - emax;
    const double v = ldexp(frac, e);
    return (splitmix(st) & 1ull) ? -v : v;
}
__global__ void __launch_bounds__(256)
ddiv_selftest_kernel(long long n, unsigned long long seed, int emax, unsigned long long *mismatch)
{
    unsigned long long bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        unsigned long long st = seed + 0x632be59bd9b4e019ull * (unsigned long long)(i + 1);
        double a = rnd_operand(st, emax);
        const double b = rnd_operand(st, emax);
        if ((i & 15) == 0) a = (i & 16) ? 0.0 : -0.0;
        const double ref = a / b;
        const double got = dartb200::ddiv(a, b);
        const dartb200::Recip R = dartb200::make_recip(b);
        const double got2 = dartb200::ddiv(a, R);
        bool same = __double_as_longlong(ref) == __double_as_longlong(got) &&
                    __double_as_longlong(ref) == __double_as_longlong(got2);
        if (a == 0.0) same = (got == 0.0 && got2 == 0.0); /* the sign of a zero quotient is not kept */
        if (!same) ++bad;
    }
    if (bad) atomicAdd(mismatch, bad);
}
}

extern "C" int dart_ddiv_selftest(int64_t n, uint64_t seed, int32_t emax, uint64_t *mismatch_dev,
                                  void *stream)
{
    if (n <= 0 || emax < 0 || emax > 300 || !mismatch_dev) return DART_E_BADARG;
    ddiv_selftest_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(n, seed, emax,
                                                                    (unsigned long long *)mismatch_dev);
    if (cudaGetLastError() != cudaSuccess) return DART_E_CUDA;
    dart_count_launch_();
    return DART_OK;
}
