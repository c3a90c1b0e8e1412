/*
 * se3mpc_kernels.cu -- sm_100a kernels + C ABI (include/dart_se3mpc.h) of the batched
 * SE(3)-MPC solve.  One problem per sub-warp (see se3mpc_core.cuh); a persistent grid walks
 * the batch with a grid-stride loop.  I/O is batch-major SoA so the 32/LANES problems of a
 * warp touch consecutive addresses of every row.
 *
 * Replaces: SE3MPCPlanner._solve_se3_mpc (se3_mpc_planner.py:230-280) for B problems.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "map_query.cuh"
#include "se3mpc_core.cuh"

using namespace dartb200;

namespace {

std::atomic<long long> g_launches{0};
thread_local char g_err[256] = "";

struct SolveArgs {
    long long B, ld;
    const double *p0, *v0, *goal;
    const unsigned char *has_goal;
    const double *x_warm;
    const unsigned char *warm_mask;
    double *x_out, *cost;
    int *nit, *nfev, *status, *task;
    double *acc, *att, *rates, *thrust;
    /* fused post-solve safety check (is_trajectory_safe on the solved positions); off when
     * first_hit == nullptr */
    dart_grid grid;
    double margin, threshold;
    int *first_hit;
    /* fused plant step of the closed-loop simulation (off when p_next == nullptr): the state is
     * advanced with the first control of the new solution; may alias p0 / v0 */
    double *p_next, *v_next;
    double plant_dt;
};

template <int LANES, int TPL, int BLOCK, int MINB, int GM>
__global__ void __launch_bounds__(BLOCK, MINB)
se3mpc_solve_kernel(const __grid_constant__ dart_se3mpc_params P, const __grid_constant__ SolveArgs A)
{
    extern __shared__ double smem_all[];
    constexpr int GPB = BLOCK / LANES; /* problems (groups) per block */
    const int gib = threadIdx.x / LANES;
    double *sm = smem_all + gib * SM_DOUBLES;
    const int N = P.horizon;
    const long long stride = (long long)gridDim.x * GPB;
    double ws[MMAX][9 * TPL], wy[MMAX][9 * TPL]; /* per-lane S / Y pairs (local memory, L1) */
    /* block-uniform trip count + a warp barrier per round: the sub-warps of a warp start every
     * problem together (a sub-warp that converged early waits instead of running ahead into
     * different code) */
    (void)stride;
    const long long rounds = (A.B + GPB - 1) / GPB;
    for (long long blk = blockIdx.x; blk < rounds; blk += gridDim.x) {
        __syncwarp();
        const long long b = blk * GPB + gib;
        if (b >= A.B) continue;
        Solver<SubWarp<LANES>, TPL, GM, (MINB >= 3)> sv(P, sm, ws, wy);
        if (GM == 2) {
            sv.obs.g = A.grid;
            sv.obs.w = P.w_obstacle;
            sv.obs.free_level = P.obstacle_free_level;
        }
        double p0[3], v0[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p0[c] = __ldg(A.p0 + c * A.ld + b);
            v0[c] = __ldg(A.v0 + c * A.ld + b);
            sv.goal[c] = __ldg(A.goal + c * A.ld + b);
        }
        sv.has_goal = A.has_goal ? (A.has_goal[b] != 0) : true;
        const bool warm = A.x_warm != nullptr && (A.warm_mask == nullptr || A.warm_mask[b] != 0);
        if (warm) {
            /* plain loads: in the closed loop x_out aliases x_warm */
            const double *xw = A.x_warm + b;
            const long long ld = A.ld;
            sv.warm_start(p0, v0, [xw, ld](int row) { return xw[(long long)row * ld]; });
        } else
            sv.cold_start(p0, v0);
        SolveStats st;
        sv.minimize(st);
        if (A.x_out) {
#pragma unroll
            for (int tt = 0; tt < TPL; ++tt)
                if (sv.act[tt]) {
#pragma unroll
                    for (int q = 0; q < 9; ++q)
                        A.x_out[(long long)sv.row_of(tt, q) * A.ld + b] = sv.x[tt * 9 + q];
                }
        }
        if (sv.grp.leader()) {
            if (A.cost) A.cost[b] = st.f;
            if (A.nit) A.nit[b] = st.nit;
            if (A.nfev) A.nfev[b] = st.nfev;
            if (A.status) A.status[b] = st.status;
            if (A.task) A.task[b] = st.task;
        }
        if (A.acc || A.att || A.rates || A.thrust) {
            const SolveArgs &a = A;
            sv.extract([&a, b](int k, double ax, double ay, double az, double r0, double r1,
                               double r2, double w0, double w1, double w2, double th) {
                const long long ld = a.ld;
                if (a.acc) {
                    a.acc[(long long)(3 * k) * ld + b] = ax;
                    a.acc[(long long)(3 * k + 1) * ld + b] = ay;
                    a.acc[(long long)(3 * k + 2) * ld + b] = az;
                }
                if (a.att) {
                    a.att[(long long)(3 * k) * ld + b] = r0;
                    a.att[(long long)(3 * k + 1) * ld + b] = r1;
                    a.att[(long long)(3 * k + 2) * ld + b] = r2;
                }
                if (a.rates) {
                    a.rates[(long long)(3 * k) * ld + b] = w0;
                    a.rates[(long long)(3 * k + 1) * ld + b] = w1;
                    a.rates[(long long)(3 * k + 2) * ld + b] = w2;
                }
                if (a.thrust) a.thrust[(long long)k * ld + b] = th;
            });
        }
        if (A.first_hit) {
            /* each lane tests its own timesteps against the map; the first colliding index is
             * the minimum over the group (explicit_geometric_mapper.py:195-219) */
            int hit = 0x7fffffff;
#pragma unroll
            for (int tt = TPL - 1; tt >= 0; --tt)
                if (sv.act[tt] && position_collides(A.grid, sv.x[tt * 9], sv.x[tt * 9 + 1], sv.x[tt * 9 + 2],
                                                    A.margin, A.threshold))
                    hit = sv.grp.lane() * TPL + tt;
            hit = sv.grp.mini(hit);
            if (sv.grp.leader()) A.first_hit[b] = (hit == 0x7fffffff) ? -1 : hit;
        }
        if (A.p_next && sv.grp.leader()) {
            /* reference planner model (se3_mpc_planner.py:430-431, :445-459) driven by T_0:
             * a = T_0/m - g e3;  p <- p + v dt + (0.5 a) dt^2;  v <- v + a dt.  Every operation
             * individually rounded (NumPy's order). */
            const double dt = A.plant_dt, dt2 = DP_MUL(dt, dt);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double a = DP_ADD(ddiv(sv.x[6 + c], P.mass), c == 2 ? -P.gravity : -0.0);
                A.p_next[c * A.ld + b] = DP_ADD(DP_ADD(p0[c], DP_MUL(v0[c], dt)), DP_MUL(DP_MUL(0.5, a), dt2));
                A.v_next[c * A.ld + b] = DP_ADD(v0[c], DP_MUL(a, dt));
            }
        }
        (void)N;
    }
}

struct KernelChoice {
    const void *fn;       /* gradient_mode 0: the reference gradient (:552-580) */
    const void *fn_exact; /* gradient_mode 1: exact gradient of :516-550       */
    const void *fn_grid;  /* gradient_mode 2: mode 0 + occupancy-grid penalty  */
    int lanes, tpl, block, minb;
    int occ_blocks; /* resident blocks per SM (queried once) */
    int regs;
    bool ready;
};

constexpr int BLOCK = 128;

template <int LANES, int TPL, int MINB, int BLK = BLOCK>
KernelChoice make_choice()
{
    KernelChoice k;
    k.fn = (const void *)se3mpc_solve_kernel<LANES, TPL, BLK, MINB, 0>;
    k.fn_exact = (const void *)se3mpc_solve_kernel<LANES, TPL, BLK, MINB, 1>;
    k.fn_grid = (const void *)se3mpc_solve_kernel<LANES, TPL, BLK, MINB, 2>;
    k.lanes = LANES;
    k.tpl = TPL;
    k.block = BLK;
    k.minb = MINB;
    k.occ_blocks = 0;
    k.regs = 0;
    k.ready = false;
    return k;
}

KernelChoice g_kernels[] = {
    make_choice<4, 1, 2>(), make_choice<8, 1, 2>(), make_choice<16, 1, 2>(),
    make_choice<32, 1, 2>(), make_choice<32, 2, 2>(),
    /* tuning variants, selected with DART_SE3MPC_VARIANT=<index> (tools/kbench.py) */
    make_choice<8, 1, 3>(), make_choice<8, 1, 4, 64>(),
};
std::mutex g_mu;
int g_sms = 0;

int smem_bytes(const KernelChoice &k) { return (k.block / k.lanes) * SM_DOUBLES * (int)sizeof(double); }

int set_err(cudaError_t e, const char *what)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return DART_E_CUDA;
}

KernelChoice *pick_kernel(int N, long long B)
{
    int idx = N <= 4 ? 0 : N <= 8 ? 1 : N <= 16 ? 2 : N <= 32 ? 3 : 4;
    /* N <= 8, many rounds of work: the 168-register build (3 resident blocks per SM) trades a
     * few spills for 50 % more warps in flight: +15 % at 64 Ki and 1 Mi problems, but -12 % on a
     * single round, where nothing waits for a free slot (profiles/README.md) */
    if (idx == 1 && N > 4 && B >= 16384) idx = 5;
    if (const char *v = getenv("DART_SE3MPC_VARIANT")) {
        const int want = atoi(v);
        const int nk = (int)(sizeof(g_kernels) / sizeof(g_kernels[0]));
        if (want >= 0 && want < nk && g_kernels[want].lanes * g_kernels[want].tpl >= N) idx = want;
    }
    return &g_kernels[idx];
}

int prepare(KernelChoice *k)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (k->ready) return DART_OK;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_err(e, "cudaGetDevice");
        return DART_E_NODEVICE;
    }
    if (g_sms == 0) {
        e = cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return set_err(e, "cudaDeviceGetAttribute");
    }
    const int smem = smem_bytes(*k);
    for (const void *fn : {k->fn, k->fn_exact, k->fn_grid}) {
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return set_err(e, "cudaFuncSetAttribute(smem)");
    }
    /* shared-memory share of the 256 KB L1: what `minb` resident blocks need (+1 KB each that
     * the driver reserves); the rest stays L1 for the per-lane S/Y pairs in local memory */
    int carve = (int)((long long)k->minb * (smem + 1024) * 100 / (228 * 1024)) + 1;
    if (carve > 100) carve = 100;
    for (const void *fn : {k->fn, k->fn_exact, k->fn_grid}) {
        e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e != cudaSuccess) return set_err(e, "cudaFuncSetAttribute(carveout)");
    }
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, k->fn);
    if (e != cudaSuccess) return set_err(e, "cudaFuncGetAttributes");
    k->regs = fa.numRegs;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k->occ_blocks, k->fn, k->block, smem);
    if (e != cudaSuccess) return set_err(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (k->occ_blocks < 1) k->occ_blocks = 1;
    k->ready = true;
    return DART_OK;
}

int check_params(const dart_se3mpc_params *p)
{
    if (!p || p->struct_size != (int)sizeof(dart_se3mpc_params)) return DART_E_BADARG;
    if (p->horizon < 1 || p->horizon > 64) return DART_E_UNSUPPORTED;
    if (p->max_corrections < 1 || p->max_corrections > MMAX) return DART_E_UNSUPPORTED;
    if (p->max_iterations < 0 || p->max_linesearch < 0) return DART_E_BADARG;
    if (p->gradient_mode < 0 || p->gradient_mode > 2) return DART_E_UNSUPPORTED;
    if (!(p->dt > 0.0) || !(p->mass > 0.0)) return DART_E_BADARG;
    return DART_OK;
}

long long grid_for(const KernelChoice &k, long long B)
{
    const long long gpb = k.block / k.lanes;
    long long need = (B + gpb - 1) / gpb;
    long long cap = (long long)g_sms * k.occ_blocks;
    return need < cap ? need : cap;
}

} /* namespace */

extern "C" {

int dart_abi_version(void) { return DART_SE3MPC_ABI_VERSION; }
const char *dart_last_cuda_error(void) { return g_err; }
int64_t dart_launch_count(void) { return (int64_t)g_launches.load(); }
void dart_count_launch_(void) { g_launches.fetch_add(1); }

void dart_se3mpc_default_params(dart_se3mpc_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->struct_size = (int)sizeof(*p);
    p->horizon = 6;             /* se3_mpc_planner.py:41 */
    p->max_iterations = 15;     /* :66 */
    p->max_corrections = 10;    /* SciPy maxcor */
    p->max_linesearch = 20;     /* SciPy maxls */
    p->max_fun = 15000;         /* SciPy maxfun */
    p->gradient_mode = 0;
    p->dt = 1.0 / 400.0;        /* timing_alignment.py:76-78 with frozen_config.py:86 */
    p->mass = 1.5;              /* :149 */
    p->gravity = 9.81;          /* :150 */
    p->pos_bound = 100.0;       /* :384 */
    p->max_velocity = 10.0;     /* :45 */
    p->max_thrust = 25.0;       /* :48 */
    p->min_thrust = 2.0;        /* :49 */
    p->tilt_thrust = 25.0 * 0.70710678118654746; /* max_thrust*sin(pi/4) (:393-397) */
    p->w_pos = 100.0;           /* :56-59 */
    p->w_vel = 10.0;
    p->w_acc = 1.0;
    p->w_thrust = 0.1;
    p->gtol = 5e-2;             /* :68, :264 */
    p->ftol = 5e-1;             /* :265 */
    p->w_obstacle = 1000.0;     /* obstacle_weight :60 (unused by the reference solve) */
    p->obstacle_free_level = 0.5; /* the mapper's prior, explicit_geometric_mapper.py:71 */
}

int dart_se3mpc_kernel_info(const dart_se3mpc_params *params, int64_t B, int32_t *lanes,
                            int32_t *block_threads, int32_t *grid_blocks, int32_t *smem,
                            int32_t *regs_per_thread)
{
    int rc = check_params(params);
    if (rc) return rc;
    KernelChoice *k = pick_kernel(params->horizon, B);
    rc = prepare(k);
    if (rc) return rc;
    if (lanes) *lanes = k->lanes;
    if (block_threads) *block_threads = k->block;
    if (grid_blocks) *grid_blocks = (int32_t)grid_for(*k, B > 0 ? B : 1);
    if (smem) *smem = smem_bytes(*k);
    if (regs_per_thread) *regs_per_thread = k->regs;
    return DART_OK;
}

static int launch_solve(const dart_se3mpc_params *params, const SolveArgs &a, void *cuda_stream)
{
    KernelChoice *k = pick_kernel(params->horizon, a.B);
    int rc = prepare(k);
    if (rc) return rc;
    dart_se3mpc_params P = *params;
    SolveArgs args_copy = a;
    void *args[] = {(void *)&P, (void *)&args_copy};
    const long long grid_blocks = grid_for(*k, a.B);
    const void *fn = params->gradient_mode == 1 ? k->fn_exact : (params->gradient_mode == 2 ? k->fn_grid : k->fn);
    cudaError_t e = cudaLaunchKernel(fn, dim3((unsigned)grid_blocks), dim3(k->block), args,
                                     smem_bytes(*k), (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) return set_err(e, "cudaLaunchKernel(se3mpc_solve)");
    g_launches.fetch_add(1);
    return DART_OK;
}

int dart_se3mpc_solve_batch_map(const dart_se3mpc_params *params, int64_t B, int64_t ld,
                                const double *p0, const double *v0, const double *goal,
                                const uint8_t *has_goal, const double *x_warm,
                                const uint8_t *warm_mask, double *x_out, double *cost, int32_t *nit,
                                int32_t *nfev, int32_t *status, int32_t *task, double *acc,
                                double *att, double *rates, double *thrust, const dart_grid *grid,
                                double margin, double threshold, int32_t *first_hit,
                                void *cuda_stream)
{
    int rc = check_params(params);
    if (rc) return rc;
    if (B < 0 || ld < B || !p0 || !v0 || !goal) return DART_E_BADARG;
    const bool need_grid = first_hit != nullptr || params->gradient_mode == 2;
    if (need_grid && (!grid || !grid->occ || grid->nx <= 0 || grid->ny <= 0 || grid->nz <= 0 ||
                      !(grid->resolution > 0.0)))
        return DART_E_BADARG;
    if (B == 0) return DART_OK;
    SolveArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.ld = ld;
    a.p0 = p0; a.v0 = v0; a.goal = goal; a.has_goal = has_goal;
    a.x_warm = x_warm; a.warm_mask = warm_mask;
    a.x_out = x_out; a.cost = cost; a.nit = nit; a.nfev = nfev; a.status = status; a.task = task;
    a.acc = acc; a.att = att; a.rates = rates; a.thrust = thrust;
    if (need_grid) a.grid = *grid;
    if (first_hit) {
        a.margin = margin;
        a.threshold = threshold;
        a.first_hit = first_hit;
    }
    return launch_solve(params, a, cuda_stream);
}

int dart_se3mpc_closed_loop_step(const dart_se3mpc_params *params, int64_t B, int64_t ld, double *p,
                                 double *v, const double *goal, const uint8_t *has_goal, double *x,
                                 int32_t warm, double *cost, int32_t *nit, int32_t *nfev,
                                 int32_t *status, double plant_dt, void *cuda_stream)
{
    int rc = check_params(params);
    if (rc) return rc;
    if (B < 0 || ld < B || !p || !v || !goal || !x || !(plant_dt > 0.0)) return DART_E_BADARG;
    if (params->gradient_mode == 2) return DART_E_UNSUPPORTED; /* no map argument here */
    if (B == 0) return DART_OK;
    SolveArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.ld = ld;
    a.p0 = p; a.v0 = v; a.goal = goal; a.has_goal = has_goal;
    a.x_warm = warm ? x : nullptr;
    a.x_out = x; a.cost = cost; a.nit = nit; a.nfev = nfev; a.status = status;
    a.p_next = p; a.v_next = v; a.plant_dt = plant_dt;
    return launch_solve(params, a, cuda_stream);
}

int dart_se3mpc_solve_batch(const dart_se3mpc_params *params, int64_t B, int64_t ld,
                            const double *p0, const double *v0, const double *goal,
                            const uint8_t *has_goal, const double *x_warm,
                            const uint8_t *warm_mask, double *x_out, double *cost, int32_t *nit,
                            int32_t *nfev, int32_t *status, int32_t *task, double *acc,
                            double *att, double *rates, double *thrust, void *cuda_stream)
{
    return dart_se3mpc_solve_batch_map(params, B, ld, p0, v0, goal, has_goal, x_warm, warm_mask, x_out,
                                       cost, nit, nfev, status, task, acc, att, rates, thrust, nullptr,
                                       0.0, 0.0, nullptr, cuda_stream);
}

/* ---- host-buffer entry: staged through a cached per-thread device workspace --------------- */
namespace {
struct HostWs {
    void *dev = nullptr;
    size_t cap = 0;
    void *pin = nullptr; /* pinned host mirror of the workspace (small batches) */
    size_t pin_cap = 0;
    cudaStream_t stream = nullptr;
};
thread_local HostWs g_ws;

/* workspace layout for row pitch Bp (elements): fp64 rows [p0 3 | v0 3 | goal 3 | x_warm 9N |
 * x 9N | cost 1 | acc 3N | att 3N | rates 3N | thrust N], int32 rows [nit nfev status task],
 * u8 row [has_goal].  Inputs are one contiguous range, outputs another. */
struct WsLayout {
    size_t Bp, in_rows, out_off_rows, out_rows, int_off, hg_off, bytes;
    WsLayout(int N, size_t Bp_) : Bp(Bp_)
    {
        in_rows = 9 + 9 * (size_t)N;
        out_off_rows = in_rows;
        out_rows = 9 * (size_t)N + 1 + 9 * (size_t)N + (size_t)N;
        int_off = (in_rows + out_rows) * Bp * 8;
        hg_off = int_off + 4 * Bp * 4;
        bytes = hg_off + ((Bp + 7) / 8) * 8;
    }
};

void rows_in(void *dst, size_t pitch, const void *src, size_t B, size_t rows, size_t esz)
{
    for (size_t r = 0; r < rows; ++r)
        memcpy((char *)dst + r * pitch * esz, (const char *)src + r * B * esz, B * esz);
}
void rows_out(void *dst, const void *src, size_t pitch, size_t B, size_t rows, size_t esz)
{
    if (!dst) return;
    for (size_t r = 0; r < rows; ++r)
        memcpy((char *)dst + r * B * esz, (const char *)src + r * pitch * esz, B * esz);
}
}

int dart_se3mpc_solve_batch_host(const dart_se3mpc_params *params, int64_t B,
                                 const double *p0, const double *v0, const double *goal,
                                 const uint8_t *has_goal, const double *x_warm, double *x_out,
                                 double *cost, int32_t *nit, int32_t *nfev, int32_t *status,
                                 double *acc, double *att, double *rates, double *thrust)
{
    int rc = check_params(params);
    if (rc) return rc;
    if (B < 0 || !p0 || !v0 || !goal) return DART_E_BADARG;
    if (B == 0) return DART_OK;
    const size_t N = (size_t)params->horizon;
    /* small batches (the drop-in planner's single solve): ONE pinned H2D copy, one launch, ONE
     * D2H copy; large batches copy rows straight from / to the caller's buffers */
    const bool staged = B <= 512;
    const size_t Bp = staged ? (size_t)((B + 3) / 4 * 4) : (size_t)((B + 31) / 32 * 32);
    const WsLayout L((int)N, Bp);
    cudaError_t e;
    if (!g_ws.stream) {
        e = cudaStreamCreateWithFlags(&g_ws.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) return set_err(e, "cudaStreamCreate");
    }
    if (g_ws.cap < L.bytes) {
        if (g_ws.dev) cudaFree(g_ws.dev);
        g_ws.dev = nullptr;
        g_ws.cap = 0;
        e = cudaMalloc(&g_ws.dev, L.bytes);
        if (e != cudaSuccess) return set_err(e, "cudaMalloc(workspace)");
        g_ws.cap = L.bytes;
    }
    if (staged && g_ws.pin_cap < L.bytes) {
        if (g_ws.pin) cudaFreeHost(g_ws.pin);
        g_ws.pin = nullptr;
        g_ws.pin_cap = 0;
        e = cudaHostAlloc(&g_ws.pin, L.bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) return set_err(e, "cudaHostAlloc(staging)");
        g_ws.pin_cap = L.bytes;
    }
    cudaStream_t s = g_ws.stream;
    char *base = (char *)g_ws.dev;
    double *d = (double *)base;
    double *d_p0 = d, *d_v0 = d_p0 + 3 * Bp, *d_goal = d_v0 + 3 * Bp, *d_xw = d_goal + 3 * Bp;
    double *d_x = d_xw + 9 * N * Bp, *d_cost = d_x + 9 * N * Bp;
    double *d_acc = d_cost + Bp, *d_att = d_acc + 3 * N * Bp, *d_rates = d_att + 3 * N * Bp;
    double *d_thr = d_rates + 3 * N * Bp;
    int32_t *d_nit = (int32_t *)(base + L.int_off), *d_nfev = d_nit + Bp, *d_status = d_nfev + Bp,
            *d_task = d_status + Bp;
    uint8_t *d_hg = (uint8_t *)(base + L.hg_off);
    const size_t in_bytes = (x_warm ? L.in_rows : 9) * Bp * 8;
    if (staged) {
        char *h = (char *)g_ws.pin;
        rows_in(h, Bp, p0, (size_t)B, 3, 8);
        rows_in(h + 3 * Bp * 8, Bp, v0, (size_t)B, 3, 8);
        rows_in(h + 6 * Bp * 8, Bp, goal, (size_t)B, 3, 8);
        if (x_warm) rows_in(h + 9 * Bp * 8, Bp, x_warm, (size_t)B, 9 * N, 8);
        e = cudaMemcpyAsync(base, h, in_bytes, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return set_err(e, "cudaMemcpyAsync H2D");
        if (has_goal) {
            memcpy(h + L.hg_off, has_goal, (size_t)B);
            e = cudaMemcpyAsync(d_hg, h + L.hg_off, (size_t)B, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return set_err(e, "cudaMemcpyAsync H2D");
        }
    } else {
#define H2D(dst, src, rows, esz)                                                              \
    do {                                                                                      \
        e = cudaMemcpy2DAsync(dst, Bp * (esz), src, (size_t)B * (esz), (size_t)B * (esz), rows, \
                              cudaMemcpyHostToDevice, s);                                     \
        if (e != cudaSuccess) return set_err(e, "cudaMemcpy2DAsync H2D");                     \
    } while (0)
        H2D(d_p0, p0, 3, 8);
        H2D(d_v0, v0, 3, 8);
        H2D(d_goal, goal, 3, 8);
        if (has_goal) H2D(d_hg, has_goal, 1, 1);
        if (x_warm) H2D(d_xw, x_warm, 9 * N, 8);
#undef H2D
    }
    /* the staged path always produces every output row (one copy back); the direct path only
     * the rows the caller asked for */
    const bool all = staged;
    rc = dart_se3mpc_solve_batch(params, B, (int64_t)Bp, d_p0, d_v0, d_goal, has_goal ? d_hg : nullptr,
                                 x_warm ? d_xw : nullptr, nullptr, (all || x_out) ? d_x : nullptr,
                                 (all || cost) ? d_cost : nullptr, (all || nit) ? d_nit : nullptr,
                                 (all || nfev) ? d_nfev : nullptr, (all || status) ? d_status : nullptr,
                                 nullptr, (all || acc) ? d_acc : nullptr, (all || att) ? d_att : nullptr,
                                 (all || rates) ? d_rates : nullptr, (all || thrust) ? d_thr : nullptr,
                                 (void *)s);
    if (rc) return rc;
    if (staged) {
        char *h = (char *)g_ws.pin;
        const size_t out_off = L.out_off_rows * Bp * 8;
        e = cudaMemcpyAsync(h + out_off, base + out_off, L.hg_off - out_off, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return set_err(e, "cudaMemcpyAsync D2H");
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return set_err(e, "cudaStreamSynchronize");
        const double *hd = (const double *)h;
        const double *h_x = hd + L.in_rows * Bp, *h_cost = h_x + 9 * N * Bp, *h_acc = h_cost + Bp,
                     *h_att = h_acc + 3 * N * Bp, *h_rates = h_att + 3 * N * Bp, *h_thr = h_rates + 3 * N * Bp;
        const int32_t *h_nit = (const int32_t *)(h + L.int_off);
        rows_out(x_out, h_x, Bp, (size_t)B, 9 * N, 8);
        rows_out(cost, h_cost, Bp, (size_t)B, 1, 8);
        rows_out(acc, h_acc, Bp, (size_t)B, 3 * N, 8);
        rows_out(att, h_att, Bp, (size_t)B, 3 * N, 8);
        rows_out(rates, h_rates, Bp, (size_t)B, 3 * N, 8);
        rows_out(thrust, h_thr, Bp, (size_t)B, N, 8);
        rows_out(nit, h_nit, Bp, (size_t)B, 1, 4);
        rows_out(nfev, h_nit + Bp, Bp, (size_t)B, 1, 4);
        rows_out(status, h_nit + 2 * Bp, Bp, (size_t)B, 1, 4);
        return DART_OK;
    }
#define D2H(dst, src, rows, esz)                                                              \
    do {                                                                                      \
        if (dst) {                                                                            \
            e = cudaMemcpy2DAsync(dst, (size_t)B * (esz), src, Bp * (esz), (size_t)B * (esz), rows, \
                                  cudaMemcpyDeviceToHost, s);                                 \
            if (e != cudaSuccess) return set_err(e, "cudaMemcpy2DAsync D2H");                 \
        }                                                                                     \
    } while (0)
    D2H(x_out, d_x, 9 * N, 8);
    D2H(cost, d_cost, 1, 8);
    D2H(nit, d_nit, 1, 4);
    D2H(nfev, d_nfev, 1, 4);
    D2H(status, d_status, 1, 4);
    D2H(acc, d_acc, 3 * N, 8);
    D2H(att, d_att, 3 * N, 8);
    D2H(rates, d_rates, 3 * N, 8);
    D2H(thrust, d_thr, N, 8);
#undef D2H
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return set_err(e, "cudaStreamSynchronize");
    return DART_OK;
}

} /* extern "C" */
