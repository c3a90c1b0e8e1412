/*
 * se3mpc_kernel.cuh -- the solve kernel template (one problem per sub-warp, see se3mpc_core.cuh)
 * and the per-configuration table its instantiation units export.  Each lane configuration is
 * compiled in its own translation unit (se3mpc_inst_*.cu) so the build runs in parallel.
 *
 * Replaces: SE3MPCPlanner._solve_se3_mpc (se3_mpc_planner.py:230-280) for B problems.
 */
#pragma once
#include <cuda_runtime.h>

#include "map_query.cuh"
#include "se3mpc_core.cuh"

namespace dartb200 {

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
    /* the caller promises that every lateral thrust entry of x_warm is exactly zero (solutions
     * this library produced from cold starts): the 7-slot instantiation then serves warm starts
     * too.  The kernel checks the promise per problem and refuses (status 3) where it is broken. */
    int no_tilt_promise;
};

template <int LANES, int TPL, int BLOCK, int MINB, int GM, bool TILT>
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
        Solver<SubWarp<LANES>, TPL, GM, (MINB >= 3), TILT> sv(P, sm, ws, wy);
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
            if (!TILT) {
                int tilted = 0;
#pragma unroll
                for (int tt = 0; tt < TPL; ++tt)
                    tilted |= (sv.x[tt * 9 + 6] != 0.0 || sv.x[tt * 9 + 7] != 0.0) ? 1 : 0;
                if (sv.grp.ori(tilted)) { /* broken promise: no solve, say so */
                    if (sv.grp.leader()) {
                        if (A.status) A.status[b] = 3;
                        if (A.nit) A.nit[b] = 0;
                        if (A.nfev) A.nfev[b] = 0;
                        if (A.cost) A.cost[b] = nan("");
                    }
                    continue;
                }
            }
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

/* the six instantiations of one lane configuration: [gradient_mode][tilt] */
struct KernelSet {
    const void *fn[3][2];
    int lanes, tpl, block, minb;
};

template <int LANES, int TPL, int MINB, int BLK>
KernelSet make_kernel_set()
{
    KernelSet k;
    k.fn[0][1] = (const void *)se3mpc_solve_kernel<LANES, TPL, BLK, MINB, 0, true>;
    k.fn[1][1] = (const void *)se3mpc_solve_kernel<LANES, TPL, BLK, MINB, 1, true>;
    k.fn[2][1] = (const void *)se3mpc_solve_kernel<LANES, TPL, BLK, MINB, 2, true>;
    k.fn[0][0] = (const void *)se3mpc_solve_kernel<LANES, TPL, BLK, MINB, 0, false>;
    k.fn[1][0] = (const void *)se3mpc_solve_kernel<LANES, TPL, BLK, MINB, 1, false>;
    k.fn[2][0] = (const void *)se3mpc_solve_kernel<LANES, TPL, BLK, MINB, 2, false>;
    k.lanes = LANES;
    k.tpl = TPL;
    k.block = BLK;
    k.minb = MINB;
    return k;
}

/* defined one per instantiation unit */
KernelSet kernel_set_l4();
KernelSet kernel_set_l8();
KernelSet kernel_set_l16();
KernelSet kernel_set_l32();
KernelSet kernel_set_l32x2();
KernelSet kernel_set_l8_occ3();
KernelSet kernel_set_l8_b64();
KernelSet kernel_set_l16_occ3();
KernelSet kernel_set_l32_occ3();

} /* namespace dartb200 */
