/*
 * tests/emu/emu_solver.cpp -- TEST INFRASTRUCTURE.
 * Compiles dart_planner_b200/csrc/se3mpc_core.cuh for the HOST with a one-lane group so the
 * CPU-only test tier can check the kernel's control flow against the oracle without a GPU.
 * The product never loads this library (the product path is CUDA only).
 */
#include <stdlib.h>
#include <string.h>

#include "../../dart_planner_b200/csrc/se3mpc_core.cuh"

using namespace dartb200;

static int g_ls_shared = 0; /* 1: the register-capped build's policies (shared line-search
                             * state, status bit sets) */
extern "C" void emu_set_ls_shared(int v) { g_ls_shared = v; }

static int g_two_phase = 0; /* 1: first iteration through iterate<true>, then the context is saved,
                             * a NEW solver object restores it (fresh shared block and pair storage)
                             * and finishes -- the two-phase schedule's context switch */
extern "C" void emu_set_two_phase(int v) { g_two_phase = v; }

static int g_cold_special = 0; /* 1: cold starts run the TILT=false instantiation */
extern "C" void emu_set_cold_special(int v) { g_cold_special = v; }

template <class SV>
static void finish_outputs(SV &sv, int N, double *x_out, double *acc, double *att, double *rates, double *thrust);

template <int TPL, int GM, bool LSS, bool TILT>
static void solve_one(const dart_se3mpc_params &P, const double *p0, const double *v0,
                      const double *goal, int has_goal, const double *xw, double *x_out,
                      double *acc, double *att, double *rates, double *thrust, SolveStats &st,
                      const dart_grid *grid)
{
    double smem[SM_DOUBLES];
    for (int i = 0; i < SM_DOUBLES; ++i) smem[i] = 0.0 / 0.0; /* NaN-poison: catches stale reads */
    static double ws[MMAX][9 * TPL], wy[MMAX][9 * TPL];
    Solver<SeqGroup, TPL, GM, LSS, TILT> sv(P, smem, ws, wy);
    const int N = P.horizon;
    if (GM == 2) {
        sv.obs.g = *grid;
        sv.obs.w = P.w_obstacle;
        sv.obs.free_level = P.obstacle_free_level;
    }
    sv.has_goal = has_goal != 0;
    for (int c = 0; c < 3; ++c) sv.goal[c] = goal[c];
    if (xw)
        sv.warm_start(p0, v0, [&](int row) { return xw[row]; });
    else
        sv.cold_start(p0, v0);
    using ST = Solver<SeqGroup, TPL, GM, LSS, TILT>;
    double smem2[SM_DOUBLES];
    static double ws2[MMAX][9 * TPL], wy2[MMAX][9 * TPL];
    ST sv2(P, smem2, ws2, wy2);
    if (g_two_phase && LSS) {
        sv.begin();
        if (sv.task == 0) sv.template iterate<true>();
        if (sv.task == 0) {
            static double ctx[ST::ctx_doubles(1)];
            for (int i = 0; i < ST::ctx_doubles(1); ++i) ctx[i] = 0.0 / 0.0;
            for (int i = 0; i < SM_DOUBLES; ++i) smem2[i] = 0.0 / 0.0;
            for (int j = 0; j < MMAX; ++j)
                for (int i = 0; i < 9 * TPL; ++i) ws2[j][i] = wy2[j][i] = 0.0 / 0.0;
            if (sv.col > 1) abort();
            sv.save_context(ctx, 1, 12345);
            if (GM == 2) sv2.obs = sv.obs;
            if (sv2.restore_context(ctx, 1) != 12345) abort();
            while (sv2.task == 0) sv2.template iterate<false>();
            sv2.finish(st);
            return finish_outputs(sv2, N, x_out, acc, att, rates, thrust);
        }
        sv.finish(st);
    } else
        sv.minimize(st);
    finish_outputs(sv, N, x_out, acc, att, rates, thrust);
}

template <class SV>
static void finish_outputs(SV &sv, int N, double *x_out, double *acc, double *att, double *rates, double *thrust)
{
    for (int k = 0; k < N; ++k)
        for (int q = 0; q < 9; ++q) x_out[(q / 3) * 3 * N + 3 * k + q % 3] = sv.x[k * 9 + q];
    sv.extract([&](int k, double ax, double ay, double az, double a0, double a1, double a2,
                   double w0, double w1, double w2, double th) {
        acc[3 * k] = ax; acc[3 * k + 1] = ay; acc[3 * k + 2] = az;
        att[3 * k] = a0; att[3 * k + 1] = a1; att[3 * k + 2] = a2;
        rates[3 * k] = w0; rates[3 * k + 1] = w1; rates[3 * k + 2] = w2;
        thrust[k] = th;
    });
}

extern "C" int emu_solve_batch(const dart_se3mpc_params *P, long B, const double *p0,
                               const double *v0, const double *goal, const unsigned char *has_goal,
                               const double *x_warm, double *x, double *cost, int *nit, int *nfev,
                               int *status, int *task, double *acc, double *att, double *rates,
                               double *thrust, const dart_grid *grid)
{
    const int N = P->horizon, n = 9 * N;
    if (N < 1 || N > 32 || P->max_corrections < 1 || P->max_corrections > MMAX) return -2;
    if (P->gradient_mode == 2 && !grid) return -1;
    for (long b = 0; b < B; ++b) {
        SolveStats st;
        const double *xw = x_warm ? x_warm + (long)n * b : nullptr;
        const int hg = has_goal ? has_goal[b] : 1;
#define CALL3(T, GM, L, TI) solve_one<T, GM, L, TI>(*P, p0 + 3 * b, v0 + 3 * b, goal + 3 * b, hg, xw, x + (long)n * b, \
                             acc + 3L * N * b, att + 3L * N * b, rates + 3L * N * b, thrust + (long)N * b, st, grid)
/* the build policies (throughput build, 7-slot cold start) are emulated at the N <= 8 size only */
#define CALLP(T, GM) do { \
        if (g_cold_special && !xw) { if (g_ls_shared) CALL3(T, GM, true, false); else CALL3(T, GM, false, false); } \
        else { if (g_ls_shared) CALL3(T, GM, true, true); else CALL3(T, GM, false, true); } } while (0)
#define CALLG(T, GM) CALL3(T, GM, false, true)
#define CALL(T, F) do { if (P->gradient_mode == 1) F(T, 1); else if (P->gradient_mode == 2) F(T, 2); else F(T, 0); } while (0)
        if (N <= 8) CALL(8, CALLP);
        else if (N <= 20) CALL(20, CALLG);
        else CALL(32, CALLG);
#undef CALL
#undef CALLG
#undef CALLP
#undef CALL3
        cost[b] = st.f; nit[b] = st.nit; nfev[b] = st.nfev; status[b] = st.status; task[b] = st.task;
    }
    return 0;
}

/* solution extraction alone (the kernel core's `extract`), T: (B, N, 3) problem-major */
template <int TPL, bool TILT>
static void extract_one(const dart_se3mpc_params &P, const double *T, double *acc, double *att,
                        double *rates, double *thrust)
{
    Solver<SeqGroup, TPL, 0, false, TILT> sv(P, nullptr, nullptr, nullptr);
    const int N = P.horizon;
    for (int k = 0; k < TPL; ++k)
        for (int q = 0; q < 9; ++q) sv.x[k * 9 + q] = (k < N && q >= 6) ? T[3 * k + q - 6] : 0.0;
    sv.extract([&](int k, double ax, double ay, double az, double a0, double a1, double a2,
                   double w0, double w1, double w2, double th) {
        acc[3 * k] = ax; acc[3 * k + 1] = ay; acc[3 * k + 2] = az;
        att[3 * k] = a0; att[3 * k + 1] = a1; att[3 * k + 2] = a2;
        rates[3 * k] = w0; rates[3 * k + 1] = w1; rates[3 * k + 2] = w2;
        thrust[k] = th;
    });
}

extern "C" int emu_extract_batch(const dart_se3mpc_params *P, long B, const double *T, double *acc,
                                 double *att, double *rates, double *thrust, int untilted)
{
    const int N = P->horizon;
    if (N < 1 || N > 32) return -2;
    for (long b = 0; b < B; ++b) {
        const double *t = T + 3L * N * b;
        double *a = acc + 3L * N * b, *at = att + 3L * N * b, *r = rates + 3L * N * b, *th = thrust + (long)N * b;
        if (N <= 8) {
            if (untilted) extract_one<8, false>(*P, t, a, at, r, th); else extract_one<8, true>(*P, t, a, at, r, th);
        } else {
            if (untilted) extract_one<32, false>(*P, t, a, at, r, th); else extract_one<32, true>(*P, t, a, at, r, th);
        }
    }
    return 0;
}
