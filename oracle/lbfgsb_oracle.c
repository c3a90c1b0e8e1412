/*
 * oracle/lbfgsb_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar fp64 restatement of L-BFGS-B 3.0 (Byrd, Lu, Nocedal, Zhu 1995; Zhu, Byrd, Lu,
 * Nocedal TOMS 778; Morales & Nocedal 2011) as it is driven by SciPy's
 * `_minimize_lbfgsb` (scipy 1.18.1, `optimize/_lbfgsb_py.py`), which is the third-party
 * routine the reference calls at `planning/se3_mpc_planner.py:256-268`.  SciPy ships
 * only the compiled routine; this file restates the published algorithm routine by
 * routine (names follow the published code: active, projgr, cauchy, hpsolb, bmv,
 * formk, cmprlb, subsm, lnsrlb, dcsrch, dcstep, matupd, formt) and is pinned against
 * the installed SciPy by tests/test_oracle_vs_scipy.py and the golden fixtures.
 *
 * Deliberate simplification (mathematically identical): the 2col x
 * 2col matrix of `formk` is rebuilt from W = [Y, theta*S] each time instead of being
 * updated incrementally.
 */
#include "se3mpc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static __thread double g_flops;
#define FL(k) (g_flops += (double)(k))

double orc_flops_get(void) { return g_flops; }
void orc_flops_reset(void) { g_flops = 0.0; }

typedef struct {
    int n, m;
    /* limited-memory matrices; column j (physical) of ws/wy is ws + j*n */
    double *ws, *wy;
    double *sy, *ss, *wt; /* m x m, element (i,j) at [i*m+j]                       */
    double *wn;           /* 2m x 2m upper-triangular factor from formk            */
    double *z, *r, *d, *t, *xp;
    double *brk;          /* breakpoints                                           */
    int *iorder, *iwhere, *isfree;
    double *p, *c, *wbp, *v; /* 2m each */
    double *wv;              /* 2m */
    int col, head, itail, iupdat, updatd;
    double theta;
} lb_ws;

static double ddot(int n, const double *a, const double *b)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    FL(2 * n);
    return s;
}

/* Cholesky A = R^T R, R upper triangular stored in the upper triangle (LINPACK dpofa).
 * a is lda x lda row-major with (i,j) at a[i*lda+j].  Returns 0 or the failing order. */
static int dpofa(double *a, int lda, int n)
{
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int k = 0; k < j; ++k) {
            double tt = a[k * lda + j];
            for (int i = 0; i < k; ++i) tt -= a[i * lda + k] * a[i * lda + j];
            tt = tt / a[k * lda + k];
            a[k * lda + j] = tt;
            s += tt * tt;
            FL(2 * k + 3);
        }
        s = a[j * lda + j] - s;
        FL(1);
        if (s <= 0.0) return j + 1;
        a[j * lda + j] = sqrt(s);
        FL(1);
    }
    return 0;
}

/* LINPACK dtrsl with an upper-triangular T: job 01 solves T x = b, job 11 solves T^T x = b */
static int dtrsl_upper(const double *tm, int lda, int n, double *b, int transposed)
{
    for (int j = 0; j < n; ++j)
        if (tm[j * lda + j] == 0.0) return j + 1;
    if (!transposed) {
        for (int j = n - 1; j >= 0; --j) {
            double s = b[j];
            for (int k = j + 1; k < n; ++k) s -= tm[j * lda + k] * b[k];
            b[j] = s / tm[j * lda + j];
            FL(2 * (n - 1 - j) + 1);
        }
    } else {
        for (int j = 0; j < n; ++j) {
            double s = b[j];
            for (int k = 0; k < j; ++k) s -= tm[k * lda + j] * b[k];
            b[j] = s / tm[j * lda + j];
            FL(2 * j + 1);
        }
    }
    return 0;
}

static void lb_reset(lb_ws *w)
{
    w->col = 0;
    w->head = 0;
    w->theta = 1.0;
    w->iupdat = 0;
    w->updatd = 0;
}

/* ---- projgr ------------------------------------------------------------------- */
static double projgr(int n, const double *l, const double *u, const int32_t *nbd,
                     const double *x, const double *g)
{
    double sbgnrm = 0.0;
    for (int i = 0; i < n; ++i) {
        double gi = g[i];
        if (nbd[i] != 0) {
            if (gi < 0.0) {
                if (nbd[i] >= 2) gi = fmax(x[i] - u[i], gi);
            } else {
                if (nbd[i] <= 2) gi = fmin(x[i] - l[i], gi);
            }
        }
        sbgnrm = fmax(sbgnrm, fabs(gi));
    }
    FL(2 * n);
    return sbgnrm;
}

/* ---- bmv: product of the 2col x 2col middle matrix with v --------------------- */
static int bmv(const lb_ws *w, const double *v, double *p)
{
    const int col = w->col, m = w->m;
    const double *sy = w->sy;
    if (col == 0) return 0;
    p[col] = v[col];
    for (int i = 1; i < col; ++i) {
        double sum = 0.0;
        for (int k = 0; k < i; ++k) sum += sy[i * m + k] * v[k] / sy[k * m + k];
        p[col + i] = v[col + i] + sum;
        FL(3 * i + 1);
    }
    if (dtrsl_upper(w->wt, m, col, p + col, 1)) return 1;
    for (int i = 0; i < col; ++i) p[i] = v[i] / sqrt(sy[i * m + i]);
    FL(2 * col);
    if (dtrsl_upper(w->wt, m, col, p + col, 0)) return 1;
    for (int i = 0; i < col; ++i) p[i] = -p[i] / sqrt(sy[i * m + i]);
    FL(2 * col);
    for (int i = 0; i < col; ++i) {
        double sum = 0.0;
        for (int k = i + 1; k < col; ++k) sum += sy[k * m + i] * p[col + k] / sy[i * m + i];
        p[i] += sum;
        FL(3 * (col - 1 - i) + 1);
    }
    return 0;
}

/* ---- hpsolb: heap extraction of the least breakpoint (1-based n as published) -- */
static void hpsolb(int n, double *t, int *iorder, int iheap)
{
    /* arrays are used 1-based: t[1..n] */
    if (iheap == 0) {
        for (int k = 2; k <= n; ++k) {
            double ddum = t[k];
            int indxin = iorder[k];
            int i = k;
            while (i > 1) {
                int j = i / 2;
                if (ddum < t[j]) {
                    t[i] = t[j];
                    iorder[i] = iorder[j];
                    i = j;
                } else
                    break;
            }
            t[i] = ddum;
            iorder[i] = indxin;
        }
    }
    if (n > 1) {
        int i = 1;
        double out = t[1];
        int indxou = iorder[1];
        double ddum = t[n];
        int indxin = iorder[n];
        for (;;) {
            int j = i + i;
            if (j <= n - 1) {
                if (t[j + 1] < t[j]) j = j + 1;
                if (t[j] < ddum) {
                    t[i] = t[j];
                    iorder[i] = iorder[j];
                    i = j;
                    continue;
                }
            }
            break;
        }
        t[i] = ddum;
        iorder[i] = indxin;
        t[n] = out;
        iorder[n] = indxou;
    }
}

/* ---- cauchy: generalised Cauchy point ------------------------------------------ */
static int cauchy(lb_ws *w, const double *x, const double *l, const double *u,
                  const int32_t *nbd, const double *g, double sbgnrm, int *nseg_out)
{
    const int n = w->n, m = w->m, col = w->col, col2 = 2 * w->col;
    const double theta = w->theta, epsmch = DBL_EPSILON;
    double *xcp = w->z, *d = w->d, *t = w->brk; /* t is 1-based */
    int *iorder = w->iorder, *iwhere = w->iwhere;
    double *p = w->p, *c = w->c, *wbp = w->wbp, *v = w->v;

    *nseg_out = 0;
    if (sbgnrm <= 0.0) {
        memcpy(xcp, x, sizeof(double) * n);
        return 0;
    }
    int bnded = 1, nfree = n + 1, nbreak = 0, ibkmin = 0;
    double bkmin = 0.0, f1 = 0.0;
    for (int i = 0; i < col2; ++i) p[i] = 0.0;

    for (int i = 0; i < n; ++i) {
        double neggi = -g[i], tl = 0.0, tu = 0.0;
        if (iwhere[i] != 3 && iwhere[i] != -1) {
            if (nbd[i] <= 2) tl = x[i] - l[i];
            if (nbd[i] >= 2) tu = u[i] - x[i];
            int xlower = nbd[i] <= 2 && tl <= 0.0;
            int xupper = nbd[i] >= 2 && tu <= 0.0;
            iwhere[i] = 0;
            if (xlower) {
                if (neggi <= 0.0) iwhere[i] = 1;
            } else if (xupper) {
                if (neggi >= 0.0) iwhere[i] = 2;
            } else {
                if (fabs(neggi) <= 0.0) iwhere[i] = -3;
            }
            FL(2);
        }
        if (iwhere[i] != 0 && iwhere[i] != -1) {
            d[i] = 0.0;
        } else {
            d[i] = neggi;
            f1 -= neggi * neggi;
            FL(2);
            for (int j = 0; j < col; ++j) {
                int ptr = (w->head + j) % m;
                p[j] += w->wy[ptr * n + i] * neggi;
                p[col + j] += w->ws[ptr * n + i] * neggi;
            }
            FL(4 * col);
            if (nbd[i] <= 2 && nbd[i] != 0 && neggi < 0.0) {
                nbreak++;
                iorder[nbreak] = i;
                t[nbreak] = tl / (-neggi);
                FL(1);
                if (nbreak == 1 || t[nbreak] < bkmin) {
                    bkmin = t[nbreak];
                    ibkmin = nbreak;
                }
            } else if (nbd[i] >= 2 && neggi > 0.0) {
                nbreak++;
                iorder[nbreak] = i;
                t[nbreak] = tu / neggi;
                FL(1);
                if (nbreak == 1 || t[nbreak] < bkmin) {
                    bkmin = t[nbreak];
                    ibkmin = nbreak;
                }
            } else {
                nfree--;
                iorder[nfree] = i;
                if (fabs(neggi) > 0.0) bnded = 0;
            }
        }
    }
    if (theta != 1.0) {
        for (int j = 0; j < col; ++j) p[col + j] *= theta;
        FL(col);
    }
    memcpy(xcp, x, sizeof(double) * n);
    if (nbreak == 0 && nfree == n + 1) return 0;

    for (int j = 0; j < col2; ++j) c[j] = 0.0;
    double f2 = -theta * f1;
    double f2_org = f2;
    FL(1);
    if (col > 0) {
        if (bmv(w, p, v)) return 1;
        f2 -= ddot(col2, v, p);
        FL(1);
    }
    double dtm = -f1 / f2;
    FL(1);
    double tsum = 0.0;
    int nseg = 1;
    int skip_to_999 = 0;

    if (nbreak > 0) {
        int nleft = nbreak, iter = 1;
        double tj = 0.0;
        for (;;) {
            double tj0 = tj;
            int ibp;
            if (iter == 1) {
                tj = bkmin;
                ibp = iorder[ibkmin];
            } else {
                if (iter == 2) {
                    if (ibkmin != nbreak) {
                        t[ibkmin] = t[nbreak];
                        iorder[ibkmin] = iorder[nbreak];
                    }
                }
                hpsolb(nleft, t, iorder, iter - 2);
                tj = t[nleft];
                ibp = iorder[nleft];
            }
            double dt = tj - tj0;
            FL(1);
            if (dtm < dt) break; /* goto 888 */

            tsum += dt;
            nleft--;
            iter++;
            double dibp = d[ibp], zibp;
            d[ibp] = 0.0;
            if (dibp > 0.0) {
                zibp = u[ibp] - x[ibp];
                xcp[ibp] = u[ibp];
                iwhere[ibp] = 2;
            } else {
                zibp = l[ibp] - x[ibp];
                xcp[ibp] = l[ibp];
                iwhere[ibp] = 1;
            }
            FL(2);
            if (nleft == 0 && nbreak == n) {
                dtm = dt;
                skip_to_999 = 1;
                break;
            }
            nseg++;
            double dibp2 = dibp * dibp;
            f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
            f2 = f2 - theta * dibp2;
            FL(9);
            if (col > 0) {
                for (int j = 0; j < col2; ++j) c[j] += dt * p[j];
                FL(2 * col2);
                for (int j = 0; j < col; ++j) {
                    int ptr = (w->head + j) % m;
                    wbp[j] = w->wy[ptr * n + ibp];
                    wbp[col + j] = theta * w->ws[ptr * n + ibp];
                }
                FL(col);
                if (bmv(w, wbp, v)) return 1;
                double wmc = ddot(col2, c, v);
                double wmp = ddot(col2, p, v);
                double wmw = ddot(col2, wbp, v);
                for (int j = 0; j < col2; ++j) p[j] -= dibp * wbp[j];
                FL(2 * col2);
                f1 = f1 + dibp * wmc;
                f2 = f2 + 2.0 * dibp * wmp - dibp2 * wmw;
                FL(7);
            }
            f2 = fmax(epsmch * f2_org, f2);
            FL(1);
            if (nleft > 0) {
                dtm = -f1 / f2;
                FL(1);
                continue;
            } else if (bnded) {
                f1 = 0.0;
                f2 = 0.0;
                dtm = 0.0;
            } else {
                dtm = -f1 / f2;
                FL(1);
            }
            break;
        }
    }
    if (!skip_to_999) {
        if (dtm <= 0.0) dtm = 0.0;
        tsum += dtm;
        for (int i = 0; i < n; ++i) xcp[i] += tsum * d[i];
        FL(2 * n + 1);
    }
    if (col > 0) {
        for (int j = 0; j < col2; ++j) c[j] += dtm * p[j];
        FL(2 * col2);
    }
    *nseg_out = nseg;
    return 0;
}

/* ---- formk (rebuilt from scratch; see header comment) --------------------------- */
static int formk(lb_ws *w)
{
    const int n = w->n, m = w->m, col = w->col, col2 = 2 * w->col, lda = 2 * m;
    const double theta = w->theta;
    double *wn = w->wn;
    /* upper triangle of
     *   WN = [ D + Y'ZZ'Y/theta     -L_a' + R_z' ]
     *        [ -L_a + R_z           theta*S'AA'S ]
     * L_a = strictly lower part of S'AA'Y, R_z = upper (incl. diagonal) part of S'ZZ'Y */
    for (int iy = 0; iy < col; ++iy) {
        const double *wyi = w->wy + ((w->head + iy) % m) * n;
        const double *wsi = w->ws + ((w->head + iy) % m) * n;
        for (int jy = 0; jy < col; ++jy) {
            const double *wyj = w->wy + ((w->head + jy) % m) * n;
            const double *wsj = w->ws + ((w->head + jy) % m) * n;
            if (jy <= iy) {
                double yzy = 0.0, sas = 0.0;
                for (int k = 0; k < n; ++k) {
                    if (w->isfree[k])
                        yzy += wyi[k] * wyj[k];
                    else
                        sas += wsi[k] * wsj[k];
                }
                FL(2 * n);
                wn[jy * lda + iy] = yzy / theta;
                wn[(col + jy) * lda + (col + iy)] = sas * theta;
                FL(2);
            }
            /* block (1,2): element (row jy, column col+iy) from s_iy . y_jy */
            double acc = 0.0;
            if (jy < iy) {
                for (int k = 0; k < n; ++k)
                    if (!w->isfree[k]) acc += wsi[k] * wyj[k];
                wn[jy * lda + (col + iy)] = -acc;
            } else {
                for (int k = 0; k < n; ++k)
                    if (w->isfree[k]) acc += wsi[k] * wyj[k];
                wn[jy * lda + (col + iy)] = acc;
            }
            FL(2 * n);
        }
        wn[iy * lda + iy] += w->sy[iy * m + iy];
        FL(1);
    }
    if (dpofa(wn, lda, col)) return -1;
    for (int js = col; js < col2; ++js) {
        /* solve L x = column js of the (1,2) block, L' stored in the upper triangle */
        double colv[64];
        for (int i = 0; i < col; ++i) colv[i] = wn[i * lda + js];
        if (dtrsl_upper(wn, lda, col, colv, 1)) return -1;
        for (int i = 0; i < col; ++i) wn[i * lda + js] = colv[i];
    }
    for (int is = col; is < col2; ++is)
        for (int js = is; js < col2; ++js) {
            double s = 0.0;
            for (int k = 0; k < col; ++k) s += wn[k * lda + is] * wn[k * lda + js];
            wn[is * lda + js] += s;
            FL(2 * col + 1);
        }
    if (dpofa(wn + col * lda + col, lda, col)) return -2;
    return 0;
}

/* ---- cmprlb: reduced gradient r = -Z'(B(xcp - x) + g) (kept in natural indexing) -- */
static int cmprlb(lb_ws *w, const double *x, const double *g, int cnstnd)
{
    const int n = w->n, m = w->m, col = w->col;
    const double theta = w->theta;
    double *r = w->r;
    if (!cnstnd && col > 0) {
        for (int i = 0; i < n; ++i) r[i] = -g[i];
        return 0;
    }
    for (int k = 0; k < n; ++k)
        if (w->isfree[k]) {
            r[k] = -theta * (w->z[k] - x[k]) - g[k];
            FL(3);
        } else
            r[k] = 0.0;
    if (bmv(w, w->c, w->p)) return -8; /* p reused as the product (wa(1..2m) in the published code) */
    for (int j = 0; j < col; ++j) {
        int ptr = (w->head + j) % m;
        double a1 = w->p[j], a2 = theta * w->p[col + j];
        FL(1);
        for (int k = 0; k < n; ++k)
            if (w->isfree[k]) {
                r[k] += w->wy[ptr * n + k] * a1 + w->ws[ptr * n + k] * a2;
                FL(4);
            }
    }
    return 0;
}

/* ---- subsm: subspace minimisation with the Morales-Nocedal projection ------------ */
static int subsm(lb_ws *w, const double *l, const double *u, const int32_t *nbd,
                 const double *xx, const double *gg, int nsub, int *iword_out)
{
    const int n = w->n, m = w->m, col = w->col, col2 = 2 * w->col, lda = 2 * m;
    const double theta = w->theta;
    double *x = w->z, *d = w->r, *xp = w->xp, *wv = w->wv;
    if (nsub <= 0) return 0;
    for (int i = 0; i < col; ++i) {
        int ptr = (w->head + i) % m;
        double temp1 = 0.0, temp2 = 0.0;
        for (int k = 0; k < n; ++k)
            if (w->isfree[k]) {
                temp1 += w->wy[ptr * n + k] * d[k];
                temp2 += w->ws[ptr * n + k] * d[k];
                FL(4);
            }
        wv[i] = temp1;
        wv[col + i] = theta * temp2;
        FL(1);
    }
    if (dtrsl_upper(w->wn, lda, col2, wv, 1)) return 1;
    for (int i = 0; i < col; ++i) wv[i] = -wv[i];
    if (dtrsl_upper(w->wn, lda, col2, wv, 0)) return 1;
    for (int jy = 0; jy < col; ++jy) {
        int ptr = (w->head + jy) % m, js = col + jy;
        for (int k = 0; k < n; ++k)
            if (w->isfree[k]) {
                d[k] = d[k] + w->wy[ptr * n + k] * wv[jy] / theta + w->ws[ptr * n + k] * wv[js];
                FL(5);
            }
    }
    {
        double s = 1.0 / theta;
        for (int k = 0; k < n; ++k)
            if (w->isfree[k]) d[k] *= s;
        FL(nsub + 1);
    }
    /* projected Newton point */
    int iword = 0;
    memcpy(xp, x, sizeof(double) * n);
    for (int k = 0; k < n; ++k) {
        if (!w->isfree[k]) continue;
        double dk = d[k], xk = x[k];
        FL(1);
        if (nbd[k] != 0) {
            if (nbd[k] == 1) {
                x[k] = fmax(l[k], xk + dk);
                if (x[k] == l[k]) iword = 1;
            } else if (nbd[k] == 2) {
                xk = fmax(l[k], xk + dk);
                x[k] = fmin(u[k], xk);
                if (x[k] == l[k] || x[k] == u[k]) iword = 1;
            } else if (nbd[k] == 3) {
                x[k] = fmin(u[k], xk + dk);
                if (x[k] == u[k]) iword = 1;
            }
        } else
            x[k] = xk + dk;
    }
    *iword_out = iword;
    if (iword == 0) return 0;

    double dd_p = 0.0;
    for (int i = 0; i < n; ++i) dd_p += (x[i] - xx[i]) * gg[i];
    FL(3 * n);
    if (dd_p > 0.0) {
        memcpy(x, xp, sizeof(double) * n);
        double alpha = 1.0, temp1 = alpha;
        int ibd = -1;
        for (int k = 0; k < n; ++k) {
            if (!w->isfree[k]) continue;
            double dk = d[k];
            if (nbd[k] != 0) {
                if (dk < 0.0 && nbd[k] <= 2) {
                    double temp2 = l[k] - x[k];
                    if (temp2 >= 0.0)
                        temp1 = 0.0;
                    else if (dk * alpha < temp2)
                        temp1 = temp2 / dk;
                } else if (dk > 0.0 && nbd[k] >= 2) {
                    double temp2 = u[k] - x[k];
                    if (temp2 <= 0.0)
                        temp1 = 0.0;
                    else if (dk * alpha > temp2)
                        temp1 = temp2 / dk;
                }
                FL(3);
                if (temp1 < alpha) {
                    alpha = temp1;
                    ibd = k;
                }
            }
        }
        if (alpha < 1.0 && ibd >= 0) {
            double dk = d[ibd];
            if (dk > 0.0) {
                x[ibd] = u[ibd];
                d[ibd] = 0.0;
            } else if (dk < 0.0) {
                x[ibd] = l[ibd];
                d[ibd] = 0.0;
            }
        }
        for (int k = 0; k < n; ++k)
            if (w->isfree[k]) x[k] += alpha * d[k];
        FL(2 * nsub);
    }
    return 0;
}

/* ---- dcstep / dcsrch: More-Thuente line search (MINPACK-2) ----------------------- */
typedef struct {
    int brackt, stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
} ls_state;

enum { LS_START = 0, LS_FG = 1, LS_CONV = 2, LS_WARN = 3, LS_ERROR = 4 };

static void dcstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy,
                   double *stp, double fp, double dp, int *brackt, double stpmin,
                   double stpmax)
{
    double gamma, p, q, r, s, sgnd, stpc, stpf, stpq, theta;
    sgnd = dp * (*dx / fabs(*dx));
    FL(40);
    if (fp > *fx) {
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp < *stx) gamma = -gamma;
        p = (gamma - *dx) + theta;
        q = ((gamma - *dx) + gamma) + dp;
        r = p / q;
        stpc = *stx + r * (*stp - *stx);
        stpq = *stx + ((*dx / ((*fx - fp) / (*stp - *stx) + *dx)) / 2.0) * (*stp - *stx);
        if (fabs(stpc - *stx) < fabs(stpq - *stx))
            stpf = stpc;
        else
            stpf = stpc + (stpq - stpc) / 2.0;
        *brackt = 1;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp > *stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + *dx;
        r = p / q;
        stpc = *stp + r * (*stx - *stp);
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (fabs(stpc - *stp) > fabs(stpq - *stp))
            stpf = stpc;
        else
            stpf = stpq;
        *brackt = 1;
    } else if (fabs(dp) < fabs(*dx)) {
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (*dx / s) * (dp / s)));
        if (*stp > *stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (*dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0)
            stpc = *stp + r * (*stx - *stp);
        else if (*stp > *stx)
            stpc = stpmax;
        else
            stpc = stpmin;
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (*brackt) {
            if (fabs(stpc - *stp) < fabs(stpq - *stp))
                stpf = stpc;
            else
                stpf = stpq;
            if (*stp > *stx)
                stpf = fmin(*stp + 0.66 * (*sty - *stp), stpf);
            else
                stpf = fmax(*stp + 0.66 * (*sty - *stp), stpf);
        } else {
            if (fabs(stpc - *stp) > fabs(stpq - *stp))
                stpf = stpc;
            else
                stpf = stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        if (*brackt) {
            theta = 3.0 * (fp - *fy) / (*sty - *stp) + *dy + dp;
            s = fmax(fabs(theta), fmax(fabs(*dy), fabs(dp)));
            gamma = s * sqrt((theta / s) * (theta / s) - (*dy / s) * (dp / s));
            if (*stp > *sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + *dy;
            r = p / q;
            stpc = *stp + r * (*sty - *stp);
            stpf = stpc;
        } else if (*stp > *stx)
            stpf = stpmax;
        else
            stpf = stpmin;
    }
    if (fp > *fx) {
        *sty = *stp;
        *fy = fp;
        *dy = dp;
    } else {
        if (sgnd < 0.0) {
            *sty = *stx;
            *fy = *fx;
            *dy = *dx;
        }
        *stx = *stp;
        *fx = fp;
        *dx = dp;
    }
    *stp = stpf;
}

static int dcsrch(double f, double g, double *stp, double ftol, double gtol, double xtol,
                  double stpmin, double stpmax, int task, ls_state *s)
{
    const double xtrapl = 1.1, xtrapu = 4.0;
    if (task == LS_START) {
        if (*stp < stpmin) return LS_ERROR;
        if (*stp > stpmax) return LS_ERROR;
        if (g >= 0.0) return LS_ERROR;
        s->brackt = 0;
        s->stage = 1;
        s->finit = f;
        s->ginit = g;
        s->gtest = ftol * s->ginit;
        s->width = stpmax - stpmin;
        s->width1 = s->width / 0.5;
        s->stx = 0.0;
        s->fx = s->finit;
        s->gx = s->ginit;
        s->sty = 0.0;
        s->fy = s->finit;
        s->gy = s->ginit;
        s->stmin = 0.0;
        s->stmax = *stp + xtrapu * *stp;
        FL(5);
        return LS_FG;
    }
    double ftest = s->finit + *stp * s->gtest;
    FL(12);
    if (s->stage == 1 && f <= ftest && g >= 0.0) s->stage = 2;
    int out = LS_FG;
    if (s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) out = LS_WARN;
    if (s->brackt && s->stmax - s->stmin <= xtol * s->stmax) out = LS_WARN;
    if (*stp == stpmax && f <= ftest && g <= s->gtest) out = LS_WARN;
    if (*stp == stpmin && (f > ftest || g >= s->gtest)) out = LS_WARN;
    if (f <= ftest && fabs(g) <= gtol * (-s->ginit)) out = LS_CONV;
    if (out == LS_WARN || out == LS_CONV) return out;

    if (s->stage == 1 && f <= s->fx && f > ftest) {
        double fm = f - *stp * s->gtest;
        double fxm = s->fx - s->stx * s->gtest;
        double fym = s->fy - s->sty * s->gtest;
        double gm = g - s->gtest;
        double gxm = s->gx - s->gtest;
        double gym = s->gy - s->gtest;
        dcstep(&s->stx, &fxm, &gxm, &s->sty, &fym, &gym, stp, fm, gm, &s->brackt, s->stmin,
               s->stmax);
        s->fx = fxm + s->stx * s->gtest;
        s->fy = fym + s->sty * s->gtest;
        s->gx = gxm + s->gtest;
        s->gy = gym + s->gtest;
        FL(13);
    } else {
        dcstep(&s->stx, &s->fx, &s->gx, &s->sty, &s->fy, &s->gy, stp, f, g, &s->brackt,
               s->stmin, s->stmax);
    }
    if (s->brackt) {
        if (fabs(s->sty - s->stx) >= 0.66 * s->width1) *stp = s->stx + 0.5 * (s->sty - s->stx);
        s->width1 = s->width;
        s->width = fabs(s->sty - s->stx);
    }
    if (s->brackt) {
        s->stmin = fmin(s->stx, s->sty);
        s->stmax = fmax(s->stx, s->sty);
    } else {
        s->stmin = *stp + xtrapl * (*stp - s->stx);
        s->stmax = *stp + xtrapu * (*stp - s->stx);
    }
    *stp = fmax(*stp, stpmin);
    *stp = fmin(*stp, stpmax);
    if ((s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) ||
        (s->brackt && s->stmax - s->stmin <= xtol * s->stmax))
        *stp = s->stx;
    return LS_FG;
}

/* ---- matupd / formt ---------------------------------------------------------------- */
static void matupd(lb_ws *w, double rr, double dr, double stp, double dtd)
{
    const int n = w->n, m = w->m;
    if (w->iupdat <= m) {
        w->col = w->iupdat;
        w->itail = (w->head + w->iupdat - 1) % m;
    } else {
        w->itail = (w->itail + 1) % m;
        w->head = (w->head + 1) % m;
    }
    memcpy(w->ws + w->itail * n, w->d, sizeof(double) * n);
    memcpy(w->wy + w->itail * n, w->r, sizeof(double) * n);
    w->theta = rr / dr;
    FL(1);
    const int col = w->col;
    if (w->iupdat > m) {
        /* drop the oldest pair: shift ss up-left (upper triangle) and sy up-left (lower) */
        for (int j = 0; j < col - 1; ++j) {
            for (int i = 0; i <= j; ++i) w->ss[i * m + j] = w->ss[(i + 1) * m + (j + 1)];
            for (int i = j; i < col - 1; ++i) w->sy[i * m + j] = w->sy[(i + 1) * m + (j + 1)];
        }
    }
    for (int j = 0; j < col - 1; ++j) {
        int ptr = (w->head + j) % m;
        w->sy[(col - 1) * m + j] = ddot(n, w->d, w->wy + ptr * n);
        w->ss[j * m + (col - 1)] = ddot(n, w->ws + ptr * n, w->d);
    }
    if (stp == 1.0)
        w->ss[(col - 1) * m + (col - 1)] = dtd;
    else {
        w->ss[(col - 1) * m + (col - 1)] = stp * stp * dtd;
        FL(2);
    }
    w->sy[(col - 1) * m + (col - 1)] = dr;
}

static int formt(lb_ws *w)
{
    const int m = w->m, col = w->col;
    const double theta = w->theta;
    for (int j = 0; j < col; ++j) w->wt[0 * m + j] = theta * w->ss[0 * m + j];
    FL(col);
    for (int i = 1; i < col; ++i)
        for (int j = i; j < col; ++j) {
            int k1 = (i < j ? i : j);
            double ddum = 0.0;
            for (int k = 0; k < k1; ++k)
                ddum += w->sy[i * m + k] * w->sy[j * m + k] / w->sy[k * m + k];
            w->wt[i * m + j] = ddum + theta * w->ss[i * m + j];
            FL(3 * k1 + 2);
        }
    if (dpofa(w->wt, m, col)) return -3;
    return 0;
}

/* ---- driver: mainlb + SciPy's _minimize_lbfgsb loop ---------------------------------- */
int orc_lbfgsb(int n, int m, double *x, const double *l, const double *u, const int32_t *nbd,
               orc_fg_fn fg, void *user, double factr, double pgtol, int maxiter,
               int maxfun, int maxls, orc_stats *st)
{
    const double epsmch = DBL_EPSILON, big = 1.0e10;
    const double tol = factr * epsmch;
    lb_ws w;
    memset(&w, 0, sizeof(w));
    w.n = n;
    w.m = m;
    size_t nd = (size_t)2 * m * n + 7 * (size_t)n + 3 * (size_t)m * m + 4 * (size_t)m * m +
                12 * (size_t)m + 16;
    double *mem = (double *)calloc(nd, sizeof(double));
    int *imem = (int *)calloc(3 * (size_t)n + 8, sizeof(int));
    double *g = (double *)calloc(2 * (size_t)n, sizeof(double));
    double *xlast = g + n;
    if (!mem || !imem || !g) {
        free(mem);
        free(imem);
        free(g);
        return -1;
    }
    double *q = mem;
    w.ws = q; q += (size_t)m * n;
    w.wy = q; q += (size_t)m * n;
    w.sy = q; q += m * m;
    w.ss = q; q += m * m;
    w.wt = q; q += m * m;
    w.wn = q; q += 4 * m * m;
    w.z = q; q += n;
    w.r = q; q += n;
    w.d = q; q += n;
    w.t = q; q += n;
    w.xp = q; q += n;
    w.brk = q; q += n + 1;
    w.p = q; q += 2 * m;
    w.c = q; q += 2 * m;
    w.wbp = q; q += 2 * m;
    w.v = q; q += 2 * m;
    w.wv = q; q += 2 * m;
    w.iorder = imem;
    w.iwhere = imem + n + 2;
    w.isfree = imem + 2 * n + 4;

    double f = 0.0, fold = 0.0, dnorm = 0.0, gd = 0.0, gdold = 0.0, stp = 0.0, stpmx = 0.0,
           sbgnrm = 0.0, dtd = 0.0;
    double flast = 0.0; /* SciPy's OptimizeResult.fun is the last f it evaluated, even when
                         * the routine restores the previous iterate (ABNORMAL) */
    int iter = 0, nfev = 0, nit = 0, nintol = 0, nskip = 0, nrestart = 0, task = 0;
    int ifun = 0, iback = 0, info = 0, nfree = n;
    ls_state ls;
    memset(&ls, 0, sizeof(ls));
    lb_reset(&w);
    const double flops0 = g_flops;

    /* SciPy wrapper: x0 = clip(x0, lb, ub) (infinite bounds leave x unchanged) */
    for (int i = 0; i < n; ++i) {
        if ((nbd[i] == 1 || nbd[i] == 2) && x[i] < l[i]) x[i] = l[i];
        if ((nbd[i] == 2 || nbd[i] == 3) && x[i] > u[i]) x[i] = u[i];
    }
    /* active */
    int cnstnd = 0, boxed = 1;
    for (int i = 0; i < n; ++i) {
        if (nbd[i] != 2) boxed = 0;
        if (nbd[i] == 0)
            w.iwhere[i] = -1;
        else {
            cnstnd = 1;
            w.iwhere[i] = (nbd[i] == 2 && u[i] - l[i] <= 0.0) ? 3 : 0;
        }
    }

    /* FG_START.  SciPy counts a fresh evaluation whenever x differs from the last x it
     * evaluated (ScalarFunction cache); the very first request is at the clipped x0. */
    f = fg(n, x, g, user);
    flast = f;
    nfev = 1;
    memcpy(xlast, x, sizeof(double) * n);

    /* Non-finite start (behaviour of the reference recorded in tests/golden/nonfinite_*.npz,
     * SciPy 1.18.1): a NaN in x0 survives the wrapper's clip, f is NaN, every trial of the first
     * line search is rejected and after maxls of them the routine restores x0 and stops with
     * ABNORMAL -- nit 0, status 2, fun NaN.  nfev = 1 + maxls, and one more when x itself
     * holds a NaN: SciPy's evaluation cache compares x by value, a NaN never equals itself, so
     * the wrapper's closing evaluation at the restored point counts as a new one. */
    {
        int xnan = 0;
        for (int i = 0; i < n; ++i) xnan |= (x[i] != x[i]);
        if (xnan || f != f) {
            task = ORC_TASK_ABNORMAL;
            nfev = 1 + maxls + (xnan ? 1 : 0);
            goto done;
        }
    }

    sbgnrm = projgr(n, l, u, nbd, x, g);
    if (sbgnrm <= pgtol) {
        task = ORC_TASK_CONV_PGTOL;
        goto done;
    }

    for (;;) { /* label 222 */
        int nseg = 0, wrk, iword = -1;
        if (!cnstnd && w.col > 0) {
            memcpy(w.z, x, sizeof(double) * n);
            wrk = w.updatd;
        } else {
            info = cauchy(&w, x, l, u, nbd, g, sbgnrm, &nseg);
            if (info != 0) {
                info = 0;
                lb_reset(&w);
                nrestart++;
                continue;
            }
            nintol += nseg;
            /* freev: only the free set at the GCP is needed by the rebuilt formk */
            nfree = 0;
            for (int i = 0; i < n; ++i) {
                w.isfree[i] = (w.iwhere[i] <= 0);
                nfree += w.isfree[i];
            }
            wrk = 1;
        }
        (void)wrk;
        if (nfree != 0 && w.col != 0) {
            info = formk(&w);
            if (info != 0) {
                info = 0;
                lb_reset(&w);
                nrestart++;
                continue;
            }
            info = cmprlb(&w, x, g, cnstnd);
            if (info == 0) info = subsm(&w, l, u, nbd, x, g, nfree, &iword);
            if (info != 0) {
                info = 0;
                lb_reset(&w);
                nrestart++;
                continue;
            }
        }
        for (int i = 0; i < n; ++i) w.d[i] = w.z[i] - x[i];
        FL(n);

        /* ---- lnsrlb ---- */
        dtd = ddot(n, w.d, w.d);
        dnorm = sqrt(dtd);
        FL(1);
        stpmx = big;
        if (cnstnd) {
            if (iter == 0)
                stpmx = 1.0;
            else {
                for (int i = 0; i < n; ++i) {
                    double a1 = w.d[i];
                    if (nbd[i] != 0) {
                        if (a1 < 0.0 && nbd[i] <= 2) {
                            double a2 = l[i] - x[i];
                            if (a2 >= 0.0)
                                stpmx = 0.0;
                            else if (a1 * stpmx < a2)
                                stpmx = a2 / a1;
                        } else if (a1 > 0.0 && nbd[i] >= 2) {
                            double a2 = u[i] - x[i];
                            if (a2 <= 0.0)
                                stpmx = 0.0;
                            else if (a1 * stpmx > a2)
                                stpmx = a2 / a1;
                        }
                        FL(2);
                    }
                }
            }
        }
        if (iter == 0 && !boxed)
            stp = fmin(1.0 / dnorm, stpmx);
        else
            stp = 1.0;
        memcpy(w.t, x, sizeof(double) * n);
        memcpy(w.r, g, sizeof(double) * n);
        fold = f;
        ifun = 0;
        iback = 0;
        int csave = LS_START;
        int ls_done = 0; /* 1: NEW_X, 2: failure */
        while (!ls_done) {
            gd = ddot(n, g, w.d);
            if (ifun == 0) {
                gdold = gd;
                if (gd >= 0.0) {
                    info = -4;
                    ls_done = 2;
                    break;
                }
            }
            csave = dcsrch(f, gd, &stp, 1.0e-3, 0.9, 0.1, 0.0, stpmx, csave, &ls);
            if (csave != LS_CONV && csave != LS_WARN) {
                if (csave == LS_ERROR) { /* published code would spin on stale state; treat as failure */
                    info = -4;
                    ls_done = 2;
                    break;
                }
                ifun++;
                iback = ifun - 1;
                if (stp == 1.0)
                    memcpy(x, w.z, sizeof(double) * n);
                else {
                    for (int i = 0; i < n; ++i) x[i] = stp * w.d[i] + w.t[i];
                    FL(2 * n);
                }
                if (iback >= maxls) {
                    ls_done = 2;
                    break;
                }
                f = fg(n, x, g, user);
                flast = f;
                if (memcmp(x, xlast, sizeof(double) * n) != 0) {
                    nfev++;
                    memcpy(xlast, x, sizeof(double) * n);
                }
            } else
                ls_done = 1;
        }
        if (ls_done == 2) {
            memcpy(x, w.t, sizeof(double) * n);
            memcpy(g, w.r, sizeof(double) * n);
            f = fold;
            if (w.col == 0) {
                task = ORC_TASK_ABNORMAL;
                iter++;
                goto done;
            }
            info = 0;
            lb_reset(&w);
            nrestart++;
            continue;
        }
        /* NEW_X */
        iter++;
        sbgnrm = projgr(n, l, u, nbd, x, g);
        nit++;
        if (nit >= maxiter) {
            task = ORC_TASK_STOP_MAXITER;
            goto done;
        }
        if (nfev > maxfun) {
            task = ORC_TASK_STOP_MAXFUN;
            goto done;
        }
        /* label 777 */
        if (sbgnrm <= pgtol) {
            task = ORC_TASK_CONV_PGTOL;
            goto done;
        }
        {
            double ddum = fmax(fabs(fold), fmax(fabs(f), 1.0));
            FL(3);
            if ((fold - f) <= tol * ddum) {
                task = ORC_TASK_CONV_FTOL;
                goto done;
            }
        }
        for (int i = 0; i < n; ++i) w.r[i] = g[i] - w.r[i];
        FL(n);
        double rr = ddot(n, w.r, w.r), dr, ddum;
        if (stp == 1.0) {
            dr = gd - gdold;
            ddum = -gdold;
        } else {
            dr = (gd - gdold) * stp;
            for (int i = 0; i < n; ++i) w.d[i] *= stp;
            ddum = -gdold * stp;
            FL(n + 2);
        }
        FL(2);
        if (dr <= epsmch * ddum) {
            nskip++;
            w.updatd = 0;
            continue;
        }
        w.updatd = 1;
        w.iupdat++;
        matupd(&w, rr, dr, stp, dtd);
        info = formt(&w);
        if (info != 0) {
            info = 0;
            lb_reset(&w);
            nrestart++;
            continue;
        }
    }

done:
    if (st) {
        st->f = flast;
        (void)f;
        st->nit = nit;
        st->nfev = nfev;
        st->task = task;
        if (task == ORC_TASK_CONV_PGTOL || task == ORC_TASK_CONV_FTOL)
            st->status = 0;
        else if (nfev > maxfun || nit >= maxiter)
            st->status = 1;
        else
            st->status = 2;
        st->nseg_total = nintol;
        st->nupdates = w.iupdat;
        st->nskip = nskip;
        st->col_final = w.col;
        st->nrestart = nrestart;
        st->flops = g_flops - flops0;
    }
    free(mem);
    free(imem);
    free(g);
    return 0;
}
