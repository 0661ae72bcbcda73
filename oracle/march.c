/*
 * oracle/march.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's vertical march
 *     bldfm.solver.ivp_solver          (/root/reference/src/bldfm/solver.py:307-374)
 * used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker for the CUDA path.  The product
 * (bldfm_b200/) never links, imports or calls anything in this directory.
 *
 * Parity pinning: oracle_ivp() is compared BITWISE with the reference's numba-JITted
 * ivp_solver on random modes (tests/golden/make_golden.py -> tests/golden/ivp_*.npz,
 * tests/test_oracle.py).  To make that possible every +,-,* below is an individually
 * rounded IEEE-754 binary64 operation: build with -ffp-contract=off (see Makefile).
 *
 * Operation order per march step (solver.py:357-368), T = tr + i*ti:
 *   Ti = -(Kx[i]*Lx**2 + Ky[i]*Ly**2) - 1j*u[i]*Lx - 1j*v[i]*Ly          :357
 *   a  = 1.0 - 0.5*Kzinv*Ti*dzi**2                                       :361 (= d, :364)
 *   b  = -Kzinv*dzi - 1.0/6.0*Kzinv**2*Ti*dzi**3                         :362 (sign quirk kept)
 *   c  = Ti*dzi - 1.0/6.0*Kzinv*Ti**2*dzi**3                             :363
 *   (p,q) <- (a*p + b*q, c*p + d*q)                                      :366-368
 * Snapshot of (p,q) BEFORE step i when i is in `levels` (:352-355) and after the
 * last step when nz-1 is in `levels` (:370-372); rows are filled in visit order.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
    double Kx, Ky, u, v;      /* profile values at level i                       */
    double s;                 /* 0.5*Kzinv                                       */
    double h, h2, h3;         /* dz, dz*dz, (dz*dz)*dz                           */
    double s6;                /* (1.0/6.0)*(Kzinv*Kzinv)                         */
    double s61;               /* (1.0/6.0)*Kzinv                                 */
    double c0;                /* (-Kzinv)*dz                                     */
} level_coef;

static void fill_level_coefs(int nz, const double *z, const double *u, const double *v,
                             const double *Kx, const double *Ky, const double *Kz,
                             level_coef *lc)
{
    for (int i = 0; i < nz - 1; ++i) {
        double kinv = 1.0 / Kz[i];            /* :358 */
        double h = z[i + 1] - z[i];           /* :341 np.diff */
        lc[i].Kx = Kx[i]; lc[i].Ky = Ky[i]; lc[i].u = u[i]; lc[i].v = v[i];
        lc[i].s = 0.5 * kinv;
        lc[i].h = h;
        lc[i].h2 = h * h;
        lc[i].h3 = (h * h) * h;
        lc[i].s6 = (1.0 / 6.0) * (kinv * kinv);
        lc[i].s61 = (1.0 / 6.0) * kinv;
        lc[i].c0 = (-kinv) * h;
    }
}

typedef struct {
    int64_t m0, m1, M;
    const double *p0, *q0, *Lx, *Ly;
    int nz;
    const level_coef *lc;
    const int *row_of;
    double *p_top, *q_top, *P, *Q;
} march_job;

/* march modes [m0, m1): one mode at a time, all levels, state in registers */
static void *march_range(void *arg)
{
    const march_job *j = (const march_job *)arg;
    const int nz = j->nz;
    const size_t M = (size_t)j->M;
    for (int64_t m = j->m0; m < j->m1; ++m) {
        const double lx = j->Lx[m], ly = j->Ly[m];
        const double lx2 = lx * lx, ly2 = ly * ly;
        double pr = j->p0[2 * m], pi = j->p0[2 * m + 1];
        double qr = j->q0[2 * m], qi = j->q0[2 * m + 1];

        for (int i = 0; i < nz - 1; ++i) {
            if (j->row_of[i] >= 0) {
                size_t o = 2 * ((size_t)j->row_of[i] * M + (size_t)m);
                j->P[o] = pr; j->P[o + 1] = pi; j->Q[o] = qr; j->Q[o + 1] = qi;
            }
            const level_coef *c = &j->lc[i];
            double tr = -(c->Kx * lx2 + c->Ky * ly2);
            double ti = -(c->u * lx) - (c->v * ly);
            double ar = 1.0 - (c->s * tr) * c->h2;
            double ai = 0.0 - (c->s * ti) * c->h2;
            double br = c->c0 - (c->s6 * tr) * c->h3;
            double bi = 0.0 - (c->s6 * ti) * c->h3;
            double t2r = tr * tr - ti * ti;
            double t2i = tr * ti + ti * tr;
            double cr = tr * c->h - (c->s61 * t2r) * c->h3;
            double ci = ti * c->h - (c->s61 * t2i) * c->h3;

            double npr = (ar * pr - ai * pi) + (br * qr - bi * qi);
            double npi = (ar * pi + ai * pr) + (br * qi + bi * qr);
            double nqr = (cr * pr - ci * pi) + (ar * qr - ai * qi);
            double nqi = (cr * pi + ci * pr) + (ar * qi + ai * qr);
            pr = npr; pi = npi; qr = nqr; qi = nqi;
        }
        if (nz >= 1 && j->row_of[nz - 1] >= 0) {
            size_t o = 2 * ((size_t)j->row_of[nz - 1] * M + (size_t)m);
            j->P[o] = pr; j->P[o + 1] = pi; j->Q[o] = qr; j->Q[o + 1] = qi;
        }
        j->p_top[2 * m] = pr; j->p_top[2 * m + 1] = pi;
        j->q_top[2 * m] = qr; j->q_top[2 * m + 1] = qi;
    }
    return NULL;
}

int oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/*
 * All complex arrays are interleaved (re,im) doubles, i.e. numpy complex128.
 *   p0,q0        [M]        initial state
 *   levels       [nlv]      int64 level indices (membership test `i in levels`)
 *   Lx,Ly        [M]        wavenumbers per mode
 *   p_top,q_top  [M]        state after the last step
 *   P,Q          [nlv][M]   snapshots (rows not visited stay zero)
 *   nthreads                pthreads used over disjoint mode ranges (modes are independent)
 * Returns 0.
 */
int oracle_ivp(int64_t M, const double *p0, const double *q0,
               int nz, const double *z,
               const double *u, const double *v,
               const double *Kx, const double *Ky, const double *Kz,
               int nlv, const int64_t *levels,
               const double *Lx, const double *Ly,
               double *p_top, double *q_top, double *P, double *Q,
               int nthreads)
{
    level_coef *lc = (level_coef *)malloc(sizeof(level_coef) * (size_t)(nz > 1 ? nz - 1 : 1));
    int *row_of = (int *)malloc(sizeof(int) * (size_t)(nz > 0 ? nz : 1));
    fill_level_coefs(nz, z, u, v, Kx, Ky, Kz, lc);

    /* row_of[i] = output row written when level i is visited, else -1 (:348-355,370-372) */
    int lvl = 0;
    for (int i = 0; i < nz; ++i) {
        int hit = 0;
        for (int k = 0; k < nlv; ++k) if (levels[k] == i) { hit = 1; break; }
        row_of[i] = hit ? lvl++ : -1;
    }
    memset(P, 0, sizeof(double) * 2 * (size_t)nlv * (size_t)M);
    memset(Q, 0, sizeof(double) * 2 * (size_t)nlv * (size_t)M);

    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > M) nthreads = M > 0 ? (int)M : 1;
    march_job jobs[256];
    pthread_t tid[256];
    for (int t = 0; t < nthreads; ++t) {
        march_job *j = &jobs[t];
        j->m0 = M * t / nthreads; j->m1 = M * (t + 1) / nthreads; j->M = M;
        j->p0 = p0; j->q0 = q0; j->Lx = Lx; j->Ly = Ly; j->nz = nz; j->lc = lc;
        j->row_of = row_of; j->p_top = p_top; j->q_top = q_top; j->P = P; j->Q = Q;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&tid[t], NULL, march_range, &jobs[t]);
    march_range(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(tid[t], NULL);

    free(lc);
    free(row_of);
    return 0;
}
