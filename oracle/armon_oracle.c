/*
 * armon_oracle.c -- CPU restatement of Armon.jl's axis-split Lagrange+remap time step (Float64).
 *
 * TEST INFRASTRUCTURE ONLY (see armon_oracle.h).  Parity status: pinned against the reference's five
 * 64-bit golden CSVs (tests/test_oracle_golden.py).
 *
 * The structure deliberately mirrors the reference CPU path: one function per reference kernel, the
 * same 16 arrays, the same loop domains (compute_steps_ranges), the same expression order (SURVEY.md
 * Appendix A), seven passes per axis sweep.  It is therefore also the "restated reference CPU path"
 * timed as cpu_baseline by bench.py (OpenMP over rows, like the reference's `@threaded` outer loop,
 * src/generic_kernel.jl:109-193).
 *
 * All paths cited below are relative to the reference root (/root/reference at survey time).
 */
#include "armon_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ROW(nx, g) ((nx) + 2 * (g))
/* 0-based offset of cell (ix, iy), 1-based real coordinates: lin_position, src/blocking/blocking.jl:129-131 */
#define IDX(ix, iy) ((long)((iy) + g - 1) * row + ((ix) + g - 1))

/* min/max with "first argument on ties", NaN-free inputs assumed (the reference's @fastmath min/max are
 * the non-NaN-propagating kind, src/generic_kernel.jl:23-27).  Signs of zero never influence a non-zero
 * result on this path, comparisons in the tests are numerical (-0 == +0). */
static inline double dmin(double a, double b) { return (b < a) ? b : a; }
static inline double dmax(double a, double b) { return (a < b) ? b : a; }
static inline double dsign(double a) { return (a > 0.0) ? 1.0 : ((a < 0.0) ? -1.0 : 0.0); }

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * EOS -- src/kernels.jl:4-13 (perfect gas), src/kernels.jl:16-55 (Bizarrium)
 * ---------------------------------------------------------------------------------------------- */
void orc_perfect_gas_EOS(int nx, int ny, int g, orc_domain dom, double gamma,
                         const double *rho, const double *E, const double *u, const double *v,
                         double *p, double *c, double *gg)
{
    const long row = ROW(nx, g);
    (void)ny;
#pragma omp parallel for schedule(static)
    for (int iy = dom.iy0; iy <= dom.iy1; iy++) {
        for (int ix = dom.ix0; ix <= dom.ix1; ix++) {
            const long i = IDX(ix, iy);
            const double e = E[i] - 0.5 * (u[i] * u[i] + v[i] * v[i]);   /* kernels.jl:9 */
            p[i] = ((gamma - 1.) * rho[i]) * e;                           /* kernels.jl:10 */
            c[i] = sqrt((gamma * p[i]) / rho[i]);                         /* kernels.jl:11 */
            gg[i] = (1. + gamma) / 2;                                     /* kernels.jl:12 */
        }
    }
}

void orc_bizarrium_EOS(int nx, int ny, int g, orc_domain dom,
                       const double *rho, const double *u, const double *v, const double *E,
                       double *p, double *c, double *gg)
{
    const long row = ROW(nx, g);
    (void)ny;
    /* kernels.jl:25-35 */
    const double rho0 = 10000., K0 = 1e+11, Cv0 = 1000., T0 = 300., eps0 = 0., G0 = 1.5, s = 1.5;
    const double q = -42080895. / 14941154., r = 727668333. / 149411540.;
#pragma omp parallel for schedule(static)
    for (int iy = dom.iy0; iy <= dom.iy1; iy++) {
        for (int ix = dom.ix0; ix <= dom.ix1; ix++) {
            const long i = IDX(ix, iy);
            const double x = rho[i] / rho0 - 1;                 /* kernels.jl:37 */
            const double G = G0 * (1 - rho0 / rho[i]);          /* kernels.jl:38 */
            const double x2 = x * x, x3 = (x * x) * x;
            const double opx = 1 + x;
            const double opx2 = opx * opx, opx3 = (opx * opx) * opx, opx4 = (opx * opx) * (opx * opx);
            const double den = 1 - s * x;

            const double f0 = (((1 + (s / 3 - 2) * x) + q * x2) + r * x3) / den;             /* :40 */
            const double f1 = ((((s / 3 - 2) + (2 * q) * x) + (3 * r) * x2) + s * f0) / den; /* :41 */
            const double f2 = (((2 * q) + (6 * r) * x) + (2 * s) * f1) / den;                /* :42 */
            const double f3 = ((6 * r) + (3 * s) * f2) / den;                                /* :43 */

            const double epsk0 = (eps0 - (Cv0 * T0) * (1 + G)) + ((0.5 * (K0 / rho0)) * x2) * f0;      /* :45 */
            const double pk0 = (((-Cv0 * T0) * G0) * rho0) + (((0.5 * K0) * x) * opx2) * (2 * f0 + x * f1); /* :46 */
            const double pk0prime = (((-0.5 * K0) * opx3) * rho0) *
                (((2 * (1 + 3 * x)) * f0 + ((2 * x) * (2 + 3 * x)) * f1) + (x2 * opx) * f2);       /* :47 */
            const double pk0second = (((0.5 * K0) * opx4) * (rho0 * rho0)) *
                ((((12 * (1 + 2 * x)) * f0 + (6 * ((1 + 6 * x) + 6 * x2)) * f1) +
                  ((6 * x) * opx) * (1 + 2 * x) * f2) + (x2 * opx2) * f3);                           /* :48-49 */

            const double e = E[i] - 0.5 * (u[i] * u[i] + v[i] * v[i]);                            /* :51 */
            p[i] = pk0 + (G0 * rho0) * (e - epsk0);                                              /* :52 */
            c[i] = sqrt((G0 * rho0) * (p[i] - pk0) - pk0prime) / rho[i];                         /* :53 */
            gg[i] = (0.5 / (((rho[i] * rho[i]) * rho[i]) * (c[i] * c[i]))) *
                    (pk0second + ((G0 * rho0) * (G0 * rho0)) * (p[i] - pk0));                    /* :54 */
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Boundary conditions -- src/halo_exchange.jl:2-36; domain border_domain(side), src/blocking/blocking.jl:148-172
 * ghost(edge + k outward) <- real(edge - k + 1 inward), k = 1..g; rows/columns of real cells only (no corners)
 * ---------------------------------------------------------------------------------------------- */
void orc_boundary_conditions(int nx, int ny, int g, int side, double u_factor, double v_factor,
                             double *rho, double *u, double *v, double *p, double *c, double *gg, double *E)
{
    const long row = ROW(nx, g);
    const int along_x = (side == ORC_SIDE_LEFT || side == ORC_SIDE_RIGHT);
    const int n_face = along_x ? ny : nx;
#pragma omp parallel for schedule(static)
    for (int f = 1; f <= n_face; f++) {
        for (int k = 1; k <= g; k++) {
            long i, ig;
            switch (side) {
            case ORC_SIDE_LEFT:   i = IDX(k, f);          ig = IDX(1 - k, f);  break;
            case ORC_SIDE_RIGHT:  i = IDX(nx - k + 1, f); ig = IDX(nx + k, f); break;
            case ORC_SIDE_BOTTOM: i = IDX(f, k);          ig = IDX(f, 1 - k);  break;
            default:              i = IDX(f, ny - k + 1); ig = IDX(f, ny + k); break;
            }
            rho[ig] = rho[i];
            u[ig] = u[i] * u_factor;
            v[ig] = v[i] * v_factor;
            p[ig] = p[i];
            c[ig] = c[i];
            gg[ig] = gg[i];
            E[ig] = E[i];
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Riemann solvers -- src/riemann_schemes.jl:21-30 (acoustic_Godunov), :33-43 (acoustic!), :55-104 (acoustic_GAD!)
 * ---------------------------------------------------------------------------------------------- */
static inline void acoustic_godunov(double rho_i, double rho_im, double c_i, double c_im,
                                    double u_i, double u_im, double p_i, double p_im,
                                    double *us, double *ps)
{
    const double rc_l = rho_im * c_im;
    const double rc_r = rho_i * c_i;
    *us = ((rc_l * u_im + rc_r * u_i) + (p_im - p_i)) / (rc_l + rc_r);
    *ps = ((rc_r * p_im + rc_l * p_i) + (rc_l * rc_r) * (u_im - u_i)) / (rc_l + rc_r);
}

void orc_acoustic(int nx, int ny, int g, orc_domain dom, int axis,
                  double *us, double *ps, const double *rho, const double *ua, const double *p, const double *c)
{
    const long row = ROW(nx, g);
    const long s = (axis == ORC_AXIS_X) ? 1 : row;   /* stride_along, blocking.jl:197 */
    (void)ny;
#pragma omp parallel for schedule(static)
    for (int iy = dom.iy0; iy <= dom.iy1; iy++) {
        for (int ix = dom.ix0; ix <= dom.ix1; ix++) {
            const long i = IDX(ix, iy);
            acoustic_godunov(rho[i], rho[i - s], c[i], c[i - s], ua[i], ua[i - s], p[i], p[i - s],
                             &us[i], &ps[i]);
        }
    }
}

/* src/limiters.jl:6-8 */
static inline double limiter(double r, int lim)
{
    switch (lim) {
    case ORC_LIMITER_MINMOD:   return dmax(0.0, dmin(1.0, r));
    case ORC_LIMITER_SUPERBEE: return dmax(dmax(0.0, dmin(2 * r, 1.0)), dmin(r, 2.0));
    default:                   return 1.0;
    }
}

void orc_acoustic_GAD(int nx, int ny, int g, orc_domain dom, int axis, double dt, double dx, int lim,
                      double *us, double *ps, const double *rho, const double *ua, const double *p, const double *c)
{
    const long row = ROW(nx, g);
    const long s = (axis == ORC_AXIS_X) ? 1 : row;
    const double *u = ua;
    (void)ny;
#pragma omp parallel for schedule(static)
    for (int iy = dom.iy0; iy <= dom.iy1; iy++) {
        for (int ix = dom.ix0; ix <= dom.ix1; ix++) {
            const long i = IDX(ix, iy);
            double us_im, ps_im, us_i, ps_i, us_ip, ps_ip;
            /* riemann_schemes.jl:65-80 */
            acoustic_godunov(rho[i - s], rho[i - 2 * s], c[i - s], c[i - 2 * s],
                             u[i - s], u[i - 2 * s], p[i - s], p[i - 2 * s], &us_im, &ps_im);
            acoustic_godunov(rho[i], rho[i - s], c[i], c[i - s], u[i], u[i - s], p[i], p[i - s], &us_i, &ps_i);
            acoustic_godunov(rho[i + s], rho[i], c[i + s], c[i], u[i + s], u[i], p[i + s], p[i], &us_ip, &ps_ip);

            /* riemann_schemes.jl:84-87 */
            double r_um = (us_ip - u[i]) / ((us_i - u[i - s]) + 1e-6);
            double r_pm = (ps_ip - p[i]) / ((ps_i - p[i - s]) + 1e-6);
            double r_up = (u[i - s] - us_im) / ((u[i] - us_i) + 1e-6);
            double r_pp = (p[i - s] - ps_im) / ((p[i] - ps_i) + 1e-6);

            r_um = limiter(r_um, lim);
            r_pm = limiter(r_pm, lim);
            r_up = limiter(r_up, lim);
            r_pp = limiter(r_pp, lim);

            /* riemann_schemes.jl:94-100 */
            const double dm_l = rho[i - s] * dx;
            const double dm_r = rho[i] * dx;
            const double Dm = (dm_l + dm_r) / 2;
            const double rc_l = rho[i - s] * c[i - s];
            const double rc_r = rho[i] * c[i];
            const double theta = 0.5 * (1 - ((rc_l + rc_r) / 2) * (dt / Dm));

            /* riemann_schemes.jl:102-103 */
            us[i] = us_i + theta * (r_up * (u[i] - us_i) - r_um * (us_i - u[i - s]));
            ps[i] = ps_i + theta * (r_pp * (p[i] - ps_i) - r_pm * (ps_i - p[i - s]));
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Lagrangian cell update -- src/kernels.jl:58-68
 * ---------------------------------------------------------------------------------------------- */
void orc_cell_update(int nx, int ny, int g, orc_domain dom, int axis, double dx, double dt,
                     const double *us, const double *ps, double *rho, double *ua, double *E)
{
    const long row = ROW(nx, g);
    const long s = (axis == ORC_AXIS_X) ? 1 : row;
    (void)ny;
#pragma omp parallel for schedule(static)
    for (int iy = dom.iy0; iy <= dom.iy1; iy++) {
        for (int ix = dom.ix0; ix <= dom.ix1; ix++) {
            const long i = IDX(ix, iy);
            const double dm = rho[i] * dx;
            rho[i] = dm / (dx + dt * (us[i + s] - us[i]));
            ua[i] = ua[i] + (dt / dm) * (ps[i] - ps[i + s]);
            E[i] = E[i] + (dt / dm) * (ps[i] * us[i] - ps[i + s] * us[i + s]);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Remap: advection fluxes -- src/projection_schemes.jl:62-78 (1st order), :92-124 (2nd order), :15-20 (slope_minmod)
 * ---------------------------------------------------------------------------------------------- */
void orc_advection_first_order(int nx, int ny, int g, orc_domain dom, int axis, double dt,
                               const double *us, const double *rho, const double *u, const double *v, const double *E,
                               double *a_rho, double *a_urho, double *a_vrho, double *a_Erho)
{
    const long row = ROW(nx, g);
    const long s = (axis == ORC_AXIS_X) ? 1 : row;
    (void)ny;
#pragma omp parallel for schedule(static)
    for (int iy = dom.iy0; iy <= dom.iy1; iy++) {
        for (int ix = dom.ix0; ix <= dom.ix1; ix++) {
            const long is = IDX(ix, iy);
            long i = is;
            const double disp = dt * us[is];
            if (disp > 0) i = i - s;
            a_rho[is] = disp * (rho[i]);
            a_urho[is] = disp * (rho[i] * u[i]);
            a_vrho[is] = disp * (rho[i] * v[i]);
            a_Erho[is] = disp * (rho[i] * E[i]);
        }
    }
}

static inline double slope_minmod(double um, double ui, double up, double rm, double rp)
{
    const double d_p = rp * (up - ui);
    const double d_m = rm * (ui - um);
    const double s = dsign(d_p);
    return s * dmax(0.0, dmin(s * d_p, s * d_m));
}

void orc_advection_second_order(int nx, int ny, int g, orc_domain dom, int axis, double dx, double dt,
                                const double *us, const double *rho, const double *u, const double *v, const double *E,
                                double *a_rho, double *a_urho, double *a_vrho, double *a_Erho)
{
    const long row = ROW(nx, g);
    const long s = (axis == ORC_AXIS_X) ? 1 : row;
    (void)ny;
#pragma omp parallel for schedule(static)
    for (int iy = dom.iy0; iy <= dom.iy1; iy++) {
        for (int ix = dom.ix0; ix <= dom.ix1; ix++) {
            const long is = IDX(ix, iy);
            long i = is;
            const double disp = dt * us[i];
            double dxe;
            if (disp > 0) {
                dxe = -(dx - dt * us[i - s]);
                i = i - s;
            } else {
                dxe = dx + dt * us[i + s];
            }

            const double dxl_m = dx + dt * (us[i] - us[i - s]);
            const double dxl = dx + dt * (us[i + s] - us[i]);
            const double dxl_p = dx + dt * (us[i + 2 * s] - us[i + s]);

            const double r_m = (2 * dxl) / (dxl + dxl_m);
            const double r_p = (2 * dxl) / (dxl + dxl_p);

            const double sl_rho = slope_minmod(rho[i - s], rho[i], rho[i + s], r_m, r_p);
            const double sl_urho = slope_minmod(rho[i - s] * u[i - s], rho[i] * u[i], rho[i + s] * u[i + s], r_m, r_p);
            const double sl_vrho = slope_minmod(rho[i - s] * v[i - s], rho[i] * v[i], rho[i + s] * v[i + s], r_m, r_p);
            const double sl_Erho = slope_minmod(rho[i - s] * E[i - s], rho[i] * E[i], rho[i + s] * E[i + s], r_m, r_p);

            const double length_factor = dxe / (2 * dxl);
            a_rho[is] = disp * (rho[i] - sl_rho * length_factor);
            a_urho[is] = disp * (rho[i] * u[i] - sl_urho * length_factor);
            a_vrho[is] = disp * (rho[i] * v[i] - sl_vrho * length_factor);
            a_Erho[is] = disp * (rho[i] * E[i] - sl_Erho * length_factor);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Remap: projection -- src/projection_schemes.jl:23-41
 * ---------------------------------------------------------------------------------------------- */
void orc_euler_projection(int nx, int ny, int g, orc_domain dom, int axis, double dx, double dt,
                          const double *us, double *rho, double *u, double *v, double *E,
                          const double *a_rho, const double *a_urho, const double *a_vrho, const double *a_Erho)
{
    const long row = ROW(nx, g);
    const long s = (axis == ORC_AXIS_X) ? 1 : row;
    (void)ny;
#pragma omp parallel for schedule(static)
    for (int iy = dom.iy0; iy <= dom.iy1; iy++) {
        for (int ix = dom.ix0; ix <= dom.ix1; ix++) {
            const long i = IDX(ix, iy);
            const double dX = dx + dt * (us[i + s] - us[i]);
            const double t_rho = (dX * rho[i] - (a_rho[i + s] - a_rho[i])) / dx;
            const double t_urho = ((dX * rho[i]) * u[i] - (a_urho[i + s] - a_urho[i])) / dx;
            const double t_vrho = ((dX * rho[i]) * v[i] - (a_vrho[i + s] - a_vrho[i])) / dx;
            const double t_Erho = ((dX * rho[i]) * E[i] - (a_Erho[i + s] - a_Erho[i])) / dx;
            rho[i] = t_rho;
            u[i] = t_urho / t_rho;
            v[i] = t_vrho / t_rho;
            E[i] = t_Erho / t_rho;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Reductions -- src/reductions.jl:14-20 + :23-53 (dt CFL, real domain, no mask), :202-259 (conservation)
 * ---------------------------------------------------------------------------------------------- */
double orc_dtCFL(int nx, int ny, int g, const double *u, const double *v, const double *c, double dx, double dy)
{
    const long row = ROW(nx, g);
    double res = INFINITY;   /* typemax(T), reductions.jl:29 */
#pragma omp parallel for schedule(static) reduction(min : res)
    for (int iy = 1; iy <= ny; iy++) {
        for (int ix = 1; ix <= nx; ix++) {
            const long i = IDX(ix, iy);
            const double tx = dx / fabs(dmax(fabs(u[i] + c[i]), fabs(u[i] - c[i])));
            const double ty = dy / fabs(dmax(fabs(v[i] + c[i]), fabs(v[i] - c[i])));
            const double cell_dt = dmin(tx, ty);
            res = dmin(res, cell_dt);
        }
    }
    return res;
}

void orc_conservation_vars(int nx, int ny, int g, const double *rho, const double *E, double ds,
                           double *mass, double *energy)
{
    const long row = ROW(nx, g);
    double m = 0.0, en = 0.0;
    /* serial, row-major order: the reference sum order with use_threading=false (reductions.jl:226-232) */
    for (int iy = 1; iy <= ny; iy++) {
        for (int ix = 1; ix <= nx; ix++) {
            const long i = IDX(ix, iy);
            m += rho[i];
            en += rho[i] * E[i];
        }
    }
    *mass = m * ds;
    *energy = en * ds;
}

/* ------------------------------------------------------------------------------------------------
 * Initialisation -- src/kernels.jl:106-145, :71-103 ; regions src/tests.jl:59-63
 * ---------------------------------------------------------------------------------------------- */
static inline int region_high(const orc_test_case *tc, double mx, double my)
{
    switch (tc->test) {
    case ORC_TEST_SOD:       return mx <= 0.5;
    case ORC_TEST_SOD_Y:     return my <= 0.5;
    case ORC_TEST_SOD_CIRC:  return ((mx - 0.5) * (mx - 0.5) + (my - 0.5) * (my - 0.5)) <= 0.09;
    case ORC_TEST_BIZARRIUM: return mx <= 0.5;
    case ORC_TEST_SEDOV:     return (mx * mx + my * my) <= tc->sedov_r * tc->sedov_r;
    default:                 return 0;
    }
}

void orc_init_test(const orc_params *p, orc_data *d)
{
    const int nx = p->nx, ny = p->ny, g = p->g;
    const long row = ROW(nx, g);
    const double dX = p->domain_size[0] / p->global_nx;   /* kernels.jl:184 */
    const double dY = p->domain_size[1] / p->global_ny;
#pragma omp parallel for schedule(static)
    for (int iy = 1 - g; iy <= ny + g; iy++) {
        for (int ix = 1 - g; ix <= nx + g; ix++) {
            const long i = IDX(ix, iy);
            /* 0-based global index: I + global_pos - 1, global_pos = N_origin - 1 (kernels.jl:119-122,181) */
            const long gIx = (long)ix + p->origin_ix - 2;
            const long gIy = (long)iy + p->origin_iy - 2;
            d->x[i] = (double)gIx * dX + p->origin[0];
            d->y[i] = (double)gIy * dY + p->origin[1];
            const int ghost = (ix < 1 || ix > nx || iy < 1 || iy > ny);
            d->mask[i] = ghost ? 0.0 : 1.0;
            const double mx = d->x[i] + dX / 2;
            const double my = d->y[i] + dY / 2;
            if (p->tc.test == ORC_TEST_DEBUG_INDEXES) {
                const double gi = (double)(gIx + gIy * (long)nx + 1);   /* kernels.jl:138 */
                d->rho[i] = d->E[i] = d->u[i] = d->v[i] = d->p[i] = d->c[i] = d->g[i] = gi;
            } else {
                if (region_high(&p->tc, mx, my)) {
                    d->rho[i] = p->tc.high_rho; d->E[i] = p->tc.high_E;
                    d->u[i] = p->tc.high_u;     d->v[i] = p->tc.high_v;
                } else {
                    d->rho[i] = p->tc.low_rho;  d->E[i] = p->tc.low_E;
                    d->u[i] = p->tc.low_u;      d->v[i] = p->tc.low_v;
                }
                d->p[i] = 0.0; d->c[i] = 0.0; d->g[i] = 0.0;
            }
            d->us[i] = 0.0; d->ps[i] = 0.0;
            d->work_1[i] = 0.0; d->work_2[i] = 0.0; d->work_3[i] = 0.0; d->work_4[i] = 0.0;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Iteration domains -- compute_steps_ranges, src/parameters.jl:984-1025
 * ---------------------------------------------------------------------------------------------- */
void orc_steps_ranges(int nx, int ny, int g, int axis, int projection,
                      orc_domain *eos, orc_domain *fluxes, orc_domain *cell_update,
                      orc_domain *advection, orc_domain *proj)
{
    const int e = (projection == ORC_PROJ_EULER_2ND) ? 2 : 1;   /* stencil_width, projection_schemes.jl:11-12 */
    (void)g;
    const orc_domain real = {1, nx, 1, ny};
    if (eos) *eos = real;
    if (proj) *proj = real;
    orc_domain f = real, cu = real, ad = real;
    if (axis == ORC_AXIS_X) {
        f.ix0 -= e;  f.ix1 += e + 1;
        cu.ix0 -= e; cu.ix1 += e;
        ad.ix1 += 1;
    } else {
        f.iy0 -= e;  f.iy1 += e + 1;
        cu.iy0 -= e; cu.iy1 += e;
        ad.iy1 += 1;
    }
    if (fluxes) *fluxes = f;
    if (cell_update) *cell_update = cu;
    if (advection) *advection = ad;
}

/* ------------------------------------------------------------------------------------------------
 * Solver
 * ---------------------------------------------------------------------------------------------- */
orc_solver *orc_solver_create(const orc_params *p)
{
    orc_solver *s = (orc_solver *)calloc(1, sizeof(orc_solver));
    if (!s) return NULL;
    s->p = *p;
    const size_t n = (size_t)(p->nx + 2 * p->g) * (size_t)(p->ny + 2 * p->g);
    double **arrs[16] = {&s->d.x, &s->d.y, &s->d.rho, &s->d.u, &s->d.v, &s->d.E, &s->d.p, &s->d.c, &s->d.g,
                         &s->d.us, &s->d.ps, &s->d.work_1, &s->d.work_2, &s->d.work_3, &s->d.work_4, &s->d.mask};
    for (int k = 0; k < 16; k++) {
        *arrs[k] = (double *)malloc(n * sizeof(double));
        if (!*arrs[k]) { orc_solver_destroy(s); return NULL; }
    }
#ifdef _OPENMP
    if (p->nthreads > 0) omp_set_num_threads(p->nthreads);
#endif
    return s;
}

void orc_solver_destroy(orc_solver *s)
{
    if (!s) return;
    double *arrs[16] = {s->d.x, s->d.y, s->d.rho, s->d.u, s->d.v, s->d.E, s->d.p, s->d.c, s->d.g,
                        s->d.us, s->d.ps, s->d.work_1, s->d.work_2, s->d.work_3, s->d.work_4, s->d.mask};
    for (int k = 0; k < 16; k++) free(arrs[k]);
    free(s);
}

/* init_test (kernels.jl:176-214) + reset!(global_dt) (solver_state.jl:58-67) */
void orc_solver_init(orc_solver *s)
{
    orc_init_test(&s->p, &s->d);
    s->t.cycle = 0;
    s->t.time = 0.0;
    s->t.current_dt = s->p.cst_dt ? s->p.Dt : 0.0;
    s->t.next_cycle_dt = INFINITY;
    s->error = 0;
}

/* src/axis_splitting.jl:24-46 */
int orc_split_axes(int splitting, int cycle, int axes[3], double factors[3])
{
    const int even = (cycle % 2) == 0;
    switch (splitting) {
    case ORC_SPLIT_SEQUENTIAL:
        axes[0] = ORC_AXIS_X; axes[1] = ORC_AXIS_Y; factors[0] = factors[1] = 1.0; return 2;
    case ORC_SPLIT_GODUNOV:
        axes[0] = even ? ORC_AXIS_X : ORC_AXIS_Y; axes[1] = even ? ORC_AXIS_Y : ORC_AXIS_X;
        factors[0] = factors[1] = 1.0; return 2;
    case ORC_SPLIT_STRANG:
        axes[0] = axes[2] = even ? ORC_AXIS_X : ORC_AXIS_Y; axes[1] = even ? ORC_AXIS_Y : ORC_AXIS_X;
        factors[0] = factors[2] = 0.5; factors[1] = 1.0; return 3;
    case ORC_SPLIT_X_ONLY: axes[0] = ORC_AXIS_X; factors[0] = 1.0; return 1;
    default:               axes[0] = ORC_AXIS_Y; factors[0] = 1.0; return 1;
    }
}

/* update_EOS!, src/kernels.jl:151-166 */
void orc_step_EOS(orc_solver *s, int axis)
{
    orc_domain eos;
    orc_steps_ranges(s->p.nx, s->p.ny, s->p.g, axis, s->p.projection, &eos, 0, 0, 0, 0);
    if (s->p.tc.eos == ORC_EOS_BIZARRIUM)
        orc_bizarrium_EOS(s->p.nx, s->p.ny, s->p.g, eos, s->d.rho, s->d.u, s->d.v, s->d.E, s->d.p, s->d.c, s->d.g);
    else
        orc_perfect_gas_EOS(s->p.nx, s->p.ny, s->p.g, eos, s->p.tc.gamma,
                            s->d.rho, s->d.E, s->d.u, s->d.v, s->d.p, s->d.c, s->d.g);
}

/* block_ghost_exchange, src/halo_exchange.jl:286-354: the two sides along `axis`; BC on global edges,
 * neighbour exchange (hook) otherwise */
void orc_step_BC(orc_solver *s, int axis)
{
    const int first = (axis == ORC_AXIS_X) ? ORC_SIDE_LEFT : ORC_SIDE_BOTTOM;
    int any_neighbour = 0;
    for (int k = 0; k < 2; k++) {
        const int side = first + k;
        if (s->p.has_neighbour[side]) { any_neighbour = 1; continue; }
        orc_boundary_conditions(s->p.nx, s->p.ny, s->p.g, side, s->p.tc.bc_u[side], s->p.tc.bc_v[side],
                                s->d.rho, s->d.u, s->d.v, s->d.p, s->d.c, s->d.g, s->d.E);
    }
    if (any_neighbour && s->halo_exchange) s->halo_exchange(s->halo_user, axis);
}

/* numerical_fluxes!, src/riemann_schemes.jl:46-52,107-123 */
void orc_step_fluxes(orc_solver *s, int axis, double dt)
{
    orc_domain fl;
    orc_steps_ranges(s->p.nx, s->p.ny, s->p.g, axis, s->p.projection, 0, &fl, 0, 0, 0);
    const double dx = s->p.domain_size[axis] / (axis == ORC_AXIS_X ? s->p.global_nx : s->p.global_ny);
    double *ua = (axis == ORC_AXIS_X) ? s->d.u : s->d.v;
    if (s->p.riemann == ORC_RIEMANN_GAD)
        orc_acoustic_GAD(s->p.nx, s->p.ny, s->p.g, fl, axis, dt, dx, s->p.limiter,
                         s->d.us, s->d.ps, s->d.rho, ua, s->d.p, s->d.c);
    else
        orc_acoustic(s->p.nx, s->p.ny, s->p.g, fl, axis, s->d.us, s->d.ps, s->d.rho, ua, s->d.p, s->d.c);
}

/* cell_update!, src/kernels.jl:217-223 */
void orc_step_cell_update(orc_solver *s, int axis, double dt)
{
    orc_domain cu;
    orc_steps_ranges(s->p.nx, s->p.ny, s->p.g, axis, s->p.projection, 0, 0, &cu, 0, 0);
    const double dx = s->p.domain_size[axis] / (axis == ORC_AXIS_X ? s->p.global_nx : s->p.global_ny);
    double *ua = (axis == ORC_AXIS_X) ? s->d.u : s->d.v;
    orc_cell_update(s->p.nx, s->p.ny, s->p.g, cu, axis, dx, dt, s->d.us, s->d.ps, s->d.rho, ua, s->d.E);
}

/* projection_remap!, src/projection_schemes.jl:148-157 */
void orc_step_remap(orc_solver *s, int axis, double dt)
{
    orc_domain ad, pr;
    orc_steps_ranges(s->p.nx, s->p.ny, s->p.g, axis, s->p.projection, 0, 0, 0, &ad, &pr);
    const double dx = s->p.domain_size[axis] / (axis == ORC_AXIS_X ? s->p.global_nx : s->p.global_ny);
    if (s->p.projection == ORC_PROJ_EULER_2ND)
        orc_advection_second_order(s->p.nx, s->p.ny, s->p.g, ad, axis, dx, dt, s->d.us,
                                   s->d.rho, s->d.u, s->d.v, s->d.E,
                                   s->d.work_1, s->d.work_2, s->d.work_3, s->d.work_4);
    else
        orc_advection_first_order(s->p.nx, s->p.ny, s->p.g, ad, axis, dt, s->d.us,
                                  s->d.rho, s->d.u, s->d.v, s->d.E,
                                  s->d.work_1, s->d.work_2, s->d.work_3, s->d.work_4);
    orc_euler_projection(s->p.nx, s->p.ny, s->p.g, pr, axis, dx, dt, s->d.us,
                         s->d.rho, s->d.u, s->d.v, s->d.E,
                         s->d.work_1, s->d.work_2, s->d.work_3, s->d.work_4);
}

/* one axis sweep of solver_cycle, src/solver.jl:300-317 */
void orc_sweep(orc_solver *s, int axis, double dt)
{
    orc_step_EOS(s, axis);
    orc_step_BC(s, axis);
    orc_step_fluxes(s, axis, dt);
    orc_step_cell_update(s, axis, dt);
    orc_step_remap(s, axis, dt);
}

/* local_time_step, src/reductions.jl:91-110: dx, dy are the GLOBAL cell sizes */
double orc_local_time_step(orc_solver *s)
{
    const double dx = s->p.domain_size[0] / s->p.global_nx;
    const double dy = s->p.domain_size[1] / s->p.global_ny;
    return orc_dtCFL(s->p.nx, s->p.ny, s->p.g, s->d.u, s->d.v, s->d.c, dx, dy);
}

/* next_time_step (grid version, src/reductions.jl:164-199) + contribute_to_dt!/update_dt!
 * (src/solver_state.jl:70-142), synchronous path.  Returns the dt to use for THIS cycle in s->t.current_dt
 * and stores the next cycle's in s->t.next_cycle_dt.  SURVEY.md section 3.3. */
int orc_next_time_step(orc_solver *s)
{
    if (s->p.cst_dt) {
        s->t.current_dt = s->p.Dt;
        s->t.next_cycle_dt = s->p.Dt;
        return 0;
    }
    double new_dt = orc_local_time_step(s);
    if (s->allreduce_min) new_dt = s->allreduce_min(s->min_user, new_dt);   /* MPI_Iallreduce(MIN), utils.jl:126-134 */
    const double previous_dt = s->t.current_dt;
    if (!isfinite(new_dt) || new_dt <= 0) {   /* solver_state.jl:123-124 */
        s->error = 1;
        return 1;
    } else if (previous_dt == 0) {
        new_dt = s->p.cfl * new_dt;           /* :125-126 */
    } else {
        new_dt = dmin(s->p.cfl * new_dt, 1.05 * previous_dt);   /* :127-130 */
    }
    s->t.next_cycle_dt = new_dt;
    if (s->t.current_dt == 0) s->t.current_dt = s->t.next_cycle_dt;   /* :134-137 */
    return 0;
}

/* solver_cycle (src/solver.jl:288-320) followed by next_cycle! (src/solver_state.jl:145-166) */
int orc_solver_cycle(orc_solver *s)
{
    if (s->t.cycle == 0) {
        orc_step_EOS(s, ORC_AXIS_X);   /* "EOS_init", solver.jl:291-295 (EOS range is the real domain on both axes) */
    }
    if (orc_next_time_step(s)) return 1;

    int axes[3]; double factors[3];
    const int n = orc_split_axes(s->p.splitting, s->t.cycle, axes, factors);
    for (int k = 0; k < n; k++) {
        const double dt = s->t.current_dt * factors[k];   /* update_solver_state!, solver_state.jl:339-345 */
        orc_sweep(s, axes[k], dt);
    }

    /* next_cycle! */
    s->t.cycle += 1;
    s->t.time += s->t.current_dt;
    if (s->p.cst_dt) {
        s->t.current_dt = s->t.next_cycle_dt = s->p.Dt;
    } else {
        s->t.current_dt = s->t.next_cycle_dt;
        s->t.next_cycle_dt = INFINITY;
    }
    return 0;
}

/* time_loop, src/solver.jl:323-403 */
int orc_time_loop(orc_solver *s)
{
    while (s->t.time < s->p.maxtime && s->t.cycle < s->p.maxcycle) {
        if (orc_solver_cycle(s)) return 1;
    }
    return 0;
}
