// step_kernels.cu -- one CUDA kernel per `@generic_kernel` of the reference's hot path (the "kernel seam",
// SURVEY.md section 8b).  These back the per-step overloads (update_EOS!, numerical_fluxes!, cell_update!,
// advection_fluxes!, euler_projection!, boundary_conditions!, dtCFL_kernel, conservation_vars, init_test) used by
// the debug / `compare=true` path, where every intermediate array of the reference must exist.  They keep the
// reference's expression order with explicitly rounded operations (type `sd`), so that each step is bit-identical
// to the strict CPU oracle.  The production path is the fused marching kernel of sweep_kernel.cuh.
#include "common.cuh"

namespace {

constexpr int TPB = 256;

struct Dom {
    int64_t ix0, iy0, nxd, nyd;   // first cell and extent of the iteration rectangle
    int64_t row, g;
};

__host__ Dom make_dom(armon_dims d, armon_domain dom)
{
    Dom D;
    D.ix0 = dom.ix0; D.iy0 = dom.iy0;
    D.nxd = dom.ix1 - dom.ix0 + 1; D.nyd = dom.iy1 - dom.iy0 + 1;
    D.row = d.nx + 2 * d.g; D.g = d.g;
    return D;
}

// thread -> cell of the rectangle; X fastest (coalesced).  Returns false for out-of-range threads.
__device__ __forceinline__ bool cell_of_thread(const Dom &D, int64_t &i)
{
    const int64_t ix = D.ix0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t iy = D.iy0 + blockIdx.y;
    if (ix >= D.ix0 + D.nxd) return false;
    i = cell_index(ix, iy, D.row, D.g);
    return true;
}

__host__ dim3 grid_of(const Dom &D) { return dim3((unsigned)((D.nxd + TPB - 1) / TPB), (unsigned)D.nyd, 1); }

// src/kernels.jl:4-13
__global__ void k_perfect_gas_EOS(Dom D, double gamma, const double *rho, const double *E, const double *u,
                                  const double *v, double *p, double *c, double *g)
{
    int64_t i;
    if (!cell_of_thread(D, i)) return;
    sd pp, cc;
    RangeFlag f;
    eos_perfect_gas<sd, DIV_IEEE>(sd(gamma), sd(rho[i]), sd(u[i]), sd(v[i]), sd(E[i]), pp, cc, f);
    p[i] = pp.v;
    c[i] = cc.v;
    g[i] = ((sd(1.) + sd(gamma)) / sd(2.)).v;
}

// src/kernels.jl:16-55
__global__ void k_bizarrium_EOS(Dom D, const double *rho, const double *u, const double *v, const double *E,
                                double *p, double *c, double *g)
{
    int64_t i;
    if (!cell_of_thread(D, i)) return;
    sd pp, cc, gg;
    RangeFlag f;
    eos_bizarrium<sd, DIV_IEEE, true>(sd(rho[i]), sd(u[i]), sd(v[i]), sd(E[i]), pp, cc, gg, f);
    p[i] = pp.v;
    c[i] = cc.v;
    g[i] = gg.v;
}

// src/halo_exchange.jl:2-29 : one thread per border cell, loops over the ghosts
__global__ void k_boundary_conditions(int64_t nx, int64_t ny, int64_t g, int side, double u_factor, double v_factor,
                                      double *rho, double *u, double *v, double *p, double *c, double *gg, double *E)
{
    const int64_t row = nx + 2 * g;
    const int64_t n_face = (side == ARMON_SIDE_LEFT || side == ARMON_SIDE_RIGHT) ? ny : nx;
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (f > n_face) return;
    for (int64_t k = 1; k <= g; k++) {
        int64_t i, ig;
        switch (side) {
        case ARMON_SIDE_LEFT:   i = cell_index(k, f, row, g);          ig = cell_index(1 - k, f, row, g);  break;
        case ARMON_SIDE_RIGHT:  i = cell_index(nx - k + 1, f, row, g); ig = cell_index(nx + k, f, row, g); break;
        case ARMON_SIDE_BOTTOM: i = cell_index(f, k, row, g);          ig = cell_index(f, 1 - k, row, g);  break;
        default:                i = cell_index(f, ny - k + 1, row, g); ig = cell_index(f, ny + k, row, g); break;
        }
        rho[ig] = rho[i];
        u[ig] = __dmul_rn(u[i], u_factor);
        v[ig] = __dmul_rn(v[i], v_factor);
        p[ig] = p[i];
        c[ig] = c[i];
        gg[ig] = gg[i];
        E[ig] = E[i];
    }
}

// src/riemann_schemes.jl:33-43
__global__ void k_acoustic(Dom D, int64_t s, double *us, double *ps, const double *rho, const double *u,
                           const double *p, const double *c)
{
    int64_t i;
    if (!cell_of_thread(D, i)) return;
    sd a, b;
    RangeFlag f;
    acoustic_godunov<sd, DIV_IEEE>(sd(rho[i - s]) * sd(c[i - s]), sd(rho[i]) * sd(c[i]), sd(u[i - s]), sd(u[i]),
                                   sd(p[i - s]), sd(p[i]), a, b, f);
    us[i] = a.v;
    ps[i] = b.v;
}

// src/riemann_schemes.jl:55-104
template <int LIMITER>
__global__ void k_acoustic_GAD(Dom D, int64_t s, double dt_, double dx_, double *us, double *ps, const double *rho,
                               const double *u, const double *p, const double *c)
{
    int64_t i;
    if (!cell_of_thread(D, i)) return;
    const sd dt(dt_), dx(dx_);
    const sd r_mm(rho[i - 2 * s]), r_m(rho[i - s]), r_0(rho[i]), r_p(rho[i + s]);
    const sd c_mm(c[i - 2 * s]), c_m(c[i - s]), c_0(c[i]), c_p(c[i + s]);
    const sd u_mm(u[i - 2 * s]), u_m(u[i - s]), u_0(u[i]), u_p(u[i + s]);
    const sd p_mm(p[i - 2 * s]), p_m(p[i - s]), p_0(p[i]), p_p(p[i + s]);

    sd us_im, ps_im, us_i, ps_i, us_ip, ps_ip;
    RangeFlag f;
    acoustic_godunov<sd, DIV_IEEE>(r_mm * c_mm, r_m * c_m, u_mm, u_m, p_mm, p_m, us_im, ps_im, f);
    acoustic_godunov<sd, DIV_IEEE>(r_m * c_m, r_0 * c_0, u_m, u_0, p_m, p_0, us_i, ps_i, f);
    acoustic_godunov<sd, DIV_IEEE>(r_0 * c_0, r_p * c_p, u_0, u_p, p_0, p_p, us_ip, ps_ip, f);

    sd r_um = (us_ip - u_0) / ((us_i - u_m) + sd(1e-6));
    sd r_pm = (ps_ip - p_0) / ((ps_i - p_m) + sd(1e-6));
    sd r_up = (u_m - us_im) / ((u_0 - us_i) + sd(1e-6));
    sd r_pp = (p_m - ps_im) / ((p_0 - ps_i) + sd(1e-6));

    r_um = limiter<sd, LIMITER>(r_um);
    r_pm = limiter<sd, LIMITER>(r_pm);
    r_up = limiter<sd, LIMITER>(r_up);
    r_pp = limiter<sd, LIMITER>(r_pp);

    const sd dm_l = r_m * dx;
    const sd dm_r = r_0 * dx;
    const sd Dm = (dm_l + dm_r) / sd(2.);
    const sd rc_l = r_m * c_m;
    const sd rc_r = r_0 * c_0;
    const sd theta = sd(0.5) * (sd(1.) - ((rc_l + rc_r) / sd(2.)) * (dt / Dm));

    us[i] = (us_i + theta * (r_up * (u_0 - us_i) - r_um * (us_i - u_m))).v;
    ps[i] = (ps_i + theta * (r_pp * (p_0 - ps_i) - r_pm * (ps_i - p_m))).v;
}

// src/kernels.jl:58-68
__global__ void k_cell_update(Dom D, int64_t s, double dx_, double dt_, const double *us, const double *ps,
                              double *rho, double *u, double *E)
{
    int64_t i;
    if (!cell_of_thread(D, i)) return;
    const sd dx(dx_), dt(dt_);
    const sd us_i(us[i]), us_p(us[i + s]), ps_i(ps[i]), ps_p(ps[i + s]);
    const sd dm = sd(rho[i]) * dx;
    rho[i] = (dm / (dx + dt * (us_p - us_i))).v;
    u[i] = (sd(u[i]) + (dt / dm) * (ps_i - ps_p)).v;
    E[i] = (sd(E[i]) + (dt / dm) * (ps_i * us_i - ps_p * us_p)).v;
}

// src/projection_schemes.jl:62-78
__global__ void k_advection_first_order(Dom D, int64_t s, double dt_, const double *us, const double *rho,
                                        const double *u, const double *v, const double *E,
                                        double *a_r, double *a_ur, double *a_vr, double *a_Er)
{
    int64_t is;
    if (!cell_of_thread(D, is)) return;
    int64_t i = is;
    const sd disp = sd(dt_) * sd(us[is]);
    if (disp.v > 0) i = i - s;
    const sd r(rho[i]);
    a_r[is] = (disp * r).v;
    a_ur[is] = (disp * (r * sd(u[i]))).v;
    a_vr[is] = (disp * (r * sd(v[i]))).v;
    a_Er[is] = (disp * (r * sd(E[i]))).v;
}

// src/projection_schemes.jl:15-20
template <class R> __device__ __forceinline__ R slope_minmod(R um, R ui, R up, R rm, R rp)
{
    const R d_p = rp * (up - ui);
    const R d_m = rm * (ui - um);
    const R s = R(d_p.v > 0.0 ? 1.0 : (d_p.v < 0.0 ? -1.0 : 0.0));
    return s * rmax(R(0.0), rmin(s * d_p, s * d_m));
}

// src/projection_schemes.jl:92-124
__global__ void k_advection_second_order(Dom D, int64_t s, double dx_, double dt_, const double *us,
                                         const double *rho, const double *u, const double *v, const double *E,
                                         double *a_r, double *a_ur, double *a_vr, double *a_Er)
{
    int64_t is;
    if (!cell_of_thread(D, is)) return;
    const sd dx(dx_), dt(dt_);
    int64_t i = is;
    const sd disp = dt * sd(us[i]);
    sd dxe;
    if (disp.v > 0) {
        dxe = -(dx - dt * sd(us[i - s]));
        i = i - s;
    } else {
        dxe = dx + dt * sd(us[i + s]);
    }
    const sd dxl_m = dx + dt * (sd(us[i]) - sd(us[i - s]));
    const sd dxl = dx + dt * (sd(us[i + s]) - sd(us[i]));
    const sd dxl_p = dx + dt * (sd(us[i + 2 * s]) - sd(us[i + s]));

    const sd r_m = (sd(2.) * dxl) / (dxl + dxl_m);
    const sd r_p = (sd(2.) * dxl) / (dxl + dxl_p);

    const sd rm(rho[i - s]), r0(rho[i]), rp(rho[i + s]);
    const sd sl_r = slope_minmod<sd>(rm, r0, rp, r_m, r_p);
    const sd sl_u = slope_minmod<sd>(rm * sd(u[i - s]), r0 * sd(u[i]), rp * sd(u[i + s]), r_m, r_p);
    const sd sl_v = slope_minmod<sd>(rm * sd(v[i - s]), r0 * sd(v[i]), rp * sd(v[i + s]), r_m, r_p);
    const sd sl_E = slope_minmod<sd>(rm * sd(E[i - s]), r0 * sd(E[i]), rp * sd(E[i + s]), r_m, r_p);

    const sd lf = dxe / (sd(2.) * dxl);
    a_r[is] = (disp * (r0 - sl_r * lf)).v;
    a_ur[is] = (disp * (r0 * sd(u[i]) - sl_u * lf)).v;
    a_vr[is] = (disp * (r0 * sd(v[i]) - sl_v * lf)).v;
    a_Er[is] = (disp * (r0 * sd(E[i]) - sl_E * lf)).v;
}

// src/projection_schemes.jl:23-41
__global__ void k_euler_projection(Dom D, int64_t s, double dx_, double dt_, const double *us, double *rho,
                                   double *u, double *v, double *E, const double *a_r, const double *a_ur,
                                   const double *a_vr, const double *a_Er)
{
    int64_t i;
    if (!cell_of_thread(D, i)) return;
    const sd dx(dx_), dt(dt_);
    const sd dX = dx + dt * (sd(us[i + s]) - sd(us[i]));
    const sd r(rho[i]);
    const sd t_r = (dX * r - (sd(a_r[i + s]) - sd(a_r[i]))) / dx;
    const sd t_ur = ((dX * r) * sd(u[i]) - (sd(a_ur[i + s]) - sd(a_ur[i]))) / dx;
    const sd t_vr = ((dX * r) * sd(v[i]) - (sd(a_vr[i + s]) - sd(a_vr[i]))) / dx;
    const sd t_Er = ((dX * r) * sd(E[i]) - (sd(a_Er[i + s]) - sd(a_Er[i]))) / dx;
    rho[i] = t_r.v;
    u[i] = (t_ur / t_r).v;
    v[i] = (t_vr / t_r).v;
    E[i] = (t_Er / t_r).v;
}

// ---- reductions -------------------------------------------------------------------------------------
// src/reductions.jl:14-20 : min over real cells; the min of doubles is order-independent, so one atomicMin on the
// order-preserving integer image of the (non-negative) per-cell value is exact.  NaN maps above +Inf and is
// replaced by 0 so that an invalid state is caught by the `dt <= 0` test of update_dt! (solver_state.jl:123).
__global__ void k_dtCFL(int64_t nx, int64_t ny, int64_t g, const double *u, const double *v, const double *c,
                        double dx, double dy, unsigned long long *result)
{
    const int64_t row = nx + 2 * g;
    const int64_t ix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int64_t iy = blockIdx.y + 1;
    double val = __longlong_as_double(0x7FF0000000000000LL);   // +Inf
    if (ix <= nx) {
        const int64_t i = cell_index(ix, iy, row, g);
        const double uu = u[i], vv = v[i], cc = c[i];
        const double tx = __ddiv_rn(dx, fabs(fmax(fabs(__dadd_rn(uu, cc)), fabs(__dsub_rn(uu, cc)))));
        const double ty = __ddiv_rn(dy, fabs(fmax(fabs(__dadd_rn(vv, cc)), fabs(__dsub_rn(vv, cc)))));
        val = fmin(tx, ty);
        if (val != val || uu != uu || vv != vv || cc != cc) val = 0.0;
    }
    unsigned long long bits = (unsigned long long)__double_as_longlong(val);
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, off);
        bits = other < bits ? other : bits;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(result, bits);
}

// src/reductions.jl:202-259 : sum(rho), sum(rho*E) over real cells.  Sums are order-dependent; a fixed two-level
// tree (per-row serial-strided partials, then a serial pass over rows) makes the result reproducible run to run.
__global__ void k_conservation_rows(int64_t nx, int64_t ny, int64_t g, const double *rho, const double *E,
                                    double *row_mass, double *row_energy)
{
    const int64_t row = nx + 2 * g;
    const int64_t iy = blockIdx.x + 1;
    __shared__ double sm[TPB], se[TPB];
    double m = 0.0, e = 0.0;
    for (int64_t ix = threadIdx.x + 1; ix <= nx; ix += TPB) {
        const int64_t i = cell_index(ix, iy, row, g);
        m = __dadd_rn(m, rho[i]);
        e = __dadd_rn(e, __dmul_rn(rho[i], E[i]));
    }
    sm[threadIdx.x] = m; se[threadIdx.x] = e;
    __syncthreads();
    for (int off = TPB / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sm[threadIdx.x] = __dadd_rn(sm[threadIdx.x], sm[threadIdx.x + off]);
            se[threadIdx.x] = __dadd_rn(se[threadIdx.x], se[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { row_mass[blockIdx.x] = sm[0]; row_energy[blockIdx.x] = se[0]; }
}

__global__ void k_conservation_final(int64_t ny, const double *row_mass, const double *row_energy, double ds,
                                     double *out)
{
    __shared__ double sm[TPB], se[TPB];
    double m = 0.0, e = 0.0;
    for (int64_t r = threadIdx.x; r < ny; r += TPB) {
        m = __dadd_rn(m, row_mass[r]);
        e = __dadd_rn(e, row_energy[r]);
    }
    sm[threadIdx.x] = m; se[threadIdx.x] = e;
    __syncthreads();
    for (int off = TPB / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sm[threadIdx.x] = __dadd_rn(sm[threadIdx.x], sm[threadIdx.x + off]);
            se[threadIdx.x] = __dadd_rn(se[threadIdx.x], se[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = __dmul_rn(sm[0], ds); out[1] = __dmul_rn(se[0], ds); }
}

// src/kernels.jl:106-145 (+ init_vars :71-103, regions src/tests.jl:59-63)
__device__ __forceinline__ bool region_high(const armon_test_case &tc, double mx, double my)
{
    switch (tc.test) {
    case ARMON_TEST_SOD:       return mx <= 0.5;
    case ARMON_TEST_SOD_Y:     return my <= 0.5;
    case ARMON_TEST_SOD_CIRC:  {
        const double ax = __dsub_rn(mx, 0.5), ay = __dsub_rn(my, 0.5);
        return __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)) <= 0.09;
    }
    case ARMON_TEST_BIZARRIUM: return mx <= 0.5;
    case ARMON_TEST_SEDOV:
        return __dadd_rn(__dmul_rn(mx, mx), __dmul_rn(my, my)) <= __dmul_rn(tc.sedov_r, tc.sedov_r);
    default:                   return false;
    }
}

struct InitArrays {
    double *x, *y, *mask, *rho, *E, *u, *v, *p, *c, *g, *us, *ps, *w1, *w2, *w3, *w4;
};

__global__ void k_init_test(int64_t nx, int64_t ny, int64_t g, int64_t origin_ix, int64_t origin_iy,
                            double dX, double dY, double ox, double oy, armon_test_case tc, InitArrays A)
{
    const int64_t row = nx + 2 * g;
    const int64_t ix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1 - g;
    const int64_t iy = (int64_t)blockIdx.y + 1 - g;
    if (ix > nx + g) return;
    const int64_t i = cell_index(ix, iy, row, g);
    const int64_t gIx = ix + origin_ix - 2;
    const int64_t gIy = iy + origin_iy - 2;
    const double x = __dadd_rn(__dmul_rn((double)gIx, dX), ox);
    const double y = __dadd_rn(__dmul_rn((double)gIy, dY), oy);
    if (A.x) A.x[i] = x;
    if (A.y) A.y[i] = y;
    const bool ghost = (ix < 1 || ix > nx || iy < 1 || iy > ny);
    if (A.mask) A.mask[i] = ghost ? 0.0 : 1.0;
    const double mx = __dadd_rn(x, __ddiv_rn(dX, 2.0));
    const double my = __dadd_rn(y, __ddiv_rn(dY, 2.0));
    if (tc.test == ARMON_TEST_DEBUG_INDEXES) {
        const double gi = (double)(gIx + gIy * nx + 1);
        A.rho[i] = gi; A.E[i] = gi; A.u[i] = gi; A.v[i] = gi;
        if (A.p) A.p[i] = gi;
        if (A.c) A.c[i] = gi;
        if (A.g) A.g[i] = gi;
    } else {
        const bool high = region_high(tc, mx, my);
        A.rho[i] = high ? tc.high_rho : tc.low_rho;
        A.E[i] = high ? tc.high_E : tc.low_E;
        A.u[i] = high ? tc.high_u : tc.low_u;
        A.v[i] = high ? tc.high_v : tc.low_v;
        if (A.p) A.p[i] = 0.0;
        if (A.c) A.c[i] = 0.0;
        if (A.g) A.g[i] = 0.0;
    }
    if (A.us) A.us[i] = 0.0;
    if (A.ps) A.ps[i] = 0.0;
    if (A.w1) A.w1[i] = 0.0;
    if (A.w2) A.w2[i] = 0.0;
    if (A.w3) A.w3[i] = 0.0;
    if (A.w4) A.w4[i] = 0.0;
}

__global__ void k_fill(double *dst, double value, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = value;
}

__global__ void k_fill_ghosts(int64_t nx, int64_t ny, int64_t g, double *arr, double value)
{
    const int64_t row = nx + 2 * g;
    const int64_t ix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1 - g;
    const int64_t iy = (int64_t)blockIdx.y + 1 - g;
    if (ix > nx + g) return;
    if (ix < 1 || ix > nx || iy < 1 || iy > ny) arr[cell_index(ix, iy, row, g)] = value;
}

int check_dims(armon_dims d)
{
    ARMON_CHECK_ARG(d.nx > 0 && d.ny > 0 && d.g >= 1, "block dimensions");
    ARMON_CHECK_ARG(d.ny + 2 * d.g <= 65535, "ny + 2g must fit gridDim.y for the per-step kernels");
    return ARMON_OK;
}

int check_dom(armon_dims d, armon_domain dom)
{
    ARMON_CHECK_ARG(dom.ix0 >= 1 - d.g && dom.ix1 <= d.nx + d.g && dom.iy0 >= 1 - d.g && dom.iy1 <= d.ny + d.g,
                    "iteration domain exceeds the block");
    ARMON_CHECK_ARG(dom.ix1 >= dom.ix0 && dom.iy1 >= dom.iy0, "empty iteration domain");
    return ARMON_OK;
}

inline int64_t stride_of(armon_dims d, int axis) { return axis == ARMON_AXIS_X ? 1 : d.nx + 2 * d.g; }

}   // namespace

#define STEP_PROLOGUE()                                      \
    ARMON_CHECK_ARG(ctx != nullptr, "null context");         \
    if (int rc = armon_ctx_activate(ctx)) return rc;         \
    if (int rc = check_dims(d)) return rc;

#define STEP_DOM_PROLOGUE()                                  \
    STEP_PROLOGUE()                                          \
    if (int rc = check_dom(d, dom)) return rc;               \
    const Dom D = make_dom(d, dom);

extern "C" {

int armon_perfect_gas_EOS(armon_ctx *ctx, armon_dims d, armon_domain dom, double gamma, const double *rho,
                          const double *E, const double *u, const double *v, double *p, double *c, double *g)
{
    STEP_DOM_PROLOGUE();
    k_perfect_gas_EOS<<<grid_of(D), TPB, 0, ctx->stream>>>(D, gamma, rho, E, u, v, p, c, g);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_bizarrium_EOS(armon_ctx *ctx, armon_dims d, armon_domain dom, const double *rho, const double *u,
                        const double *v, const double *E, double *p, double *c, double *g)
{
    STEP_DOM_PROLOGUE();
    k_bizarrium_EOS<<<grid_of(D), TPB, 0, ctx->stream>>>(D, rho, u, v, E, p, c, g);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_boundary_conditions(armon_ctx *ctx, armon_dims d, int side, double u_factor, double v_factor, double *rho,
                              double *u, double *v, double *p, double *c, double *g, double *E)
{
    STEP_PROLOGUE();
    ARMON_CHECK_ARG(side >= 0 && side < 4, "side");
    const int64_t n_face = (side == ARMON_SIDE_LEFT || side == ARMON_SIDE_RIGHT) ? d.ny : d.nx;
    k_boundary_conditions<<<(unsigned)((n_face + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(
        d.nx, d.ny, d.g, side, u_factor, v_factor, rho, u, v, p, c, g, E);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_acoustic(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double *us, double *ps,
                   const double *rho, const double *ua, const double *p, const double *c)
{
    STEP_DOM_PROLOGUE();
    k_acoustic<<<grid_of(D), TPB, 0, ctx->stream>>>(D, stride_of(d, axis), us, ps, rho, ua, p, c);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_acoustic_GAD(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dt, double dx, int limiter_,
                       double *us, double *ps, const double *rho, const double *ua, const double *p, const double *c)
{
    STEP_DOM_PROLOGUE();
    const int64_t s = stride_of(d, axis);
    switch (limiter_) {
    case ARMON_LIMITER_NONE:
        k_acoustic_GAD<ARMON_LIMITER_NONE><<<grid_of(D), TPB, 0, ctx->stream>>>(D, s, dt, dx, us, ps, rho, ua, p, c);
        break;
    case ARMON_LIMITER_MINMOD:
        k_acoustic_GAD<ARMON_LIMITER_MINMOD><<<grid_of(D), TPB, 0, ctx->stream>>>(D, s, dt, dx, us, ps, rho, ua, p, c);
        break;
    case ARMON_LIMITER_SUPERBEE:
        k_acoustic_GAD<ARMON_LIMITER_SUPERBEE><<<grid_of(D), TPB, 0, ctx->stream>>>(D, s, dt, dx, us, ps, rho, ua, p, c);
        break;
    default:
        armon_set_error("unknown limiter %d", limiter_);
        return ARMON_ERR_INVALID;
    }
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_cell_update(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dx, double dt,
                      const double *us, const double *ps, double *rho, double *ua, double *E)
{
    STEP_DOM_PROLOGUE();
    k_cell_update<<<grid_of(D), TPB, 0, ctx->stream>>>(D, stride_of(d, axis), dx, dt, us, ps, rho, ua, E);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_advection_first_order(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dt, const double *us,
                                const double *rho, const double *u, const double *v, const double *E,
                                double *a_r, double *a_ur, double *a_vr, double *a_Er)
{
    STEP_DOM_PROLOGUE();
    k_advection_first_order<<<grid_of(D), TPB, 0, ctx->stream>>>(D, stride_of(d, axis), dt, us, rho, u, v, E,
                                                                 a_r, a_ur, a_vr, a_Er);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_advection_second_order(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dx, double dt,
                                 const double *us, const double *rho, const double *u, const double *v,
                                 const double *E, double *a_r, double *a_ur, double *a_vr, double *a_Er)
{
    STEP_DOM_PROLOGUE();
    k_advection_second_order<<<grid_of(D), TPB, 0, ctx->stream>>>(D, stride_of(d, axis), dx, dt, us, rho, u, v, E,
                                                                  a_r, a_ur, a_vr, a_Er);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_euler_projection(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dx, double dt,
                           const double *us, double *rho, double *u, double *v, double *E, const double *a_r,
                           const double *a_ur, const double *a_vr, const double *a_Er)
{
    STEP_DOM_PROLOGUE();
    k_euler_projection<<<grid_of(D), TPB, 0, ctx->stream>>>(D, stride_of(d, axis), dx, dt, us, rho, u, v, E,
                                                            a_r, a_ur, a_vr, a_Er);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_dtCFL(armon_ctx *ctx, armon_dims d, const double *u, const double *v, const double *c, double dx, double dy,
                double *result)
{
    STEP_PROLOGUE();
    ARMON_CHECK_ARG(result != nullptr, "null result");
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(ctx->scratch);
    const unsigned long long inf_bits = 0x7FF0000000000000ULL;
    ARMON_CUDA(cudaMemcpyAsync(acc, &inf_bits, sizeof(inf_bits), cudaMemcpyHostToDevice, ctx->stream));
    const dim3 grid((unsigned)((d.nx + TPB - 1) / TPB), (unsigned)d.ny, 1);
    k_dtCFL<<<grid, TPB, 0, ctx->stream>>>(d.nx, d.ny, d.g, u, v, c, dx, dy, acc);
    ARMON_LAUNCH_CHECK(ctx);
    ARMON_CUDA(cudaMemcpyAsync(ctx->pinned, acc, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ARMON_CUDA(cudaStreamSynchronize(ctx->stream));
    *result = ctx->pinned[0];
    return ARMON_OK;
}

int armon_conservation_vars(armon_ctx *ctx, armon_dims d, const double *rho, const double *E, double ds, double *mass,
                            double *energy)
{
    STEP_PROLOGUE();
    ARMON_CHECK_ARG(mass != nullptr && energy != nullptr, "null result");
    ARMON_CHECK_ARG((size_t)(2 * d.ny + 2) <= ctx->scratch_elems, "ny too large for the reduction scratch");
    double *row_mass = ctx->scratch + 2, *row_energy = ctx->scratch + 2 + d.ny;
    k_conservation_rows<<<(unsigned)d.ny, TPB, 0, ctx->stream>>>(d.nx, d.ny, d.g, rho, E, row_mass, row_energy);
    ARMON_LAUNCH_CHECK(ctx);
    k_conservation_final<<<1, TPB, 0, ctx->stream>>>(d.ny, row_mass, row_energy, ds, ctx->scratch);
    ARMON_LAUNCH_CHECK(ctx);
    ARMON_CUDA(cudaMemcpyAsync(ctx->pinned, ctx->scratch, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ARMON_CUDA(cudaStreamSynchronize(ctx->stream));
    *mass = ctx->pinned[0];
    *energy = ctx->pinned[1];
    return ARMON_OK;
}

int armon_init_test(armon_ctx *ctx, armon_dims d, int64_t origin_ix, int64_t origin_iy, int64_t global_nx,
                    int64_t global_ny, const double domain_size[2], const double origin[2], const armon_test_case *tc,
                    double *x, double *y, double *mask, double *rho, double *E, double *u, double *v, double *p,
                    double *c, double *g, double *us, double *ps, double *work_1, double *work_2, double *work_3,
                    double *work_4)
{
    STEP_PROLOGUE();
    ARMON_CHECK_ARG(tc && rho && E && u && v, "rho, E, u, v and the test case are required");
    ARMON_CHECK_ARG(global_nx > 0 && global_ny > 0, "global grid");
    const double dX = domain_size[0] / (double)global_nx;
    const double dY = domain_size[1] / (double)global_ny;
    InitArrays A{x, y, mask, rho, E, u, v, p, c, g, us, ps, work_1, work_2, work_3, work_4};
    const dim3 grid((unsigned)((d.nx + 2 * d.g + TPB - 1) / TPB), (unsigned)(d.ny + 2 * d.g), 1);
    k_init_test<<<grid, TPB, 0, ctx->stream>>>(d.nx, d.ny, d.g, origin_ix, origin_iy, dX, dY, origin[0], origin[1],
                                               *tc, A);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_fill(armon_ctx *ctx, double *dst_dev, double value, uint64_t n_elems)
{
    ARMON_CHECK_ARG(ctx != nullptr, "null context");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    if (n_elems == 0) return ARMON_OK;
    const unsigned blocks = (unsigned)((n_elems + TPB - 1) / TPB < 148 * 16 ? (n_elems + TPB - 1) / TPB : 148 * 16);
    k_fill<<<blocks, TPB, 0, ctx->stream>>>(dst_dev, value, n_elems);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

int armon_fill_ghosts(armon_ctx *ctx, armon_dims d, double *arr, double value)
{
    STEP_PROLOGUE();
    const dim3 grid((unsigned)((d.nx + 2 * d.g + TPB - 1) / TPB), (unsigned)(d.ny + 2 * d.g), 1);
    k_fill_ghosts<<<grid, TPB, 0, ctx->stream>>>(d.nx, d.ny, d.g, arr, value);
    ARMON_LAUNCH_CHECK(ctx);
    return ARMON_OK;
}

}   // extern "C"
