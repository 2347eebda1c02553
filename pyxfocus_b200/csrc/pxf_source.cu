// Ray sources on the device (sources.py:20-53 pointsource, :56-88 circularbeam, :91-127 annulus,
// :130-170 subannulus).  Either from caller-supplied uniforms (numpy's legacy MT19937 stream
// uploaded by the Python layer: "identical ray seeds") or from a counter-based Philox4x32-10
// generator for bundles too large to draw on the host.  Writes all ten rows once: 80 B/ray.
#include "pxf_internal.h"
#include "pxf_ray.cuh"

namespace pxf {

struct RowPtrs { double *p[10]; };

struct u32x4 { unsigned a, b, c, d; };

PXF_DEV u32x4 philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    u32x4 o = {c0, c1, c2, c3};
    return o;
}

// numpy's random_sample: ((a>>5)*2^26 + (b>>6)) / 2^53
PXF_DEV double u53(unsigned a, unsigned b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

struct SourceP { int kind; double a, b, c, d; double pi; };

PXF_DEV void make_ray(Ray &r, const SourceP &p, double u1, double u2)
{
    r.opd = 0.; r.x = 0.; r.y = 0.; r.z = 0.; r.l = 0.; r.m = 0.; r.n = 0.; r.ux = 0.; r.uy = 0.; r.uz = 0.;
    double s, c;
    if (p.kind == 0) {            // subannulus(rin=a, rout=b, dphi=c, zhat=d)
        double rho = sqrt(p.a * p.a + u1 * (p.b * p.b - p.a * p.a));
        double theta = u2 * p.c - p.c / 2.;
        sincos(theta, &s, &c);
        r.x = rho * c; r.y = rho * s; r.n = p.d;
    } else if (p.kind == 1) {     // circularbeam(rad=a)
        double rho = sqrt(u1) * p.a;
        double theta = u2 * 2 * p.pi;          // rand*2*np.pi
        sincos(theta, &s, &c);
        r.x = rho * c; r.y = rho * s; r.n = 1.;
    } else if (p.kind == 2) {     // pointsource(ang=a): b = sin(ang) folded on the host
        double rho = sqrt(u1) * p.b;
        double theta = u2 * 2 * p.pi;
        sincos(theta, &s, &c);
        r.l = rho * c; r.m = rho * s;
        r.n = sqrt(1. - r.l * r.l - r.m * r.m);
    } else {                      // annulus(rin=a, rout=b, zhat=d)
        double rho = sqrt(p.a * p.a + u1 * (p.b * p.b - p.a * p.a));
        double theta = u2 * 2 * p.pi;
        sincos(theta, &s, &c);
        r.x = rho * c; r.y = rho * s; r.n = p.d;
    }
}

template <bool PHILOX>
__global__ void __launch_bounds__(PXF_BLOCK)
k_source(const RowPtrs P, int64_t num, int64_t first, unsigned long long seed,
         const double *__restrict__ u1, const double *__restrict__ u2, const SourceP p)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < num; i += nthr) {
        double a, b;
        if (PHILOX) {
            unsigned long long g = (unsigned long long)(first + i);
            u32x4 o = philox4x32_10((unsigned)g, (unsigned)(g >> 32), 0u, 0u, (unsigned)seed, (unsigned)(seed >> 32));
            a = u53(o.a, o.b);
            b = u53(o.c, o.d);
        } else {
            a = u1[i]; b = u2[i];
        }
        Ray r;
        make_ray(r, p, a, b);
        if (P.p[0]) P.p[0][i] = r.opd;
        P.p[1][i] = r.x; P.p[2][i] = r.y; P.p[3][i] = r.z;
        P.p[4][i] = r.l; P.p[5][i] = r.m; P.p[6][i] = r.n;
        P.p[7][i] = r.ux; P.p[8][i] = r.uy; P.p[9][i] = r.uz;
    }
}

// One launch for a nested assembly: segment s = rays [seg_start[s], seg_start[s+1]) drawn from the source
// `kind` with its own parameters params[4*s .. 4*s+3] = (a,b,c,d).  Counter = first + global ray index, so the
// stream is the same as per-segment launches with first = seg_start[s].
#define SRC_TILE (PXF_BLOCK * 8)
__global__ void __launch_bounds__(PXF_BLOCK)
k_source_seg(const RowPtrs P, int64_t num, int64_t first, unsigned long long seed, int kind,
             const long long *__restrict__ seg_start, const double *__restrict__ params, int nseg, double pi)
{
    for (int64_t t0 = (int64_t)blockIdx.x * SRC_TILE; t0 < num; t0 += (int64_t)gridDim.x * SRC_TILE) {
        const int64_t t1 = t0 + SRC_TILE < num ? t0 + SRC_TILE : num;
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seg_start[mid] <= t0) lo = mid; else hi = mid - 1;
        }
        int seg = lo;
        int64_t pos = t0;
        while (pos < t1 && seg < nseg) {
            int64_t send = seg_start[seg + 1];
            if (send <= pos) { seg++; continue; }
            if (send > t1) send = t1;
            SourceP p;
            p.kind = kind; p.pi = pi;
            p.a = params[4 * seg]; p.b = params[4 * seg + 1]; p.c = params[4 * seg + 2]; p.d = params[4 * seg + 3];
            for (int64_t i = pos + threadIdx.x; i < send; i += blockDim.x) {
                unsigned long long g = (unsigned long long)(first + i);
                u32x4 o = philox4x32_10((unsigned)g, (unsigned)(g >> 32), 0u, 0u, (unsigned)seed, (unsigned)(seed >> 32));
                Ray r;
                make_ray(r, p, u53(o.a, o.b), u53(o.c, o.d));
                if (P.p[0]) P.p[0][i] = r.opd;
                P.p[1][i] = r.x; P.p[2][i] = r.y; P.p[3][i] = r.z;
                P.p[4][i] = r.l; P.p[5][i] = r.m; P.p[6][i] = r.n;
                P.p[7][i] = r.ux; P.p[8][i] = r.uy; P.p[9][i] = r.uz;
            }
            pos = send;
            seg++;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The set-up sources (sources.py:173-471).  Grid sources (xslit, rectArray, fanBeam, circFan) are functions of
// the ray index alone; beam sources (convergingbeam, convergingbeam2, rectbeam, gaussianBeam) take two or three
// draws per ray, either uploaded (numpy's stream, in the reference's order of draws) or from Philox blocks
// (counter word 2 = block number).  Every expression keeps the reference's order of operations (no contraction:
// the library is built with -fmad=false).
struct Lin { double start, stop, step, delta, div; long long n; int step_zero; };

// numpy.linspace(start, stop, n)[i]: i*step + start with step = (stop-start)/(n-1), the last point pinned to stop;
// when the step underflows to zero numpy divides first ((i/div)*delta + start); n == 1 gives start.
template <class Int>
PXF_DEV double lin_at(const Lin &q, Int i)
{
    if (q.n > 1 && (long long)i == q.n - 1) return q.stop;
    const double t = (double)i;
    if (q.n <= 1) return t * q.delta + q.start;
    if (q.step_zero) return (t / q.div) * q.delta + q.start;
    return t * q.step + q.start;
}

struct GridP { int kind; Lin u, v; long long nu; double zhat; };

__global__ void __launch_bounds__(PXF_BLOCK)
k_source_grid(const RowPtrs P, int64_t num, int64_t first, const GridP p)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    // (row, column) of the meshgrid carried along the grid stride: one 64-bit division per thread, not per ray
    // (32-bit: the host refuses meshgrids with 2^31 or more rows or columns; 64-bit integer -> double conversions and
    // compares are slow instructions)
    int c = (int)((first + tid) % p.nu), r = (int)((first + tid) / p.nu);
    const int nu = (int)p.nu, dc = (int)(nthr % p.nu), dr = (int)(nthr / p.nu);
    for (int64_t i = tid; i < num; i += nthr, c += dc, r += dr) {
        if (c >= nu) { c -= nu; r++; }
        const long long g = first + i;
        double x = 0., y = 0., l = 0., m = 0., n;
        if (p.kind == PXF_SRC_XSLIT) {                 // sources.py:173-207
            x = lin_at(p.u, g);
            n = p.zhat;
        } else {
            // np.meshgrid(u, v) flattened row-major: column index runs fastest
            const double a = lin_at(p.u, c), b = lin_at(p.v, r);
            if (p.kind == PXF_SRC_RECTARRAY) {         // sources.py:210-247
                x = a; y = b; n = 1.;
            } else {
                double xa = a, ya = b;
                if (p.kind == PXF_SRC_CIRCFAN) {       // sources.py:444-471: a = ring radius, b = arm azimuth
                    double sb, cb;
                    sincos(b, &sb, &cb);
                    const double sa = sin(a);
                    xa = sa * cb; ya = sa * sb;
                }
                l = sin(xa); m = sin(ya);              // sources.py:418-442 (fanBeam) and the tail of circFan
                n = sqrt(1. - l * l - m * m);
            }
        }
        if (P.p[0]) P.p[0][i] = 0.;
        P.p[1][i] = x; P.p[2][i] = y; P.p[3][i] = 0.;
        P.p[4][i] = l; P.p[5][i] = m; P.p[6][i] = n;
        P.p[7][i] = 0.; P.p[8][i] = 0.; P.p[9][i] = 0.;
    }
}

// a..f: the source's scalars, partly folded on the host (see beam_params)
struct BeamP { int kind; double a, b, c, d, e, f, g; double pi; };

PXF_DEV void make_beam_ray(Ray &r, const BeamP &p, double d1, double d2, double d3)
{
    r.opd = 0.; r.x = 0.; r.y = 0.; r.z = 0.; r.l = 0.; r.m = 0.; r.n = 0.; r.ux = 0.; r.uy = 0.; r.uz = 0.;
    if (p.kind == PXF_SRC_RECTBEAM) {                  // sources.py:348-379: (rand-.5)*2*halfwidth
        r.x = (d1 - .5) * 2 * p.a;
        r.y = (d2 - .5) * 2 * p.b;
        r.n = 1.;
        return;
    }
    if (p.kind == PXF_SRC_GAUSSIAN) {                  // sources.py:381-416: randn*sin(ang)/sqrt(2)
        r.l = d1 * p.a / p.b;
        r.m = d2 * p.a / p.b;
        r.n = sqrt(1. - r.l * r.l - r.m * r.m);
        return;
    }
    double rho, theta, s, c;
    if (p.kind == PXF_SRC_CONVERGING) {                // sources.py:250-296: a=zset b=rin^2 c=rout^2-rin^2 d=tmin e=tmax-tmin
        rho = sqrt(p.b + d1 * p.c);
        theta = p.d + d2 * p.e;
        sincos(theta, &s, &c);
        r.x = rho * c; r.y = rho * s;
    } else {                                           // sources.py:299-345: b=xmin c=xmax-xmin d=ymin e=ymax-ymin
        r.x = p.b + d1 * p.c;
        r.y = p.d + d2 * p.e;
        rho = sqrt(r.x * r.x + r.y * r.y);
        theta = atan2(r.y, r.x);
        sincos(theta, &s, &c);
    }
    r.z = p.a;
    double ls = p.f * tan((d3 - .5) * p.pi);           // f = lscat
    ls = ls / p.g * p.pi / 180.;                       // g = 60**2
    r.n = -cos(atan(rho / p.a) + ls);
    const double t = sqrt(1. - r.n * r.n);
    r.l = -t * c;
    r.m = -t * s;
}

template <bool PHILOX>
__global__ void __launch_bounds__(PXF_BLOCK)
k_source_beam(const RowPtrs P, int64_t num, int64_t first, unsigned long long seed,
              const double *__restrict__ d1, const double *__restrict__ d2, const double *__restrict__ d3,
              const BeamP p)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const bool three = p.kind == PXF_SRC_CONVERGING || p.kind == PXF_SRC_CONVERGING2;
    for (int64_t i = tid; i < num; i += nthr) {
        double a, b, c = 0.;
        if (PHILOX) {
            const unsigned long long g = (unsigned long long)(first + i);
            u32x4 o = philox4x32_10((unsigned)g, (unsigned)(g >> 32), 0u, 0u, (unsigned)seed, (unsigned)(seed >> 32));
            a = u53(o.a, o.b);
            b = u53(o.c, o.d);
            if (three) {
                o = philox4x32_10((unsigned)g, (unsigned)(g >> 32), 1u, 0u, (unsigned)seed, (unsigned)(seed >> 32));
                c = u53(o.a, o.b);
            }
            if (p.kind == PXF_SRC_GAUSSIAN) {
                // Box-Muller on (1-a, b): a is in [0,1), so the logarithm's argument is in (0,1]
                const double rad = sqrt(-2. * log(1. - a));
                double sn, cs;
                sincospi(2. * b, &sn, &cs);
                a = rad * cs; b = rad * sn;
            }
        } else {
            a = d1[i]; b = d2[i];
            if (three) c = d3[i];
        }
        Ray r;
        make_beam_ray(r, p, a, b, c);
        if (P.p[0]) P.p[0][i] = r.opd;
        P.p[1][i] = r.x; P.p[2][i] = r.y; P.p[3][i] = r.z;
        P.p[4][i] = r.l; P.p[5][i] = r.m; P.p[6][i] = r.n;
        P.p[7][i] = r.ux; P.p[8][i] = r.uy; P.p[9][i] = r.uz;
    }
}

static Lin make_lin(double start, double stop, long long n)
{
    // numpy/_core/function_base.py linspace: div = n-1; delta = stop-start; step = delta/div
    Lin q;
    q.start = start; q.stop = stop; q.n = n;
    q.div = (double)(n - 1);
    q.delta = stop - start;
    q.step = n > 1 ? q.delta / q.div : 0.;
    q.step_zero = n > 1 && q.step == 0.;
    return q;
}

static int rows_ok(RowPtrs &P, double *const rays[10], const char *who)
{
    if (!rays) { set_error("%s: null bundle", who); return PXF_ERR_INVALID; }
    for (int k = 0; k < 10; k++) {
        P.p[k] = rays[k];
        if (k > 0 && !rays[k]) { set_error("%s: null row", who); return PXF_ERR_INVALID; }
    }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    return PXF_OK;
}

static int beam_params(BeamP &p, int kind, const double *par)
{
    if (!par) return -1;
    p.kind = kind; p.pi = 3.141592653589793;
    p.a = p.b = p.c = p.d = p.e = p.f = 0.; p.g = 3600.;       // 60**2
    switch (kind) {
    case PXF_SRC_CONVERGING:    // (zset, rin, rout, tmin, tmax, lscat)
        p.a = par[0]; p.b = par[1] * par[1]; p.c = par[2] * par[2] - p.b; p.d = par[3]; p.e = par[4] - par[3]; p.f = par[5];
        return 0;
    case PXF_SRC_CONVERGING2:   // (zset, xmin, xmax, ymin, ymax, lscat)
        p.a = par[0]; p.b = par[1]; p.c = par[2] - par[1]; p.d = par[3]; p.e = par[4] - par[3]; p.f = par[5];
        return 0;
    case PXF_SRC_RECTBEAM:      // (xhalfwidth, yhalfwidth)
        p.a = par[0]; p.b = par[1];
        return 0;
    case PXF_SRC_GAUSSIAN:      // (ang): np.sin(ang), np.sqrt(2)
        p.a = sin(par[0]); p.b = sqrt(2.);
        return 0;
    }
    return -1;
}

static int source_launch(int kind, double *const rays[10], int64_t num, int64_t first, uint64_t seed,
                         const double *u1, const double *u2, bool philox, double a, double b, double c, double d,
                         cudaStream_t s)
{
    if (num < 0 || !rays || kind < 0 || kind > 3) { set_error("pxf_source: bad argument"); return PXF_ERR_INVALID; }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    RowPtrs P;
    for (int k = 0; k < 10; k++) {
        P.p[k] = rays[k];
        if (k > 0 && !rays[k]) { set_error("pxf_source: null row"); return PXF_ERR_INVALID; }
    }
    if (!philox && (!u1 || !u2)) { set_error("pxf_source: null uniforms"); return PXF_ERR_INVALID; }
    if (num == 0) return PXF_OK;
    SourceP p;
    p.kind = kind; p.a = a; p.b = b; p.c = c; p.d = d;
    p.pi = 3.141592653589793;   // np.pi
    if (kind == 2) p.b = sin(a);
    int grid = grid_for(num, PXF_BLOCK, 8);
    if (philox) k_source<true><<<grid, PXF_BLOCK, 0, s>>>(P, num, first, seed, u1, u2, p);
    else k_source<false><<<grid, PXF_BLOCK, 0, s>>>(P, num, first, seed, u1, u2, p);
    count_launch();
    return check_launch("k_source");
}

}  // namespace pxf

using namespace pxf;

extern "C" {

int pxf_source(int32_t kind, double *const rays[10], int64_t num, int64_t first, uint64_t seed,
               double a, double b, double c, double d, pxf_stream_t stream)
{
    return source_launch(kind, rays, num, first, seed, nullptr, nullptr, true, a, b, c, d,
                         reinterpret_cast<cudaStream_t>(stream));
}

int pxf_source_segmented(int32_t kind, double *const rays[10], int64_t num, int64_t first, uint64_t seed,
                         int32_t nseg, const int64_t *seg_start_dev, const double *params_dev, pxf_stream_t stream)
{
    if (num < 0 || !rays || kind < 0 || kind > 3 || kind == 2 || nseg < 1 || !seg_start_dev || !params_dev) {
        set_error("pxf_source_segmented: bad argument (pointsource is not supported: its sin(ang) is folded per call)");
        return PXF_ERR_INVALID;
    }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    RowPtrs P;
    for (int k = 0; k < 10; k++) {
        P.p[k] = rays[k];
        if (k > 0 && !rays[k]) { set_error("pxf_source_segmented: null row"); return PXF_ERR_INVALID; }
    }
    if (num == 0) return PXF_OK;
    const int grid = grid_for(num, SRC_TILE, 8);
    k_source_seg<<<grid, PXF_BLOCK, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        P, num, first, seed, kind, reinterpret_cast<const long long *>(seg_start_dev), params_dev, nseg, 3.141592653589793);
    count_launch();
    return check_launch("k_source_seg");
}

int pxf_source_from_uniform(int32_t kind, double *const rays[10], int64_t num, const double *u1,
                            const double *u2, double a, double b, double c, double d, pxf_stream_t stream)
{
    return source_launch(kind, rays, num, 0, 0, u1, u2, false, a, b, c, d, reinterpret_cast<cudaStream_t>(stream));
}

int pxf_source_grid(int32_t kind, double *const rays[10], int64_t num, int64_t first, int64_t n1, int64_t n2,
                    double a, double b, double c, pxf_stream_t stream)
{
    RowPtrs P;
    int rc = rows_ok(P, rays, "pxf_source_grid");
    if (rc != PXF_OK) return rc;
    GridP p;
    p.kind = kind; p.zhat = c; p.nu = n1 > 0 ? n1 : 1;
    int64_t total;
    switch (kind) {
    case PXF_SRC_XSLIT:     p.u = make_lin(a, b, n1); p.v = p.u; total = n1; break;
    case PXF_SRC_RECTARRAY:
    case PXF_SRC_FANBEAM:   p.u = make_lin(-a, a, n1); p.v = make_lin(-b, b, n1); total = n1 * n1; break;
    case PXF_SRC_CIRCFAN: {
        // rad = linspace(0, halfang, rings); az = linspace(0, 2*np.pi, arms+1)[:-1] (the pinned endpoint is dropped)
        p.u = make_lin(0., a, n1);
        p.v = make_lin(0., 2 * 3.141592653589793, n2 + 1);   // arm r < arms never is the pinned last point
        total = n1 * n2;
        break;
    }
    default: set_error("pxf_source_grid: bad kind %d", kind); return PXF_ERR_INVALID;
    }
    if (kind != PXF_SRC_XSLIT && (n1 >= (1ll << 31) - 1 || n2 >= (1ll << 31) - 2)) {
        set_error("pxf_source_grid: grid too large");
        return PXF_ERR_INVALID;
    }
    if (n1 < 0 || n2 < 0 || num < 0 || first < 0 || first + num > total) {
        set_error("pxf_source_grid: rays [%lld, %lld) are not inside the %lld-ray source", (long long)first,
                  (long long)(first + num), (long long)total);
        return PXF_ERR_INVALID;
    }
    if (num == 0) return PXF_OK;
    k_source_grid<<<grid_for(num, PXF_BLOCK, 8), PXF_BLOCK, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, num, first, p);
    count_launch();
    return check_launch("k_source_grid");
}

int pxf_source_beam(int32_t kind, double *const rays[10], int64_t num, int64_t first, uint64_t seed,
                    const double *par, pxf_stream_t stream)
{
    RowPtrs P;
    int rc = rows_ok(P, rays, "pxf_source_beam");
    if (rc != PXF_OK) return rc;
    BeamP p;
    if (num < 0 || beam_params(p, kind, par) < 0) { set_error("pxf_source_beam: bad argument"); return PXF_ERR_INVALID; }
    if (num == 0) return PXF_OK;
    k_source_beam<true><<<grid_for(num, PXF_BLOCK, 8), PXF_BLOCK, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        P, num, first, seed, nullptr, nullptr, nullptr, p);
    count_launch();
    return check_launch("k_source_beam");
}

int pxf_source_beam_from_draws(int32_t kind, double *const rays[10], int64_t num, const double *d1, const double *d2,
                               const double *d3, const double *par, pxf_stream_t stream)
{
    RowPtrs P;
    int rc = rows_ok(P, rays, "pxf_source_beam_from_draws");
    if (rc != PXF_OK) return rc;
    BeamP p;
    if (num < 0 || beam_params(p, kind, par) < 0) { set_error("pxf_source_beam_from_draws: bad argument"); return PXF_ERR_INVALID; }
    const bool three = kind == PXF_SRC_CONVERGING || kind == PXF_SRC_CONVERGING2;
    if (num > 0 && (!d1 || !d2 || (three && !d3))) { set_error("pxf_source_beam_from_draws: null draws"); return PXF_ERR_INVALID; }
    if (num == 0) return PXF_OK;
    k_source_beam<false><<<grid_for(num, PXF_BLOCK, 8), PXF_BLOCK, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        P, num, 0, 0, d1, d2, d3, p);
    count_launch();
    return check_launch("k_source_beam");
}

}  // extern "C"
