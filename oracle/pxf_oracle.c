/*
 * pxf_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the PyXFocus f2py Fortran hot path.  It exists only
 * to check the CUDA engine (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline / --impl reference legs).  The product path never links,
 * imports or executes anything in oracle/.
 *
 * PARITY STATUS: pinned against the Fortran SOURCE TEXT, not against a compiled
 * build of it.  The reference ships no tests, golden vectors or fixtures, and no
 * Fortran compiler exists in the build container, so there is no oracle/_ref.
 * Instead oracle/f95run.py EXECUTES the reference's .f95 files: a translator for
 * the Fortran subset they use with gfortran's arithmetic rules (REAL*4 literals
 * and implicit typing, kind promotion, truncating integer division, __powidf2
 * powers, glibc libm, by-reference arguments).  tests/golden/make_f95_golden.py
 * ran every subroutine of transformationsf / surfacesf / woltsurf / zernsurf /
 * reconstruct (57 cases, on- and off-axis) through it and this file reproduces
 * every output bit for bit (tests/test_f95_source.py; fixture
 * tests/golden/f95_source.npz).  What that does NOT cover: a compiler-specific
 * deviation of a real gfortran build from those rules (none is known for
 * baseline x86-64 without -ffast-math).  Further pins: (a) physics known-answer
 * checks (tests/test_oracle_kat.py) and (b) golden vectors produced by the
 * reference's own *Python* layer driving this file (tests/golden/) -- every
 * one of which re-derives identically with the Fortran text itself in the
 * f2py slots (oracle/f95mods.py; make_golden.py --fortran-source).
 *
 * Conventions restated from the Fortran (all citations relative to the
 * reference tree):
 *   - Fortran evaluates a*b*c left to right; x**2 is x*x; x**3 is (x*x)*x;
 *     x**4 is (x*x)*(x*x); real exponents go through pow().
 *   - Default-real literals are REAL*4: 1.e-8, 1.e-10, acos(-1.), sqrt(0.5)
 *     are single precision values promoted to double.
 *   - Build with -ffp-contract=off: gfortran for baseline x86-64 emits no FMA.
 *   - OpenMP races in the reference (woltsurf.f95:404 'c', transformationsf
 *     .f95:92 'dt') are resolved to single-thread semantics: every loop-local
 *     below is private.
 *   - The reference's Newton loops have no iteration cap (a non-converging
 *     ray hangs the process).  PXF_NEWTON_CAP bounds them here and in the
 *     CUDA engine identically; it is never reached by converging rays.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* -DPXF_ORACLE_CR_LIBM (liboracle_cr.so): every libm call goes through binary128 (libquadmath) and is rounded
 * to double ONCE -- correctly rounded results (up to the astronomically rare 113-bit ties).  The reference's own
 * libm is unpinned (whatever glibc gfortran linked): glibc's double routines are within 1 ulp but
 * not always the nearest double, and WHICH double they return differs between glibc versions.  For the chaotic
 * Wolter-Schwarzschild rays beyond the graze angle that last bit decides discrete outcomes; the correctly rounded
 * value is the one answer every conforming libm is an approximation of, so this variant is the canonical form the
 * GPU's long-trip re-trace (which evaluates the same functions correctly rounded) is held to bit for bit. */
#ifdef PXF_ORACLE_CR_LIBM
#include <quadmath.h>
static inline double cr_sin(double x) { return (double)sinq((__float128)x); }
static inline double cr_cos(double x) { return (double)cosq((__float128)x); }
static inline double cr_tan(double x) { return (double)tanq((__float128)x); }
static inline double cr_asin(double x) { return (double)asinq((__float128)x); }
static inline double cr_acos(double x) { return (double)acosq((__float128)x); }
static inline double cr_atan(double x) { return (double)atanq((__float128)x); }
static inline double cr_atan2(double y, double x) { return (double)atan2q((__float128)y, (__float128)x); }
static inline double cr_pow(double x, double y) { return (double)powq((__float128)x, (__float128)y); }
#define sin cr_sin
#define cos cr_cos
#define tan cr_tan
#define asin cr_asin
#define acos cr_acos
#define atan cr_atan
#define atan2 cr_atan2
#define pow cr_pow
#endif

#define PXF_NEWTON_CAP 1000

/* REAL*4 literals promoted to REAL*8 */
#define TOL_1EM8  ((double)1.e-8f)
#define TOL_1EM10 ((double)1.e-10f)
#define TOL_1EM7  ((double)1.e-7f)

static inline double sq(double a) { return a * a; }
static inline double cube(double a) { return (a * a) * a; }
static inline double pow4(double a) { double t = a * a; return t * t; }

/* REAL*4 acos(-1.) promoted to double: 3.1415927410125732 */
static inline double pi32(void) { return (double)acosf(-1.0f); }

double pxfo_legendre(double x, int n);
double pxfo_legendrep(double x, int n);

/* ------------------------------------------------------------------ */
/* transformationsf.f95                                               */
/* ------------------------------------------------------------------ */

/* transformationsf.f95:3-28 */
static inline void rotatevector(double *x, double *y, double *z, double theta, int axis)
{
    double o1, o2, o3;
    if (axis == 1) {
        o1 = *x;
        o2 = cos(theta) * (*y) - sin(theta) * (*z);
        o3 = sin(theta) * (*y) + cos(theta) * (*z);
    } else if (axis == 2) {
        o1 = cos(theta) * (*x) + sin(theta) * (*z);
        o2 = *y;
        o3 = -sin(theta) * (*x) + cos(theta) * (*z);
    } else {
        o1 = cos(theta) * (*x) - sin(theta) * (*y);
        o2 = sin(theta) * (*x) + cos(theta) * (*y);
        o3 = *z;
    }
    *x = o1; *y = o2; *z = o3;
}

/* transformationsf.f95:32-55 (normalises the axis in place) */
static inline void rotateaxis(double *x, double *y, double *z, double theta,
                              double *ux, double *uy, double *uz)
{
    double mag = sqrt(sq(*ux) + sq(*uy) + sq(*uz));
    *ux = *ux / mag; *uy = *uy / mag; *uz = *uz / mag;
    double c = cos(theta), s = sin(theta);
    double a = *ux, b = *uy, g = *uz;
    double o1 = (c + sq(a) * (1 - c)) * (*x) + (a * b * (1 - c) - g * s) * (*y) + (a * g * (1 - c) + b * s) * (*z);
    double o2 = (b * a * (1 - c) + g * s) * (*x) + (c + sq(b) * (1 - c)) * (*y) + (b * g * (1 - c) - a * s) * (*z);
    double o3 = (g * a * (1 - c) - b * s) * (*x) + (g * b * (1 - c) + a * s) * (*y) + (c + sq(g) * (1 - c)) * (*z);
    *x = o1; *y = o2; *z = o3;
}

/* transformationsf.f95:60-79 */
void pxfo_reflect(double *l, double *m, double *n, double *ux, double *uy, double *uz, int64_t num)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double dot = ux[i] * l[i] + uy[i] * m[i] + uz[i] * n[i];
        l[i] = l[i] - 2 * dot * ux[i];
        m[i] = m[i] - 2 * dot * uy[i];
        n[i] = n[i] - 2 * dot * uz[i];
    }
}

/* transformationsf.f95:82-130 */
void pxfo_refract(double *l, double *m, double *n, double *ux, double *uy, double *uz,
                  int64_t num, double n1, double n2)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double dot = l[i] * ux[i] + m[i] * uy[i] + n[i] * uz[i];
        if (dot < 0) {
            ux[i] = -ux[i]; uy[i] = -uy[i]; uz[i] = -uz[i];
            dot = -dot;
        }
        if (dot == 1) continue;
        double t1 = acos(dot);
        double t2 = asin((n1 / n2) * sin(t1));
        double cx = uy[i] * n[i] - m[i] * uz[i];
        double cy = l[i] * uz[i] - ux[i] * n[i];
        double cz = ux[i] * m[i] - l[i] * uy[i];
        double dt = t2 - t1;
        rotateaxis(&l[i], &m[i], &n[i], dt, &cx, &cy, &cz);
        double alpha = sqrt(sq(l[i]) + sq(m[i]) + sq(n[i]));
        l[i] = l[i] / alpha;
        m[i] = m[i] / alpha;
        n[i] = n[i] / alpha;
    }
}

/* transformationsf.f95:134-163 */
void pxfo_transform(double *x, double *y, double *z, double *l, double *m, double *n,
                    double *ux, double *uy, double *uz, int64_t num,
                    double tx, double ty, double tz, double rx, double ry, double rz)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        x[i] = x[i] + tx;
        y[i] = y[i] + ty;
        z[i] = z[i] + tz;
        rotatevector(&x[i], &y[i], &z[i], rx, 1);
        rotatevector(&l[i], &m[i], &n[i], rx, 1);
        rotatevector(&ux[i], &uy[i], &uz[i], rx, 1);
        rotatevector(&x[i], &y[i], &z[i], ry, 2);
        rotatevector(&l[i], &m[i], &n[i], ry, 2);
        rotatevector(&ux[i], &uy[i], &uz[i], ry, 2);
        rotatevector(&x[i], &y[i], &z[i], rz, 3);
        rotatevector(&l[i], &m[i], &n[i], rz, 3);
        rotatevector(&ux[i], &uy[i], &uz[i], rz, 3);
    }
}

/* transformationsf.f95:168-201 */
void pxfo_itransform(double *x, double *y, double *z, double *l, double *m, double *n,
                     double *ux, double *uy, double *uz, int64_t num,
                     double tx, double ty, double tz, double rx, double ry, double rz)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double tmp = -rz;
        rotatevector(&x[i], &y[i], &z[i], tmp, 3);
        rotatevector(&l[i], &m[i], &n[i], tmp, 3);
        rotatevector(&ux[i], &uy[i], &uz[i], tmp, 3);
        tmp = -ry;
        rotatevector(&x[i], &y[i], &z[i], tmp, 2);
        rotatevector(&l[i], &m[i], &n[i], tmp, 2);
        rotatevector(&ux[i], &uy[i], &uz[i], tmp, 2);
        tmp = -rx;
        rotatevector(&x[i], &y[i], &z[i], tmp, 1);
        rotatevector(&l[i], &m[i], &n[i], tmp, 1);
        rotatevector(&ux[i], &uy[i], &uz[i], tmp, 1);
        x[i] = x[i] - tx;
        y[i] = y[i] - ty;
        z[i] = z[i] - tz;
    }
}

/* transformationsf.f95:205-238 -- SERIAL in the reference (no omp directive) */
void pxfo_radgrat(const double *x, const double *y, double *l, double *m, double *n,
                  double wave, int64_t num, double dpermm, double order)
{
    double pi = pi32();
    for (int64_t i = 0; i < num; i++) {
        double sn = n[i] / fabs(n[i]);
        double d = dpermm * sqrt(sq(y[i]) + sq(x[i]));
        double yaw = -pi / 2 - atan2(x[i], y[i]);
        l[i] = l[i] + sin(yaw) * order * wave / d;
        m[i] = m[i] - cos(yaw) * order * wave / d;
        n[i] = sn * sqrt(1. - sq(l[i]) - sq(m[i]));
    }
}

/* transformationsf.f95:242-272 -- SERIAL; sign taken from y (:258) */
void pxfo_radgratw(const double *x, const double *y, double *l, double *m, double *n,
                   const double *wave, int64_t num, double dpermm, double order)
{
    double pi = pi32();
    for (int64_t i = 0; i < num; i++) {
        double d = dpermm * sqrt(sq(y[i]) + sq(x[i]));
        double sn = y[i] / fabs(y[i]);
        double yaw = -pi / 2 - atan2(x[i], y[i]);
        l[i] = l[i] + sin(yaw) * order * wave[i] / d;
        m[i] = m[i] - cos(yaw) * order * wave[i] / d;
        n[i] = sn * sqrt(1. - sq(l[i]) - sq(m[i]));
    }
}

/* transformationsf.f95:277-305; 'sn' is implicitly REAL*4 (:292), value is +-1 or NaN */
void pxfo_grat(const double *x, const double *y, double *l, double *m, double *n,
               int64_t num, double d, const double *order, const double *wave)
{
    (void)x; (void)y;
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        float sn = (float)(n[i] / fabs(n[i]));
        l[i] = l[i] - order[i] * wave[i] / d;
        n[i] = (double)sn * sqrt(1 - sq(l[i]) - sq(m[i]));
        if ((sq(l[i]) + sq(m[i])) > 1) {
            l[i] = 0.; m[i] = 0.; n[i] = 0.;
        }
    }
}

/* ------------------------------------------------------------------ */
/* surfacesf.f95                                                      */
/* ------------------------------------------------------------------ */

/* surfacesf.f95:4-29 -- no 'implicit none': delta is REAL*4 */
void pxfo_flat(double *x, double *y, double *z, const double *l, const double *m, const double *n,
               double *ux, double *uy, double *uz, int64_t num)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        float delta = (float)(-z[i] / n[i]);
        z[i] = 0.;
        x[i] = x[i] + (double)delta * l[i];
        y[i] = y[i] + (double)delta * m[i];
        ux[i] = 0.; uy[i] = 0.; uz[i] = 1.;
    }
}

/* surfacesf.f95:32-53 */
void pxfo_flatopd(double *x, double *y, double *z, const double *l, const double *m, const double *n,
                  double *ux, double *uy, double *uz, double *opd, int64_t num, double nr)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        float delta = (float)(-z[i] / n[i]);
        z[i] = 0.;
        x[i] = x[i] + (double)delta * l[i];
        y[i] = y[i] + (double)delta * m[i];
        ux[i] = 0.; uy[i] = 0.; uz[i] = 1.;
        opd[i] = opd[i] + (double)delta * nr;
    }
}

/* surfacesf.f95:302-360 (opd==NULL) and :366-420 (opd!=NULL; note the
 * different association of 'c' at :386) */
static void conic_impl(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                       double *ux, double *uy, double *uz, int64_t num, double R, double K, double nr)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double s = 0., denom, b, c, disc, s1, s2;
        if (K == -1 && fabs(n[i]) == 1.) {
            s = (sq(x[i]) + sq(y[i]) - 2 * R * z[i]) / (2 * R * n[i]);
        } else {
            denom = sq(l[i]) + sq(m[i]) + (K + 1) * sq(n[i]);
            b = x[i] * l[i] + y[i] * m[i] + ((K + 1) * z[i] - R) * n[i];
            b = b / denom;
            if (opd)
                c = sq(x[i]) + sq(y[i]) + (K + 1) * sq(z[i]) - 2 * R * z[i];
            else
                c = sq(x[i]) + sq(y[i]) - 2 * R * z[i] + (K + 1) * sq(z[i]);
            c = c / denom;
            disc = sq(b) - c;
            if (disc >= 0.) {
                s1 = -b + sqrt(disc);
                s2 = -b - sqrt(disc);
                if (fabs(s1) <= fabs(s2)) s = s1; else s = s2;
            }
        }
        if (s == 0.) {
            l[i] = 0.; m[i] = 0.; n[i] = 0.;
        } else {
            x[i] = x[i] + l[i] * s;
            y[i] = y[i] + m[i] * s;
            z[i] = z[i] + n[i] * s;
            if (opd) opd[i] = opd[i] + s * nr;
            denom = sqrt(sq(R) - K * (sq(x[i]) + sq(y[i])));
            ux[i] = -x[i] / denom;
            uy[i] = -y[i] / denom;
            uz[i] = -R / fabs(R) * sqrt(sq(R) - (K + 1) * (sq(x[i]) + sq(y[i])));
            uz[i] = -uz[i] / denom;
        }
    }
}

void pxfo_conic(double *x, double *y, double *z, double *l, double *m, double *n,
                double *ux, double *uy, double *uz, int64_t num, double R, double K)
{
    conic_impl(NULL, x, y, z, l, m, n, ux, uy, uz, num, R, K, 0.);
}

void pxfo_conicopd(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                   double *ux, double *uy, double *uz, int64_t num, double R, double K, double nr)
{
    conic_impl(opd, x, y, z, l, m, n, ux, uy, uz, num, R, K, nr);
}

/* ------------------------------------------------------------------ */
/* woltsurf.f95                                                       */
/* ------------------------------------------------------------------ */

/* Van Speybroeck parameters, woltsurf.f95:18-23 */
static void vanspeybroeck(double r0, double z0, double psi, double *p, double *d, double *e)
{
    double alpha = .25 * atan(r0 / z0);
    double thetah = 2 * (1 + 2 * psi) / (1 + psi) * alpha;
    double thetap = 2 * psi / (1 + psi) * alpha;
    *p = z0 * tan(4 * alpha) * tan(thetap);
    *d = z0 * tan(4 * alpha) * tan(4 * alpha - thetah);
    *e = cos(4 * alpha) * (1 + tan(4 * alpha) * tan(thetah));
}

/* woltsurf.f95:7-54 (opd==NULL, tol 1.e-8) and :60-108 (opd, tol 1.e-10) */
static void wolterprimary_impl(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                               double *ux, double *uy, double *uz, int64_t num,
                               double r0, double z0, double psi, double nr)
{
    double p, d, e;
    vanspeybroeck(r0, z0, psi, &p, &d, &e);
    const double Fz = 2 * p;
    const double tol = opd ? TOL_1EM10 : TOL_1EM8;
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fp;
        int it = 0;
        while (fabs(delt) > tol && it++ < PXF_NEWTON_CAP) {
            F = 2 * p * z[i] + sq(p) + 4 * sq(e) * p * d / (sq(e) - 1) - sq(x[i]) - sq(y[i]);
            Fx = -2. * x[i];
            Fy = -2. * y[i];
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
            if (opd) opd[i] = opd[i] + nr * delt;
        }
        Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
        ux[i] = Fx / Fp;
        uy[i] = Fy / Fp;
        uz[i] = Fz / Fp;
    }
}

void pxfo_wolterprimary(double *x, double *y, double *z, double *l, double *m, double *n,
                        double *ux, double *uy, double *uz, int64_t num, double r0, double z0, double psi)
{
    wolterprimary_impl(NULL, x, y, z, l, m, n, ux, uy, uz, num, r0, z0, psi, 0.);
}

void pxfo_wolterprimaryopd(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                           double *ux, double *uy, double *uz, int64_t num,
                           double r0, double z0, double psi, double nr)
{
    wolterprimary_impl(opd, x, y, z, l, m, n, ux, uy, uz, num, r0, z0, psi, nr);
}

/* woltsurf.f95:114-161 */
void pxfo_woltersecondary(double *x, double *y, double *z, double *l, double *m, double *n,
                          double *ux, double *uy, double *uz, int64_t num, double r0, double z0, double psi)
{
    double p, d, e;
    vanspeybroeck(r0, z0, psi, &p, &d, &e);
    (void)p;
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp;
        int it = 0;
        while (fabs(delt) > TOL_1EM8 && it++ < PXF_NEWTON_CAP) {
            F = sq(e) * sq(d + z[i]) - sq(z[i]) - sq(x[i]) - sq(y[i]);
            Fx = -2. * x[i];
            Fy = -2. * y[i];
            Fz = 2 * sq(e) * (d + z[i]) - 2 * z[i];
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
        }
        Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
        ux[i] = Fx / Fp;
        uy[i] = Fy / Fp;
        uz[i] = Fz / Fp;
    }
}

/* woltsurf.f95:167-215; 2*acos(-1.) is REAL*4 arithmetic (:190,:194) */
void pxfo_woltersine(double *x, double *y, double *z, double *l, double *m, double *n,
                     double *ux, double *uy, double *uz, int64_t num,
                     double r0, double z0, double amp, double freq)
{
    double alpha = .25 * atan(r0 / z0);
    double thetah = 3. * alpha;
    double thetap = alpha;
    double p = z0 * tan(4 * alpha) * tan(thetap);
    double d = z0 * tan(4 * alpha) * tan(4 * alpha - thetah);
    double e = cos(4 * alpha) * (1 + tan(4 * alpha) * tan(thetah));
    const double twopi = (double)(2 * acosf(-1.0f));   /* 2*acos(-1.) in REAL*4 */
    const double pi = pi32();
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp, rad;
        int it = 0;
        while (fabs(delt) > TOL_1EM10 && it++ < PXF_NEWTON_CAP) {
            rad = sqrt(sq(x[i]) + sq(y[i])) + amp * sin(twopi * freq * z[i]);
            F = 2 * p * z[i] + sq(p) + 4 * sq(e) * p * d / (sq(e) - 1) - sq(rad);
            Fx = -2. * x[i];
            Fy = -2. * y[i];
            /* 2.*p - 2*rad*amp*2*acos(-1.)*freq*cos(2*acos(-1.)*freq*z) */
            Fz = 2. * p - 2 * rad * amp * 2 * pi * freq * cos(twopi * freq * z[i]);
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
        }
        Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
        ux[i] = Fx / Fp;
        uy[i] = Fy / Fp;
        uz[i] = Fz / Fp;
    }
}

/* woltsurf.f95:387-476.  'c' made private (single-thread semantics). */
void pxfo_wsprimary(double *x, double *y, double *z, double *l, double *m, double *n,
                    double *ux, double *uy, double *uz, int64_t num,
                    double alpha, double z0, double psi)
{
    const double betas = 4 * alpha;
    const double ff = z0 / cos(betas);
    const double g = ff / psi;
    const double k = sq(tan(betas / 2));
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp, Fb, kterm, beta, dbdx, dbdy, r;
        int flag, c = 0;
        double xi = x[i], yi = y[i], zi = z[i];
        while (fabs(delt) > TOL_1EM8) {
            beta = asin(sqrt(sq(x[i]) + sq(y[i])) / ff);
            flag = 0;
            if (beta <= betas) {
                beta = betas;
                flag = 1;
                kterm = 0.;
            } else {
                kterm = (1 / k) * sq(tan(beta / 2)) - 1;
            }
            F = -z[i] - ff * sq(sin(betas / 2)) +
                sq(ff) * sq(sin(beta)) / (4 * ff * sq(sin(betas / 2))) +
                g * pow4(cos(beta / 2)) * pow(kterm, 1 - k);
            Fb = sq(ff) * sin(beta) * cos(beta) / (2 * ff * sq(sin(betas / 2))) -
                 2 * g * cube(cos(beta / 2)) * sin(beta / 2) * pow(kterm, 1 - k) +
                 g * (1 - k) * cos(beta / 2) * sin(beta / 2) * pow(kterm, -k) * (1 / k);
            Fz = -1.;
            if (flag == 1) {
                r = sqrt(sq(x[i]) + sq(y[i]));
                Fb = sq(ff) * sin(betas) * cos(betas) / (2 * ff * sq(sin(betas / 2))) +
                     g * (1 - k) * cos(betas / 2) * sin(betas / 2) * (1 / k);
                F = F + (r - ff * sin(betas)) * z[i] / (sq(r) + sq(z[i])) * Fb;
                Fz = Fz + (r - ff * sin(betas)) * (sq(r) - sq(z[i])) / sq(sq(r) + sq(z[i])) * Fb;
            }
            dbdx = x[i] / sqrt(1 - (sq(x[i]) + sq(y[i])) / sq(ff)) / ff / sqrt(sq(x[i]) + sq(y[i]));
            dbdy = y[i] / sqrt(1 - (sq(x[i]) + sq(y[i])) / sq(ff)) / ff / sqrt(sq(x[i]) + sq(y[i]));
            Fx = Fb * dbdx;
            Fy = Fb * dbdy;
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
            if (c > 25 || isnan(delt)) {
                delt = 0.;
                x[i] = xi; y[i] = yi; z[i] = zi;
                c = 1000;
            }
            c = c + 1;
        }
        if (c < 26) {
            Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
            ux[i] = -Fx / Fp;
            uy[i] = -Fy / Fp;
            uz[i] = -Fz / Fp;
        }
    }
}

/* woltsurf.f95:484-588 */
void pxfo_wssecondary(double *x, double *y, double *z, double *l, double *m, double *n,
                      double *ux, double *uy, double *uz, int64_t num,
                      double alpha, double z0, double psi)
{
    const double betas = 4 * alpha;
    const double ff = z0 / cos(betas);
    const double g = ff / psi;
    const double k = sq(tan(betas / 2));
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp, Fb, kterm, beta, dbdx, dbdy, dbdz;
        double a, dadbs, dbdzs, gam, dadb;
        int flag, c = 0;
        double xi = x[i], yi = y[i], zi = z[i];
        while (fabs(delt) > TOL_1EM8) {
            beta = atan2(sqrt(sq(x[i]) + sq(y[i])), z[i]);
            flag = 0;
            if (beta <= betas) {
                beta = betas;
                kterm = 0;
                a = 1 / ff;
                flag = 1;
            } else {
                kterm = (1 / k) * sq(tan(beta / 2)) - 1;
                a = (1 - cos(beta)) / (1 - cos(betas)) / ff +
                    (1 + cos(beta)) / (2 * g) * pow(kterm, 1 + k);
            }
            F = -z[i] + cos(beta) / a;
            if (flag == 1) {
                Fb = 0.;
                dadbs = sin(betas) / ff / (1 - cos(betas)) +
                        (k + 1) * (cos(betas) + 1) * tan(betas / 2) / sq(cos(betas / 2)) / 2 / g / k;
                dbdzs = -sq(sin(betas)) / sqrt(sq(x[i]) + sq(y[i]));
                gam = (-ff * sin(betas) - sq(ff) * cos(betas) * dadbs) * dbdzs;
                F = F + gam * (z[i] - sqrt(sq(x[i]) + sq(y[i])) / tan(betas));
                Fx = -2. / tan(betas) * x[i] / sqrt(sq(x[i]) + sq(y[i]));
                Fy = -2. / tan(betas) * y[i] / sqrt(sq(x[i]) + sq(y[i]));
                Fz = gam - 1.;
            } else {
                dadb = sin(beta) / ff / (1 - cos(betas)) -
                       sin(beta) / (2 * g) * pow(kterm, 1 + k) +
                       (k + 1) * (cos(beta) + 1) * tan(beta / 2) * pow(kterm, k) / 2 / g / k / sq(cos(beta / 2));
                Fb = -sin(beta) / a - cos(beta) / sq(a) * dadb;
                dbdx = x[i] * z[i] / (sq(x[i]) + sq(y[i]) + sq(z[i])) / sqrt(sq(x[i]) + sq(y[i]));
                dbdy = y[i] * z[i] / (sq(x[i]) + sq(y[i]) + sq(z[i])) / sqrt(sq(x[i]) + sq(y[i]));
                dbdz = -sqrt(sq(x[i]) + sq(y[i])) / (sq(x[i]) + sq(y[i]) + sq(z[i]));
                Fx = Fb * dbdx;
                Fy = Fb * dbdy;
                Fz = -1. + Fb * dbdz;
            }
            (void)Fb;
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
            if (c > 25 || isnan(delt)) {
                delt = 0.;
                x[i] = xi; y[i] = yi; z[i] = zi;
                c = 1000;
            }
            c = c + 1;
        }
        if (c < 26) {
            Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
            ux[i] = Fx / Fp;
            uy[i] = Fy / Fp;
            uz[i] = Fz / Fp;
        }
    }
}

/* woltsurf.f95:591-638 */
void pxfo_spocone(double *x, double *y, double *z, double *l, double *m, double *n,
                  double *ux, double *uy, double *uz, int64_t num, double R0, double tg)
{
    const double sl = tan(tg);
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double A = sq(n[i]) * sq(sl) - sq(m[i]) - sq(l[i]);
        double B = 2 * n[i] * sl * R0 + 2 * z[i] * sq(sl) * n[i] - 2 * x[i] * l[i] - 2 * y[i] * m[i];
        double C = sq(R0) + 2 * sl * R0 * z[i] + sq(z[i]) * sq(sl) - sq(x[i]) - sq(y[i]);
        double det = sq(B) - 4 * A * C;
        if (det >= 0) {
            double t1 = (-B + sqrt(det)) / (2 * A);
            double t2 = (-B - sqrt(det)) / (2 * A);
            if (fabs(t2) < fabs(t1)) t1 = t2;
            x[i] = x[i] + t1 * l[i];
            y[i] = y[i] + t1 * m[i];
            z[i] = z[i] + t1 * n[i];
            ux[i] = -x[i] / sqrt(sq(x[i]) + sq(y[i])) * cos(tg);
            uy[i] = -y[i] / sqrt(sq(x[i]) + sq(y[i])) * cos(tg);
            uz[i] = sin(tg);
        } else {
            l[i] = 0.; m[i] = 0.; n[i] = 0.;
        }
    }
}

/* ------------------------------------------------------------------ */
/* surfacesf.f95 -- remaining closed-form / Newton surfaces */
/* ------------------------------------------------------------------ */

/* surfacesf.f95:57-101 / :104-149.  x**2. is folded to x*x by gfortran. */
static void tracesphere_impl(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                             double *ux, double *uy, double *uz, int64_t num, double rad, double nr)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double dotol = l[i] * x[i] + m[i] * y[i] + n[i] * z[i];
        double mago = sq(x[i]) + sq(y[i]) + sq(z[i]);
        double determinant = sq(dotol) - mago + sq(rad);
        if (determinant < 0) {
            x[i] = 0.; y[i] = 0.; z[i] = 0.; l[i] = 0.; m[i] = 0.; n[i] = 0.;
        } else {
            double d1 = -dotol + sqrt(determinant);
            double d2 = -dotol - sqrt(determinant);
            if (fabs(d2) < fabs(d1)) d1 = d2;
            x[i] = x[i] + d1 * l[i];
            y[i] = y[i] + d1 * m[i];
            z[i] = z[i] + d1 * n[i];
            if (opd) opd[i] = opd[i] + d1 * nr;
        }
        mago = sqrt(sq(x[i]) + sq(y[i]) + sq(z[i]));
        ux[i] = x[i] / mago;
        uy[i] = y[i] / mago;
        uz[i] = z[i] / mago;
    }
}
void pxfo_tracesphere(double *x, double *y, double *z, double *l, double *m, double *n,
                      double *ux, double *uy, double *uz, int64_t num, double rad)
{
    tracesphere_impl(NULL, x, y, z, l, m, n, ux, uy, uz, num, rad, 0.);
}
void pxfo_tracesphereopd(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                         double *ux, double *uy, double *uz, int64_t num, double rad, double nr)
{
    tracesphere_impl(opd, x, y, z, l, m, n, ux, uy, uz, num, rad, nr);
}

/* surfacesf.f95:153-197 / :201-246 */
static void tracecyl_impl(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                          double *ux, double *uy, double *uz, int64_t num, double rad, double nr)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double a = sq(l[i]) + sq(n[i]);
        double b = 2 * (x[i] * l[i] + z[i] * n[i]);
        double c = sq(x[i]) + sq(z[i]) - sq(rad);
        double det = sq(b) - 4 * a * c;
        if (det < 0) {
            x[i] = 0.; y[i] = 0.; z[i] = 0.; l[i] = 0.; m[i] = 0.; n[i] = 0.;
        } else {
            double d1 = (-b + sqrt(det)) / 2 / a;
            double d2 = (-b - sqrt(det)) / 2 / a;
            if (fabs(d2) < fabs(d1)) d1 = d2;
            x[i] = x[i] + l[i] * d1;
            y[i] = y[i] + m[i] * d1;
            z[i] = z[i] + n[i] * d1;
            if (opd) opd[i] = opd[i] + d1 * nr;
        }
        double mag = sqrt(sq(x[i]) + sq(z[i]));
        ux[i] = x[i] / mag;
        uz[i] = z[i] / mag;
        uy[i] = 0.;
    }
}
void pxfo_tracecyl(double *x, double *y, double *z, double *l, double *m, double *n,
                   double *ux, double *uy, double *uz, int64_t num, double rad)
{
    tracecyl_impl(NULL, x, y, z, l, m, n, ux, uy, uz, num, rad, 0.);
}
void pxfo_tracecylopd(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                      double *ux, double *uy, double *uz, int64_t num, double rad, double nr)
{
    tracecyl_impl(opd, x, y, z, l, m, n, ux, uy, uz, num, rad, nr);
}

/* surfacesf.f95:251-296 (rad is the curvature; tol 1.e-10; uz = 0) */
void pxfo_cylconic(double *x, double *y, double *z, double *l, double *m, double *n,
                   double *ux, double *uy, double *uz, int64_t num, double rad, double k)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fp, low, high, dL, dH;
        int it = 0;
        while (fabs(delt) > TOL_1EM10 && it++ < PXF_NEWTON_CAP) {
            low = 1 + sqrt(1 - (1 + k) * sq(rad) * sq(x[i]));
            high = rad * sq(x[i]);
            dL = -(1 + k) * sq(rad) * x[i] / sqrt(1 - (1 + k) * sq(rad) * sq(x[i]));
            dH = 2 * rad * x[i];
            F = y[i] - high / low;
            Fx = (high * dL - low * dH) / sq(low);
            Fy = 1.;
            Fp = Fx * l[i] + Fy * m[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
        }
        Fp = sqrt(Fx * Fx + Fy * Fy);
        ux[i] = Fx / Fp;
        uy[i] = Fy / Fp;
        uz[i] = 0.;
    }
}

/* surfacesf.f95:423-440 / :443-460 */
void pxfo_paraxial(double *x, double *y, double *z, double *l, double *m, double *n,
                   double *ux, double *uy, double *uz, int64_t num, double F)
{
    (void)z; (void)n; (void)ux; (void)uy; (void)uz;
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        l[i] = l[i] - x[i] / F;
        m[i] = m[i] - y[i] / F;
    }
}
void pxfo_paraxialy(double *x, double *y, double *z, double *l, double *m, double *n,
                    double *ux, double *uy, double *uz, int64_t num, double F)
{
    (void)x; (void)z; (void)l; (void)n; (void)ux; (void)uy; (void)uz;
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) m[i] = m[i] - y[i] / F;
}

/* surfacesf.f95:468-508.  The normal is divided by sqrt(Fx**2+Fy**2) -- Fz is NOT in the norm
 * (:499), kept as written. */
void pxfo_torus(double *x, double *y, double *z, double *l, double *m, double *n,
                double *ux, double *uy, double *uz, int64_t num, double rin, double rout)
{
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp;
        int it = 0;
        while (fabs(delt) > TOL_1EM10 && it++ < PXF_NEWTON_CAP) {
            F = sq(sq(z[i] + rin + rout) + sq(y[i]) + sq(x[i]) + sq(rout) - sq(rin)) -
                (4 * sq(rout) * (sq(y[i]) + sq(z[i] + rin + rout)));
            Fx = 4 * x[i] * (-sq(rin) + sq(rin + rout + z[i]) + sq(rout) + sq(x[i]) + sq(y[i]));
            Fy = 4 * y[i] * (2 * rin * (rout + z[i]) + 2 * rout * z[i] + sq(z[i]) + sq(y[i]) + sq(x[i]));
            Fz = 4 * (rout + rin + z[i]) * (2 * rin * (rout + z[i]) + 2 * rout * z[i] + sq(z[i]) + sq(y[i]) + sq(x[i]));
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
        }
        Fp = sqrt(Fx * Fx + Fy * Fy);
        ux[i] = Fx / Fp;
        uy[i] = Fy / Fp;
        uz[i] = Fz / Fp;
    }
}

static double powi(double x, int m);

/* surfacesf.f95:514-572 / :578-638.  rad**(2*j) has a run-time INTEGER exponent: libgcc's
 * __powidf2 (binary exponentiation), restated in powi(). */
static void conicplus_impl(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                           double *ux, double *uy, double *uz, int64_t num, double R, double K,
                           const double *p, int Np, double nr)
{
    const double Fz = 1.;
    const double c = 1 / R;
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., rad, a0, a1, F, Fr, Fx = 0, Fy = 0, Fp, denom;
        int it = 0;
        while (fabs(delt) > TOL_1EM10 && it++ < PXF_NEWTON_CAP) {
            rad = sqrt(sq(x[i]) + sq(y[i]));
            a0 = 0.;
            a1 = 0.;
            for (int j = 1; j <= Np; j++) {
                a0 = a0 + p[j - 1] * powi(rad, 2 * j);
                a1 = a1 + p[j - 1] * (double)(2 * j) * powi(rad, 2 * j - 1);
            }
            denom = sqrt(1 - (K + 1) * sq(c) * sq(rad)) + 1;
            F = z[i] - c * sq(rad) / denom + a0;
            Fr = -(2 * c * rad / denom + ((K + 1) * cube(rad) * cube(c)) / (sq(denom) * sqrt(1 - (K + 1) * sq(c) * sq(rad)))) + a1;
            Fx = Fr * x[i] / rad;
            Fy = Fr * y[i] / rad;
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
            if (opd) opd[i] = opd[i] + delt * nr;
        }
        Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
        ux[i] = Fx / Fp;
        uy[i] = Fy / Fp;
        uz[i] = Fz / Fp;
    }
}
void pxfo_conicplus(double *x, double *y, double *z, double *l, double *m, double *n,
                    double *ux, double *uy, double *uz, int64_t num, double R, double K, const double *p, int32_t Np)
{
    conicplus_impl(NULL, x, y, z, l, m, n, ux, uy, uz, num, R, K, p, Np, 0.);
}
void pxfo_conicplusopd(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                       double *ux, double *uy, double *uz, int64_t num, double R, double K, const double *p,
                       int32_t Np, double nr)
{
    conicplus_impl(opd, x, y, z, l, m, n, ux, uy, uz, num, R, K, p, Np, nr);
}

/* surfacesf.f95:642-668 */
void pxfo_legsurf(double *x, double *y, double *z, double *l, double *m, double *n,
                  double *ux, double *uy, double *uz, double xwidth, double ywidth, double order,
                  const double *coeff, const int32_t *xo, const int32_t *yo, int32_t Nc, int64_t num)
{
    (void)z; (void)ux; (void)uy; (void)uz;
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double dphidx = 0, dphidy = 0;
        for (int j = 0; j < Nc; j++) {
            dphidx = dphidx + coeff[j] * pxfo_legendre(y[i] / ywidth, yo[j]) * pxfo_legendrep(x[i] / xwidth, xo[j]);
            dphidy = dphidy + coeff[j] * pxfo_legendrep(y[i] / ywidth, yo[j]) * pxfo_legendre(x[i] / xwidth, xo[j]);
        }
        l[i] = l[i] + dphidx * order / xwidth;
        m[i] = m[i] + dphidy * order / ywidth;
        n[i] = n[i] / fabs(n[i]) * sqrt(1. - sq(l[i]) - sq(m[i]));
    }
}

/* ------------------------------------------------------------------ */
/* woltsurf.f95 -- Wolter-Schwarzschild back surfaces (:726-933)       */
/* ------------------------------------------------------------------ */

/* woltsurf.f95:726-815: wsprimary with the radius reduced by `thick` */
void pxfo_wsprimaryback(double *x, double *y, double *z, double *l, double *m, double *n,
                        double *ux, double *uy, double *uz, int64_t num,
                        double alpha, double z0, double psi, double thick)
{
    const double betas = 4 * alpha;
    const double ff = z0 / cos(betas);
    const double g = ff / psi;
    const double k = sq(tan(betas / 2));
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp, Fb, kterm, beta, dbdx, dbdy, r, theta, x2, y2;
        int flag, c = 0;
        double xi = x[i], yi = y[i], zi = z[i];
        while (fabs(delt) > TOL_1EM8) {
            r = sqrt(sq(x[i]) + sq(y[i]));
            theta = atan2(y[i], x[i]);
            x2 = (r - thick) * cos(theta);
            y2 = (r - thick) * sin(theta);
            beta = asin(sqrt(sq(x2) + sq(y2)) / ff);
            flag = 0;
            if (beta <= betas) {
                beta = betas;
                flag = 1;
                kterm = 0.;
            } else {
                kterm = (1 / k) * sq(tan(beta / 2)) - 1;
            }
            F = -z[i] - ff * sq(sin(betas / 2)) +
                sq(ff) * sq(sin(beta)) / (4 * ff * sq(sin(betas / 2))) +
                g * pow4(cos(beta / 2)) * pow(kterm, 1 - k);
            Fb = sq(ff) * sin(beta) * cos(beta) / (2 * ff * sq(sin(betas / 2))) -
                 2 * g * cube(cos(beta / 2)) * sin(beta / 2) * pow(kterm, 1 - k) +
                 g * (1 - k) * cos(beta / 2) * sin(beta / 2) * pow(kterm, -k) * (1 / k);
            Fz = -1.;
            if (flag == 1) {
                r = sqrt(sq(x2) + sq(y2));
                Fb = sq(ff) * sin(betas) * cos(betas) / (2 * ff * sq(sin(betas / 2))) +
                     g * (1 - k) * cos(betas / 2) * sin(betas / 2) * (1 / k);
                F = F + (r - ff * sin(betas)) * z[i] / (sq(r) + sq(z[i])) * Fb;
                Fz = Fz + (r - ff * sin(betas)) * (sq(r) - sq(z[i])) / sq(sq(r) + sq(z[i])) * Fb;
            }
            dbdx = x2 / sqrt(1 - (sq(x2) + sq(y2)) / sq(ff)) / ff / sqrt(sq(x2) + sq(y2));
            dbdy = y2 / sqrt(1 - (sq(x2) + sq(y2)) / sq(ff)) / ff / sqrt(sq(x2) + sq(y2));
            Fx = Fb * dbdx;
            Fy = Fb * dbdy;
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
            if (c > 25 || isnan(delt)) {
                delt = 0.;
                x[i] = xi; y[i] = yi; z[i] = zi;
                c = 1000;
            }
            c = c + 1;
        }
        if (c < 26) {
            Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
            ux[i] = -Fx / Fp;
            uy[i] = -Fy / Fp;
            uz[i] = -Fz / Fp;
        }
    }
}

/* woltsurf.f95:824-933 */
void pxfo_wssecondaryback(double *x, double *y, double *z, double *l, double *m, double *n,
                          double *ux, double *uy, double *uz, int64_t num,
                          double alpha, double z0, double psi, double thick)
{
    const double betas = 4 * alpha;
    const double ff = z0 / cos(betas);
    const double g = ff / psi;
    const double k = sq(tan(betas / 2));
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp, Fb, kterm, beta, dbdx, dbdy, dbdz;
        double a, dadbs, dbdzs, gam, dadb, r, theta, x2, y2;
        int flag, c = 0;
        double xi = x[i], yi = y[i], zi = z[i];
        while (fabs(delt) > TOL_1EM8) {
            r = sqrt(sq(x[i]) + sq(y[i]));
            theta = atan2(y[i], x[i]);
            x2 = (r - thick) * cos(theta);
            y2 = (r - thick) * sin(theta);
            beta = atan2(sqrt(sq(x2) + sq(y2)), z[i]);
            flag = 0;
            if (beta <= betas) {
                beta = betas;
                kterm = 0;
                a = 1 / ff;
                flag = 1;
            } else {
                kterm = (1 / k) * sq(tan(beta / 2)) - 1;
                a = (1 - cos(beta)) / (1 - cos(betas)) / ff +
                    (1 + cos(beta)) / (2 * g) * pow(kterm, 1 + k);
            }
            F = -z[i] + cos(beta) / a;
            if (flag == 1) {
                Fb = 0.;
                dadbs = sin(betas) / ff / (1 - cos(betas)) +
                        (k + 1) * (cos(betas) + 1) * tan(betas / 2) / sq(cos(betas / 2)) / 2 / g / k;
                dbdzs = -sq(sin(betas)) / sqrt(sq(x2) + sq(y2));
                gam = (-ff * sin(betas) - sq(ff) * cos(betas) * dadbs) * dbdzs;
                F = F + gam * (z[i] - sqrt(sq(x2) + sq(y2)) / tan(betas));
                Fx = -2. / tan(betas) * x2 / sqrt(sq(x2) + sq(y2));
                Fy = -2. / tan(betas) * y2 / sqrt(sq(x2) + sq(y2));
                Fz = gam - 1.;
            } else {
                dadb = sin(beta) / ff / (1 - cos(betas)) -
                       sin(beta) / (2 * g) * pow(kterm, 1 + k) +
                       (k + 1) * (cos(beta) + 1) * tan(beta / 2) * pow(kterm, k) / 2 / g / k / sq(cos(beta / 2));
                Fb = -sin(beta) / a - cos(beta) / sq(a) * dadb;
                dbdx = x2 * z[i] / (sq(x2) + sq(y2) + sq(z[i])) / sqrt(sq(x2) + sq(y2));
                dbdy = y2 * z[i] / (sq(x2) + sq(y2) + sq(z[i])) / sqrt(sq(x2) + sq(y2));
                dbdz = -sqrt(sq(x2) + sq(y2)) / (sq(x2) + sq(y2) + sq(z[i]));
                Fx = Fb * dbdx;
                Fy = Fb * dbdy;
                Fz = -1. + Fb * dbdz;
            }
            (void)Fb;
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
            if (c > 25 || isnan(delt)) {
                delt = 0.;
                x[i] = xi; y[i] = yi; z[i] = zi;
                c = 1000;
            }
            c = c + 1;
        }
        if (c < 26) {
            Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
            ux[i] = Fx / Fp;
            uy[i] = Fy / Fp;
            uz[i] = Fz / Fp;
        }
    }
}

/* ------------------------------------------------------------------ */
/* specialFunctions.f95                                               */
/* ------------------------------------------------------------------ */

/* specialFunctions.f95:142-232.  rnm/rprime are radnum x radnum scratch
 * tables (Fortran ALLOCATEs them per call, :165-166).  Only same-parity
 * (i,j) cells are ever written or read. */
#define ZMAXRAD 64
static int zern_radnum(int znum)
{
    int tznum = 1, radnum = 1;
    while (tznum < znum) { tznum = tznum + (radnum + 1); radnum = radnum + 1; }
    return radnum;
}

void pxfo_zernset(double rho, double theta, const int32_t *rorder, const int32_t *aorder, int znum,
                  double *polyout, double *derrho, double *dertheta)
{
    int radnum = zern_radnum(znum);
    double *rnm = (double *)calloc((size_t)(radnum + 1) * (radnum + 1), sizeof(double));
    double *rprime = (double *)calloc((size_t)(radnum + 1) * (radnum + 1), sizeof(double));
#define RNM(i, j) rnm[(i) * (radnum + 1) + (j)]
#define RPR(i, j) rprime[(i) * (radnum + 1) + (j)]
    for (int i = 1; i <= radnum; i++) {
        double n = (double)i - 1;
        for (int j = (int)n + 1; j >= 1; j -= 2) {
            double m = (double)j - 1;
            if (rho == 0) {
                if (m == 1) {
                    RPR(i, j) = pow(-1., (n - 1) / 2) * (n + 1) / 2;
                    RNM(i, j) = 0.;
                } else if (m == 0) {
                    RPR(i, j) = 0.;
                    RNM(i, j) = pow(-1., n / 2);
                } else {
                    RPR(i, j) = 0.;
                    RNM(i, j) = 0.;
                }
            } else {
                if (n == m) {
                    RNM(i, j) = pow(rho, n);
                    RPR(i, j) = n * pow(rho, n - 1);
                } else if (m == n - 2) {
                    RNM(i, j) = n * RNM(i, i) - (n - 1) * RNM(i - 2, i - 2);
                    RPR(i, j) = n * RPR(i, i) - (n - 1) * RPR(i - 2, i - 2);
                } else {
                    double h3 = -4 * (m + 2) * (m + 1) / (n + m + 2) / (n - m);
                    double h2 = h3 * (n + m + 4) * (n - m - 2) / 4. / (m + 3) + (m + 2);
                    double h1 = .5 * (m + 4) * (m + 3) - (m + 4) * h2 + h3 * (n + m + 6) * (n - m - 4) / 8.;
                    RNM(i, j) = h1 * RNM(i, j + 4) + (h2 + h3 / sq(rho)) * RNM(i, j + 2);
                    RPR(i, j) = h1 * RPR(i, j + 4) + (h2 + h3 / sq(rho)) * RPR(i, j + 2) -
                                2 * h3 / cube(rho) * RNM(i, j + 2);
                }
            }
        }
    }
    const double sqrthalf = (double)sqrtf(0.5f);   /* sqrt(0.5) is REAL*4 (:225-226) */
    for (int i = 0; i < znum; i++) {
        double n = rorder[i];
        double mm = aorder[i];
        double m = fabs(mm);
        double norm = sqrt(2 * (n + 1));
        int in = (int)n + 1, im = (int)m + 1;
        if (mm < 0) {
            polyout[i] = norm * RNM(in, im) * sin(m * theta);
            derrho[i] = norm * RPR(in, im) * sin(m * theta);
            dertheta[i] = norm * RNM(in, im) * cos(m * theta) * m;
        } else if (mm > 0) {
            polyout[i] = norm * RNM(in, im) * cos(m * theta);
            derrho[i] = norm * RPR(in, im) * cos(m * theta);
            dertheta[i] = -norm * RNM(in, im) * sin(m * theta) * m;
        } else {
            polyout[i] = norm * sqrthalf * RNM(in, im);
            derrho[i] = norm * sqrthalf * RPR(in, im);
            dertheta[i] = 0.;
        }
    }
#undef RNM
#undef RPR
    free(rnm);
    free(rprime);
}

/* specialFunctions.f95:2-14 */
static double factorial(int n)
{
    double ans = 1;
    for (int i = 1; i <= n; i++) ans = ans * (double)i;
    return ans;
}

/* Fortran x**k with integer k (real base): libgcc __powidf2 / gfortran
 * inline expansion (binary exponentiation). */
static double powi(double x, int m)
{
    unsigned int n = m < 0 ? -(unsigned int)m : (unsigned int)m;
    double y = n % 2 ? x : 1;
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y *= x;
    }
    return m < 0 ? 1 / y : y;
}

/* specialFunctions.f95:17-38 (closed form; used only as a K10 cross-check) */
double pxfo_radialpoly(double rho, int n, int m)
{
    double output = 0.;
    int am = m < 0 ? -m : m;
    if (rho <= 1) {
        for (int j = 0; j <= (n - am) / 2; j++) {
            double sign = (j % 2) ? -1. : 1.;
            double cst = sign * factorial(n - j) / factorial(j) / factorial((n + m) / 2 - j) / factorial((n - m) / 2 - j);
            output = output + cst * powi(rho, n - 2 * j);
        }
    }
    return output;
}

/* specialFunctions.f95:337-360 */
double pxfo_legendre(double x, int n)
{
    double x2 = (fabs(x) > 1.) ? x / fabs(x) : x;
    double leg = 0.;
    if (n == 0) return 1.;
    for (int i = 0; i <= n / 2; i++) {
        double sign = (i % 2) ? -1. : 1.;
        /* (-1)**i*f(2n-2i)/f(i)/f(n-i)/f(n-2i)/2**n*x2**(n-2i); 2**n is INTEGER */
        leg = leg + sign * factorial(2 * n - 2 * i) / factorial(i) / factorial(n - i) / factorial(n - 2 * i) /
                        (double)(1 << n) * powi(x2, n - 2 * i);
    }
    return leg;
}

/* specialFunctions.f95:363-388 */
double pxfo_legendrep(double x, int n)
{
    double lp = 0.;
    if (n == 0) lp = 0.;
    else if (n == 1) lp = 1.;
    else if (x == 0. && (n % 2) == 0) lp = 0.;
    else {
        for (int i = 0; i <= n / 2; i++) {
            double sign = (i % 2) ? -1. : 1.;
            lp = lp + sign * factorial(2 * n - 2 * i) / factorial(i) / factorial(n - i) / factorial(n - 2 * i) /
                          (double)(1 << n) * (double)(n - 2 * i) * powi(x, n - 2 * i - 1);
        }
    }
    if (fabs(x) > 1.) lp = 0.;
    return lp;
}

/* ------------------------------------------------------------------ */
/* woltsurf.f95 -- Legendre-Legendre deformed shells */
/* ------------------------------------------------------------------ */

/* The Legendre additive block shared by wolterprimLL (:247-261), woltersecLL (:323-343) and
 * ellipsoidWoltLL (:674-688); evaluation order as written there. */
static void ll_terms(double x, double y, double z, double zmax, double zmin, double dphi,
                     const double *coeff, const int32_t *axial, const int32_t *az, int cnum,
                     double *add, double *addx, double *addy, double *addz)
{
    double ang = atan2(y, x);
    double zarg = (z - ((zmax + zmin) / 2.)) / ((zmax - zmin) / 2.);
    double targ = 2 * ang / dphi;
    double a0 = 0., ax = 0., ay = 0., azz = 0.;
    for (int a = 0; a < cnum; a++) {
        a0 = a0 + coeff[a] * pxfo_legendre(zarg, axial[a]) * pxfo_legendre(targ, az[a]);
        ax = ax - coeff[a] * pxfo_legendre(zarg, axial[a]) * pxfo_legendrep(targ, az[a]) * (2 / dphi) * (y / (sq(y) + sq(x)));
        ay = ay + coeff[a] * pxfo_legendre(zarg, axial[a]) * pxfo_legendrep(targ, az[a]) * (2 / dphi) * (x / (sq(y) + sq(x)));
        azz = azz + coeff[a] * pxfo_legendrep(zarg, axial[a]) * pxfo_legendre(targ, az[a]) * 2 / (zmax - zmin);
    }
    *add = a0; *addx = ax; *addy = ay; *addz = azz;
}

/* woltsurf.f95:219-288 (thetah = 3.*alpha, thetap = alpha; tol 1.e-10) */
void pxfo_wolterprimll(double *x, double *y, double *z, double *l, double *m, double *n,
                       double *ux, double *uy, double *uz, int64_t num, double r0, double z0,
                       double zmax, double zmin, double dphi, const double *coeff,
                       const int32_t *axial, const int32_t *az, int cnum)
{
    double alpha = .25 * atan(r0 / z0);
    double thetah = 3. * alpha;
    double thetap = alpha;
    double p = z0 * tan(4 * alpha) * tan(thetap);
    double d = z0 * tan(4 * alpha) * tan(4 * alpha - thetah);
    double e = cos(4 * alpha) * (1 + tan(4 * alpha) * tan(thetah));
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp, add, addx, addy, addz, G;
        int it = 0;
        while (fabs(delt) > TOL_1EM10 && it++ < PXF_NEWTON_CAP) {
            ll_terms(x[i], y[i], z[i], zmax, zmin, dphi, coeff, axial, az, cnum, &add, &addx, &addy, &addz);
            G = sqrt(sq(x[i]) + sq(y[i])) + add;
            F = -(sq(G) - sq(p) - 2 * p * z[i] - 4 * sq(e) * p * d / (sq(e) - 1));
            Fx = -2 * G * (x[i] / sqrt(sq(x[i]) + sq(y[i])) + addx);
            Fy = -2 * G * (y[i] / sqrt(sq(x[i]) + sq(y[i])) + addy);
            Fz = 2 * p - 2 * G * addz;
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
        }
        Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
        ux[i] = Fx / Fp; uy[i] = Fy / Fp; uz[i] = Fz / Fp;
    }
}

/* woltsurf.f95:293-379 (general psi; tol 1.e-7) */
void pxfo_woltersecll(double *x, double *y, double *z, double *l, double *m, double *n,
                      double *ux, double *uy, double *uz, int64_t num, double r0, double z0, double psi,
                      double zmax, double zmin, double dphi, const double *coeff,
                      const int32_t *axial, const int32_t *az, int cnum)
{
    double p, d, e;
    vanspeybroeck(r0, z0, psi, &p, &d, &e);
    (void)p;
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp, add, addx, addy, addz, G;
        int it = 0;
        while (fabs(delt) > TOL_1EM7 && it++ < PXF_NEWTON_CAP) {
            ll_terms(x[i], y[i], z[i], zmax, zmin, dphi, coeff, axial, az, cnum, &add, &addx, &addy, &addz);
            G = sqrt(sq(x[i]) + sq(y[i])) + add;
            F = -(sq(G) - sq(e) * sq(d + z[i]) + sq(z[i]));
            Fx = -2 * G * (x[i] / sqrt(sq(x[i]) + sq(y[i])) + addx);
            Fy = -2 * G * (y[i] / sqrt(sq(x[i]) + sq(y[i])) + addy);
            Fz = 2 * sq(e) * (d + z[i]) - 2 * z[i] - 2 * G * addz;
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
        }
        Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
        ux[i] = Fx / Fp; uy[i] = Fy / Fp; uz[i] = Fz / Fp;
    }
}

/* woltsurf.f95:643-718 (ellipsoid primary with L-L terms; tol 1.e-10) */
void pxfo_ellipsoidwoltll(double *x, double *y, double *z, double *l, double *m, double *n,
                          double *ux, double *uy, double *uz, int64_t num, double r0, double z0, double psi,
                          double S, double zmax, double zmin, double dphi, const double *coeff,
                          const int32_t *axial, const int32_t *az, int cnum)
{
    double P = r0 / sin((psi * asin(r0 / z0) - asin(r0 / S)) / (1 + psi));
    double ff = (S + P) / 2.;
    double bq = -(sq(r0) + sq(ff - P) + sq(ff));
    double cq = sq(ff) * sq(ff - P);
    double aa = sqrt((-bq + sqrt(sq(bq) - 4 * cq)) / 2.);
    double bb = sqrt(sq(aa) - sq(ff));
    double zfoc = ff - P + z0;
    #pragma omp parallel for
    for (int64_t i = 0; i < num; i++) {
        double delt = 100., F, Fx = 0, Fy = 0, Fz = 0, Fp, add, addx, addy, addz, G;
        int it = 0;
        while (fabs(delt) > TOL_1EM10 && it++ < PXF_NEWTON_CAP) {
            ll_terms(x[i], y[i], z[i], zmax, zmin, dphi, coeff, axial, az, cnum, &add, &addx, &addy, &addz);
            G = sqrt(sq(x[i]) + sq(y[i])) + add;
            F = sq(z[i] - zfoc) / sq(aa) + sq(G) / sq(bb) - 1.;
            Fx = 2 * G / sq(bb) * (x[i] / sqrt(sq(x[i]) + sq(y[i])) + addx);
            Fy = 2 * G / sq(bb) * (y[i] / sqrt(sq(x[i]) + sq(y[i])) + addy);
            Fz = 2 * (z[i] - zfoc) / sq(aa) + (2 * G / sq(bb)) * (addz);
            Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
            delt = -F / Fp;
            x[i] = x[i] + l[i] * delt;
            y[i] = y[i] + m[i] * delt;
            z[i] = z[i] + n[i] * delt;
        }
        Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
        ux[i] = Fx / Fp; uy[i] = Fy / Fp; uz[i] = Fz / Fp;
    }
}

/* ------------------------------------------------------------------ */
/* zernsurf.f95                                                       */
/* ------------------------------------------------------------------ */

/* zernsurf.f95:8-101 (opd==NULL) and :108-203 (opd) */
static void tracezern_impl(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                           double *ux, double *uy, double *uz, int64_t num,
                           const double *coeff, const int32_t *rorder, const int32_t *aorder,
                           int arrsize, double rad, double nr)
{
    #pragma omp parallel
    {
        double *zern = (double *)malloc(sizeof(double) * (size_t)arrsize);
        double *rhoder = (double *)malloc(sizeof(double) * (size_t)arrsize);
        double *thetader = (double *)malloc(sizeof(double) * (size_t)arrsize);
        #pragma omp for
        for (int64_t i = 0; i < num; i++) {
            double t = 0., delta = 100.;
            double F, Frho, Ftheta, Frhox, Frhoy, Fthetax, Fthetay, Fx = 0, Fy = 0, Fz = 0, Fp, rho, theta;
            int it = 0;
            while (fabs(delta) > TOL_1EM10 && it++ < PXF_NEWTON_CAP) {
                rho = sqrt(sq(x[i]) + sq(y[i]));
                theta = atan2(y[i], x[i]);
                F = z[i];
                Frho = 0.;
                Ftheta = 0.;
                pxfo_zernset(rho / rad, theta, rorder, aorder, arrsize, zern, rhoder, thetader);
                for (int c = 0; c < arrsize; c++) {
                    F = F - coeff[c] * zern[c];
                    Frho = Frho - coeff[c] * rhoder[c] / rad;
                    Ftheta = Ftheta - coeff[c] * thetader[c];
                }
                Frhox = (x[i] / rho) * Frho;
                Frhoy = (y[i] / rho) * Frho;
                Fthetax = (-y[i] / rho) * Ftheta / rho;
                Fthetay = (x[i] / rho) * Ftheta / rho;
                Fx = Frhox + Fthetax;
                Fy = Frhoy + Fthetay;
                Fz = 1.;
                Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
                delta = -F / Fp;
                x[i] = x[i] + l[i] * delta;
                y[i] = y[i] + m[i] * delta;
                z[i] = z[i] + n[i] * delta;
                t = t + delta;
            }
            Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
            ux[i] = Fx / Fp;
            uy[i] = Fy / Fp;
            uz[i] = Fz / Fp;
            if (opd) opd[i] = opd[i] + t * nr;
        }
        free(zern); free(rhoder); free(thetader);
    }
}

/* zernsurf.f95:206-250 (serial in the reference) */
void pxfo_zernphase(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                    double *ux, double *uy, double *uz, int64_t num,
                    const double *coeff, const int32_t *rorder, const int32_t *aorder, int arrsize,
                    double rad, double wave)
{
    (void)z; (void)ux; (void)uy; (void)uz;
    #pragma omp parallel
    {
        double *zern = (double *)malloc(sizeof(double) * (size_t)arrsize);
        double *rhoder = (double *)malloc(sizeof(double) * (size_t)arrsize);
        double *thetader = (double *)malloc(sizeof(double) * (size_t)arrsize);
        #pragma omp for
        for (int64_t i = 0; i < num; i++) {
            double rho = sqrt(sq(x[i]) + sq(y[i]));
            double theta = atan2(y[i], x[i]);
            pxfo_zernset(rho / rad, theta, rorder, aorder, arrsize, zern, rhoder, thetader);
            double F = 0., Frho = 0., Ftheta = 0.;
            for (int c = 0; c < arrsize; c++) {
                F = F + coeff[c] * zern[c];
                Frho = Frho + coeff[c] * rhoder[c] / rad;
                Ftheta = Ftheta + coeff[c] * thetader[c];
            }
            double Frhox = (x[i] / rho) * Frho;
            double Frhoy = (y[i] / rho) * Frho;
            double Fthetax = (-y[i] / rho) * Ftheta / rho;
            double Fthetay = (x[i] / rho) * Ftheta / rho;
            double Fx = Frhox + Fthetax;
            double Fy = Frhoy + Fthetay;
            l[i] = l[i] + Fx * wave;
            m[i] = m[i] + Fy * wave;
            n[i] = copysign(sqrt(1. - sq(l[i]) - sq(m[i])), n[i]);
            opd[i] = opd[i] + F * wave;
        }
        free(zern); free(rhoder); free(thetader);
    }
}

/* zernsurf.f95:257-359 */
void pxfo_tracezernrot(double *x, double *y, double *z, double *l, double *m, double *n,
                       double *ux, double *uy, double *uz, int64_t num,
                       const double *coeff1, const int32_t *rorder1, const int32_t *aorder1, int arrsize1,
                       const double *coeff2, const int32_t *rorder2, const int32_t *aorder2, int arrsize2,
                       double rad, double rot)
{
    #pragma omp parallel
    {
        int mx = arrsize1 > arrsize2 ? arrsize1 : arrsize2;
        double *zern = (double *)malloc(sizeof(double) * (size_t)mx);
        double *rhoder = (double *)malloc(sizeof(double) * (size_t)mx);
        double *thetader = (double *)malloc(sizeof(double) * (size_t)mx);
        #pragma omp for
        for (int64_t i = 0; i < num; i++) {
            double t = 0., delta = 100.;
            double F, Frho, Ftheta, Frhox, Frhoy, Fthetax, Fthetay, Fx = 0, Fy = 0, Fz = 0, Fp, rho, theta;
            int it = 0;
            while (fabs(delta) > TOL_1EM10 && it++ < PXF_NEWTON_CAP) {
                rho = sqrt(sq(x[i]) + sq(y[i]));
                theta = atan2(y[i], x[i]);
                F = z[i];
                Frho = 0.;
                Ftheta = 0.;
                pxfo_zernset(rho / rad, theta, rorder1, aorder1, arrsize1, zern, rhoder, thetader);
                for (int c = 0; c < arrsize1; c++) {
                    F = F - coeff1[c] * zern[c];
                    Frho = Frho - coeff1[c] * rhoder[c] / rad;
                    Ftheta = Ftheta - coeff1[c] * thetader[c];
                }
                pxfo_zernset(rho / rad, theta + rot, rorder2, aorder2, arrsize2, zern, rhoder, thetader);
                for (int c = 0; c < arrsize2; c++) {
                    F = F - coeff2[c] * zern[c];
                    Frho = Frho - coeff2[c] * rhoder[c] / rad;
                    Ftheta = Ftheta - coeff2[c] * thetader[c];
                }
                Frhox = (x[i] / rho) * Frho;
                Frhoy = (y[i] / rho) * Frho;
                Fthetax = (-y[i] / rho) * Ftheta / rho;
                Fthetay = (x[i] / rho) * Ftheta / rho;
                Fx = Frhox + Fthetax;
                Fy = Frhoy + Fthetay;
                Fz = 1.;
                Fp = Fx * l[i] + Fy * m[i] + Fz * n[i];
                delta = -F / Fp;
                x[i] = x[i] + l[i] * delta;
                y[i] = y[i] + m[i] * delta;
                z[i] = z[i] + n[i] * delta;
                t = t + delta;
            }
            Fp = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
            ux[i] = Fx / Fp;
            uy[i] = Fy / Fp;
            uz[i] = Fz / Fp;
        }
        free(zern); free(rhoder); free(thetader);
    }
}

void pxfo_tracezern(double *x, double *y, double *z, double *l, double *m, double *n,
                    double *ux, double *uy, double *uz, int64_t num,
                    const double *coeff, const int32_t *rorder, const int32_t *aorder, int arrsize, double rad)
{
    tracezern_impl(NULL, x, y, z, l, m, n, ux, uy, uz, num, coeff, rorder, aorder, arrsize, rad, 0.);
}

void pxfo_tracezernopd(double *opd, double *x, double *y, double *z, double *l, double *m, double *n,
                       double *ux, double *uy, double *uz, int64_t num,
                       const double *coeff, const int32_t *rorder, const int32_t *aorder, int arrsize,
                       double rad, double nr)
{
    tracezern_impl(opd, x, y, z, l, m, n, ux, uy, uz, num, coeff, rorder, aorder, arrsize, rad, nr);
}

/* Newton step counter for diagnostics (tests only): number of iterations the
 * woltersecondary loop takes for one ray; mirrors woltsurf.f95:136-150. */
int pxfo_woltersecondary_steps(double x, double y, double z, double l, double m, double n,
                               double r0, double z0, double psi)
{
    double p, d, e;
    vanspeybroeck(r0, z0, psi, &p, &d, &e);
    (void)p;
    double delt = 100.;
    int it = 0;
    while (fabs(delt) > TOL_1EM8 && it < PXF_NEWTON_CAP) {
        double F = sq(e) * sq(d + z) - sq(z) - sq(x) - sq(y);
        double Fx = -2. * x, Fy = -2. * y, Fz = 2 * sq(e) * (d + z) - 2 * z;
        double Fp = Fx * l + Fy * m + Fz * n;
        delt = -F / Fp;
        x = x + l * delt; y = y + m * delt; z = z + n * delt;
        it++;
    }
    return it;
}

/* ======================================================================== reconstruct.f95
 * Southwell wavefront reconstruction.  Arrays are Fortran column-major:
 * A(xi,yi) = a[(xi-1) + (yi-1)*xdim], xi fastest. */
#define RA(a, xi, yi) (a)[((xi) - 1) + (int64_t)((yi) - 1) * xdim]

/* reconstruct.f95:1-128.  Successive over-relaxation in lexicographic Gauss-Seidel order (xi outer,
 * yi inner); 100. marks an invalid lenslet.  Returns the number of sweeps executed (the Fortran
 * returns nothing; used by the tests). */
int64_t pxfo_reconstruct(double *xang, double *yang, int32_t xdim, int32_t ydim, double criteria, double h,
                         double *phase, double *phasec, int32_t maxiter)
{
    const double pi = (double)3.1415926535897931f;        /* :13 default-real literal: REAL*4 */
    const double w = 2 / (1 + sin(pi / (sqrt((double)xdim * (double)ydim) + 1)));   /* :14 */
    const int64_t ncell = (int64_t)xdim * ydim;
    int64_t sweeps = 0;
    int counter = 0;
    memcpy(phasec, phase, (size_t)ncell * sizeof(double));                         /* :17 */
    for (;;) {
        double rms = 0.;
        int nannum = 0;
        for (int xi = 2; xi <= xdim - 1; xi++) {
            for (int yi = 2; yi <= ydim - 1; yi++) {
                int compute = 1;
                double yplus = 0, yneg = 0, xplus = 0, xneg = 0, pyplus, pyneg, pxplus, pxneg, bk = 0, psum = 0, goodpix = 4.;
                if (RA(phasec, xi, yi) == 100.) compute = 0;                       /* :33-35 */
                if (compute == 1) {
                    yplus = RA(yang, xi, yi + 1); if (yplus == 100.) yplus = -RA(yang, xi, yi);   /* :39-42 */
                    yneg = RA(yang, xi, yi - 1);  if (yneg == 100.) yneg = -RA(yang, xi, yi);
                    xplus = RA(xang, xi + 1, yi); if (xplus == 100.) xplus = -RA(xang, xi, yi);
                    xneg = RA(xang, xi - 1, yi);  if (xneg == 100.) xneg = -RA(xang, xi, yi);
                    bk = .5 * (yplus - yneg + xplus - xneg) * h;                   /* :57 */
                    pyplus = RA(phasec, xi, yi + 1); if (pyplus == 100.) { pyplus = 0.; goodpix = goodpix - 1; }
                    pyneg = RA(phasec, xi, yi - 1);  if (pyneg == 100.) { pyneg = 0.; goodpix = goodpix - 1; }
                    pxplus = RA(phasec, xi + 1, yi); if (pxplus == 100.) { pxplus = 0.; goodpix = goodpix - 1; }
                    pxneg = RA(phasec, xi - 1, yi);  if (pxneg == 100.) { pxneg = 0.; goodpix = goodpix - 1; }
                    psum = pyplus + pyneg + pxplus + pxneg;                        /* :85 */
                    if (goodpix == 0.) {                                           /* :88-94 */
                        RA(phasec, xi, yi) = 100.;
                        RA(phase, xi, yi) = 100.;
                        RA(xang, xi, yi) = 100.;
                        RA(yang, xi, yi) = 100.;
                        compute = 0;
                    }
                }
                if (compute == 1) {
                    RA(phasec, xi, yi) = RA(phasec, xi, yi) + w * ((psum + bk) / goodpix - RA(phasec, xi, yi));   /* :99 */
                    nannum = nannum + 1;
                    double dlt = RA(phasec, xi, yi) - RA(phase, xi, yi);
                    rms = rms + dlt * dlt;                                         /* :103 */
                }
            }
        }
        sweeps++;
        rms = sqrt(rms / nannum);                                                  /* :112 (0/0 = NaN when no cell) */
        memcpy(phase, phasec, (size_t)ncell * sizeof(double));                     /* :114 */
        if (rms < criteria) break;
        counter = counter + 1;
        if (counter > maxiter) break;
    }
    return sweeps;
}

/* reconstruct.f95:136-187.  xang, yang, phase are intent(out) and accumulated into without being
 * cleared (:165-166): the routine relies on zero-filled arrays; restated with an explicit clear.  Rays
 * whose bin falls outside the array corrupt memory in the reference; they are skipped here.  The
 * reference's OpenMP loop races on the bins; single-thread semantics (ray order) are restated. */
void pxfo_southwellbin(const double *x, const double *y, const double *l, const double *m, int64_t num,
                       double binsize, double *xang, double *yang, double *phase, int32_t xdim, int32_t ydim)
{
    const int64_t ncell = (int64_t)xdim * ydim;
    int *accum = (int *)calloc((size_t)ncell, sizeof(int));
    memset(xang, 0, (size_t)ncell * sizeof(double));
    memset(yang, 0, (size_t)ncell * sizeof(double));
    memset(phase, 0, (size_t)ncell * sizeof(double));
    for (int64_t i = 0; i < num; i++) {
        int xb, yb;
        if (xdim % 2 == 0) xb = (int)floor(x[i] / binsize) + xdim / 2 + 1 + 1;     /* :155-156 */
        else xb = (int)floor((x[i] + binsize / 2) / binsize) + (xdim - 1) / 2 + 1;
        if (ydim % 2 == 0) yb = (int)floor(y[i] / binsize) + ydim / 2 + 1 + 1;
        else yb = (int)floor((y[i] + binsize / 2) / binsize) + (ydim - 1) / 2 + 1;
        if (xb < 1 || xb > xdim || yb < 1 || yb > ydim) continue;
        RA(xang, xb, yb) = RA(xang, xb, yb) + l[i];
        RA(yang, xb, yb) = RA(yang, xb, yb) + m[i];
        accum[(xb - 1) + (int64_t)(yb - 1) * xdim] += 1;
    }
    for (int64_t c = 0; c < ncell; c++) {
        if (accum[c] == 0) { phase[c] = 100.; xang[c] = 100.; yang[c] = 100.; }
        else { xang[c] = tan(asin(xang[c] / accum[c])); yang[c] = tan(asin(yang[c] / accum[c])); }   /* :180-181 */
    }
    free(accum);
}
