// Host-side check of pyxfocus_b200/csrc/pxf_crmath.cuh (the correctly rounded functions of the W-S long-trip
// re-trace) against binary128 libquadmath rounded once to double.  Built and run by tests/test_crmath.py:
//   g++ -O2 -ffp-contract=off tests/native/crmath_check.cpp -o ... -lquadmath
// Prints, per function, the number of arguments tried and the number whose result is not the correctly rounded one.
#include <quadmath.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../pyxfocus_b200/csrc/pxf_crmath.cuh"

static uint64_t s = 88172645463325252ull;
static double urand() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char **argv)
{
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    long bad[6] = {0, 0, 0, 0, 0, 0};
    double worst[6] = {0, 0, 0, 0, 0, 0};
    for (long i = 0; i < n; i++) {
        // the W-S ranges (beta ~ 0.02, kterm in (0, few), exponents +-k, 1+-k with k ~ 1e-4) and wide ranges
        const bool wide = (i & 3) == 3;
        double a = wide ? (urand() - .5) * 200. : urand() * .1;
        double x = wide ? (urand() - .5) * 1.9 : urand() * .05;
        double yy = wide ? (urand() - .5) * 1e4 : urand() * 300.;
        double xx = wide ? (urand() - .5) * 1e4 : 9000. + urand() * 2000.;
        double pb = wide ? urand() * 1e3 + 1e-6 : urand() * 3. + 1e-9;
        double pe = wide ? (urand() - .5) * 20. : ((i & 4) ? 1. : 0.) + (urand() - .5) * 4e-4;
        double got[6] = {pxfcr::cr_sin(a), pxfcr::cr_cos(a), pxfcr::cr_tan(a), pxfcr::cr_asin(x), pxfcr::cr_atan2(yy, xx),
                         pxfcr::cr_pow(pb, pe)};
        double want[6] = {(double)sinq(a), (double)cosq(a), (double)tanq(a), (double)asinq(x), (double)atan2q(yy, xx),
                          (double)powq(pb, pe)};
        for (int k = 0; k < 6; k++)
            if (got[k] != want[k]) {
                bad[k]++;
                double e = fabs(got[k] - want[k]) / fabs(want[k]);
                if (e > worst[k]) worst[k] = e;
            }
    }
    const char *nm[6] = {"sin", "cos", "tan", "asin", "atan2", "pow"};
    for (int k = 0; k < 6; k++) printf("%s %ld %ld %.3e\n", nm[k], n, bad[k], worst[k]);
    return 0;
}
