/* TEST INFRASTRUCTURE (oracle build only) -- shim for R's standalone math library.
 * The reference includes "Rmath.h" (genetics/genotype/common_genotype_func.h:36-37) and calls only
 * pchisq(x, df, lower_tail, log_p) (algorithms/epistasis_func.cpp:242,294,340; src/test/pairwise.c:44).
 * Rmath is an un-vendored system dependency (no version pinned by the reference). For integer df the
 * chi-square upper tail has a closed form; df in {1,2,4} is all the reference can request (it uses 4). */
#ifndef ORACLE_SHIM_RMATH_H
#define ORACLE_SHIM_RMATH_H
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif
static inline double oracle_shim_chisq_upper(double x, double df) {
    if (x <= 0.0) return 1.0;
    if (df == 1.0) return erfc(sqrt(0.5 * x));
    if (df == 2.0) return exp(-0.5 * x);
    if (df == 4.0) return exp(-0.5 * x) * (1.0 + 0.5 * x);
    return NAN;
}
static inline double pchisq(double x, double df, int lower_tail, int log_p) {
    double q = oracle_shim_chisq_upper(x, df);
    double p = lower_tail ? 1.0 - q : q;
    return log_p ? log(p) : p;
}
#ifdef __cplusplus
}
#endif
#endif
