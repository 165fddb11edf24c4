/* TEST INFRASTRUCTURE (oracle build only) -- opaque GSL RNG stand-in. GSL is linked by the reference
 * only for gsl_rng_uniform in a random-ID helper (algorithms/genetic_data_func.cpp:36-65), never on
 * the association path. */
#ifndef ORACLE_SHIM_GSL_RNG_H
#define ORACLE_SHIM_GSL_RNG_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { int dummy; } gsl_rng_type;
typedef struct { unsigned long long s; } gsl_rng;
extern const gsl_rng_type *gsl_rng_default;
extern const gsl_rng_type *gsl_rng_mt19937;
const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(const gsl_rng *r, unsigned long int seed);
double gsl_rng_uniform(const gsl_rng *r);
unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n);
void gsl_rng_free(gsl_rng *r);
#ifdef __cplusplus
}
#endif
#endif
