/* TEST INFRASTRUCTURE (oracle build only) -- definitions for shims/gsl/gsl_rng.h. */
#include <stdlib.h>
#include "gsl/gsl_rng.h"
static const gsl_rng_type the_type = {0};
const gsl_rng_type *gsl_rng_default = &the_type;
const gsl_rng_type *gsl_rng_mt19937 = &the_type;
const gsl_rng_type *gsl_rng_env_setup(void) { return &the_type; }
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) { (void)T; gsl_rng *r = (gsl_rng *)calloc(1, sizeof *r); r->s = 88172645463325252ULL; return r; }
void gsl_rng_set(const gsl_rng *r, unsigned long int seed) { ((gsl_rng *)r)->s = seed ? seed : 1; }
static unsigned long long step(gsl_rng *r) { unsigned long long x = r->s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return r->s = x; }
double gsl_rng_uniform(const gsl_rng *r) { return (double)(step((gsl_rng *)r) >> 11) / 9007199254740992.0; }
unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n) { return n ? step((gsl_rng *)r) % n : 0; }
void gsl_rng_free(gsl_rng *r) { free(r); }
