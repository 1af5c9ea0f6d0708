// oracle/oracle_prims.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
// Probes that expose the math helpers of halide_math.h to the tests.
#include <omp.h>

#include "halide_math.h"
#include "oracle.h"

extern "C" {
float oracle_halide_exp(float x) { return hl::halide_exp(x); }
float oracle_halide_log(float x) { return hl::halide_log(x); }
float oracle_halide_pow(float x, float y) { return hl::halide_pow(x, y); }
float oracle_fast_exp(float x) { return hl::fast_exp(x); }
int oracle_div_floor(int a, int b) { return hl::div_floor(a, b); }
int oracle_mod_floor(int a, int b) { return hl::mod_floor(a, b); }
int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
}
