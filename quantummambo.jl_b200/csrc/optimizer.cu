// optimizer.cu -- placeholder translation unit; the BFGS / multistart scheduler lands here.
#include <string>
#include "../../include/mambo_csa.h"
extern "C" {
int mambo_csa_optimize(mambo_csa*, const double*, double*, const mambo_opt_options_t*, mambo_opt_result_t*) { return MAMBO_ESTATE; }
int mambo_multistart_run(mambo_csa**, int, int, const double*, double*, const mambo_opt_options_t*, mambo_opt_result_t*, int*) { return MAMBO_ESTATE; }
}
