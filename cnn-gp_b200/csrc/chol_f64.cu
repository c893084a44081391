// chol_f64.cu -- blocked fp64 Cholesky / triangular solves / prediction (placeholder).
#include "plan.h"

extern "C" {

int cnngp_potrf_upper_f64(double *, int64_t, int64_t, int32_t *, void *) {
    cnngp::set_error("cnngp_potrf_upper_f64: not implemented yet");
    return 100;
}
int cnngp_potrs_upper_f64(const double *, int64_t, int64_t, double *, int32_t, int64_t, void *) {
    cnngp::set_error("cnngp_potrs_upper_f64: not implemented yet");
    return 100;
}
int cnngp_predict_argmax(const float *, int64_t, int64_t, int64_t, const double *, int32_t, int64_t *, double *,
                         void *) {
    cnngp::set_error("cnngp_predict_argmax: not implemented yet");
    return 100;
}
}
