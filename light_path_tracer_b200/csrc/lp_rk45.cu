// lp_rk45.cu — kernel (1b): batched 8-D Hamiltonian tracer with scipy-RK45 semantics.
// (placeholder until the kernel lands: the entry points exist and report UNSUPPORTED)
#include "lp_internal.cuh"

extern "C" int lp_schw_rk45_trace_batch(const double *, int64_t, double, double, double,
                                        double, double, double, double, double, double,
                                        double *, double *, int8_t *, int32_t *, void *)
{
    return LP_ERR_UNSUPPORTED;
}

extern "C" int lp_schw_rk45_trace_path(double, double, double, double, double, double, double, double,
                                       double, double, double *, int32_t, int32_t *, int8_t *, int32_t *, void *)
{
    return LP_ERR_UNSUPPORTED;
}
