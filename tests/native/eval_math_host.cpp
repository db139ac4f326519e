// Host build of arvae_b200/csrc/eval_math.cuh for the CPU tests (the same source the device code compiles).
#include "../../arvae_b200/csrc/eval_math.cuh"

extern "C" double evalmath_student_t_two_sided(double t, double dof) { return arvae::student_t_two_sided(t, dof); }
extern "C" double evalmath_correlation_t(double rho, double dof) { return arvae::correlation_t(rho, dof); }
