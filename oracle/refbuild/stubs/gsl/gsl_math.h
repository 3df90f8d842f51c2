// Stand-in for the GSL headers PBSM3D.hpp:43-47 includes.  GSL is only reached on the use_subgrid_topo* branches
// (PBSM3D.cpp:538-640), which the B200 path refuses; the harness aborts if one of them is ever called.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
struct gsl_function { double (*function)(double x, void* params); void* params; };
struct gsl_integration_workspace { int unused; };
[[noreturn]] inline void gsl_refharness_unsupported(const char* what)
{
    std::fprintf(stderr, "oracle/refbuild: %s is not available in the reference harness (optional PBSM3D branch)\n", what);
    std::abort();
}
inline double gsl_ran_gaussian_pdf(double, double) { gsl_refharness_unsupported("gsl_ran_gaussian_pdf"); }
inline double gsl_ran_gamma_pdf(double, double, double) { gsl_refharness_unsupported("gsl_ran_gamma_pdf"); }
inline double gsl_cdf_gaussian_Q(double, double) { gsl_refharness_unsupported("gsl_cdf_gaussian_Q"); }
inline double gsl_cdf_gamma_P(double, double, double) { gsl_refharness_unsupported("gsl_cdf_gamma_P"); }
inline double gsl_cdf_gamma_Q(double, double, double) { gsl_refharness_unsupported("gsl_cdf_gamma_Q"); }
inline double gsl_sf_gamma(double) { gsl_refharness_unsupported("gsl_sf_gamma"); }
inline gsl_integration_workspace* gsl_integration_workspace_alloc(size_t) { gsl_refharness_unsupported("gsl_integration_workspace_alloc"); }
inline void gsl_integration_workspace_free(gsl_integration_workspace*) {}
inline int gsl_integration_qags(const gsl_function*, double, double, double, double, size_t, gsl_integration_workspace*, double*, double*)
{
    gsl_refharness_unsupported("gsl_integration_qags");
}
