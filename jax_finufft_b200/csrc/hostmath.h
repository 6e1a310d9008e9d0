// hostmath.h -- host-side plan arithmetic of libb200nufft (no CUDA types).
#pragma once
#include <stdint.h>
#include <cstdio>

#include "../../include/b200nufft.h"

namespace b2n {
constexpr double PI_D = 3.141592653589793238462643383279502884;
constexpr int MAX_NQUAD_H = 100;
constexpr int MAX_NCOEF_H = 24;

int setup_spreader(double eps, double upsampfac, int kerevalmeth, bool is_double, int *ns, double *beta);
int64_t next235beven(int64_t n, int64_t b);
int64_t set_nf_type12(int64_t ms, double upsampfac, int ns);
void gaussquad(int n, double *x, double *w);
double es_kernel(double x, int ns, double beta);
int kernel_quadrature(int ns, double beta, bool type3, int64_t nf, double *f, double *zp);
void fseries_host(int64_t nf, int ns, double beta, double *out);
void set_nhg_type3(double S, double X, double upsampfac, int ns, bool is_double, int64_t *nf, double *h, double *gam);
void widcen(double lo, double hi, bool is_double, double *w, double *c);
int horner_fit(int ns, double beta, bool is_double, double *coef);
}  // namespace b2n
