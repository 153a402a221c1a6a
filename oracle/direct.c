/* TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * float64 direct convolution: the definition the partitioned convolver must equal
 * (SURVEY.md 8.A "Identity").  O(nh) per output sample, so tests use it on short
 * signals and on spot-checked windows of long ones.
 */
#include "oracle.h"

void orc_direct_convolve(const double* x, unsigned nx, const double* h, unsigned nh, unsigned n0, unsigned count,
                         double* y, int nthreads) {
  long i;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (i = 0; i < (long)count; i++) {
    unsigned n = n0 + (unsigned)i, j;
    unsigned jmax = (n < nh - 1) ? n : nh - 1; /* x[n-j] needs n-j >= 0 */
    double acc = 0.0, comp = 0.0;
    for (j = 0; j <= jmax; j++) {
      unsigned k = n - j;
      if (k >= nx) continue;
      /* Kahan summation keeps the truth far below fp32 noise for 144k-tap IRs */
      double term = h[j] * x[k] - comp;
      double t = acc + term;
      comp = (t - acc) - term;
      acc = t;
    }
    y[i] = acc;
  }
}
