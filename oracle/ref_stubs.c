/* ref_stubs.c -- link-time stand-ins for the compiled reference (oracle/_ref/libyael_ref.so).
 *
 * TEST INFRASTRUCTURE ONLY.  gmm.c is compiled unmodified for its E-step (gmm_compute_p, the part
 * SURVEY.md 8(f)-N4 names); its learning / Fisher-vector code also calls fmat_mul_tr from
 * matrix.c, which needs a LAPACK the recipe does not link.  Those entry points are outside the
 * path and are never called by the tests, so the symbol resolves to a loud abort. */
#include <stdio.h>
#include <stdlib.h>

void fmat_mul_tr(const float *left, const float *right, int m, int n, int k, float *result) {
  (void)left; (void)right; (void)m; (void)n; (void)k; (void)result;
  fprintf(stderr, "oracle/_ref: fmat_mul_tr (yael/matrix.c) is not part of this build\n");
  abort();
}
