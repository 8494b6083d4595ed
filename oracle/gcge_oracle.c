/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the leaf kernels of the GCG hot path.
 *
 * A CPU restatement of what the reference computes at the OPS slot level, written
 * from the reference's behaviour (file:line cited per function), with no BLAS: the
 * loops below are the checker for the CUDA kernels when oracle/_ref (the compiled
 * reference) is not available, and are themselves pinned against oracle/_ref in
 * tests/test_oracle.py.  Compiled with -ffp-contract=off so a*b+c is two roundings,
 * like the reference's gcc -O2 x86-64 build.
 *
 * Multi-vectors here are column-major n x ncols with leading dimension n, exactly
 * LAPACKVEC (reference app/app_lapack.h:17-20).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* y[:,s1:e1] = A x[:,s0:e0], CCS scatter -- reference app/app_ccs.c:116-131 */
void oracle_ccs_spmm(int n, const int *j_col, const int *i_row, const double *data,
		const double *x, double *y, int ncols)
{
	for (int c = 0; c < ncols; ++c) {
		const double *dx = x + (size_t)n * c;
		double *dy = y + (size_t)n * c;
		memset(dy, 0, sizeof(double) * (size_t)n);
		for (int j = 0; j < n; ++j)
			for (int e = j_col[j]; e < j_col[j + 1]; ++e)
				dy[i_row[e]] += data[e] * dx[j];
	}
}

/* y = alpha x + beta y on n*ncols contiguous entries -- reference app/app_lapack.c:334-395
 * (memset for beta==0, dscal unless beta==1, then daxpy; x==NULL => scale only) */
void oracle_axpby(size_t len, double alpha, const double *x, double beta, double *y)
{
	if (beta == 0.0) memset(y, 0, sizeof(double) * len);
	else if (beta != 1.0) for (size_t i = 0; i < len; ++i) y[i] *= beta;
	if (x) for (size_t i = 0; i < len; ++i) y[i] += alpha * x[i];
}

/* C(p x q, ldc) = X^T Y -- reference app/app_lapack.c:136-181 ('N'), :116-134 ('S'),
 * :67-115 ('D': C[ldc*i] = x_i . y_i) */
void oracle_gram(char mode, int n, int p, int q, const double *x, const double *y, double *c, int ldc)
{
	if (mode == 'D') {
		for (int i = 0; i < p; ++i) {
			double s = 0.0;
			for (int r = 0; r < n; ++r) s += x[(size_t)n * i + r] * y[(size_t)n * i + r];
			c[(size_t)ldc * i] = s;
		}
		return;
	}
	for (int j = 0; j < q; ++j)
		for (int i = 0; i < p; ++i) {
			if (mode == 'S' && i < j) { c[(size_t)ldc * j + i] = c[(size_t)ldc * i + j]; continue; }
			double s = 0.0;
			for (int r = 0; r < n; ++r) s += x[(size_t)n * i + r] * y[(size_t)n * j + r];
			c[(size_t)ldc * j + i] = s;
		}
}

/* y(n x q) = x(n x p) coef + y diag(beta) -- reference app/app_lapack.c:463-534 */
void oracle_linear_comb(int n, int p, int q, const double *x, const double *coef, int ldc,
		const double *beta, int incb, double *y)
{
	for (int j = 0; j < q; ++j) {
		double *yc = y + (size_t)n * j;
		if (beta) {
			double b = beta[(size_t)incb * j];
			if (b != 1.0) for (int r = 0; r < n; ++r) yc[r] *= b;
		}
		if (x && coef) {
			if (!beta) memset(yc, 0, sizeof(double) * (size_t)n);
			for (int k = 0; k < p; ++k) {
				const double ck = coef[(size_t)ldc * j + k];
				const double *xc = x + (size_t)n * k;
				for (int r = 0; r < n; ++r) yc[r] += xc[r] * ck;
			}
		}
	}
}

/* x[i] = rand()/(RAND_MAX+1.0), column-major order -- reference app/app_lapack.c:322-333 */
void oracle_fill_random(double *x, size_t len)
{
	for (size_t i = 0; i < len; ++i) x[i] = ((double)rand()) / ((double)RAND_MAX + 1);
}
void oracle_srand(unsigned seed) { srand(seed); }
