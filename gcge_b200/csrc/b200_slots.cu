// Slot-level entry points that carry HOST operands across the boundary: inner products /
// Gram blocks (device -> host), QtAP (SpMM + Gram, with the mv_ws side effect), linear
// combinations (host coefficients -> device).  The fused L3 providers use the b200k_*
// device-pointer primitives directly and never come through here.
#include "b200_internal.h"

// copy a compact device column-major (rows x cols) result to a strided host matrix
static int fetch_to_host(const double *dev, int rows, int cols, double *host, int ld)
{
	const size_t cnt = (size_t)rows * cols;
	double *pin = (double *)b200_pinned(0, sizeof(double) * cnt);
	if (!pin) return 1;
	B200_CUDA(cudaMemcpyAsync(pin, dev, sizeof(double) * cnt, cudaMemcpyDeviceToHost, g_b200.stream));
	B200_CUDA(cudaStreamSynchronize(g_b200.stream));
	for (int j = 0; j < cols; ++j) memcpy(host + (size_t)j * ld, pin + (size_t)j * rows, sizeof(double) * rows);
	return 0;
}

static int inner_prod_impl(char nsd, const double *x, int ldx, const double *y, int ldy, long long n,
                           int p, int q, double *host, int ld, int dist)
{
	if (p <= 0 || q <= 0) return 0;
	if (nsd == 'D') {
		B200_CHECK(p == q, "inner_prod 'D': %d x %d is not square", p, q);
		double *dev = (double *)b200_scratch(3, sizeof(double) * (size_t)p);
		if (!dev) return 1;
		if (b200k_gram('D', n, p, q, 1.0, x, ldx, y, ldy, dev, 1, 0, dist)) return 1;
		double *pin = (double *)b200_pinned(0, sizeof(double) * (size_t)p);
		if (!pin) return 1;
		B200_CUDA(cudaMemcpyAsync(pin, dev, sizeof(double) * (size_t)p, cudaMemcpyDeviceToHost, g_b200.stream));
		B200_CUDA(cudaStreamSynchronize(g_b200.stream));
		for (int i = 0; i < p; ++i) host[(size_t)ld * i] = pin[i];
		return 0;
	}
	if (nsd == 'S') B200_CHECK(p == q, "inner_prod 'S': %d x %d is not square", p, q);
	B200_CHECK(ld >= p, "inner_prod: ld %d < %d rows", ld, p);
	double *dev = (double *)b200_scratch(3, sizeof(double) * (size_t)p * q);
	if (!dev) return 1;
	if (b200k_gram(nsd == 'S' ? 'S' : 'N', n, p, q, 1.0, x, ldx, y, ldy, dev, 1, p, dist)) return 1;
	return fetch_to_host(dev, p, q, host, ld);
}

extern "C" int b200_mv_inner_prod(char nsd, const b200_mv *x, const b200_mv *y,
                                  const int *start, const int *end, double *inner_prod, int ld)
{
	B200_REQUIRE_INIT();
	B200_CHECK(x && y && start && end && inner_prod, "b200_mv_inner_prod: bad arguments");
	B200_CHECK(nsd == 'N' || nsd == 'S' || nsd == 'D', "b200_mv_inner_prod: mode '%c'", nsd);
	B200_CHECK(x->nrows == y->nrows, "b200_mv_inner_prod: row counts differ");
	B200_CHECK(start[0] >= 0 && end[0] <= x->ncols && start[1] >= 0 && end[1] <= y->ncols,
	           "b200_mv_inner_prod: column range out of bounds");
	// several ranks: the result is the GLOBAL inner product on every rank -- the reference, built
	// without MPI, expects that from MultiVecLocalInnerProd and MultiVecInnerProd alike
	return inner_prod_impl(nsd, x->d + start[0], x->ld, y->d + start[1], y->ld, x->nrows,
	                       end[0] - start[0], end[1] - start[1], inner_prod, ld, x->dist);
}

extern "C" int b200_mv_qtap(char ntsA, char ntsdQAP, const b200_mv *Q, const b200_mat *A, const b200_mv *P,
                            const int *start, const int *end, double *qAp, int ldQAP, b200_mv *ws)
{
	B200_REQUIRE_INIT();
	B200_CHECK(Q && P && start && end && qAp, "b200_mv_qtap: bad arguments");
	const int nr = end[0] - start[0], nc = end[1] - start[1];
	if (nr <= 0 || nc <= 0) return 0;   // reference src/ops_multi_vec.c:358
	B200_CHECK(start[0] >= 0 && end[0] <= Q->ncols && start[1] >= 0 && end[1] <= P->ncols,
	           "b200_mv_qtap: column range out of bounds");
	const double *pd; int pld; long long n;
	if (A) {
		B200_CHECK(ws && ws->ncols >= nc, "b200_mv_qtap: workspace needs %d columns", nc);
		const int tr = (ntsA == 'T');
		const int out_rows = tr ? A->ncols : A->nrows, in_rows = tr ? A->nrows : A->ncols;
		B200_CHECK(P->nrows == in_rows && ws->nrows == out_rows && Q->nrows == out_rows,
		           "b200_mv_qtap: shapes do not match the matrix");
		if (b200k_spmm_check_halo(A, P)) return 1;
		int rc = b200k_spmm(A, tr ? 1 : 0, P->d + start[1], P->ld, ws->d, ws->ld, nc, nullptr);
		if (rc) return rc;
		pd = ws->d; pld = ws->ld; n = ws->nrows;
	} else {
		B200_CHECK(P->nrows == Q->nrows, "b200_mv_qtap: row counts differ");
		pd = P->d + start[1]; pld = P->ld; n = P->nrows;
	}
	if (ntsdQAP == 'T')   // store the transpose: (A P)^T Q, nc x nr (reference src/ops_multi_vec.c:394-398)
		return inner_prod_impl('N', pd, pld, Q->d + start[0], Q->ld, n, nc, nr, qAp, ldQAP, Q->dist);
	return inner_prod_impl(ntsdQAP, Q->d + start[0], Q->ld, pd, pld, n, nr, nc, qAp, ldQAP, Q->dist);
}

extern "C" int b200_mv_linear_comb(const b200_mv *x, b200_mv *y, const int *start, const int *end,
                                   const double *coef, int ldc, const double *beta, int incb)
{
	B200_REQUIRE_INIT();
	B200_CHECK(y && start && end, "b200_mv_linear_comb: bad arguments");
	const int p = end[0] - start[0], q = end[1] - start[1];
	if (p == 0 || q == 0 || y->nrows == 0) return 0;  // reference app/app_lapack.c:473-478
	B200_CHECK(p > 0 && q > 0 && start[1] >= 0 && end[1] <= y->ncols, "b200_mv_linear_comb: bad ranges");
	const bool gemm = (x != nullptr && coef != nullptr);
	if (gemm) {
		B200_CHECK(x->nrows == y->nrows, "b200_mv_linear_comb: row counts differ");
		B200_CHECK(start[0] >= 0 && end[0] <= x->ncols && ldc >= p, "b200_mv_linear_comb: bad x range / ldc");
		if (x == y) {
			const bool overlap = start[0] < end[1] && start[1] < end[0];
			B200_CHECK(!overlap, "b200_mv_linear_comb: x and y column ranges overlap on one multi-vector");
		}
	}
	// stage coef (compact p x q) and beta (q) through pinned memory
	const size_t ncoef = gemm ? (size_t)p * q : 0;
	const size_t nbeta = beta ? (size_t)q : 0;
	double *c_dev = nullptr, *b_dev = nullptr;
	if (ncoef + nbeta) {
		B200_CUDA(cudaStreamSynchronize(g_b200.stream));   // pinned staging may still be in flight
		double *pin = (double *)b200_pinned(0, sizeof(double) * (ncoef + nbeta));
		double *dev = (double *)b200_scratch(1, sizeof(double) * (ncoef + nbeta));
		if (!pin || !dev) return 1;
		// staged ROW-major (element (i, j) at i*q + j): the transposition costs nothing extra here and
		// lets the TMA-fed kernel take the call (it wants contiguous coefficient rows)
		for (int j = 0; gemm && j < q; ++j)
			for (int i = 0; i < p; ++i) pin[(size_t)i * q + j] = coef[(size_t)j * ldc + i];
		for (int j = 0; beta && j < q; ++j) pin[ncoef + j] = beta[(size_t)incb * j];
		B200_CUDA(cudaMemcpyAsync(dev, pin, sizeof(double) * (ncoef + nbeta), cudaMemcpyHostToDevice,
		                          g_b200.stream));
		if (gemm) c_dev = dev;
		if (beta) b_dev = dev + ncoef;
	}
	return b200k_lincomb(y->nrows, p, q, gemm ? x->d + start[0] : nullptr, gemm ? x->ld : 0, c_dev, q, 1,
	                     b_dev, 1, y->d + start[1], y->ld);
}
