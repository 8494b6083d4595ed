// Device matrix: CCS (caller) -> CSR (device) conversion, round trip, MatAxpby.
//
// The reference multiplies y = A x with a column-major scatter over the CCS arrays
// (reference app/app_ccs.c:116-131): for column j ascending, y[i_row[e]] += data[e]*x[j].
// Row r of y therefore accumulates its entries in ascending column order.  The device CSR
// keeps exactly that order inside each row, so a sequential gather reproduces the
// reference's floating-point sum bit for bit.
#include "b200_internal.h"
#include <vector>

extern "C" int b200_mat_create_from_ccs(int nrows, int ncols, const int *j_col, const int *i_row,
                                        const double *data, b200_mat **out)
{
	B200_REQUIRE_INIT();
	B200_CHECK(out && j_col && nrows >= 0 && ncols >= 0, "b200_mat_create_from_ccs: bad arguments");
	const int nnz = j_col[ncols];
	B200_CHECK(nnz >= 0 && (nnz == 0 || (i_row && data)), "b200_mat_create_from_ccs: bad CCS arrays");
	for (int j = 0; j < ncols; ++j)
		B200_CHECK(j_col[j] <= j_col[j + 1], "b200_mat_create_from_ccs: j_col not monotone at %d", j);

	// host transpose (counting sort by row, columns visited ascending)
	std::vector<int> rp((size_t)nrows + 1, 0), ci((size_t)nnz);
	std::vector<double> va((size_t)nnz);
	for (int e = 0; e < nnz; ++e) {
		B200_CHECK(i_row[e] >= 0 && i_row[e] < nrows, "b200_mat_create_from_ccs: row index %d out of range at %d",
		           i_row[e], e);
		++rp[(size_t)i_row[e] + 1];
	}
	for (int r = 0; r < nrows; ++r) rp[r + 1] += rp[r];
	{
		std::vector<int> next(rp.begin(), rp.end() - 1);
		for (int j = 0; j < ncols; ++j)
			for (int e = j_col[j]; e < j_col[j + 1]; ++e) {
				const int pos = next[i_row[e]]++;
				ci[pos] = j; va[pos] = data[e];
			}
	}
	// identical images (symmetric matrix, sorted columns) => share storage
	int shared = (nrows == ncols);
	if (shared) shared = (0 == memcmp(rp.data(), j_col, sizeof(int) * ((size_t)nrows + 1)));
	if (shared && nnz) shared = (0 == memcmp(ci.data(), i_row, sizeof(int) * (size_t)nnz));
	if (shared && nnz) shared = (0 == memcmp(va.data(), data, sizeof(double) * (size_t)nnz));

	b200_mat *A = (b200_mat *)calloc(1, sizeof(b200_mat));
	A->nrows = nrows; A->ncols = ncols; A->nnz = nnz; A->t_shared = shared; A->row0 = 0;
	for (int r = 0; r < nrows; ++r) if (rp[r + 1] - rp[r] > A->max_row_nnz) A->max_row_nnz = rp[r + 1] - rp[r];
	for (int j = 0; j < ncols; ++j) if (j_col[j + 1] - j_col[j] > A->t_max_row_nnz) A->t_max_row_nnz = j_col[j + 1] - j_col[j];
	cudaStream_t st = g_b200.stream;
	const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
	B200_CUDA(cudaMalloc(&A->rp, sizeof(int) * ((size_t)nrows + 1)));
	B200_CUDA(cudaMalloc(&A->ci, sizeof(int) * nz));
	B200_CUDA(cudaMalloc(&A->va, sizeof(double) * nz));
	B200_CUDA(cudaMemcpyAsync(A->rp, rp.data(), sizeof(int) * ((size_t)nrows + 1), cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(A->ci, ci.data(), sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(A->va, va.data(), sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	if (shared) {
		A->t_rp = A->rp; A->t_ci = A->ci; A->t_va = A->va;
	} else {
		B200_CUDA(cudaMalloc(&A->t_rp, sizeof(int) * ((size_t)ncols + 1)));
		B200_CUDA(cudaMalloc(&A->t_ci, sizeof(int) * nz));
		B200_CUDA(cudaMalloc(&A->t_va, sizeof(double) * nz));
		B200_CUDA(cudaMemcpyAsync(A->t_rp, j_col, sizeof(int) * ((size_t)ncols + 1), cudaMemcpyHostToDevice, st));
		B200_CUDA(cudaMemcpyAsync(A->t_ci, i_row, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
		B200_CUDA(cudaMemcpyAsync(A->t_va, data, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	}
	B200_CUDA(cudaStreamSynchronize(st));
	*out = A;
	return 0;
}

extern "C" int b200_mat_destroy(b200_mat *A)
{
	if (!A) return 0;
	if (g_b200.initialised) cudaStreamSynchronize(g_b200.stream);
	if (!A->t_shared) { cudaFree(A->t_rp); cudaFree(A->t_ci); cudaFree(A->t_va); }
	cudaFree(A->rp); cudaFree(A->ci); cudaFree(A->va);
	free(A);
	return 0;
}

extern "C" int b200_mat_shape(const b200_mat *A, int *nrows, int *ncols, int *nnz)
{
	B200_CHECK(A, "b200_mat_shape: NULL matrix");
	if (nrows) *nrows = A->nrows;
	if (ncols) *ncols = A->ncols;
	if (nnz) *nnz = A->nnz;
	return 0;
}

extern "C" int b200_mat_to_ccs(const b200_mat *A, int *j_col, int *i_row, double *data)
{
	B200_REQUIRE_INIT();
	B200_CHECK(A && j_col, "b200_mat_to_ccs: bad arguments");
	cudaStream_t st = g_b200.stream;
	B200_CUDA(cudaMemcpyAsync(j_col, A->t_rp, sizeof(int) * ((size_t)A->ncols + 1), cudaMemcpyDeviceToHost, st));
	if (A->nnz) {
		B200_CUDA(cudaMemcpyAsync(i_row, A->t_ci, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaMemcpyAsync(data, A->t_va, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost, st));
	}
	B200_CUDA(cudaStreamSynchronize(st));
	return 0;
}

// values only: Y = alpha X + beta Y elementwise over the nnz array (identical patterns)
__global__ void mat_axpby_kernel(int nnz, double alpha, const double *__restrict__ x, double beta,
                                 double *__restrict__ y)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += gridDim.x * blockDim.x) {
		// same operation order as a BLAS dscal followed by daxpy
		double v = (beta == 1.0) ? y[i] : beta * y[i];
		y[i] = fma(alpha, x[i], v);
	}
}

extern "C" int b200_mat_axpby(double alpha, const b200_mat *X, double beta, b200_mat *Y)
{
	B200_REQUIRE_INIT();
	B200_CHECK(X && Y, "b200_mat_axpby: NULL matrix");
	B200_CHECK(X->nrows == Y->nrows && X->ncols == Y->ncols && X->nnz == Y->nnz,
	           "b200_mat_axpby: shape/nnz mismatch (identical sparsity patterns required)");
	if (Y->nnz == 0) return 0;
	const int threads = 256;
	const int blocks = b200_ceil_div(Y->nnz, threads) < g_b200.num_sms * 8 ? b200_ceil_div(Y->nnz, threads)
	                                                                      : g_b200.num_sms * 8;
	mat_axpby_kernel<<<blocks, threads, 0, g_b200.stream>>>(Y->nnz, alpha, X->va, beta, Y->va);
	B200_KERNEL_CHECK();
	if (!Y->t_shared) {
		B200_CHECK(!X->t_shared || X->t_va, "b200_mat_axpby: internal");
		mat_axpby_kernel<<<blocks, threads, 0, g_b200.stream>>>(Y->nnz, alpha, X->t_va, beta, Y->t_va);
		B200_KERNEL_CHECK();
	}
	return 0;
}
