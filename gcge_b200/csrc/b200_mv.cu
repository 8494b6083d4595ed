// Device multi-vector store (row-major n x ld) and the HBM-bound elementwise kernels:
// axpby / scale / copy on column ranges, column-major <-> row-major staging, RNG fill.
#include "b200_internal.h"

static int mv_ld_for(int ncols)
{
	// rows start on 32-byte sector boundaries once there are enough columns to matter
	if (ncols <= 2) return ncols < 1 ? 1 : ncols;
	return (ncols + 3) & ~3;
}

extern "C" void b200_partition_range(long long n, int rank, int nranks, long long *lo, long long *hi);

// nrows is the GLOBAL row count.  On several ranks only this rank's row block is allocated,
// followed by room for the SpMM halo rows of the widest-halo matrix with that many columns
// created so far (matrices are created before their multi-vectors: MultiVecCreateByMat).
extern "C" int b200_mv_create(int nrows, int ncols, b200_mv **out)
{
	B200_REQUIRE_INIT();
	B200_CHECK(out && nrows >= 0 && ncols >= 0, "b200_mv_create: bad arguments");
	b200_mv *x = (b200_mv *)calloc(1, sizeof(b200_mv));
	long long lo = 0, hi = nrows;
	int halo = 0;
	if (b200_multi()) {
		b200_partition_range(nrows, g_b200.rank, g_b200.nranks, &lo, &hi);
		for (int i = 0; i < 8; ++i) if (g_b200.halo_n[i] == nrows) halo = g_b200.halo_cap[i];
		x->dist = 1;
	}
	x->nrows_global = nrows; x->row0 = lo; x->halo_cap = halo;
	nrows = (int)(hi - lo);
	x->nrows = nrows; x->ncols = ncols; x->ld = mv_ld_for(ncols); x->owner = 1;
	size_t bytes = sizeof(double) * (size_t)(nrows + 2 * halo > 0 ? nrows + 2 * halo : 1) * (size_t)x->ld;
	cudaError_t e = cudaMalloc(&x->alloc, bytes);
	if (e != cudaSuccess) {
		free(x);
		return b200_fail("b200_mv_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
	}
	x->d = x->alloc + (size_t)halo * x->ld;
	B200_CUDA(cudaMemsetAsync(x->alloc, 0, bytes, g_b200.stream));
	*out = x;
	return 0;
}

extern "C" int b200_mv_destroy(b200_mv *x)
{
	if (!x) return 0;
	if (x->owner) {
		if (g_b200.initialised) cudaStreamSynchronize(g_b200.stream);
		cudaFree(x->alloc);
	}
	free(x);
	return 0;
}

extern "C" int b200_mv_view(const b200_mv *x, int start, int end, b200_mv **view)
{
	B200_CHECK(x && view && start >= 0 && start <= end && end <= x->ncols, "b200_mv_view: bad arguments");
	b200_mv *v = (b200_mv *)calloc(1, sizeof(b200_mv));
	*v = *x;
	v->ncols = end - start; v->d = x->d + start; v->owner = 0; v->alloc = nullptr;
	*view = v;
	return 0;
}

extern "C" int b200_mv_shape(const b200_mv *x, int *nrows, int *ncols)
{
	B200_CHECK(x, "b200_mv_shape: NULL multi-vector");
	if (nrows) *nrows = x->nrows_global;
	if (ncols) *ncols = x->ncols;
	return 0;
}

extern "C" int b200_mv_local_range(const b200_mv *x, int *row0, int *nrows_local)
{
	B200_CHECK(x, "b200_mv_local_range: NULL multi-vector");
	if (row0) *row0 = (int)x->row0;
	if (nrows_local) *nrows_local = x->nrows;
	return 0;
}

// ------------------------------------------------------------------ transposes
// cm: column-major with leading dimension ld_cm (elements); rm: row-major with ld_rm.
// 32x32 tiles through shared memory so both sides are coalesced.
__global__ void cm_to_rm_kernel(long long n, int k, const double *__restrict__ cm, long long ld_cm,
                                double *__restrict__ rm, int ld_rm)
{
	__shared__ double tile[32][33];
	const long long r0 = (long long)blockIdx.x * 32;
	const int c0 = blockIdx.y * 32;
	for (int j = threadIdx.y; j < 32; j += blockDim.y) {
		const long long r = r0 + threadIdx.x; const int c = c0 + j;
		if (r < n && c < k) tile[j][threadIdx.x] = cm[(size_t)c * ld_cm + r];
	}
	__syncthreads();
	for (int j = threadIdx.y; j < 32; j += blockDim.y) {
		const long long r = r0 + j; const int c = c0 + threadIdx.x;
		if (r < n && c < k) rm[(size_t)r * ld_rm + c] = tile[threadIdx.x][j];
	}
}

__global__ void rm_to_cm_kernel(long long n, int k, const double *__restrict__ rm, int ld_rm,
                                double *__restrict__ cm, long long ld_cm)
{
	__shared__ double tile[32][33];
	const long long r0 = (long long)blockIdx.x * 32;
	const int c0 = blockIdx.y * 32;
	for (int j = threadIdx.y; j < 32; j += blockDim.y) {
		const long long r = r0 + j; const int c = c0 + threadIdx.x;
		if (r < n && c < k) tile[j][threadIdx.x] = rm[(size_t)r * ld_rm + c];
	}
	__syncthreads();
	for (int j = threadIdx.y; j < 32; j += blockDim.y) {
		const long long r = r0 + threadIdx.x; const int c = c0 + j;
		if (r < n && c < k) cm[(size_t)c * ld_cm + r] = tile[threadIdx.x][j];
	}
}

int b200k_cm_to_rm(long long n, int k, const double *cm, long long ld_cm, double *rm, int ld_rm)
{
	if (n <= 0 || k <= 0) return 0;
	dim3 grid((unsigned)((n + 31) / 32), (unsigned)((k + 31) / 32)), block(32, 8);
	cm_to_rm_kernel<<<grid, block, 0, g_b200.stream>>>(n, k, cm, ld_cm, rm, ld_rm);
	B200_KERNEL_CHECK();
	return 0;
}

int b200k_rm_to_cm(long long n, int k, const double *rm, int ld_rm, double *cm, long long ld_cm)
{
	if (n <= 0 || k <= 0) return 0;
	dim3 grid((unsigned)((n + 31) / 32), (unsigned)((k + 31) / 32)), block(32, 8);
	rm_to_cm_kernel<<<grid, block, 0, g_b200.stream>>>(n, k, rm, ld_rm, cm, ld_cm);
	B200_KERNEL_CHECK();
	return 0;
}

// host column-major <-> device, staged through scratch[2] in column chunks
static const size_t kStageBytes = (size_t)256 << 20;

static int mv_upload_rows(b200_mv *x, int start, int end, const double *host, int ld);
static int mv_download_rows(const b200_mv *x, int start, int end, double *host, int ld);

extern "C" int b200_mv_upload(b200_mv *x, int start, int end, const double *host, int ld)
{
	B200_REQUIRE_INIT();
	B200_CHECK(x && host && start >= 0 && end <= x->ncols && start <= end && ld >= x->nrows_global,
	           "b200_mv_upload: bad arguments");
	// host is the GLOBAL column-major block; this rank takes its rows [row0, row0 + nrows)
	return mv_upload_rows(x, start, end, host + x->row0, ld);
}

// host holds this rank's rows only (nrows_local x (end-start), column-major, ld >= nrows_local)
extern "C" int b200_mv_upload_local(b200_mv *x, int start, int end, const double *host, int ld)
{
	B200_REQUIRE_INIT();
	B200_CHECK(x && host && start >= 0 && end <= x->ncols && start <= end && ld >= x->nrows,
	           "b200_mv_upload_local: bad arguments");
	return mv_upload_rows(x, start, end, host, ld);
}

static int mv_upload_rows(b200_mv *x, int start, int end, const double *host, int ld)
{
	const long long n = x->nrows;
	if (n == 0 || end == start) return 0;
	int chunk = (int)(kStageBytes / (sizeof(double) * (size_t)n));
	if (chunk < 1) chunk = 1;
	for (int c0 = start; c0 < end; c0 += chunk) {
		const int k = (end - c0 < chunk) ? end - c0 : chunk;
		double *stage = (double *)b200_scratch(2, sizeof(double) * (size_t)n * k);
		if (!stage) return 1;
		B200_CUDA(cudaMemcpy2DAsync(stage, sizeof(double) * n, host + (size_t)(c0 - start) * ld,
		                            sizeof(double) * (size_t)ld, sizeof(double) * n, k,
		                            cudaMemcpyHostToDevice, g_b200.stream));
		if (b200k_cm_to_rm(n, k, stage, n, x->d + c0, x->ld)) return 1;
		// the staging buffer is reused by the next chunk and host memory is pageable
		B200_CUDA(cudaStreamSynchronize(g_b200.stream));
	}
	return 0;
}

extern "C" int b200_mv_download(const b200_mv *x, int start, int end, double *host, int ld)
{
	B200_REQUIRE_INIT();
	B200_CHECK(x && host && start >= 0 && end <= x->ncols && start <= end && ld >= x->nrows_global,
	           "b200_mv_download: bad arguments");
	// host is the GLOBAL column-major block; this rank fills its rows [row0, row0 + nrows) only
	return mv_download_rows(x, start, end, host + x->row0, ld);
}

extern "C" int b200_mv_download_local(const b200_mv *x, int start, int end, double *host, int ld)
{
	B200_REQUIRE_INIT();
	B200_CHECK(x && host && start >= 0 && end <= x->ncols && start <= end && ld >= x->nrows,
	           "b200_mv_download_local: bad arguments");
	return mv_download_rows(x, start, end, host, ld);
}

static int mv_download_rows(const b200_mv *x, int start, int end, double *host, int ld)
{
	const long long n = x->nrows;
	if (n == 0 || end == start) return 0;
	int chunk = (int)(kStageBytes / (sizeof(double) * (size_t)n));
	if (chunk < 1) chunk = 1;
	for (int c0 = start; c0 < end; c0 += chunk) {
		const int k = (end - c0 < chunk) ? end - c0 : chunk;
		double *stage = (double *)b200_scratch(2, sizeof(double) * (size_t)n * k);
		if (!stage) return 1;
		if (b200k_rm_to_cm(n, k, x->d + c0, x->ld, stage, n)) return 1;
		B200_CUDA(cudaMemcpy2DAsync(host + (size_t)(c0 - start) * ld, sizeof(double) * (size_t)ld, stage,
		                            sizeof(double) * n, sizeof(double) * n, k, cudaMemcpyDeviceToHost,
		                            g_b200.stream));
		B200_CUDA(cudaStreamSynchronize(g_b200.stream));
	}
	return 0;
}

// ------------------------------------------------------------------ axpby
// One CTA handles ROWS_PER_CTA consecutive rows of the n x k block; threads walk the
// block's elements in memory order (column fastest), so each row contributes one
// contiguous k*8-byte segment per warp request.
template <bool HAS_X, bool HAS_Y>
__global__ void axpby_kernel(long long n, int k, int rows_per_cta, double alpha, const double *x, int ldx,
                             double beta, double *y, int ldy)
{
	const long long r0 = (long long)blockIdx.x * rows_per_cta;
	long long nr = n - r0; if (nr > rows_per_cta) nr = rows_per_cta;
	const int total = (int)nr * k;
	for (int i = threadIdx.x; i < total; i += blockDim.x) {
		const int r = i / k, c = i - r * k;
		const size_t yo = (size_t)(r0 + r) * ldy + c;
		double v = 0.0;
		if (HAS_Y) { v = y[yo]; if (beta != 1.0) v *= beta; }
		if (HAS_X) v = fma(alpha, x[(size_t)(r0 + r) * ldx + c], v);
		y[yo] = v;
	}
}

int b200k_axpby(long long n, int k, double alpha, const double *x, int ldx, double beta, double *y, int ldy)
{
	if (n <= 0 || k <= 0) return 0;
	const bool has_x = (x != nullptr);
	const bool has_y = (beta != 0.0);
	if (!has_x && has_y && beta == 1.0) return 0;
	B200Prof prof(B200_PROF_AXPBY, 8.0 * n * k * (1 + (has_x ? 1 : 0) + (has_y ? 1 : 0)), 2.0 * n * k);
	int rows = 4096 / k; if (rows < 1) rows = 1;
	const unsigned grid = (unsigned)((n + rows - 1) / rows);
	const int threads = 256;
	cudaStream_t st = g_b200.stream;
	if (has_x && has_y)       axpby_kernel<true, true><<<grid, threads, 0, st>>>(n, k, rows, alpha, x, ldx, beta, y, ldy);
	else if (has_x && !has_y) axpby_kernel<true, false><<<grid, threads, 0, st>>>(n, k, rows, alpha, x, ldx, beta, y, ldy);
	else if (!has_x && has_y) axpby_kernel<false, true><<<grid, threads, 0, st>>>(n, k, rows, alpha, x, ldx, beta, y, ldy);
	else                      axpby_kernel<false, false><<<grid, threads, 0, st>>>(n, k, rows, alpha, x, ldx, beta, y, ldy);
	B200_KERNEL_CHECK();
	return 0;
}

extern "C" int b200_mv_axpby(double alpha, const b200_mv *x, double beta, b200_mv *y,
                             const int *start, const int *end)
{
	B200_REQUIRE_INIT();
	B200_CHECK(y && start && end, "b200_mv_axpby: bad arguments");
	const int k = end[1] - start[1];
	B200_CHECK(end[0] - start[0] == k, "b200_mv_axpby: column counts differ (%d vs %d)", end[0] - start[0], k);
	if (k == 0 || y->nrows == 0) return 0;
	B200_CHECK(start[1] >= 0 && end[1] <= y->ncols, "b200_mv_axpby: y range [%d,%d) outside %d columns",
	           start[1], end[1], y->ncols);
	if (x) {
		B200_CHECK(x->nrows == y->nrows, "b200_mv_axpby: row counts differ");
		B200_CHECK(start[0] >= 0 && end[0] <= x->ncols, "b200_mv_axpby: x range [%d,%d) outside %d columns",
		           start[0], end[0], x->ncols);
		if (x == y) {
			// same multi-vector: ranges must not overlap unless identical (reference uses
			// disjoint column copies, src/ops_orth.c:70,302)
			const bool overlap = start[0] < end[1] && start[1] < end[0];
			B200_CHECK(!overlap || start[0] == start[1], "b200_mv_axpby: overlapping column ranges on one multi-vector");
		}
	}
	return b200k_axpby(y->nrows, k, alpha, x ? x->d + start[0] : nullptr, x ? x->ld : 0, beta,
	                   y->d + start[1], y->ld);
}
