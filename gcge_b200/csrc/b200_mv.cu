// Device multi-vector store (row-major n x ld) and the HBM-bound elementwise kernels:
// axpby / scale / copy on column ranges, column-major <-> row-major staging, RNG fill.
#include "b200_internal.h"
#include "b200_stream.cuh"

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
	if (g_b200.initialised && g_b200.pending) b200k_pending_flush();
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

static int axpby_stream(long long n, int k, double alpha, const double *x, int ldx, double beta, double *y, int ldy);

int b200k_axpby(long long n, int k, double alpha, const double *x, int ldx, double beta, double *y, int ldy)
{
	if (n <= 0 || k <= 0) return 0;
	const bool has_x = (x != nullptr);
	const bool has_y = (beta != 0.0);
	if (!has_x && has_y && beta == 1.0) return 0;
	// with an x operand: the streaming geometry (16-byte accesses, several rows of every stream in flight); the kernel
	// below (8-byte accesses, an integer division per element: 3.0 TB/s in the solve) keeps the scale-only calls
	if (has_x) return axpby_stream(n, k, alpha, x, ldx, beta, y, ldy);
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

// ---- batching of narrow axpby calls ----------------------------------------------------------------------------
// The reference's BlockPCG updates x, r and p ONE COLUMN PER CALL, each with its own alpha / beta
// (src/ops_lin_sol.c:256-405): through OPS_B200_Set that is ~84 000 single-column launches per solve at n = 2 M,
// nev = 100, each touching 8 bytes of every k*8-byte row -- 79 % of the time of the reference's GCG over the device
// slots (profiles/tiers_prof_r2g.log).  The slot has no host-visible result, so a narrow call is DEFERRED: calls on
// adjacent columns of the same two blocks are collected (alpha and beta per column) and launched as one kernel on the
// streaming geometry -- when the next call does not extend the batch, when the batch is full, or at the start of any
// other entry point (B200_REQUIRE_INIT).  Per element the arithmetic is the unbatched kernel's, so results are bit-identical.
constexpr int AXB_MAX_COLS = 64;              // columns per batch (alpha/beta travel as kernel parameters)
constexpr int AXB_MAX_CALL = 8;               // widest call that is deferred
constexpr int AXB_SLOTS = 4;                  // batches open at a time (BlockPCG interleaves x += a p and r -= a w column by column)
struct AxpbyCols { double alpha[AXB_MAX_COLS], beta[AXB_MAX_COLS]; };
struct AxpbyBatch {
	const double *x; double *y; int ldx, ldy, count; long long n;
	AxpbyCols c;
};
static AxpbyBatch g_axb[AXB_SLOTS];

// LOADY == false: every beta is 0, y is only written (column copies and overwrites: two streams instead of three)
template <int VEC, bool LOADY>
__global__ void __launch_bounds__(ST_THREADS)
axpby_cols_kernel(long long n, int k, StreamGeom g, const __grid_constant__ AxpbyCols prm, const double *x, int ldx,
                  double *y, int ldy)
{
	const StreamThread t = stream_thread<VEC>(g);
	if (!t.active) return;
	double a[VEC], b[VEC];
#pragma unroll
	for (int i = 0; i < VEC; ++i) { a[i] = prm.alpha[t.c + i]; b[i] = prm.beta[t.c + i]; }
	const long long r_begin = (long long)blockIdx.x * g.rows_per_chunk;
	long long r_end = r_begin + g.rows_per_chunk; if (r_end > n) r_end = n;
	for (long long row0 = r_begin + t.rl; row0 < r_end; row0 += (long long)ST_UNROLL * g.rp) {
		StV<VEC> xv[ST_UNROLL], yv[ST_UNROLL];
#pragma unroll
		for (int u = 0; u < ST_UNROLL; ++u) {
			const long long row = row0 + (long long)u * g.rp;
			if (row < r_end) {
				xv[u] = st_ld<VEC>(x + (size_t)row * ldx + t.c);
				if (LOADY) yv[u] = st_ld<VEC>(y + (size_t)row * ldy + t.c);
			}
		}
#pragma unroll
		for (int u = 0; u < ST_UNROLL; ++u) {
			const long long row = row0 + (long long)u * g.rp;
			if (row < r_end) {
#pragma unroll
				for (int i = 0; i < VEC; ++i) {
					// as axpby_kernel: beta == 0 overwrites (NaN-safe), beta == 1 does not multiply, then one fma
					double v = 0.0;
					if (LOADY && b[i] != 0.0) { v = yv[u].v[i]; if (b[i] != 1.0) v *= b[i]; }
					yv[u].v[i] = fma(a[i], xv[u].v[i], v);
				}
				st_st<VEC>(y + (size_t)row * ldy + t.c, yv[u]);
			}
		}
	}
}

// y[:, c] = alpha[c] x[:, c] + beta[c] y[:, c] on k <= AXB_MAX_COLS adjacent columns, one launch on the streaming geometry
static int axpby_cols_launch(long long n, int k, const AxpbyCols &c, const double *x, int ldx, double *y, int ldy)
{
	bool load_y = false;
	for (int i = 0; i < k; ++i) load_y = load_y || c.beta[i] != 0.0;
	B200Prof prof(B200_PROF_AXPBY, (load_y ? 24.0 : 16.0) * n * k, 2.0 * n * k);
	const StreamGeom g = stream_geometry(n, k, stream_aligned16(x, ldx) && stream_aligned16(y, ldy));
	if (load_y) ST_DISPATCH_VEC(g, (axpby_cols_kernel<VEC, true><<<g.chunks, ST_THREADS, 0, g_b200.stream>>>(n, k, g, c, x, ldx, y, ldy)));
	else        ST_DISPATCH_VEC(g, (axpby_cols_kernel<VEC, false><<<g.chunks, ST_THREADS, 0, g_b200.stream>>>(n, k, g, c, x, ldx, y, ldy)));
	B200_KERNEL_CHECK();
	return 0;
}

static int axpby_stream(long long n, int k, double alpha, const double *x, int ldx, double beta, double *y, int ldy)
{
	AxpbyCols c;
	for (int i = 0; i < AXB_MAX_COLS; ++i) { c.alpha[i] = alpha; c.beta[i] = beta; }
	for (int c0 = 0; c0 < k; c0 += AXB_MAX_COLS) {
		const int kc = k - c0 < AXB_MAX_COLS ? k - c0 : AXB_MAX_COLS;
		if (axpby_cols_launch(n, kc, c, x + c0, ldx, y + c0, ldy)) return 1;
	}
	return 0;
}

static int axpby_batch_launch(AxpbyBatch &B)
{
	const int k = B.count;
	B.count = 0;
	if (k <= 0) return 0;
	return axpby_cols_launch(B.n, k, B.c, B.x, B.ldx, B.y, B.ldy);
}

// the open batches are independent of one another (axpby_defer keeps them so): any order is the program's order
extern "C" int b200k_pending_flush(void)
{
	g_b200.pending = 0;
	int rc = 0;
	for (int i = 0; i < AXB_SLOTS; ++i) rc |= axpby_batch_launch(g_axb[i]);
	return rc;
}

// do the column blocks (a: ca columns, leading dimension lda; b likewise; n rows each) share an element?  Blocks of
// different allocations never do; blocks of one allocation start in its row 0, so with equal leading dimensions the
// answer is whether their column intervals meet; anything else that overlaps in memory counts as shared
static bool axpby_blocks_meet(const double *a, int ca, int lda, const double *b, int cb, int ldb, long long n)
{
	if (a + (size_t)(n - 1) * lda + ca <= b || b + (size_t)(n - 1) * ldb + cb <= a) return false;
	if (lda != ldb) return true;
	return a < b + cb && b < a + ca;
}

// take a narrow call into a batch; *taken = 0: the caller launches it itself (everything deferred has been launched)
static int axpby_defer(double alpha, const double *x, int ldx, double beta, double *y, int ldy, long long n, int k, int *taken)
{
	*taken = 0;
	int slot = -1;
	for (int i = 0; i < AXB_SLOTS && slot < 0; ++i) {
		const AxpbyBatch &B = g_axb[i];
		if (B.count > 0 && x == B.x + B.count && y == B.y + B.count && ldx == B.ldx && ldy == B.ldy && n == B.n &&
		    B.count + k <= AXB_MAX_COLS) slot = i;
	}
	// the call reads x and (if beta != 0) y and writes y: what it touches must not meet what any OTHER open batch writes,
	// and what it writes must not meet what another batch reads -- and x must not meet y inside its own batch (a later
	// column of x could be an earlier column of y: column copies inside one multi-vector go one by one in the reference,
	// src/ops_orth.c:70,302)
	const double *bx = slot >= 0 ? g_axb[slot].x : x; const double *by = slot >= 0 ? g_axb[slot].y : y;
	const int bc = (slot >= 0 ? g_axb[slot].count : 0) + k;
	bool clash = axpby_blocks_meet(bx, bc, ldx, by, bc, ldy, n);
	for (int i = 0; i < AXB_SLOTS && !clash; ++i) {
		const AxpbyBatch &B = g_axb[i];
		if (i == slot || B.count == 0) continue;
		clash = axpby_blocks_meet(x, k, ldx, B.y, B.count, B.ldy, B.n > n ? B.n : n) ||
		        axpby_blocks_meet(y, k, ldy, B.y, B.count, B.ldy, B.n > n ? B.n : n) ||
		        axpby_blocks_meet(y, k, ldy, B.x, B.count, B.ldx, B.n > n ? B.n : n);
	}
	if (clash) {
		if (b200k_pending_flush()) return 1;
		if (axpby_blocks_meet(x, k, ldx, y, k, ldy, n)) return 0;
		slot = -1;
	}
	if (slot < 0) {
		for (int i = 0; i < AXB_SLOTS && slot < 0; ++i) if (g_axb[i].count == 0) slot = i;
		if (slot < 0) { if (b200k_pending_flush()) return 1; slot = 0; }
		AxpbyBatch &B = g_axb[slot];
		B.x = x; B.y = y; B.ldx = ldx; B.ldy = ldy; B.n = n; B.count = 0;
	}
	AxpbyBatch &B = g_axb[slot];
	for (int i = 0; i < k; ++i) { B.c.alpha[B.count + i] = alpha; B.c.beta[B.count + i] = beta; }
	B.count += k;
	g_b200.pending = 1;
	*taken = 1;
	return 0;
}

extern "C" int b200_mv_axpby(double alpha, const b200_mv *x, double beta, b200_mv *y,
                             const int *start, const int *end)
{
	if (!g_b200.initialised) { int rc_ = b200_init(-1); if (rc_) return rc_; }      // (no flush: this call may extend the batch)
	if (!(y && start && end) && g_b200.pending && b200k_pending_flush()) return 1;
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
	if (x && x != y && k <= AXB_MAX_CALL && !b200_opt(B200_OPT_NO_AXPBY_BATCH)) {
		int taken = 0;
		if (axpby_defer(alpha, x->d + start[0], x->ld, beta, y->d + start[1], y->ld, y->nrows, k, &taken)) return 1;
		if (taken) return 0;
	}
	if (g_b200.pending && b200k_pending_flush()) return 1;
	return b200k_axpby(y->nrows, k, alpha, x ? x->d + start[0] : nullptr, x ? x->ld : 0, beta,
	                   y->d + start[1], y->ld);
}
