// Small device helpers of the GCG driver: residual norms for CheckConvergence and the
// element-wise assembly kernels of the projected (N x N) problem.
#include "b200_stream.cuh"

// ax <- ax - lam[c]*bx ; res[c] = ||ax[:,c]||_2   (reference src/ops_eig_sol_gcg.c:214-224)
template <int VEC>
__global__ void __launch_bounds__(ST_THREADS)
residual_kernel(long long n, int k, StreamGeom g, double *__restrict__ ax, int ldax, const double *__restrict__ bx, int ldbx,
                const double *__restrict__ lam, double *res, double *part, unsigned *ticket, int defer)
{
	const StreamThread t = stream_thread<VEC>(g);
	const long long r_begin = (long long)blockIdx.x * g.rows_per_chunk;
	long long r_end = r_begin + g.rows_per_chunk; if (r_end > n) r_end = n;
	double acc[1][VEC], l[VEC];
#pragma unroll
	for (int i = 0; i < VEC; ++i) { acc[0][i] = 0.0; l[i] = lam[t.c + i]; }
	if (t.active) {
		for (long long row0 = r_begin + t.rl; row0 < r_end; row0 += (long long)ST_UNROLL * g.rp) {
			StV<VEC> av[ST_UNROLL], bv[ST_UNROLL];
#pragma unroll
			for (int u = 0; u < ST_UNROLL; ++u) {
				const long long row = row0 + (long long)u * g.rp;
				if (row < r_end) { av[u] = st_ld<VEC>(ax + (size_t)row * ldax + t.c); bv[u] = st_ld<VEC>(bx + (size_t)row * ldbx + t.c); }
			}
#pragma unroll
			for (int u = 0; u < ST_UNROLL; ++u) {
				const long long row = row0 + (long long)u * g.rp;
				if (row < r_end) {
#pragma unroll
					for (int i = 0; i < VEC; ++i) {
						// lambda*Bx is rounded first, then subtracted: the reference scales Bx (dscal) and
						// then forms Ax - (lambda Bx) (daxpy), src/ops_eig_sol_gcg.c:214-220
						const double v = av[u].v[i] - __dmul_rn(l[i], bv[u].v[i]);
						av[u].v[i] = v;
						acc[0][i] = fma(v, v, acc[0][i]);
					}
					st_st<VEC>(ax + (size_t)row * ldax + t.c, av[u]);
				}
			}
		}
	}
	if (!stream_reduce_and_elect<VEC, 1>(acc, k, g, t, part, ticket)) return;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int c = warp; c < k; c += ST_THREADS / 32) {
		const double s = stream_total<1>(part, gridDim.x, k, 0, c);
		if (lane == 0) res[c] = defer ? s : sqrt(s);      // several ranks: sum first, root after the allreduce
	}
}

__global__ void sqrt_inplace_kernel(int k, double *v)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < k) v[i] = sqrt(v[i]);
}

extern "C" int b200k_residual_norms(long long n, int k, double *ax, int ldax, const double *bx, int ldbx,
                                    const double *lam_dev, double *res_dev)
{
	if (k <= 0) return 0;
	B200_CHECK(k <= 128, "residual norms: %d columns (<=128 per call)", k);
	B200Prof prof(B200_PROF_DOTS, 24.0 * n * k, 4.0 * n * k);
	const StreamGeom g = stream_geometry(n, k, stream_aligned16(ax, ldax) && stream_aligned16(bx, ldbx));
	char *base = (char *)b200_scratch(0, sizeof(double) * (size_t)(g.chunks + 1) * k + 64);
	if (!base) return 1;
	double *part = (double *)base;
	unsigned *ticket = (unsigned *)(part + (size_t)(g.chunks + 1) * k);
	B200_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), g_b200.stream));
	ST_DISPATCH_VEC(g, (residual_kernel<VEC><<<g.chunks, ST_THREADS, 0, g_b200.stream>>>(n, k, g, ax, ldax, bx, ldbx, lam_dev, res_dev, part, ticket, b200_multi() ? 1 : 0)));
	B200_KERNEL_CHECK();
	if (b200_multi()) {
		if (b200k_allreduce_sum(res_dev, (size_t)k)) return 1;
		sqrt_inplace_kernel<<<1, 128, 0, g_b200.stream>>>(k, res_dev);
		B200_KERNEL_CHECK();
	}
	return 0;
}

// ---- small strided matrix helpers (N x N objects; latency-bound, one launch each) ----------
__global__ void copy2d_kernel(int rows, int cols, const double *src, int s_rs, int s_cs, double *dst, int d_rs, int d_cs)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= rows * cols) return;
	const int r = idx / cols, c = idx - r * cols;
	dst[(size_t)r * d_rs + (size_t)c * d_cs] = src[(size_t)r * s_rs + (size_t)c * s_cs];
}
extern "C" int b200k_copy2d(int rows, int cols, const double *src, int s_rs, int s_cs, double *dst, int d_rs, int d_cs)
{
	if (rows <= 0 || cols <= 0) return 0;
	copy2d_kernel<<<b200_ceil_div((long long)rows * cols, 256), 256, 0, g_b200.stream>>>(rows, cols, src, s_rs, s_cs, dst, d_rs, d_cs);
	B200_KERNEL_CHECK();
	return 0;
}

__global__ void set_diag_kernel(int n, double *a, int lda, const double *d, double shift)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	a[(size_t)i * lda + i] = (d ? d[i] : a[(size_t)i * lda + i]) + shift;
}
extern "C" int b200k_set_diag(int n, double *a, int lda, const double *d, double shift)
{
	if (n <= 0) return 0;
	set_diag_kernel<<<b200_ceil_div(n, 128), 128, 0, g_b200.stream>>>(n, a, lda, d, shift);
	B200_KERNEL_CHECK();
	return 0;
}

__global__ void zero_rows_kernel(int nidx, const int *idx, int cols, double *a, int rs, int cs)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nidx * cols) return;
	const int i = t / cols, c = t - i * cols;
	a[(size_t)idx[i] * rs + (size_t)c * cs] = 0.0;
}
extern "C" int b200k_zero_rows(int nidx, const int *idx_dev, int cols, double *a, int rs, int cs)
{
	if (nidx <= 0 || cols <= 0) return 0;
	zero_rows_kernel<<<b200_ceil_div((long long)nidx * cols, 256), 256, 0, g_b200.stream>>>(nidx, idx_dev, cols, a, rs, cs);
	B200_KERNEL_CHECK();
	return 0;
}

__global__ void gather_cols_kernel(int rows, int nidx, const int *idx, const double *src, int s_rs, int s_cs,
                                   double *dst, int d_rs, int d_cs)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= rows * nidx) return;
	const int r = t / nidx, i = t - r * nidx;
	dst[(size_t)r * d_rs + (size_t)i * d_cs] = src[(size_t)r * s_rs + (size_t)idx[i] * s_cs];
}
extern "C" int b200k_gather_cols(int rows, int nidx, const int *idx_dev, const double *src, int s_rs, int s_cs,
                                 double *dst, int d_rs, int d_cs)
{
	if (rows <= 0 || nidx <= 0) return 0;
	gather_cols_kernel<<<b200_ceil_div((long long)rows * nidx, 256), 256, 0, g_b200.stream>>>(rows, nidx, idx_dev, src, s_rs, s_cs, dst, d_rs, d_cs);
	B200_KERNEL_CHECK();
	return 0;
}

__global__ void symmetrize_upper_kernel(int n, double *a, int lda)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= n * n) return;
	const int i = idx / n, j = idx - i * n;
	if (i > j) a[(size_t)i * lda + j] = a[(size_t)j * lda + i];
}
extern "C" int b200k_symmetrize_upper(int n, double *a, int lda)
{
	if (n <= 1) return 0;
	symmetrize_upper_kernel<<<b200_ceil_div((long long)n * n, 256), 256, 0, g_b200.stream>>>(n, a, lda);
	B200_KERNEL_CHECK();
	return 0;
}
