// K5: fused BlockPCG step kernels.  Replaces the per-column MultiVecAxpby / 'D' inner-product
// chain of the reference's BlockPCG (src/ops_lin_sol.c:256-405: ~6 single-column BLAS-1
// calls per column per iteration, every scalar on the host) by three streaming kernels per
// CG iteration that work on the whole column block at once, keep rho/alpha/beta, the residual
// norms, the per-column convergence masks and the iteration counter in HBM, and fold the
// global reductions into the streaming kernels (last-CTA reduction, b200_reduce.cuh).
// Nothing crosses to the host inside the CG loop.
//
// Per iteration, per column block k (algorithmic HBM traffic):
//   update_p : read r,p      write p          3 * 8nk
//   SpMM     : matrix + read p, write w       nnz*12 + 2 * 8nk
//   ptw      : read p,w                       2 * 8nk     (+ shift: read z, write w)
//   update_xr: read p,w,x,r  write x,r        6 * 8nk
#include "b200_reduce.cuh"

extern "C" int b200k_bpcg_state(int k, b200_bpcg_state *st)
{
	const int chunks_max = g_b200.num_sms * 4 + 8;
	const size_t dbl = (size_t)6 * k + (size_t)chunks_max * 2 * k;
	const size_t bytes = sizeof(double) * dbl + sizeof(int) * ((size_t)k + 8) + 64;
	char *base = (char *)b200_scratch(4, bytes);
	if (!base) return 1;
	double *d = (double *)base;
	st->k = k;
	st->norm_b = d; st->rho1 = d + k; st->rho2 = d + 2 * k; st->ptw = d + 3 * k;
	st->init_res = d + 4 * k; st->last_res = d + 5 * k;
	st->partials = d + 6 * k;
	int *ip = (int *)(d + dbl);
	st->active = ip; st->counters = ip + k; st->tickets = (unsigned *)(ip + k + 4);
	B200_CUDA(cudaMemsetAsync(st->counters, 0, sizeof(int) * 8, g_b200.stream));
	return 0;
}

// ---------------------------------------------------------------------------- begin
// r <- b - r ; acc0 = r.r ; acc1 = b.b (rel only)
template <int CPT>
__global__ void __launch_bounds__(RED_THREADS)
bpcg_begin_kernel(long long n, int k, long long rows_per_chunk, const double *__restrict__ b, int ldb, double *__restrict__ r, int ldr,
                  double tol, int rel, b200_bpcg_state st)
{
	extern __shared__ double sm[];
	const int cx = blockDim.x, ry = blockDim.y;
	const long long r_begin = (long long)blockIdx.x * rows_per_chunk;
	long long r_end = r_begin + rows_per_chunk; if (r_end > n) r_end = n;
	double acc[2][CPT];
#pragma unroll
	for (int i = 0; i < CPT; ++i) { acc[0][i] = 0.0; acc[1][i] = 0.0; }
#pragma unroll 4
	for (long long row = r_begin + threadIdx.y; row < r_end; row += ry) {
#pragma unroll
		for (int i = 0; i < CPT; ++i) {
			const int c = threadIdx.x + i * cx;
			if (c < k) {
				const double bv = b[(size_t)row * ldb + c];
				const double rv = bv - r[(size_t)row * ldr + c];
				r[(size_t)row * ldr + c] = rv;
				acc[0][i] = fma(rv, rv, acc[0][i]);
				acc[1][i] = fma(bv, bv, acc[1][i]);
			}
		}
	}
	if (!red_block_and_elect<CPT, 2>(acc, k, sm, st.partials, st.tickets)) return;
	const int tid = threadIdx.y * cx + threadIdx.x, warp = tid >> 5, lane = tid & 31;
	__shared__ int n_active;
	if (tid == 0) n_active = 0;
	__syncthreads();
	for (int c = warp; c < k; c += RED_THREADS / 32) {
		const double rr = red_total<2>(st.partials, gridDim.x, k, 0, c);
		const double bb = red_total<2>(st.partials, gridDim.x, k, 1, c);
		if (lane == 0) {
			const double nb = rel ? sqrt(bb) : 1.0;
			const double res = sqrt(rr);
			st.norm_b[c] = nb; st.rho2[c] = rr; st.rho1[c] = rr;
			st.init_res[c] = res; st.last_res[c] = res;
			const int act = (res > tol * nb) ? 1 : 0;      // reference src/ops_lin_sol.c:239-247
			st.active[c] = act;
			if (act) atomicAdd(&n_active, 1);
		}
	}
	__syncthreads();
	if (tid == 0) { st.counters[0] = n_active; st.counters[1] = 0; }
}

// ------------------------------------------------------------------------- update_p
__global__ void bpcg_update_p_kernel(long long n, int k, int rows_per_cta, const double *__restrict__ r, int ldr, double *__restrict__ p,
                                     int ldp, int first, b200_bpcg_state st)
{
	if (st.counters[0] == 0) return;
	const long long r0 = (long long)blockIdx.x * rows_per_cta;
	long long nr = n - r0; if (nr > rows_per_cta) nr = rows_per_cta;
	const int total = (int)nr * k;
#pragma unroll 4
	for (int i = threadIdx.x; i < total; i += blockDim.x) {
		const int rr = i / k, c = i - rr * k;
		if (!st.active[c]) continue;
		const double rv = r[(size_t)(r0 + rr) * ldr + c];
		double *pp = p + (size_t)(r0 + rr) * ldp + c;
		// reference src/ops_lin_sol.c:271-284: p = r + beta p, beta = rho2/rho1 (0 on iteration 0)
		*pp = first ? rv : fma(st.rho2[c] / st.rho1[c], *pp, rv);
	}
}

// ------------------------------------------------------------------------------ ptw
template <int CPT>
__global__ void __launch_bounds__(RED_THREADS)
bpcg_ptw_kernel(long long n, int k, long long rows_per_chunk, const double *p, int ldp, double *__restrict__ w, int ldw,
                double shift, const double *z, int ldz, b200_bpcg_state st)
{
	if (st.counters[0] == 0) return;
	extern __shared__ double sm[];
	const int cx = blockDim.x, ry = blockDim.y;
	const long long r_begin = (long long)blockIdx.x * rows_per_chunk;
	long long r_end = r_begin + rows_per_chunk; if (r_end > n) r_end = n;
	double acc[1][CPT];
#pragma unroll
	for (int i = 0; i < CPT; ++i) acc[0][i] = 0.0;
#pragma unroll 4
	for (long long row = r_begin + threadIdx.y; row < r_end; row += ry) {
#pragma unroll
		for (int i = 0; i < CPT; ++i) {
			const int c = threadIdx.x + i * cx;
			if (c < k) {
				double wv = w[(size_t)row * ldw + c];
				if (z) {       // w = (A + shift B) p, reference src/ops_eig_sol_gcg.c:63-96
					wv = fma(shift, z[(size_t)row * ldz + c], wv);
					w[(size_t)row * ldw + c] = wv;
				}
				acc[0][i] = fma(p[(size_t)row * ldp + c], wv, acc[0][i]);
			}
		}
	}
	if (!red_block_and_elect<CPT, 1>(acc, k, sm, st.partials, st.tickets)) return;
	const int tid = threadIdx.y * cx + threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int c = warp; c < k; c += RED_THREADS / 32) {
		const double s = red_total<1>(st.partials, gridDim.x, k, 0, c);
		if (lane == 0 && st.active[c]) st.ptw[c] = s;
	}
}

// ------------------------------------------------------------------------ update_xr
template <int CPT>
__global__ void __launch_bounds__(RED_THREADS)
bpcg_update_xr_kernel(long long n, int k, long long rows_per_chunk, const double *__restrict__ p, int ldp, const double *__restrict__ w,
                      int ldw, double *__restrict__ x, int ldx, double *__restrict__ r, int ldr, double rate, double tol,
                      b200_bpcg_state st)
{
	if (st.counters[0] == 0) return;
	extern __shared__ double sm[];
	const int cx = blockDim.x, ry = blockDim.y;
	const long long r_begin = (long long)blockIdx.x * rows_per_chunk;
	long long r_end = r_begin + rows_per_chunk; if (r_end > n) r_end = n;
	double acc[1][CPT], alpha[CPT];
	bool act[CPT];
#pragma unroll
	for (int i = 0; i < CPT; ++i) {
		const int c = threadIdx.x + i * cx;
		acc[0][i] = 0.0;
		act[i] = (c < k) && st.active[c];
		alpha[i] = act[i] ? st.rho2[c] / st.ptw[c] : 0.0;       // reference src/ops_lin_sol.c:328
	}
#pragma unroll 4
	for (long long row = r_begin + threadIdx.y; row < r_end; row += ry) {
#pragma unroll
		for (int i = 0; i < CPT; ++i) {
			if (act[i]) {
				const int c = threadIdx.x + i * cx;
				const double pv = p[(size_t)row * ldp + c], wv = w[(size_t)row * ldw + c];
				double *xp = x + (size_t)row * ldx + c, *rp = r + (size_t)row * ldr + c;
				*xp = fma(alpha[i], pv, *xp);
				const double rv = fma(-alpha[i], wv, *rp);
				*rp = rv;
				acc[0][i] = fma(rv, rv, acc[0][i]);
			}
		}
	}
	if (!red_block_and_elect<CPT, 1>(acc, k, sm, st.partials, st.tickets)) return;
	const int tid = threadIdx.y * cx + threadIdx.x, warp = tid >> 5, lane = tid & 31;
	__shared__ int n_active;
	if (tid == 0) n_active = 0;
	__syncthreads();
	for (int c = warp; c < k; c += RED_THREADS / 32) {
		const double rr = red_total<1>(st.partials, gridDim.x, k, 0, c);
		if (lane == 0 && st.active[c]) {
			st.rho1[c] = st.rho2[c];
			st.rho2[c] = rr;
			const double res = sqrt(rr);
			st.last_res[c] = res;
			// reference src/ops_lin_sol.c:383-394: stay active while res > rate*res0 AND res > tol*||b||
			const int still = (res > rate * st.init_res[c]) && (res > tol * st.norm_b[c]);
			st.active[c] = still;
			if (still) atomicAdd(&n_active, 1);
		}
	}
	__syncthreads();
	if (tid == 0) { st.counters[0] = n_active; st.counters[1] += 1; }
}

// ------------------------------------------------------------------------- launchers
#define BPCG_DISPATCH_CPT(k, CALL)                         \
	do {                                                   \
		if ((k) <= 32) { constexpr int CPT = 1; CALL; }    \
		else if ((k) <= 64) { constexpr int CPT = 2; CALL; } \
		else { constexpr int CPT = 4; CALL; }              \
	} while (0)

extern "C" int b200k_bpcg_begin(long long n, const b200_bpcg_state *st, const double *b, int ldb,
                                double *r, int ldr, double tol, int rel)
{
	const int k = st->k;
	B200_CHECK(k >= 1 && k <= 128, "BlockPCG: %d columns (1..128 supported per block)", k);
	B200Prof prof(B200_PROF_BPCG, 24.0 * n * k, 5.0 * n * k);
	const RedGeom g = red_geometry(n, k);
	const size_t smem = sizeof(double) * (size_t)g.ry * k;
	BPCG_DISPATCH_CPT(k, (bpcg_begin_kernel<CPT><<<g.chunks, dim3(g.cx, g.ry), smem, g_b200.stream>>>(
		n, k, g.rows_per_chunk, b, ldb, r, ldr, tol, rel, *st)));
	B200_KERNEL_CHECK();
	return 0;
}

extern "C" int b200k_bpcg_update_p(long long n, const b200_bpcg_state *st, const double *r, int ldr,
                                   double *p, int ldp, int first)
{
	const int k = st->k;
	B200Prof prof(B200_PROF_BPCG, 24.0 * n * k, 2.0 * n * k);
	int rows = 4096 / k; if (rows < 1) rows = 1;
	bpcg_update_p_kernel<<<(unsigned)((n + rows - 1) / rows), 256, 0, g_b200.stream>>>(n, k, rows, r, ldr, p, ldp,
	                                                                                 first, *st);
	B200_KERNEL_CHECK();
	return 0;
}

extern "C" int b200k_bpcg_ptw(long long n, const b200_bpcg_state *st, const double *p, int ldp,
                              double *w, int ldw, double shift, const double *z, int ldz)
{
	const int k = st->k;
	B200Prof prof(B200_PROF_BPCG, (z ? 32.0 : 16.0) * n * k, (z ? 4.0 : 2.0) * n * k);
	const RedGeom g = red_geometry(n, k);
	const size_t smem = sizeof(double) * (size_t)g.ry * k;
	BPCG_DISPATCH_CPT(k, (bpcg_ptw_kernel<CPT><<<g.chunks, dim3(g.cx, g.ry), smem, g_b200.stream>>>(
		n, k, g.rows_per_chunk, p, ldp, w, ldw, shift, z, ldz, *st)));
	B200_KERNEL_CHECK();
	return 0;
}

extern "C" int b200k_bpcg_update_xr(long long n, const b200_bpcg_state *st, const double *p, int ldp,
                                    const double *w, int ldw, double *x, int ldx, double *r, int ldr,
                                    double rate, double tol)
{
	const int k = st->k;
	B200Prof prof(B200_PROF_BPCG, 48.0 * n * k, 6.0 * n * k);
	const RedGeom g = red_geometry(n, k);
	const size_t smem = sizeof(double) * (size_t)g.ry * k;
	BPCG_DISPATCH_CPT(k, (bpcg_update_xr_kernel<CPT><<<g.chunks, dim3(g.cx, g.ry), smem, g_b200.stream>>>(
		n, k, g.rows_per_chunk, p, ldp, w, ldw, x, ldx, r, ldr, rate, tol, *st)));
	B200_KERNEL_CHECK();
	return 0;
}
