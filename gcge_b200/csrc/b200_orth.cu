// K6 (panel part): the self-orthogonalisation recurrence of the reference's OrthSelf
// (src/ops_orth.c:45-118) carried out on the k x k Gram matrix instead of on the n x k
// panel.  OrthSelf walks the columns one by one -- norm, scale, project out of the later
// columns -- and every step streams the panel through memory.  The same recurrence applied
// to G = X^T B X is a right-looking Cholesky factorisation; it yields the k x k map T with
// X_orth = X T, so the panel is read once for G and once for the update.  The drop rule is
// the reference's: r_k < orth_zero_tol => the last live column is swapped in and the block
// shrinks (src/ops_orth.c:64-73).  One CTA; k <= 112 (G and T live in shared memory).
#include "b200_internal.h"

__global__ void __launch_bounds__(256)
chol_drop_kernel(int k, double *g, double zero_tol, double *t, int *n_live_out, const double *scale_in,
                 double *scale_out)
{
	extern __shared__ double sm[];
	const int S = k + 1;
	double *G = sm, *T = sm + (size_t)k * S, *cv = T + (size_t)k * S, *sc = cv + k;
	const int tid = threadIdx.x, nt = blockDim.x;
	for (int i = tid; i < k; i += nt) sc[i] = scale_in ? scale_in[i] : 1.0;
	for (int i = tid; i < k * k; i += nt) {
		const int r = i / k, c = i - r * k;
		G[r * S + c] = g[i];
		T[r * S + c] = (r == c) ? 1.0 : 0.0;
	}
	__syncthreads();
	int pos = 0, n_live = k;
	while (pos < n_live) {
		const double gkk = G[pos * S + pos];
		const double rk = gkk > 0.0 ? sqrt(gkk) : 0.0;
		// sc[pos] is the factor by which earlier passes already scaled this column up, so
		// rk*sc[pos] is its remaining norm in the caller's original scaling
		if (rk * sc[pos] < zero_tol) {
			const int last = n_live - 1;
			if (pos < last) {
				__syncthreads();
				if (tid == 0) { const double a = sc[pos]; sc[pos] = sc[last]; sc[last] = a; }
				for (int i = tid; i < k; i += nt) {      // swap rows pos,last of G
					const double a = G[pos * S + i]; G[pos * S + i] = G[last * S + i]; G[last * S + i] = a;
				}
				__syncthreads();
				for (int i = tid; i < k; i += nt) {      // swap columns of G and of T
					double a = G[i * S + pos]; G[i * S + pos] = G[i * S + last]; G[i * S + last] = a;
					a = T[i * S + pos]; T[i * S + pos] = T[i * S + last]; T[i * S + last] = a;
				}
			}
			--n_live;
			__syncthreads();
			continue;
		}
		const double inv = 1.0 / rk;
		__syncthreads();                                 // everyone has read G[pos][pos]
		for (int i = tid; i < k; i += nt) {
			T[i * S + pos] *= inv;
			if (i != pos) { G[pos * S + i] *= inv; G[i * S + pos] *= inv; }
		}
		if (tid == 0) { G[pos * S + pos] = 1.0; sc[pos] *= rk; }
		__syncthreads();
		const int m = n_live - pos - 1;
		for (int i = tid; i < m; i += nt) cv[i] = G[pos * S + pos + 1 + i];     // q^T B x_j
		__syncthreads();
		for (int i = tid; i < k * m; i += nt) {          // T[:, j] -= T[:, pos] c_j
			const int r = i / m, j = i - r * m;
			T[r * S + pos + 1 + j] -= T[r * S + pos] * cv[j];
		}
		for (int i = tid; i < m * m; i += nt) {          // G[i][j] -= c_i c_j
			const int r = i / m, j = i - r * m;
			G[(pos + 1 + r) * S + pos + 1 + j] -= cv[r] * cv[j];
		}
		__syncthreads();
		for (int i = tid; i < k; i += nt)
			if (i != pos) { G[pos * S + i] = 0.0; G[i * S + pos] = 0.0; }
		__syncthreads();
		++pos;
	}
	__syncthreads();
	for (int i = tid; i < k * k; i += nt) {
		const int r = i / k, c = i - r * k;
		t[i] = T[r * S + c];
	}
	if (scale_out) for (int i = tid; i < k; i += nt) scale_out[i] = sc[i];
	if (tid == 0) *n_live_out = n_live;
}

extern "C" int b200k_chol_drop(int k, double *g_dev, double zero_tol, double *t_dev, int *n_live_dev,
                               const double *scale_in, double *scale_out)
{
	B200_CHECK(k >= 1 && k <= 112, "orth panel: %d columns (1..112 supported)", k);
	B200Prof prof(B200_PROF_PANEL, 16.0 * k * k, 2.0 * k * k * k / 3.0);
	const size_t smem = sizeof(double) * ((size_t)2 * k * (k + 1) + 2 * k);
	static bool attr_set = false;
	if (!attr_set) {
		B200_CUDA(cudaFuncSetAttribute(chol_drop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
		attr_set = true;
	}
	chol_drop_kernel<<<1, 256, smem, g_b200.stream>>>(k, g_dev, zero_tol, t_dev, n_live_dev, scale_in, scale_out);
	B200_KERNEL_CHECK();
	return 0;
}

// *out = max over a row-major rows x cols coefficient block of |c[i][j]| * scale[j] (scale == nullptr: 1): one CTA.
// scale[j] is the norm column j had in the caller's scaling before the panels normalised it (chol_drop_kernel's
// scale_out), so the value is the coefficient the reference would have seen on its never-normalised column.
__global__ void __launch_bounds__(256)
absmax_kernel(int rows, int cols, const double *__restrict__ c, const double *__restrict__ scale, double *out)
{
	__shared__ double red[256];
	double m = 0.0;
	const long long count = (long long)rows * cols;
	for (long long i = threadIdx.x; i < count; i += 256) {
		const double v = fabs(c[i]);
		m = fmax(m, scale ? v * scale[i % cols] : v);
	}
	red[threadIdx.x] = m;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
		__syncthreads();
	}
	if (threadIdx.x == 0) *out = red[0];
}

extern "C" int b200k_absmax(int rows, int cols, const double *c_dev, const double *scale_dev, double *out_dev)
{
	absmax_kernel<<<1, 256, 0, g_b200.stream>>>(rows, cols, c_dev, scale_dev, out_dev);
	B200_KERNEL_CHECK();
	return 0;
}

// OrthSelfEVP coefficient block (reference src/ops_orth.c:183-192): t[i*nk + j] = z[i*n + lin_dep + j] / sqrt(w[lin_dep + j]),
// nk = n - lin_dep; z row-major with eigenvector j in column j
__global__ void evp_coef_kernel(int n, int lin_dep, const double *__restrict__ w, const double *__restrict__ z, double *__restrict__ t)
{
	const int nk = n - lin_dep;
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= n * nk) return;
	const int i = idx / nk, j = idx - i * nk;
	t[idx] = z[(size_t)i * n + lin_dep + j] * (1.0 / sqrt(w[lin_dep + j]));
}

extern "C" int b200k_evp_coef(int n, int lin_dep, const double *w_dev, const double *z_dev, double *t_dev)
{
	const int nk = n - lin_dep;
	if (nk <= 0) return 0;
	evp_coef_kernel<<<b200_ceil_div((long long)n * nk, 256), 256, 0, g_b200.stream>>>(n, lin_dep, w_dev, z_dev, t_dev);
	B200_KERNEL_CHECK();
	return 0;
}
