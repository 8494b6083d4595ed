// K7: on-device eigen-solve of the projected Rayleigh-Ritz problem (N <= 3*nev, a few hundred).
// Replaces the host dsyevx call of the reference (src/ops_eig_sol_gcg.c:1201-1204) so the
// projected matrix and the Ritz coefficients never leave HBM.
//
// Parallel-order cyclic two-sided Jacobi in one cooperative kernel.  N (made even by a dummy
// index) "players" meet in N-1 rounds per sweep (circle method): in round r player N-1 meets
// r, and (r+k) meets (r-k) mod N-1.  The matrix is kept in POSITION space: the two members of
// pair k sit at positions 2k and 2k+1, so every 2x2 block a thread rotates is contiguous; the
// results are scattered to the positions the players take in the NEXT round (double
// buffered), which costs nothing extra and needs one grid-wide barrier per round.
// Rotation angles are recomputed by each thread from the two diagonal 2x2 blocks it needs
// (a few flops) instead of being broadcast.  Rotations are skipped under the relative
// criterion |a_pq| <= eps*sqrt(|a_pp a_qq|); a sweep that applies no rotation ends the
// iteration.  All decisions are integer counts, so the result is run-to-run deterministic.
#include "b200_internal.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ int jac_player(int pos, int r, int N)
{
	if (pos == 0) return N - 1;
	if (pos == 1) return r;
	const int k = pos >> 1, M = N - 1;       // r in [0,M), k in [1,N/2)
	int v = (pos & 1) ? r - k : r + k;
	if (v < 0) v += M;
	if (v >= M) v -= M;
	return v;
}
__device__ __forceinline__ int jac_slot(int player, int r, int N)
{
	const int M = N - 1;
	if (player == M) return 0;
	int d = player - r; if (d < 0) d += M;
	if (d == 0) return 1;
	if (d <= N / 2 - 1) return 2 * d;
	return 2 * (M - d) + 1;
}

__device__ __forceinline__ bool jac_rotation(double app, double aqq, double apq, double &c, double &s)
{
	const double eps = 2.220446049250313e-16;
	const double thr = eps * sqrt(fabs(app) * fabs(aqq));
	if (fabs(apq) <= thr || fabs(apq) < 1e-300) { c = 1.0; s = 0.0; return false; }
	const double tau = (aqq - app) / (2.0 * apq);
	const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
	c = 1.0 / sqrt(1.0 + t * t);
	s = t * c;
	return true;
}

// a0/a1: N x N position-space matrices (row-major, ld N); v0/v1: n x N eigenvector
// accumulators (row i = original coordinate, column = position).  On exit *sweeps = number
// of sweeps done, w ascending, z[i*ldz + j] = component i of eigenvector j.
__global__ void __launch_bounds__(256)
syev_jacobi_kernel(int n, int N, double *a0, double *a1, double *v0, double *v1, double *w, double *z, int ldz,
                   int max_sweeps, int *rot_count, int *sweeps_out)
{
	cg::grid_group grid = cg::this_grid();
	const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
	const int gsz = gridDim.x * blockDim.x;
	const int m = N / 2;
	double *cur = a0, *nxt = a1, *vc = v0, *vn = v1;
	int sweep = 0;
	for (; sweep < max_sweeps; ++sweep) {
		for (int r = 0; r < N - 1; ++r) {
			const int rn = (r + 1 == N - 1) ? 0 : r + 1;      // next round (wraps into the next sweep)
			// ---- A' = J^T A J, one thread per (pair a, pair b) block
			for (int idx = gtid; idx < m * m; idx += gsz) {
				const int a = idx / m, b = idx - a * m;
				const double *ra0 = cur + (size_t)(2 * a) * N, *ra1 = ra0 + N;
				const double *rb0 = cur + (size_t)(2 * b) * N, *rb1 = rb0 + N;
				double ca, sa, cb, sb;
				const bool rot = jac_rotation(ra0[2 * a], ra1[2 * a + 1], ra0[2 * a + 1], ca, sa);
				jac_rotation(rb0[2 * b], rb1[2 * b + 1], rb0[2 * b + 1], cb, sb);
				if (b == 0 && rot) atomicAdd(rot_count, 1);
				const double b00 = ra0[2 * b], b01 = ra0[2 * b + 1], b10 = ra1[2 * b], b11 = ra1[2 * b + 1];
				// left: rows (p,q) <- (c p - s q, s p + c q)
				const double l00 = ca * b00 - sa * b10, l01 = ca * b01 - sa * b11;
				const double l10 = sa * b00 + ca * b10, l11 = sa * b01 + ca * b11;
				// right: cols (p,q) <- (c p - s q, s p + c q)
				double n00 = cb * l00 - sb * l01, n01 = sb * l00 + cb * l01;
				double n10 = cb * l10 - sb * l11, n11 = sb * l10 + cb * l11;
				if (a == b) { n01 = 0.0; n10 = 0.0; }           // annihilated exactly
				const int pa = jac_player(2 * a, r, N), qa = jac_player(2 * a + 1, r, N);
				const int pb = jac_player(2 * b, r, N), qb = jac_player(2 * b + 1, r, N);
				const int spa = jac_slot(pa, rn, N), sqa = jac_slot(qa, rn, N);
				const int spb = jac_slot(pb, rn, N), sqb = jac_slot(qb, rn, N);
				nxt[(size_t)spa * N + spb] = n00; nxt[(size_t)spa * N + sqb] = n01;
				nxt[(size_t)sqa * N + spb] = n10; nxt[(size_t)sqa * N + sqb] = n11;
			}
			// ---- V' = V J, one thread per (row i, pair b)
			for (int idx = gtid; idx < n * m; idx += gsz) {
				const int i = idx / m, b = idx - i * m;
				const double *rb0 = cur + (size_t)(2 * b) * N, *rb1 = rb0 + N;
				double cb, sb;
				jac_rotation(rb0[2 * b], rb1[2 * b + 1], rb0[2 * b + 1], cb, sb);
				const double x0 = vc[(size_t)i * N + 2 * b], x1 = vc[(size_t)i * N + 2 * b + 1];
				const int pb = jac_player(2 * b, r, N), qb = jac_player(2 * b + 1, r, N);
				vn[(size_t)i * N + jac_slot(pb, rn, N)] = cb * x0 - sb * x1;
				vn[(size_t)i * N + jac_slot(qb, rn, N)] = sb * x0 + cb * x1;
			}
			grid.sync();
			double *tp = cur; cur = nxt; nxt = tp;
			tp = vc; vc = vn; vn = tp;
		}
		// end of sweep: did anything rotate?  (count was completed before the last grid.sync)
		const int rotated = *((volatile int *)rot_count);
		grid.sync();
		if (gtid == 0) *rot_count = 0;
		grid.sync();
		if (rotated == 0) { ++sweep; break; }
	}
	// players are back in round-0 arrangement.  Rank the diagonal (dummy excluded) and emit.
	for (int idx = gtid; idx < N; idx += gsz) {
		const int pos = idx;
		const int player = jac_player(pos, 0, N);
		if (player >= n) continue;                               // dummy index of an odd problem
		const double d = cur[(size_t)pos * N + pos];
		int rank = 0;
		for (int q = 0; q < N; ++q) {
			const int pl = jac_player(q, 0, N);
			if (pl >= n) continue;
			const double dq = cur[(size_t)q * N + q];
			if (dq < d || (dq == d && pl < player)) ++rank;
		}
		w[rank] = d;
		// stash rank in the (now free) next buffer's first row for the copy-out below
		((int *)nxt)[pos] = rank;
	}
	grid.sync();
	for (int idx = gtid; idx < n * N; idx += gsz) {
		const int i = idx / N, pos = idx - i * N;
		if (jac_player(pos, 0, N) >= n) continue;
		z[(size_t)i * ldz + ((int *)nxt)[pos]] = vc[(size_t)i * N + pos];
	}
	if (gtid == 0) *sweeps_out = sweep;
}

// position-space initialisation: a0[slot(i)][slot(j)] = A[i][j], v0 = identity (same permutation)
__global__ void syev_init_kernel(int n, int N, const double *a, int lda, double *a0, double *v0)
{
	const long long gsz = (long long)gridDim.x * blockDim.x;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)N * N; idx += gsz) {
		const int pi = (int)(idx / N), pj = (int)(idx - (long long)pi * N);
		const int i = jac_player(pi, 0, N), j = jac_player(pj, 0, N);
		a0[idx] = (i < n && j < n) ? a[(size_t)i * lda + j] : 0.0;
		if (i < n) v0[(size_t)i * N + pj] = (i == j) ? 1.0 : 0.0;
	}
}

extern "C" int b200k_syev_jacobi(int n, double *a_dev, int lda, double *w_dev, double *z_dev, int ldz,
                                 int *sweeps_host)
{
	B200_CHECK(n >= 1, "syev: n = %d", n);
	B200Prof prof(B200_PROF_SYEV, 24.0 * n * n, 0.0);
	cudaStream_t st = g_b200.stream;
	int N = (n + 1) & ~1; if (N < 2) N = 2;
	const size_t nn = (size_t)N * N, vn = (size_t)n * N;
	char *base = (char *)b200_scratch(5, sizeof(double) * (2 * nn + 2 * vn) + 64);
	if (!base) return 1;
	double *a0 = (double *)base, *a1 = a0 + nn, *v0 = a1 + nn, *v1 = v0 + vn;
	int *ctr = (int *)(v1 + vn);
	B200_CUDA(cudaMemsetAsync(ctr, 0, 2 * sizeof(int), st));
	syev_init_kernel<<<b200_ceil_div((long long)N * N, 256), 256, 0, st>>>(n, N, a_dev, lda, a0, v0);
	B200_KERNEL_CHECK();
	int per_sm = 0;
	B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, syev_jacobi_kernel, 256, 0));
	B200_CHECK(per_sm >= 1, "syev: kernel does not fit on an SM");
	long long want = ((long long)(N / 2) * (N / 2) + (long long)n * (N / 2) + 255) / 256;
	long long cap = (long long)per_sm * g_b200.num_sms;
	int blocks = (int)(want < cap ? want : cap);
	if (blocks > g_b200.num_sms) blocks = g_b200.num_sms;     // one CTA per SM keeps the barrier cheap
	if (blocks < 1) blocks = 1;
	int max_sweeps = 40;
	int *rot = ctr, *sw = ctr + 1;
	void *args[] = {&n, &N, &a0, &a1, &v0, &v1, &w_dev, &z_dev, &ldz, &max_sweeps, &rot, &sw};
	B200_CUDA(cudaLaunchCooperativeKernel((void *)syev_jacobi_kernel, dim3(blocks), dim3(256), args, 0, st));
	B200_LAUNCHED();
	if (sweeps_host) {
		if (b200k_d2h(sweeps_host, sw, sizeof(int))) return 1;
	}
	return 0;
}

// host-matrix convenience entry (tests, tier A users): see include/gcge_b200.h
extern "C" int b200_dense_syev(int n, const double *a, int lda, double *w, double *z, int ldz, int *sweeps)
{
	B200_REQUIRE_INIT();
	B200_CHECK(n >= 1 && a && w && z && lda >= n && ldz >= n, "b200_dense_syev: bad arguments");
	const size_t nn = (size_t)n * n;
	double *dev = (double *)b200_scratch(3, sizeof(double) * (2 * nn + n));
	if (!dev) return 1;
	double *da = dev, *dz = dev + nn, *dw = dz + nn;
	double *pin = (double *)b200_pinned(1, sizeof(double) * nn);
	if (!pin) return 1;
	B200_CUDA(cudaStreamSynchronize(g_b200.stream));
	// symmetrise from the caller's upper triangle (dsyevx UPLO='U' semantics)
	for (int j = 0; j < n; ++j)
		for (int i = 0; i < n; ++i)
			pin[(size_t)i * n + j] = (i <= j) ? a[(size_t)j * lda + i] : a[(size_t)i * lda + j];
	B200_CUDA(cudaMemcpyAsync(da, pin, sizeof(double) * nn, cudaMemcpyHostToDevice, g_b200.stream));
	if (b200k_syev_jacobi(n, da, n, dw, dz, n, sweeps)) return 1;
	if (b200k_d2h(w, dw, sizeof(double) * n)) return 1;
	if (b200k_d2h(pin, dz, sizeof(double) * nn)) return 1;   // dz row-major (i,j) -> host col-major
	for (int j = 0; j < n; ++j)
		for (int i = 0; i < n; ++i) z[(size_t)j * ldz + i] = pin[(size_t)i * n + j];
	return 0;
}
