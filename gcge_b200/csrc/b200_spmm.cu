// K1: CSR SpMM  y[:, 0:k] = A x[:, 0:k]  on row-major multi-vector blocks.
//
// Replaces the reference's column-parallel CCS scatter loop (app/app_ccs.c:116-131), which
// streams the whole matrix once PER COLUMN.  Here the matrix is read once for all k columns:
// a group of G lanes owns one row, lane l accumulates columns l, l+G, ... so every gathered
// x row is one contiguous k*8-byte segment (row-major store) and the matrix entry is a
// broadcast load inside the group.  HBM-bound: algorithmic bytes per launch are
//   nnz*(8+4) + (n+1)*4 + 8*n*k (x, read once) + 8*n*k (y, written once)       
//
// Arithmetic: separate multiply and add (no FMA contraction), entries of a row visited in
// ascending column order -- the same operation sequence as the reference's serial scatter,
// so the result is bit-identical to the reference on every input.
#include "b200_internal.h"

// Matrix entries are fetched cooperatively: the G lanes of a row group load G consecutive
// (value, column) pairs with one coalesced request each and hand them round with shuffles,
// instead of every lane issuing the same broadcast load per entry -- the L1 wavefront count
// per entry drops from 2 + ceil(8k/128) to ceil(8k/128) (+2/G), and L1TEX issue, not HBM, is
// what bounds this kernel (profiles/).  Control flow is kept warp-uniform (loop bounds are
// warp maxima) so the shuffles are always fully converged.
template <int G, int CPL>
__global__ void __launch_bounds__(256)
spmm_csr_kernel(int nrows, const int *__restrict__ rp, const int *__restrict__ ci,
                const double *__restrict__ va, const double *x, int ldx, double *y, int ldy, int k,
                const int *__restrict__ gate)
{
	if (gate != nullptr && *gate == 0) return;
	const int gl = threadIdx.x % G;                          // lane inside the row group
	const int groups_per_cta = 256 / G;
	const long long row = (long long)blockIdx.x * groups_per_cta + threadIdx.x / G;
	const bool live = row < nrows;
	const int e0 = live ? __ldg(rp + row) : 0, e1 = live ? __ldg(rp + row + 1) : 0;
	for (int cbase = 0; cbase < k; cbase += G * CPL) {
		double acc[CPL];
		bool on[CPL];
#pragma unroll
		for (int i = 0; i < CPL; ++i) { acc[i] = 0.0; on[i] = live && (cbase + gl + i * G) < k; }
		if (G >= 8) {
			const int nchunk = (e1 - e0 + G - 1) / G;
			const int nchunk_max = __reduce_max_sync(0xffffffffu, nchunk);
			for (int ch = 0; ch < nchunk_max; ++ch) {
				const int eb = e0 + ch * G;
				int cnt = e1 - eb; cnt = cnt < 0 ? 0 : (cnt > G ? G : cnt);
				const int cnt_max = __reduce_max_sync(0xffffffffu, cnt);
				double a_l = 0.0; int c_l = 0;
				if (gl < cnt) { a_l = __ldg(va + eb + gl); c_l = __ldg(ci + eb + gl); }
#pragma unroll 4
				for (int j = 0; j < cnt_max; ++j) {
					const double a = __shfl_sync(0xffffffffu, a_l, j, G);
					const int col = __shfl_sync(0xffffffffu, c_l, j, G);
					if (j < cnt) {
						const double *xr = x + (size_t)col * ldx + cbase + gl;
#pragma unroll
						for (int i = 0; i < CPL; ++i)
							if (on[i]) acc[i] = __dadd_rn(acc[i], __dmul_rn(a, xr[i * G]));
					}
				}
			}
		} else {
			for (int e = e0; e < e1; ++e) {
				const double a = __ldg(va + e);
				const double *xr = x + (size_t)__ldg(ci + e) * ldx + cbase + gl;
#pragma unroll
				for (int i = 0; i < CPL; ++i)
					if (on[i]) acc[i] = __dadd_rn(acc[i], __dmul_rn(a, xr[i * G]));
			}
		}
		if (live) {
			double *yr = y + (size_t)row * ldy + cbase + gl;
#pragma unroll
			for (int i = 0; i < CPL; ++i)
				if (on[i]) yr[i * G] = acc[i];
		}
	}
}

// 128-bit variant (even k): lane l of a row group owns the column PAIR (2l, 2l+1), so one
// LDG.128 per matrix entry covers 2G columns -- half the load and address instructions of the
// scalar kernel for the same bytes.  The scalar kernel is instruction-issue bound (~380 warp
// instructions per 15-entry row at k = 40, profiles/), so the inner loop is kept branch-free:
// lanes past the last column pair redo the last pair (same addresses, no extra wavefront, no
// store), gathers are unconditional (padding entries point at row 0 with a zero value that is
// never accumulated) and only the two accumulate statements are predicated.  Needs 16-byte
// aligned row segments (even column offsets / leading dimensions); the launcher falls back
// to the scalar kernel otherwise.
template <int G>
__global__ void __launch_bounds__(256)
spmm_csr_v2_kernel(int nrows, const int *__restrict__ rp, const int *__restrict__ ci,
                   const double *__restrict__ va, const double *x, int ldx, double *y, int ldy, int k,
                   const int *__restrict__ gate)
{
	if (gate != nullptr && *gate == 0) return;
	const int gl = threadIdx.x % G;
	const int groups_per_cta = 256 / G;
	const long long row = (long long)blockIdx.x * groups_per_cta + threadIdx.x / G;
	const bool live = row < nrows;
	const int e0 = live ? __ldg(rp + row) : 0, e1 = live ? __ldg(rp + row + 1) : 0;
	const size_t ldxb = (size_t)ldx;
	for (int cbase = 0; cbase < k; cbase += 2 * G) {
		int c = cbase + 2 * gl;
		const bool store = live && (c + 1 < k);
		if (c > k - 2) c = k - 2;
		const double *xb = x + c;
		double acc0 = 0.0, acc1 = 0.0;
		// Loads are issued in explicit batches of U (all shuffles, then all gathers, then the
		// in-order accumulation): a warp keeps U independent 128-bit gathers in flight instead
		// of one, which is what this latency-bound loop needs (one request per ~12 cycles per SM
		// before batching, independent of the request width -- profiles/).
		constexpr int U = 8;
		if (G >= 8) {
			const int nchunk = (e1 - e0 + G - 1) / G;
			const int nchunk_max = (G == 32) ? nchunk : __reduce_max_sync(0xffffffffu, nchunk);
			for (int ch = 0; ch < nchunk_max; ++ch) {
				const int eb = e0 + ch * G;
				int cnt = e1 - eb; cnt = cnt < 0 ? 0 : (cnt > G ? G : cnt);
				const int cnt_max = (G == 32) ? cnt : __reduce_max_sync(0xffffffffu, cnt);
				double a_l = 0.0; int c_l = 0;
				if (gl < cnt) { a_l = __ldg(va + eb + gl); c_l = __ldg(ci + eb + gl); }
				for (int j0 = 0; j0 < cnt_max; j0 += U) {
					double a[U]; int col[U]; double2 v[U];
#pragma unroll
					for (int u = 0; u < U; ++u) {
						a[u] = __shfl_sync(0xffffffffu, a_l, j0 + u, G);
						col[u] = __shfl_sync(0xffffffffu, c_l, j0 + u, G);
					}
#pragma unroll
					for (int u = 0; u < U; ++u)
						v[u] = *reinterpret_cast<const double2 *>(xb + (size_t)col[u] * ldxb);
#pragma unroll
					for (int u = 0; u < U; ++u) {
						const double t0 = __dadd_rn(acc0, __dmul_rn(a[u], v[u].x));
						const double t1 = __dadd_rn(acc1, __dmul_rn(a[u], v[u].y));
						if (j0 + u < cnt) { acc0 = t0; acc1 = t1; }
					}
				}
			}
		} else {
			for (int eb = e0; eb < e1; eb += U) {
				double a[U]; int col[U]; double2 v[U];
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const bool ok = eb + u < e1;
					a[u] = ok ? __ldg(va + eb + u) : 0.0;
					col[u] = ok ? __ldg(ci + eb + u) : 0;
				}
#pragma unroll
				for (int u = 0; u < U; ++u)
					v[u] = *reinterpret_cast<const double2 *>(xb + (size_t)col[u] * ldxb);
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const double t0 = __dadd_rn(acc0, __dmul_rn(a[u], v[u].x));
					const double t1 = __dadd_rn(acc1, __dmul_rn(a[u], v[u].y));
					if (eb + u < e1) { acc0 = t0; acc1 = t1; }
				}
			}
		}
		if (store) *reinterpret_cast<double2 *>(y + (size_t)row * ldy + c) = make_double2(acc0, acc1);
	}
}

template <int G>
static int launch_spmm_v2(int nrows, const int *rp, const int *ci, const double *va,
                          const double *x, int ldx, double *y, int ldy, int k, const int *gate)
{
	const int groups_per_cta = 256 / G;
	const unsigned grid = (unsigned)(((long long)nrows + groups_per_cta - 1) / groups_per_cta);
	spmm_csr_v2_kernel<G><<<grid, 256, 0, g_b200.stream>>>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	B200_KERNEL_CHECK();
	return 0;
}

template <int G, int CPL>
static int launch_spmm(int nrows, const int *rp, const int *ci, const double *va,
                       const double *x, int ldx, double *y, int ldy, int k, const int *gate)
{
	const int groups_per_cta = 256 / G;
	const unsigned grid = (unsigned)(((long long)nrows + groups_per_cta - 1) / groups_per_cta);
	spmm_csr_kernel<G, CPL><<<grid, 256, 0, g_b200.stream>>>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	B200_KERNEL_CHECK();
	return 0;
}

int b200k_spmm(const b200_mat *M, int trans, const double *x, int ldx, double *y, int ldy, int k,
               const int *gate)
{
	const int nrows = trans ? M->ncols : M->nrows;
	const int *rp = trans ? M->t_rp : M->rp, *ci = trans ? M->t_ci : M->ci;
	const double *va = trans ? M->t_va : M->va;
	if (nrows <= 0 || k <= 0) return 0;
	B200Prof prof(B200_PROF_SPMM, 12.0 * M->nnz + 4.0 * (nrows + 1) + 8.0 * k * ((double)M->nrows + M->ncols),
	              2.0 * M->nnz * k);
	const bool al16 = (((uintptr_t)x | (uintptr_t)y) % 16 == 0) && (ldx % 2 == 0) && (ldy % 2 == 0);
	if (al16 && k >= 2 && k % 2 == 0) {
		if (k <= 2)       return launch_spmm_v2<1>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
		else if (k <= 4)  return launch_spmm_v2<2>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
		else if (k <= 8)  return launch_spmm_v2<4>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
		else if (k <= 16) return launch_spmm_v2<8>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
		else if (k <= 32) return launch_spmm_v2<16>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
		return launch_spmm_v2<32>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	}
	if (k == 1)       return launch_spmm<1, 1>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	else if (k == 2)  return launch_spmm<2, 1>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	else if (k <= 4)  return launch_spmm<4, 1>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	else if (k <= 8)  return launch_spmm<8, 1>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	else if (k <= 16) return launch_spmm<16, 1>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	else if (k <= 32) return launch_spmm<32, 1>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	else if (k <= 64) return launch_spmm<32, 2>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
	return launch_spmm<32, 4>(nrows, rp, ci, va, x, ldx, y, ldy, k, gate);
}

extern "C" int b200_mat_dot_multivec(const b200_mat *A, int trans, const b200_mv *x, b200_mv *y,
                                     const int *start, const int *end)
{
	B200_REQUIRE_INIT();
	B200_CHECK(x && y && start && end, "b200_mat_dot_multivec: bad arguments");
	const int k = end[0] - start[0];
	B200_CHECK(k == end[1] - start[1], "b200_mat_dot_multivec: column counts differ (%d vs %d)", k,
	           end[1] - start[1]);
	if (k == 0) return 0;
	B200_CHECK(start[0] >= 0 && end[0] <= x->ncols && start[1] >= 0 && end[1] <= y->ncols,
	           "b200_mat_dot_multivec: column range out of bounds");
	if (!A) {
		// reference app/app_ccs.c:134-137: NULL matrix means copy
		B200_CHECK(x->nrows == y->nrows, "b200_mat_dot_multivec: row counts differ");
		return b200k_axpby(y->nrows, k, 1.0, x->d + start[0], x->ld, 0.0, y->d + start[1], y->ld);
	}
	const int out_rows = trans ? A->ncols : A->nrows, in_rows = trans ? A->nrows : A->ncols;
	B200_CHECK(x->nrows == in_rows && y->nrows == out_rows,
	           "b200_mat_dot_multivec: shapes do not match the matrix (%d x %d)", A->nrows, A->ncols);
	if (x == y) {
		const bool overlap = start[0] < end[1] && start[1] < end[0];
		B200_CHECK(!overlap, "b200_mat_dot_multivec: x and y column ranges overlap on one multi-vector");
	}
	return b200k_spmm(A, trans, x->d + start[0], x->ld, y->d + start[1], y->ld, k, nullptr);
}
