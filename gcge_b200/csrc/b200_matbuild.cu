// Device-side construction of a b200_mat from the caller's CCS arrays.
//
// The host construction (b200_partition_build + dia_build in b200_mat.cu) walks the nnz entries
// several times on one core: counting sort CCS -> CSR, symmetry check, distinct diagonals, the
// diagonal image.  At the headline size (nnz = 1.2e8 per matrix) that is ~2.5 s per matrix --
// on EVERY rank, because every rank is handed the whole CCS and cuts out its slab -- a quarter
// of the single-GPU end-to-end time and more than the whole solve at 8 GPUs.  Here the three
// CCS arrays go to the device as they are (one bulk copy each) and everything else is a handful
// of kernels:
//
//   count / extents   one thread per CCS column: rows of the slab counted with atomics, per-rank
//                     column extents (what the halo plan needs) with atomicMin/Max
//   scan              row pointers (cub::DeviceScan)
//   scatter           (column, source index) pairs into the slab's rows, any order
//   sort rows         one thread per row, insertion sort by (column, source index): the order
//                     of the host's counting sort, so the CSR image is bit-identical to it
//   symmetry          CSR image == CCS image, entry by entry
//   diagonals         distinct col - row offsets into a 64-slot table; then the image itself
//
// Whatever this path does not cover -- halos that are not two contiguous ranges, rows longer
// than 64 entries, malformed input (the host path owns the error messages) -- returns 2 and the
// caller runs the host construction.  tests/test_slots_gpu.py compares the two paths bit for bit.
#include "b200_internal.h"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <vector>

namespace {

constexpr int MB_MAX_ROW = 64;
constexpr int MB_TABLE = 64;
constexpr long long MB_EMPTY = 0x7fffffffffffffffLL;

struct DevFlags {
	int bad_input, row_too_long, not_symmetric, too_many_offsets, duplicate;
	int max_row_nnz, max_col_nnz, pad;
};

__device__ __forceinline__ int mb_owner(long long r, long long n, int nranks)
{
	int q = (int)((r * nranks) / n);
	if (q >= nranks) q = nranks - 1;
	while ((n * q) / nranks > r) --q;
	while ((n * (q + 1)) / nranks <= r) ++q;
	return q;
}

__global__ void mb_count_kernel(int nrows, int ncols, const int *__restrict__ jc, const int *__restrict__ ir, long long lo,
                                long long hi, int nranks, int *cnt, long long *cmin, long long *cmax, DevFlags *fl)
{
	int maxc = 0;
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += gridDim.x * blockDim.x) {
		const int e0 = jc[j], e1 = jc[j + 1];
		if (e0 > e1 || e0 < 0) { fl->bad_input = 1; continue; }
		if (e1 - e0 > maxc) maxc = e1 - e0;
		for (int e = e0; e < e1; ++e) {
			const int r = ir[e];
			if (r < 0 || r >= nrows) { fl->bad_input = 1; continue; }
			if (r >= lo && r < hi) atomicAdd(cnt + (r - lo), 1);
			if (nranks > 1) {
				const int q = mb_owner(r, nrows, nranks);
				if (j < cmin[q]) atomicMin(cmin + q, (long long)j);
				if (j > cmax[q]) atomicMax(cmax + q, (long long)j);
			}
		}
	}
	maxc = __reduce_max_sync(0xffffffffu, maxc);
	if ((threadIdx.x & 31) == 0 && maxc > fl->max_col_nnz) atomicMax(&fl->max_col_nnz, maxc);
}

__global__ void mb_scatter_kernel(int ncols, const int *__restrict__ jc, const int *__restrict__ ir, long long lo, long long hi,
                                  const int *__restrict__ rp, int *cursor, int *ci, int *src)
{
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += gridDim.x * blockDim.x) {
		for (int e = jc[j]; e < jc[j + 1]; ++e) {
			const int r = ir[e];
			if (r >= lo && r < hi) {
				const int pos = rp[r - lo] + atomicAdd(cursor + (r - lo), 1);
				ci[pos] = j; src[pos] = e;
			}
		}
	}
}

// one thread per row: (column, source) pairs ascending; then the values; ci -> ci - shift
__global__ void mb_sort_rows_kernel(int nloc, const int *__restrict__ rp, int *ci, int *src, const double *__restrict__ da,
                                    double *va, int shift, DevFlags *fl)
{
	int maxr = 0;
	for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
		const int e0 = rp[r], L = rp[r + 1] - e0;
		if (L > maxr) maxr = L;
		if (L > MB_MAX_ROW) { fl->row_too_long = 1; continue; }
		for (int i = 1; i < L; ++i) {
			const int c = ci[e0 + i], s = src[e0 + i];
			int k = i - 1;
			while (k >= 0 && (ci[e0 + k] > c || (ci[e0 + k] == c && src[e0 + k] > s))) {
				ci[e0 + k + 1] = ci[e0 + k]; src[e0 + k + 1] = src[e0 + k]; --k;
			}
			ci[e0 + k + 1] = c; src[e0 + k + 1] = s;
		}
		for (int i = 0; i < L; ++i) { va[e0 + i] = da[src[e0 + i]]; ci[e0 + i] -= shift; }
	}
	maxr = __reduce_max_sync(0xffffffffu, maxr);
	if ((threadIdx.x & 31) == 0 && maxr > fl->max_row_nnz) atomicMax(&fl->max_row_nnz, maxr);
}

// CSR rows of the slab (global columns = ci + shift) against the caller's CCS columns [lo, hi)
__global__ void mb_symmetry_kernel(int nloc, long long lo, const int *__restrict__ rp, const int *__restrict__ ci,
                                   const double *__restrict__ va, int shift, const int *__restrict__ jc, const int *__restrict__ ir,
                                   const double *__restrict__ da, DevFlags *fl)
{
	for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
		const int e0 = rp[r], L = rp[r + 1] - e0;
		const int f0 = jc[lo + r], Lc = jc[lo + r + 1] - f0;
		// same row pointers relative to the slab start  <=>  same lengths for every row
		bool same = (L == Lc) && (e0 == f0 - jc[lo]);
		for (int i = 0; same && i < L; ++i)
			same = (ci[e0 + i] + shift == ir[f0 + i]) &&
			       (__double_as_longlong(va[e0 + i]) == __double_as_longlong(da[f0 + i]));
		if (!same) fl->not_symmetric = 1;
	}
}

__device__ __forceinline__ void mb_table_insert(long long *table, long long d, DevFlags *fl)
{
	unsigned h = (unsigned)((unsigned long long)d * 0x9E3779B97F4A7C15ULL >> 58);      // 6 bits
	for (int probe = 0; probe < MB_TABLE; ++probe) {
		const unsigned s = (h + probe) & (MB_TABLE - 1);
		const long long cur = table[s];
		if (cur == d) return;
		if (cur == MB_EMPTY) {
			const long long old = (long long)atomicCAS((unsigned long long *)(table + s), (unsigned long long)MB_EMPTY,
			                                           (unsigned long long)d);
			if (old == MB_EMPTY || old == d) return;
		}
	}
	fl->too_many_offsets = 1;
}

// distinct (global column - global row) over the slab
__global__ void mb_offsets_kernel(int nloc, const int *__restrict__ rp, const int *__restrict__ ci, long long *table,
                                  DevFlags *fl)
{
	long long last = MB_EMPTY;
	for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x)
		for (int e = rp[r]; e < rp[r + 1]; ++e) {
			const long long d = (long long)ci[e] - r;           // local column - local row == global difference
			if (d != last) { mb_table_insert(table, d, fl); last = d; }
		}
}

// val[r*ndp + slot(d)] = va[e]; the image was zeroed before
__global__ void mb_dia_fill_kernel(int nloc, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va,
                                   int nd, const int *__restrict__ offs, const int *__restrict__ slot_of, int ndp, double *val,
                                   DevFlags *fl)
{
	__shared__ int offs_s[32], slot_s[32];
	if (threadIdx.x < nd) { offs_s[threadIdx.x] = offs[threadIdx.x]; slot_s[threadIdx.x] = slot_of[threadIdx.x]; }
	__syncthreads();
	for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
		unsigned mask = 0;
		for (int e = rp[r]; e < rp[r + 1]; ++e) {
			const int d = ci[e] - r;
			int s = 0;
			while (s < nd && offs_s[s] != d) ++s;
			if (s == nd) { fl->too_many_offsets = 1; continue; }
			if (mask & (1u << s)) { fl->duplicate = 1; continue; }
			mask |= 1u << s;
			val[(size_t)r * ndp + slot_s[s]] = va[e];
		}
	}
}

struct Temps {
	std::vector<void *> ptrs;
	~Temps() { for (void *p : ptrs) cudaFree(p); }
	template <typename T> T *get(size_t count)
	{
		void *p = nullptr;
		if (cudaMalloc(&p, sizeof(T) * (count > 0 ? count : 1)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
		ptrs.push_back(p);
		return (T *)p;
	}
	void release(void *p) { ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), p), ptrs.end()); }
};

int grid_for(long long items)
{
	long long g = (items + 255) / 256;
	const long long cap = (long long)g_b200.num_sms * 16;
	if (g > cap) g = cap;
	if (g < 1) g = 1;
	return (int)g;
}

// The image itself, from the table of distinct offsets mb_offsets_kernel filled (read back as table_h) and the
// local CSR in A->rp / ci / va: runs of at most 3 consecutive offsets, every run on an even slot.  Leaves
// A->dia_nd == 0 when the matrix does not qualify (too many or too sparsely filled diagonals, duplicate entries).
int dia_image_from_table(b200_mat *A, const std::vector<long long> &table_h, Temps &tmp, DevFlags *fl)
{
	cudaStream_t st = g_b200.stream;
	const int nloc = A->nrows;
	std::vector<long long> offs;
	for (long long d : table_h) if (d != MB_EMPTY) offs.push_back(d);
	std::sort(offs.begin(), offs.end());
	const int nd = (int)offs.size();
	if (A->nrows_global != A->ncols_global) return 0;      // the diagonal kernels take the x window to be as long as the rows
	if (nd < 1 || nd > 32 || (double)A->nnz < 0.45 * (double)nd * nloc) return 0;
	for (long long d : offs) if (d > 0x3fffffff || d < -0x3fffffff) return 0;
	int ng = 0, ndp = 0;
	std::vector<int> slot_of((size_t)nd), offs_i((size_t)nd);
	for (int s0 = 0; s0 < nd;) {
		int w = 1;
		while (w < 3 && s0 + w < nd && offs[s0 + w] == offs[s0] + w) ++w;
		A->dia_grp_h[2 * ng] = ndp; A->dia_grp_h[2 * ng + 1] = w;
		A->dia_off_h[ng] = (int)offs[s0];
		for (int j = 0; j < w; ++j) slot_of[s0 + j] = ndp + j;
		ndp += (w + 1) & ~1; ++ng; s0 += w;
	}
	for (int i = 0; i < nd; ++i) offs_i[i] = (int)offs[i];
	const size_t npad = (((size_t)nloc + B200_DIA_PAD - 1) / B200_DIA_PAD) * B200_DIA_PAD + B200_DIA_PAD;
	int *offs_d = tmp.get<int>(64), *slot_d = offs_d ? offs_d + 32 : nullptr;
	double *val = nullptr;
	if (!offs_d || cudaMalloc(&val, sizeof(double) * npad * ndp) != cudaSuccess) { cudaGetLastError(); return 0; }
	B200_CUDA(cudaMemcpyAsync(offs_d, offs_i.data(), sizeof(int) * nd, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(slot_d, slot_of.data(), sizeof(int) * nd, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemsetAsync(val, 0, sizeof(double) * npad * ndp, st));
	mb_dia_fill_kernel<<<grid_for(nloc), 256, 0, st>>>(nloc, A->rp, A->ci, A->va, nd, offs_d, slot_d, ndp, val, fl);
	B200_KERNEL_CHECK();
	DevFlags fh;
	B200_CUDA(cudaMemcpyAsync(&fh, fl, sizeof(DevFlags), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	if (fh.duplicate || fh.too_many_offsets) { cudaFree(val); return 0; }      // keep the CSR semantics
	B200_CUDA(cudaMalloc(&A->dia_off, sizeof(int) * 32));
	B200_CUDA(cudaMalloc(&A->dia_grp, sizeof(int) * 64));
	B200_CUDA(cudaMemcpyAsync(A->dia_off, A->dia_off_h, sizeof(int) * ng, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(A->dia_grp, A->dia_grp_h, sizeof(int) * 2 * ng, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaStreamSynchronize(st));
	A->dia_val = val; A->dia_ndp = ndp; A->dia_nd = nd; A->dia_ng = ng;
	return 0;
}

}  // namespace

void b200_note_halo_capacity(long long n_global, int nhalo);

// ---- slab-local input (b200_mat_create_from_local_rows): the caller's arrays are uploaded as they are; these
// kernels check them (row pointers monotone, columns ascending and in range), record the longest row and turn them
// into the local CSR (row pointers from 0, columns minus the slab's first row)
__global__ void mb_local_rows_kernel(int nloc, int nrows_global, long long lo, int rp0, int *__restrict__ rp, int *__restrict__ ci,
                                     DevFlags *fl)
{
	for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += (long long)gridDim.x * blockDim.x) {
		const int e0 = rp[r] - rp0, e1 = rp[r + 1] - rp0;
		if (e1 < e0) { fl->bad_input = 1; continue; }
		atomicMax(&fl->max_row_nnz, e1 - e0);
		int prev = -1;
		for (int e = e0; e < e1; ++e) {
			const int c = ci[e];
			if (c < 0 || c >= nrows_global || c <= prev) fl->bad_input = 1;
			prev = c;
			ci[e] = c - (int)lo;
		}
	}
}
__global__ void mb_shift_rp_kernel(int count, int rp0, int *rp)
{
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) rp[i] -= rp0;
}

// A->rp / ci / va are already the slab's device arrays (A->nrows, nnz set); on return bad != 0 if the input was malformed
int b200k_local_rows_finish(b200_mat *A, int rp0, int *bad)
{
	cudaStream_t st = g_b200.stream;
	Temps tmp;
	DevFlags *fl = tmp.get<DevFlags>(1);
	if (!fl) return b200_fail("b200_mat_create_from_local_rows: out of device memory");
	B200_CUDA(cudaMemsetAsync(fl, 0, sizeof(DevFlags), st));
	mb_local_rows_kernel<<<grid_for(A->nrows), 256, 0, st>>>(A->nrows, A->nrows_global, (long long)A->row0, rp0, A->rp, A->ci, fl);
	B200_KERNEL_CHECK();
	mb_shift_rp_kernel<<<grid_for(A->nrows + 1), 256, 0, st>>>(A->nrows + 1, rp0, A->rp);
	B200_KERNEL_CHECK();
	DevFlags fh;
	B200_CUDA(cudaMemcpyAsync(&fh, fl, sizeof(DevFlags), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	*bad = fh.bad_input;
	A->max_row_nnz = fh.max_row_nnz; A->t_max_row_nnz = fh.max_row_nnz;
	return 0;
}

// The diagonal image of a matrix whose local CSR is on the device (the rules of dia_build in b200_mat.cu); leaves
// A->dia_nd == 0 when the matrix does not qualify.
int b200k_dia_build_device(b200_mat *A)
{
	A->dia_nd = 0;
	if (A->nrows <= 0 || A->nnz <= 0 || b200_opt(B200_OPT_NO_DIA)) return 0;
	cudaStream_t st = g_b200.stream;
	const int nloc = A->nrows;
	Temps tmp;
	long long *table = tmp.get<long long>(MB_TABLE);
	DevFlags *fl = tmp.get<DevFlags>(1);
	if (!table || !fl) return 0;
	std::vector<long long> table_h(MB_TABLE, MB_EMPTY);
	B200_CUDA(cudaMemcpyAsync(table, table_h.data(), sizeof(long long) * MB_TABLE, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemsetAsync(fl, 0, sizeof(DevFlags), st));
	mb_offsets_kernel<<<grid_for(nloc), 256, 0, st>>>(nloc, A->rp, A->ci, table, fl);
	B200_KERNEL_CHECK();
	DevFlags fh;
	B200_CUDA(cudaMemcpyAsync(&fh, fl, sizeof(DevFlags), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaMemcpyAsync(table_h.data(), table, sizeof(long long) * MB_TABLE, cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	if (fh.too_many_offsets || fh.row_too_long) return 0;
	if (dia_image_from_table(A, table_h, tmp, fl)) return 1;
	return 0;
}

// 0: A is complete (device arrays, plan, diagonal image); 1: error; 2: not applicable
int b200k_mat_build_device(int nrows, int ncols, const int *j_col, const int *i_row, const double *data, int rank,
                           int nranks, b200_mat *A)
{
	if (b200_opt(B200_OPT_HOST_BUILD)) return 2;
	if (!j_col || nrows <= 0 || ncols <= 0) return 2;
	if (nranks > 1 && nrows != ncols) return 2;
	const int nnz = j_col[ncols];
	if (nnz <= 0 || !i_row || !data) return 2;
	cudaStream_t st = g_b200.stream;
	long long lo = 0, hi = nrows;
	if (nranks > 1) b200_partition_range(nrows, rank, nranks, &lo, &hi);
	const int nloc = (int)(hi - lo);
	if (nloc <= 0) return 2;

	Temps tmp;
	int *d_jc = tmp.get<int>((size_t)ncols + 1), *d_ir = tmp.get<int>((size_t)nnz);
	double *d_da = tmp.get<double>((size_t)nnz);
	int *cnt = tmp.get<int>((size_t)nloc + 1), *cursor = tmp.get<int>((size_t)nloc);
	long long *ext = tmp.get<long long>((size_t)2 * nranks + MB_TABLE);
	DevFlags *fl = tmp.get<DevFlags>(1);
	int *rp = tmp.get<int>((size_t)nloc + 1);
	if (!d_jc || !d_ir || !d_da || !cnt || !cursor || !ext || !fl || !rp) return 2;
	long long *cmin_d = ext, *cmax_d = ext + nranks, *table = ext + 2 * nranks;
	B200_CUDA(cudaMemcpyAsync(d_jc, j_col, sizeof(int) * ((size_t)ncols + 1), cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(d_ir, i_row, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(d_da, data, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)nloc + 1), st));
	B200_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)nloc, st));
	B200_CUDA(cudaMemsetAsync(fl, 0, sizeof(DevFlags), st));
	std::vector<long long> ext_h((size_t)2 * nranks + MB_TABLE);
	for (int q = 0; q < nranks; ++q) {
		long long a, b;
		b200_partition_range(nrows, q, nranks, &a, &b);
		ext_h[q] = a; ext_h[nranks + q] = b - 1;          // a rank's own rows are always inside its extent
	}
	if (nranks == 1) { ext_h[0] = 0; ext_h[1] = ncols - 1; }
	for (int i = 0; i < MB_TABLE; ++i) ext_h[2 * nranks + i] = MB_EMPTY;
	B200_CUDA(cudaMemcpyAsync(ext, ext_h.data(), sizeof(long long) * ext_h.size(), cudaMemcpyHostToDevice, st));

	mb_count_kernel<<<grid_for(ncols), 256, 0, st>>>(nrows, ncols, d_jc, d_ir, lo, hi, nranks, cnt, cmin_d, cmax_d, fl);
	B200_KERNEL_CHECK();
	{
		size_t tb = 0;
		B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, rp, nloc + 1, st));
		void *tbuf = tmp.get<char>(tb);
		if (!tbuf) return 2;
		B200_CUDA(cub::DeviceScan::ExclusiveSum(tbuf, tb, cnt, rp, nloc + 1, st));
	}
	int nnz_loc = 0;
	DevFlags fh;
	B200_CUDA(cudaMemcpyAsync(&nnz_loc, rp + nloc, sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaMemcpyAsync(&fh, fl, sizeof(DevFlags), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaMemcpyAsync(ext_h.data(), ext, sizeof(long long) * (size_t)2 * nranks, cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	if (fh.bad_input || nnz_loc <= 0) return 2;

	// halo plan from the column extents (the rule of b200_partition_build: contiguous iff no rank's
	// two ranges together exceed its own slab)
	int shift = 0;
	if (nranks > 1) {
		const long long *cmin = ext_h.data(), *cmax = ext_h.data() + nranks;
		std::vector<long long> rlo((size_t)nranks), rhi((size_t)nranks);
		for (int q = 0; q < nranks; ++q) b200_partition_range(nrows, q, nranks, &rlo[q], &rhi[q]);
		for (int q = 0; q < nranks; ++q)
			if ((rlo[q] - cmin[q]) + (cmax[q] + 1 - rhi[q]) > (rhi[q] - rlo[q])) return 2;
		A->halo_contiguous = 1;
		A->halo_below = (int)(lo - cmin[rank]);
		std::vector<int> halo;
		std::vector<std::vector<int>> send((size_t)nranks);
		for (long long c = cmin[rank]; c < lo; ++c) halo.push_back((int)c);
		for (long long c = hi; c <= cmax[rank]; ++c) halo.push_back((int)c);
		for (int q = 0; q < nranks; ++q) {
			if (q == rank) continue;
			long long a0 = std::max(cmin[q], lo), a1 = std::min(rlo[q], hi);
			for (long long c = a0; c < a1; ++c) send[q].push_back((int)(c - lo));
			a0 = std::max(rhi[q], lo); a1 = std::min(cmax[q] + 1, hi);
			for (long long c = a0; c < a1; ++c) send[q].push_back((int)(c - lo));
		}
		A->nhalo = (int)halo.size();
		std::vector<int> recv_cnt((size_t)nranks, 0);
		{
			int q = 0;
			for (int c : halo) {
				while (c < rlo[q] || c >= rhi[q]) q = (c < rlo[q]) ? q - 1 : q + 1;
				++recv_cnt[q];
			}
		}
		std::vector<int> nbr;
		for (int q = 0; q < nranks; ++q) if (q != rank && (recv_cnt[q] || !send[q].empty())) nbr.push_back(q);
		A->nnbr = (int)nbr.size();
		A->nbr = (int *)malloc(sizeof(int) * (nbr.size() + 1));
		A->recv_off = (int *)malloc(sizeof(int) * (nbr.size() + 1));
		A->send_off = (int *)malloc(sizeof(int) * (nbr.size() + 1));
		A->halo_cols = (int *)malloc(sizeof(int) * (halo.size() + 1));
		memcpy(A->halo_cols, halo.data(), sizeof(int) * halo.size());
		int so = 0, ro = 0;
		for (size_t i = 0; i < nbr.size(); ++i) {
			A->nbr[i] = nbr[i]; A->recv_off[i] = ro; A->send_off[i] = so;
			ro += recv_cnt[nbr[i]]; so += (int)send[nbr[i]].size();
		}
		A->recv_off[nbr.size()] = ro; A->send_off[nbr.size()] = so;
		A->send_rows = (int *)malloc(sizeof(int) * (size_t)(so > 0 ? so : 1));
		for (size_t i = 0; i < nbr.size(); ++i)
			memcpy(A->send_rows + A->send_off[i], send[nbr[i]].data(), sizeof(int) * send[nbr[i]].size());
		shift = (int)lo;
	}
	auto drop_plan = [&]() {
		free(A->nbr); free(A->recv_off); free(A->send_off); free(A->halo_cols); free(A->send_rows);
		A->nbr = A->recv_off = A->send_off = A->halo_cols = A->send_rows = nullptr;
		A->nnbr = A->nhalo = A->halo_contiguous = A->halo_below = 0;
	};

	int *ci = tmp.get<int>((size_t)nnz_loc), *src = tmp.get<int>((size_t)nnz_loc);
	double *va = tmp.get<double>((size_t)nnz_loc);
	if (!ci || !src || !va) { drop_plan(); return 2; }
	mb_scatter_kernel<<<grid_for(ncols), 256, 0, st>>>(ncols, d_jc, d_ir, lo, hi, rp, cursor, ci, src);
	B200_KERNEL_CHECK();
	mb_sort_rows_kernel<<<grid_for(nloc), 256, 0, st>>>(nloc, rp, ci, src, d_da, va, shift, fl);
	B200_KERNEL_CHECK();
	if (nrows == ncols) {
		mb_symmetry_kernel<<<grid_for(nloc), 256, 0, st>>>(nloc, lo, rp, ci, va, shift, d_jc, d_ir, d_da, fl);
		B200_KERNEL_CHECK();
	}
	const bool try_dia = !b200_opt(B200_OPT_NO_DIA);
	if (try_dia) {
		mb_offsets_kernel<<<grid_for(nloc), 256, 0, st>>>(nloc, rp, ci, table, fl);
		B200_KERNEL_CHECK();
	}
	std::vector<long long> table_h(MB_TABLE);
	B200_CUDA(cudaMemcpyAsync(&fh, fl, sizeof(DevFlags), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaMemcpyAsync(table_h.data(), table, sizeof(long long) * MB_TABLE, cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	if (fh.row_too_long) { drop_plan(); return 2; }

	A->nrows = nloc; A->ncols = (nranks == 1) ? ncols : nloc; A->nnz = nnz_loc;
	A->nrows_global = nrows; A->ncols_global = ncols; A->nnz_global = nnz; A->row0 = (int)lo; A->t_col0 = lo;
	A->max_row_nnz = fh.max_row_nnz; A->t_max_row_nnz = fh.max_col_nnz;
	A->symmetric = (nrows == ncols) && !fh.not_symmetric;
	A->rp = rp; A->ci = ci; A->va = va;
	tmp.release(rp); tmp.release(ci); tmp.release(va);
	if (nranks == 1) {
		if (A->symmetric) { A->t_shared = 1; A->t_rp = A->rp; A->t_ci = A->ci; A->t_va = A->va; }
		else {
			// the caller's CCS arrays verbatim are already on the device
			A->t_shared = 0; A->t_rp = d_jc; A->t_ci = d_ir; A->t_va = d_da;
			tmp.release(d_jc); tmp.release(d_ir); tmp.release(d_da);
		}
	} else {
		const int ns = A->send_off[A->nnbr];
		B200_CUDA(cudaMalloc(&A->send_rows_dev, sizeof(int) * (size_t)(ns > 0 ? ns : 1)));
		B200_CUDA(cudaMemcpyAsync(A->send_rows_dev, A->send_rows, sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice, st));
		b200_note_halo_capacity(A->ncols_global, A->nhalo);
	}

	// ---- diagonal image (the rules of dia_build in b200_mat.cu)
	A->dia_nd = 0;
	if (try_dia && !fh.too_many_offsets && dia_image_from_table(A, table_h, tmp, fl)) return 1;
	B200_CUDA(cudaStreamSynchronize(st));
	return 0;
}
