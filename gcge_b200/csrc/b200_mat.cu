// Device matrix: CCS (caller) -> CSR (device) conversion, round trip, MatAxpby.
//
// The reference multiplies y = A x with a column-major scatter over the CCS arrays
// (reference app/app_ccs.c:116-131): for column j ascending, y[i_row[e]] += data[e]*x[j].
// Row r of y therefore accumulates its entries in ascending column order.  The device CSR
// keeps exactly that order inside each row, so a sequential gather reproduces the
// reference's floating-point sum bit for bit.
#include "b200_internal.h"
#include <vector>
#include <algorithm>

extern "C" void b200_partition_range(long long n, int rank, int nranks, long long *lo, long long *hi);

static int owner_of(long long row, long long n, int nranks)
{
	// largest g with floor(g n / G) <= row
	int g = (int)(((__int128)(row + 1) * nranks - 1) / n);
	if (g >= nranks) g = nranks - 1;
	long long lo, hi;
	b200_partition_range(n, g, nranks, &lo, &hi);
	while (row < lo) { --g; b200_partition_range(n, g, nranks, &lo, &hi); }
	while (row >= hi) { ++g; b200_partition_range(n, g, nranks, &lo, &hi); }
	return g;
}

// Host-only: the local CSR slab of rank `rank`, its halo list and the send lists.  No device
// call in here, so tests can run it on a CPU-only box for any (rank, nranks).
extern "C" int b200_partition_build(int nrows, int ncols, const int *j_col, const int *i_row, const double *data,
                                    int rank, int nranks, b200_mat *A, int **rp_h, int **ci_h, double **va_h)
{
	B200_CHECK(A && j_col && nrows >= 0 && ncols >= 0 && rp_h && ci_h && va_h, "partition: bad arguments");
	if (nranks < 1) nranks = 1;
	B200_CHECK(nranks == 1 || nrows == ncols, "partition: a %d x %d matrix cannot be row-partitioned together with "
	           "its multi-vectors (square matrices only across ranks)", nrows, ncols);
	const int nnz = j_col[ncols];
	B200_CHECK(nnz >= 0 && (nnz == 0 || (i_row && data)), "b200_mat_create_from_ccs: bad CCS arrays");
	for (int j = 0; j < ncols; ++j)
		B200_CHECK(j_col[j] <= j_col[j + 1], "b200_mat_create_from_ccs: j_col not monotone at %d", j);
	long long lo = 0, hi = nrows;
	if (nranks > 1) b200_partition_range(nrows, rank, nranks, &lo, &hi);
	const int nloc = (int)(hi - lo);
	// local CSR by counting sort over the slab's rows, columns visited ascending
	int *rp = (int *)calloc((size_t)nloc + 1, sizeof(int));
	B200_CHECK(rp, "partition: out of host memory");
	for (int e = 0; e < nnz; ++e) {
		const int r = i_row[e];
		if (r < 0 || r >= nrows) { free(rp); return b200_fail("b200_mat_create_from_ccs: row index %d out of range at %d", r, e); }
		if (r >= lo && r < hi) ++rp[r - lo + 1];
	}
	for (int r = 0; r < nloc; ++r) rp[r + 1] += rp[r];
	const int nnz_loc = rp[nloc];
	int *ci = (int *)malloc(sizeof(int) * (size_t)(nnz_loc > 0 ? nnz_loc : 1));
	double *va = (double *)malloc(sizeof(double) * (size_t)(nnz_loc > 0 ? nnz_loc : 1));
	if (!ci || !va) { free(rp); free(ci); free(va); return b200_fail("partition: out of host memory"); }
	{
		std::vector<int> next(rp, rp + nloc);
		for (int j = 0; j < ncols; ++j)
			for (int e = j_col[j]; e < j_col[j + 1]; ++e) {
				const int r = i_row[e];
				if (r >= lo && r < hi) { const int pos = next[r - lo]++; ci[pos] = j; va[pos] = data[e]; }
			}
	}
	A->nrows = nloc; A->ncols = (nranks == 1) ? ncols : nloc; A->nnz = nnz_loc;
	A->nrows_global = nrows; A->ncols_global = ncols; A->nnz_global = nnz; A->row0 = (int)lo; A->t_col0 = lo;
	for (int r = 0; r < nloc; ++r) if (rp[r + 1] - rp[r] > A->max_row_nnz) A->max_row_nnz = rp[r + 1] - rp[r];
	// symmetric: the slab's CSR rows (global columns) are the caller's CCS columns [lo, hi) bit for bit
	int sym = (nrows == ncols);
	if (sym) {
		const int base = j_col[lo];
		for (int r = 0; r <= nloc && sym; ++r) sym = (rp[r] == j_col[lo + r] - base);
		if (sym && nnz_loc) sym = (0 == memcmp(ci, i_row + base, sizeof(int) * (size_t)nnz_loc));
		if (sym && nnz_loc) sym = (0 == memcmp(va, data + base, sizeof(double) * (size_t)nnz_loc));
	}
	A->symmetric = sym;
	if (nranks > 1) {
		// Column extent of every rank's slab: [cmin_q, cmax_q] over the entries of its rows.  One
		// pass over the whole CCS (every rank holds it here), identical on all ranks.
		std::vector<long long> cmin((size_t)nranks), cmax((size_t)nranks), rlo((size_t)nranks), rhi((size_t)nranks);
		for (int q = 0; q < nranks; ++q) { b200_partition_range(nrows, q, nranks, &rlo[q], &rhi[q]); cmin[q] = rlo[q]; cmax[q] = rhi[q] - 1; }
		{
			int q = 0;
			for (int j = 0; j < ncols; ++j)
				for (int e = j_col[j]; e < j_col[j + 1]; ++e) {
					const int r = i_row[e];
					if (r < rlo[q] || r >= rhi[q]) q = owner_of(r, nrows, nranks);
					if (j < cmin[q]) cmin[q] = j;
					if (j > cmax[q]) cmax[q] = j;
				}
		}
		// Banded matrices (stencils, FEM on lattices): take the halo as the two CONTIGUOUS column
		// ranges [cmin, lo) and [hi, cmax] -- at most a few unreferenced rows travel, and the x block
		// a slab multiplies with is then piecewise contiguous, which the diagonal-storage SpMM needs.
		// The rule is global (every rank must pick the same mode): contiguous iff no rank's two
		// ranges together exceed its own slab.
		bool contiguous = true;
		for (int q = 0; q < nranks; ++q)
			if ((rlo[q] - cmin[q]) + (cmax[q] + 1 - rhi[q]) > (rhi[q] - rlo[q])) contiguous = false;
		A->halo_contiguous = contiguous ? 1 : 0;
		A->halo_below = contiguous ? (int)(lo - cmin[rank]) : 0;
		std::vector<int> halo;
		std::vector<std::vector<int>> send((size_t)nranks);
		if (contiguous) {
			for (long long c = cmin[rank]; c < lo; ++c) halo.push_back((int)c);
			for (long long c = hi; c <= cmax[rank]; ++c) halo.push_back((int)c);
			for (int q = 0; q < nranks; ++q) {
				if (q == rank) continue;
				// my rows inside q's two ranges, ascending
				long long a0 = std::max(cmin[q], lo), a1 = std::min(rlo[q], hi);
				for (long long c = a0; c < a1; ++c) send[q].push_back((int)(c - lo));
				a0 = std::max(rhi[q], lo); a1 = std::min(cmax[q] + 1, hi);
				for (long long c = a0; c < a1; ++c) send[q].push_back((int)(c - lo));
			}
		} else {
			// exact list: off-slab columns, ascending, unique
			std::vector<unsigned char> mark((size_t)ncols, 0);
			for (int e = 0; e < nnz_loc; ++e) if (ci[e] < lo || ci[e] >= hi) mark[ci[e]] = 1;
			for (int c = 0; c < ncols; ++c) if (mark[c]) halo.push_back(c);
			// what the other ranks need from me: my columns j with an entry in a row they own
			for (long long j = lo; j < hi; ++j)
				for (int e = j_col[j]; e < j_col[j + 1]; ++e) {
					const int r = i_row[e];
					if (r >= lo && r < hi) continue;
					const int q = owner_of(r, nrows, nranks);
					if (send[q].empty() || send[q].back() != (int)(j - lo)) send[q].push_back((int)(j - lo));
				}
		}
		A->nhalo = (int)halo.size();
		for (int e = 0; e < nnz_loc; ++e) {
			const int c = ci[e];
			if (contiguous || (c >= lo && c < hi)) ci[e] = c - (int)lo;      // may be negative / >= nloc: halo rows
			else ci[e] = nloc + (int)(std::lower_bound(halo.begin(), halo.end(), c) - halo.begin());
		}
		std::vector<int> recv_cnt((size_t)nranks, 0);
		for (int c : halo) ++recv_cnt[owner_of(c, nrows, nranks)];
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
	}
	*rp_h = rp; *ci_h = ci; *va_h = va;
	return 0;
}

// ---- host-only view of the partition plan (tests; needs no device) ---------------------------
struct b200_plan_ { b200_mat m; int *rp, *ci; double *va; };

extern "C" int b200_plan_create(int nrows, int ncols, const int *j_col, const int *i_row, const double *data,
                                int rank, int nranks, b200_plan **out)
{
	B200_CHECK(out, "b200_plan_create: bad arguments");
	b200_plan *p = (b200_plan *)calloc(1, sizeof(b200_plan));
	if (b200_partition_build(nrows, ncols, j_col, i_row, data, rank, nranks, &p->m, &p->rp, &p->ci, &p->va)) { free(p); return 1; }
	*out = p;
	return 0;
}

extern "C" int b200_plan_sizes(const b200_plan *p, int *row0, int *nrows_local, int *nnz_local, int *nhalo, int *nnbr,
                               int *nsend, int *symmetric, int *halo_contiguous, int *halo_below)
{
	B200_CHECK(p, "b200_plan_sizes: NULL plan");
	if (row0) *row0 = p->m.row0;
	if (nrows_local) *nrows_local = p->m.nrows;
	if (nnz_local) *nnz_local = p->m.nnz;
	if (nhalo) *nhalo = p->m.nhalo;
	if (nnbr) *nnbr = p->m.nnbr;
	if (nsend) *nsend = p->m.nnbr ? p->m.send_off[p->m.nnbr] : 0;
	if (halo_contiguous) *halo_contiguous = p->m.halo_contiguous;
	if (halo_below) *halo_below = p->m.halo_below;
	if (symmetric) *symmetric = p->m.symmetric;
	return 0;
}

extern "C" int b200_plan_copy(const b200_plan *p, int *rp, int *ci, double *va, int *halo_cols, int *nbr,
                              int *recv_off, int *send_off, int *send_rows)
{
	B200_CHECK(p, "b200_plan_copy: NULL plan");
	const b200_mat *m = &p->m;
	if (rp) memcpy(rp, p->rp, sizeof(int) * ((size_t)m->nrows + 1));
	if (ci) memcpy(ci, p->ci, sizeof(int) * (size_t)m->nnz);
	if (va) memcpy(va, p->va, sizeof(double) * (size_t)m->nnz);
	if (halo_cols && m->nhalo) memcpy(halo_cols, m->halo_cols, sizeof(int) * (size_t)m->nhalo);
	if (m->nnbr) {
		if (nbr) memcpy(nbr, m->nbr, sizeof(int) * (size_t)m->nnbr);
		if (recv_off) memcpy(recv_off, m->recv_off, sizeof(int) * ((size_t)m->nnbr + 1));
		if (send_off) memcpy(send_off, m->send_off, sizeof(int) * ((size_t)m->nnbr + 1));
		if (send_rows) memcpy(send_rows, m->send_rows, sizeof(int) * (size_t)m->send_off[m->nnbr]);
	}
	return 0;
}

extern "C" int b200_plan_destroy(b200_plan *p)
{
	if (!p) return 0;
	free(p->rp); free(p->ci); free(p->va);
	free(p->m.nbr); free(p->m.halo_cols); free(p->m.recv_off); free(p->m.send_off); free(p->m.send_rows);
	free(p);
	return 0;
}

// ---- diagonal image -----------------------------------------------------------------------------
// A lattice operator in natural ordering has its entries on a handful of diagonals
// (col - row in a fixed set: 7 / 15 / 27 offsets for the 7-point, P1-Kuhn and 27-point
// operators).  Storing it by diagonals drops the column indices (8 nd + 4 bytes per row instead
// of 12 per entry) and, more importantly, lets the SpMM kernel slide along consecutive rows:
// rows r and r+1 share the x rows of adjacent offsets, so a thread that walks a block of rows
// loads every x row once instead of once per matrix row that touches it (b200_spmm.cu).
// rp/ci/va: the local CSR slab with REMAPPED columns (see b200_partition_build).
int b200k_mat_build_device(int nrows, int ncols, const int *j_col, const int *i_row, const double *data, int rank,
                           int nranks, b200_mat *A);
int b200k_local_rows_finish(b200_mat *A, int rp0, int *bad);
int b200k_dia_build_device(b200_mat *A);
void b200_note_halo_capacity(long long n_global, int nhalo);

static int dia_build(b200_mat *A, const int *rp, const int *ci, const double *va, int nranks)
{
	const int nloc = A->nrows;
	A->dia_nd = 0;
	if (nloc <= 0 || A->nnz <= 0 || b200_opt(B200_OPT_NO_DIA)) return 0;
	if (A->nrows_global != A->ncols_global) return 0;      // the diagonal kernels take the x window to be as long as the rows
	if (nranks > 1 && !A->halo_contiguous) return 0;
	const long long lo = A->row0;
	auto gcol = [&](int c) -> long long {
		return (A->halo_contiguous || c < nloc) ? lo + c : (long long)A->halo_cols[c - nloc];
	};
	// distinct offsets (give up beyond 32)
	std::vector<long long> offs;
	for (int r = 0; r < nloc; ++r)
		for (int e = rp[r]; e < rp[r + 1]; ++e) {
			const long long d = gcol(ci[e]) - (lo + r);
			auto it = std::lower_bound(offs.begin(), offs.end(), d);
			if (it == offs.end() || *it != d) {
				if (offs.size() >= 32) return 0;
				offs.insert(it, d);
			}
		}
	const int nd = (int)offs.size();
	// too sparse on its diagonals: below ~45 % fill the zero slots cost more than the CSR gathers save (P1 on the
	// bisected cube4 mesh in lattice order fills 49.8 % of its 27 diagonals and is still worth it)
	if ((double)A->nnz < 0.45 * (double)nd * nloc) return 0;
	for (long long d : offs) if (d > 0x3fffffff || d < -0x3fffffff) return 0;
	// runs of consecutive offsets, at most 3 wide; every run starts on an even slot of the image
	// (one padding slot after a run of odd width) so its values are one aligned 128-bit load
	int ng = 0, ndp = 0;
	std::vector<int> slot_of((size_t)nd);
	for (int s0 = 0; s0 < nd;) {
		int w = 1;
		while (w < 3 && s0 + w < nd && offs[s0 + w] == offs[s0] + w) ++w;
		A->dia_grp_h[2 * ng] = ndp; A->dia_grp_h[2 * ng + 1] = w;
		A->dia_off_h[ng] = (int)offs[s0];                          // first offset of the run
		for (int j = 0; j < w; ++j) slot_of[s0 + j] = ndp + j;
		ndp += (w + 1) & ~1; ++ng; s0 += w;
	}
	// padded behind the last row by at least one row block of the SpMM kernels (zero values; their
	// block heights need not divide the row count), so their bulk copies of a block's image are
	// always full size and in bounds
	const size_t npad = (((size_t)nloc + B200_DIA_PAD - 1) / B200_DIA_PAD) * B200_DIA_PAD + B200_DIA_PAD;
	std::vector<double> val(npad * ndp, 0.0);
	std::vector<unsigned> mask((size_t)nloc, 0u);
	for (int r = 0; r < nloc; ++r)
		for (int e = rp[r]; e < rp[r + 1]; ++e) {
			const long long d = gcol(ci[e]) - (lo + r);
			const int s = (int)(std::lower_bound(offs.begin(), offs.end(), d) - offs.begin());
			if (mask[r] & (1u << s)) return 0;                    // duplicate entry: keep the CSR semantics
			val[(size_t)r * ndp + slot_of[s]] = va[e]; mask[r] |= 1u << s;
		}
	cudaStream_t st = g_b200.stream;
	B200_CUDA(cudaMalloc(&A->dia_off, sizeof(int) * 32));
	B200_CUDA(cudaMalloc(&A->dia_grp, sizeof(int) * 64));
	B200_CUDA(cudaMalloc(&A->dia_val, sizeof(double) * npad * ndp));
	B200_CUDA(cudaMemcpyAsync(A->dia_off, A->dia_off_h, sizeof(int) * ng, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(A->dia_grp, A->dia_grp_h, sizeof(int) * 2 * ng, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(A->dia_val, val.data(), sizeof(double) * npad * ndp, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaStreamSynchronize(st));       // val is a local
	A->dia_ndp = ndp;
	A->dia_nd = nd; A->dia_ng = ng;
	return 0;
}

void b200_note_halo_capacity(long long n_global, int nhalo)
{
	for (int i = 0; i < 8; ++i) {
		if (g_b200.halo_n[i] == n_global || g_b200.halo_n[i] == 0) {
			g_b200.halo_n[i] = n_global;
			if (nhalo > g_b200.halo_cap[i]) g_b200.halo_cap[i] = nhalo;
			return;
		}
	}
}

// collective: size the halo mailboxes of the copy-engine exchange for this matrix (b200_comm.cu)
// collective: A is symmetric only if every rank's slab is (a rank that saw an unsymmetric slab would otherwise
// refuse a transposed multiply while the others enter its halo exchange)
static int symmetric_across_ranks(b200_mat *A)
{
	double *flag = (double *)b200_scratch(3, 64);
	if (!flag) return 1;
	const double mine = A->symmetric ? 1.0 : 0.0;
	if (b200k_h2d(flag, &mine, sizeof(double))) return 1;
	if (b200k_allreduce_sum(flag, 1)) return 1;
	double sum = 0.0;
	if (b200k_d2h(&sum, flag, sizeof(double))) return 1;
	A->symmetric = (sum > g_b200.nranks - 0.5) ? 1 : 0;
	return 0;
}

static int p2p_register_for(b200_mat *A)
{
	int rows = 0;
	bool shape_ok = A->halo_contiguous && A->dia_nd > 0 && A->nnbr <= 2;
	for (int i = 0; i < A->nnbr && shape_ok; ++i) {
		if (A->nbr[i] != g_b200.rank - 1 && A->nbr[i] != g_b200.rank + 1) shape_ok = false;
		const int nr = A->recv_off[i + 1] - A->recv_off[i];
		if (nr > rows) rows = nr;
	}
	return b200k_p2p_register(shape_ok ? rows : -1, &A->p2p_ok);
}

extern "C" int b200_mat_create_from_ccs(int nrows, int ncols, const int *j_col, const int *i_row,
                                        const double *data, b200_mat **out)
{
	B200_REQUIRE_INIT();
	B200_CHECK(out && j_col && nrows >= 0 && ncols >= 0, "b200_mat_create_from_ccs: bad arguments");
	const int nranks = g_b200.nranks > 1 ? g_b200.nranks : 1;
	b200_mat *A = (b200_mat *)calloc(1, sizeof(b200_mat));
	{
		// fast path: the whole construction on the device (b200_matbuild.cu); 2 = not applicable
		const int rc = b200k_mat_build_device(nrows, ncols, j_col, i_row, data, g_b200.rank, nranks, A);
		if (rc == 0) {
			if (b200k_lat_detect(A)) { b200_mat_destroy(A); return 1; }
			if (nranks > 1 && (symmetric_across_ranks(A) || p2p_register_for(A))) { b200_mat_destroy(A); return 1; }
			*out = A; return 0;
		}
		if (rc == 1) { free(A); return 1; }
		memset(A, 0, sizeof(*A));
	}
	int *rp = nullptr, *ci = nullptr; double *va = nullptr;
	if (b200_partition_build(nrows, ncols, j_col, i_row, data, g_b200.rank, nranks, A, &rp, &ci, &va)) { free(A); return 1; }
	const int nnz = A->nnz, nloc = A->nrows;
	for (int j = 0; j < ncols; ++j) if (j_col[j + 1] - j_col[j] > A->t_max_row_nnz) A->t_max_row_nnz = j_col[j + 1] - j_col[j];
	// single GPU, symmetric matrix: the CCS arrays ARE the CSR image => share storage
	const int shared = (nranks == 1) && A->symmetric;
	A->t_shared = shared;
	cudaStream_t st = g_b200.stream;
	const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
	B200_CUDA(cudaMalloc(&A->rp, sizeof(int) * ((size_t)nloc + 1)));
	B200_CUDA(cudaMalloc(&A->ci, sizeof(int) * nz));
	B200_CUDA(cudaMalloc(&A->va, sizeof(double) * nz));
	B200_CUDA(cudaMemcpyAsync(A->rp, rp, sizeof(int) * ((size_t)nloc + 1), cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(A->ci, ci, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(A->va, va, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	if (shared) {
		A->t_rp = A->rp; A->t_ci = A->ci; A->t_va = A->va;
	} else if (nranks == 1) {
		// the caller's CCS arrays verbatim: CSR of A^T, and the bit-exact round trip
		B200_CUDA(cudaMalloc(&A->t_rp, sizeof(int) * ((size_t)ncols + 1)));
		B200_CUDA(cudaMalloc(&A->t_ci, sizeof(int) * nz));
		B200_CUDA(cudaMalloc(&A->t_va, sizeof(double) * nz));
		B200_CUDA(cudaMemcpyAsync(A->t_rp, j_col, sizeof(int) * ((size_t)ncols + 1), cudaMemcpyHostToDevice, st));
		B200_CUDA(cudaMemcpyAsync(A->t_ci, i_row, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
		B200_CUDA(cudaMemcpyAsync(A->t_va, data, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	}
	if (nranks > 1) {
		const int ns = A->send_off[A->nnbr];
		B200_CUDA(cudaMalloc(&A->send_rows_dev, sizeof(int) * (size_t)(ns > 0 ? ns : 1)));
		B200_CUDA(cudaMemcpyAsync(A->send_rows_dev, A->send_rows, sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice, st));
		b200_note_halo_capacity(A->ncols_global, A->nhalo);
	}
	B200_CUDA(cudaStreamSynchronize(st));
	const int rc_dia = dia_build(A, rp, ci, va, nranks);
	free(rp); free(ci); free(va);
	if (rc_dia || b200k_lat_detect(A)) { b200_mat_destroy(A); return 1; }
	if (nranks > 1 && (symmetric_across_ranks(A) || p2p_register_for(A))) { b200_mat_destroy(A); return 1; }
	*out = A;
	return 0;
}

// ---- slab-local construction ---------------------------------------------------------------------------------
// Every rank hands over ONLY its row block (the reference's distributed back ends own only their rows:
// app/app_phg.c:292-357, app/app_slepc.c): CSR-style arrays of rows [row0, row0 + nrows_local) with GLOBAL column
// indices ascending inside a row -- for the symmetric matrices of the eigenproblem these are the CCS arrays of the
// columns [row0, row0 + nrows_local).  The halo plan needs the column extent of every slab: one small allreduce.
// Banded matrices only (every slab's off-slab columns lie in its two neighbouring ranges, as for
// b200_mat_create_from_ccs's contiguous mode); the matrix is taken to be symmetric, as the reference assumes for
// every matrix (app/app_ccs.c:140-150) -- a rank cannot check that on its rows alone.
extern "C" int b200_mat_create_from_local_rows(int nrows_global, int row0, int nrows_local, const int *rp_in,
                                               const int *ci_in, const double *va_in, b200_mat **out)
{
	B200_REQUIRE_INIT();
	B200_CHECK(out && rp_in && nrows_global >= 0 && nrows_local >= 0, "b200_mat_create_from_local_rows: bad arguments");
	const int nranks = g_b200.nranks > 1 ? g_b200.nranks : 1, rank = nranks > 1 ? g_b200.rank : 0;
	long long lo = 0, hi = nrows_global;
	if (nranks > 1) b200_partition_range(nrows_global, rank, nranks, &lo, &hi);
	B200_CHECK(row0 == lo && nrows_local == hi - lo, "b200_mat_create_from_local_rows: rank %d owns rows [%lld, %lld), got [%d, %d)",
	           rank, lo, hi, row0, row0 + nrows_local);
	const int nloc = nrows_local, nnz = rp_in[nloc] - rp_in[0];
	B200_CHECK(nnz >= 0 && (nnz == 0 || (ci_in && va_in)), "b200_mat_create_from_local_rows: bad arrays");
	// column extent of the slab from the first / last entry of every row (columns ascend inside a row; the device
	// checks that, and everything else about the arrays, once they are uploaded)
	// Malformed input must fail on EVERY rank (the construction below is collective): what this rank finds wrong with
	// its own arrays is summed over the ranks before anybody gives up.
	long long cmin = lo, cmax = hi - 1;
	int bad_rows = 0;
	for (int r = 0; r < nloc && !bad_rows; ++r) {
		const int e0 = rp_in[r] - rp_in[0], e1 = rp_in[r + 1] - rp_in[0];
		if (!(e0 >= 0 && e1 >= e0 && e1 <= nnz)) { bad_rows = 1; break; }
		if (e1 > e0) {
			const int *cr = ci_in + rp_in[0];
			if (cr[e0] < cmin) cmin = cr[e0];
			if (cr[e1 - 1] > cmax) cmax = cr[e1 - 1];
		}
	}
	if (!(cmin >= 0 && cmax < nrows_global)) bad_rows |= 2;
	if (nranks == 1) {
		B200_CHECK(!(bad_rows & 1), "b200_mat_create_from_local_rows: row pointers not monotone");
		B200_CHECK(!(bad_rows & 2), "b200_mat_create_from_local_rows: column index out of range");
	}
	if (bad_rows) { cmin = lo; cmax = hi - 1; }
	b200_mat *A = (b200_mat *)calloc(1, sizeof(b200_mat));
	A->nrows = nloc; A->ncols = (nranks == 1) ? nrows_global : nloc; A->nnz = nnz;
	A->nrows_global = nrows_global; A->ncols_global = nrows_global; A->row0 = (int)lo; A->t_col0 = lo;
	A->symmetric = 1;
	std::vector<long long> ext_min((size_t)nranks, 0), ext_max((size_t)nranks, 0), rlo((size_t)nranks), rhi((size_t)nranks);
	long long nnz_global = nnz;
	if (nranks > 1) {
		// extents and the global entry count of all ranks: one allreduce over a zero-padded array
		double *buf = (double *)b200_scratch(3, sizeof(double) * (2 * (size_t)nranks + 4));
		if (!buf) { free(A); return 1; }
		std::vector<double> h(2 * (size_t)nranks + 2, 0.0);
		h[2 * rank] = (double)cmin; h[2 * rank + 1] = (double)cmax; h[2 * nranks] = (double)nnz;
		h[2 * nranks + 1] = bad_rows ? 1.0 : 0.0;
		if (b200k_h2d(buf, h.data(), sizeof(double) * h.size()) || b200k_allreduce_sum(buf, h.size()) ||
		    b200k_d2h(h.data(), buf, sizeof(double) * h.size())) { free(A); return 1; }
		if (h[2 * nranks + 1] != 0.0) {
			free(A);
			return b200_fail("b200_mat_create_from_local_rows: %s", bad_rows ? ((bad_rows & 1) ? "row pointers not monotone" :
			                 "column index out of range") : "another rank's rows are malformed");
		}
		for (int q = 0; q < nranks; ++q) {
			ext_min[q] = (long long)h[2 * q]; ext_max[q] = (long long)h[2 * q + 1];
			b200_partition_range(nrows_global, q, nranks, &rlo[q], &rhi[q]);
		}
		nnz_global = (long long)h[2 * nranks];
		bool contiguous = true;
		for (int q = 0; q < nranks; ++q)
			if ((rlo[q] - ext_min[q]) + (ext_max[q] + 1 - rhi[q]) > (rhi[q] - rlo[q])) contiguous = false;
		if (!contiguous) {
			free(A);
			return b200_fail("b200_mat_create_from_local_rows: the off-slab columns of some rank exceed its own slab (not a banded "
			                 "matrix in this partition); hand the whole matrix to b200_mat_create_from_ccs instead");
		}
		A->halo_contiguous = 1;
		A->halo_below = (int)(lo - ext_min[rank]);
		A->nhalo = (int)((lo - ext_min[rank]) + (ext_max[rank] + 1 - hi));
		std::vector<int> nbr, recv_cnt, send_cnt;
		std::vector<std::vector<int>> send((size_t)nranks);
		std::vector<int> rc((size_t)nranks, 0);
		for (int q = 0; q < nranks; ++q) {
			if (q == rank) continue;
			// rows of q inside my two ranges
			long long a0 = std::max(ext_min[rank], rlo[q]), a1 = std::min(lo, rhi[q]);
			if (a1 > a0) rc[q] += (int)(a1 - a0);
			a0 = std::max(hi, rlo[q]); a1 = std::min(ext_max[rank] + 1, rhi[q]);
			if (a1 > a0) rc[q] += (int)(a1 - a0);
			// my rows inside q's two ranges, ascending
			a0 = std::max(ext_min[q], lo); a1 = std::min(rlo[q], hi);
			for (long long c = a0; c < a1; ++c) send[q].push_back((int)(c - lo));
			a0 = std::max(rhi[q], lo); a1 = std::min(ext_max[q] + 1, hi);
			for (long long c = a0; c < a1; ++c) send[q].push_back((int)(c - lo));
		}
		for (int q = 0; q < nranks; ++q) if (q != rank && (rc[q] || !send[q].empty())) nbr.push_back(q);
		A->nnbr = (int)nbr.size();
		A->nbr = (int *)malloc(sizeof(int) * (nbr.size() + 1));
		A->recv_off = (int *)malloc(sizeof(int) * (nbr.size() + 1));
		A->send_off = (int *)malloc(sizeof(int) * (nbr.size() + 1));
		A->halo_cols = (int *)malloc(sizeof(int) * ((size_t)A->nhalo + 1));
		{
			int h2 = 0;
			for (long long c = ext_min[rank]; c < lo; ++c) A->halo_cols[h2++] = (int)c;
			for (long long c = hi; c <= ext_max[rank]; ++c) A->halo_cols[h2++] = (int)c;
		}
		int so = 0, ro = 0;
		for (size_t i = 0; i < nbr.size(); ++i) {
			A->nbr[i] = nbr[i]; A->recv_off[i] = ro; A->send_off[i] = so;
			ro += rc[nbr[i]]; so += (int)send[nbr[i]].size();
		}
		A->recv_off[nbr.size()] = ro; A->send_off[nbr.size()] = so;
		A->send_rows = (int *)malloc(sizeof(int) * (size_t)(so > 0 ? so : 1));
		for (size_t i = 0; i < nbr.size(); ++i)
			memcpy(A->send_rows + A->send_off[i], send[nbr[i]].data(), sizeof(int) * send[nbr[i]].size());
	}
	B200_CHECK(nnz_global < 0x7fffffffLL, "b200_mat_create_from_local_rows: %lld entries in all (32-bit counts)", nnz_global);
	A->nnz_global = (int)nnz_global;
	// the caller's arrays go to the device as they are; kernels validate them and turn them into the local CSR
	// (row pointers from 0, columns minus row0: halo rows in front of / behind the local ones)
	cudaStream_t st = g_b200.stream;
	const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
	B200_CUDA(cudaMalloc(&A->rp, sizeof(int) * ((size_t)nloc + 1)));
	B200_CUDA(cudaMalloc(&A->ci, sizeof(int) * nz));
	B200_CUDA(cudaMalloc(&A->va, sizeof(double) * nz));
	B200_CUDA(cudaMemcpyAsync(A->rp, rp_in, sizeof(int) * ((size_t)nloc + 1), cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(A->ci, ci_in + rp_in[0], sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	B200_CUDA(cudaMemcpyAsync(A->va, va_in + rp_in[0], sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
	if (nranks == 1) { A->t_rp = A->rp; A->t_ci = A->ci; A->t_va = A->va; A->t_shared = 1; }   // symmetric by contract
	if (nranks > 1) {
		const int ns = A->send_off[A->nnbr];
		B200_CUDA(cudaMalloc(&A->send_rows_dev, sizeof(int) * (size_t)(ns > 0 ? ns : 1)));
		B200_CUDA(cudaMemcpyAsync(A->send_rows_dev, A->send_rows, sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice, st));
		b200_note_halo_capacity(A->ncols_global, A->nhalo);
	}
	int bad = 0;
	if (b200k_local_rows_finish(A, rp_in[0], &bad)) { b200_mat_destroy(A); return 1; }
	int bad_any = bad;
	if (nranks > 1) {
		double *buf = (double *)b200_scratch(3, sizeof(double) * 2), hb = bad ? 1.0 : 0.0;
		if (!buf || b200k_h2d(buf, &hb, sizeof(double)) || b200k_allreduce_sum(buf, 1) || b200k_d2h(&hb, buf, sizeof(double))) {
			b200_mat_destroy(A); return 1;
		}
		bad_any = hb != 0.0;
	}
	if (bad_any) {
		b200_mat_destroy(A);
		return b200_fail(bad ? "b200_mat_create_from_local_rows: columns of some row not ascending / out of range, or row pointers not monotone"
		                     : "b200_mat_create_from_local_rows: another rank's rows are malformed");
	}
	if (b200k_dia_build_device(A) || b200k_lat_detect(A)) { b200_mat_destroy(A); return 1; }
	if (nranks > 1 && p2p_register_for(A)) { b200_mat_destroy(A); return 1; }
	*out = A;
	return 0;
}

extern "C" int b200_mat_destroy(b200_mat *A)
{
	if (!A) return 0;
	if (g_b200.initialised) cudaStreamSynchronize(g_b200.stream);
	if (!A->t_shared && A->t_rp) { cudaFree(A->t_rp); cudaFree(A->t_ci); cudaFree(A->t_va); }
	cudaFree(A->rp); cudaFree(A->ci); cudaFree(A->va);
	if (A->send_rows_dev) cudaFree(A->send_rows_dev);
	if (A->dia_off) cudaFree(A->dia_off);
	if (A->dia_grp) cudaFree(A->dia_grp);
	if (A->dia_val) cudaFree(A->dia_val);
	free(A->nbr); free(A->halo_cols); free(A->recv_off); free(A->send_off); free(A->send_rows);
	free(A);
	return 0;
}

extern "C" int b200_mat_shape(const b200_mat *A, int *nrows, int *ncols, int *nnz)
{
	B200_CHECK(A, "b200_mat_shape: NULL matrix");
	if (nrows) *nrows = A->nrows_global;
	if (ncols) *ncols = A->ncols_global;
	if (nnz) *nnz = A->nnz_global;
	return 0;
}

extern "C" int b200_mat_storage(const b200_mat *A, int *dia_nd, int *lat_s1, int *lat_s2, int *lat_const)
{
	B200_CHECK(A, "b200_mat_storage: NULL matrix");
	if (dia_nd) *dia_nd = A->dia_nd;
	if (lat_s1) *lat_s1 = A->lat_s1;
	if (lat_s2) *lat_s2 = A->lat_s2;
	if (lat_const) *lat_const = A->lat_const;
	return 0;
}

extern "C" int b200_mat_local_range(const b200_mat *A, int *row0, int *nrows_local, int *nnz_local, int *nhalo)
{
	B200_CHECK(A, "b200_mat_local_range: NULL matrix");
	if (row0) *row0 = A->row0;
	if (nrows_local) *nrows_local = A->nrows;
	if (nnz_local) *nnz_local = A->nnz;
	if (nhalo) *nhalo = A->nhalo;
	return 0;
}

// This rank's CSR row slab from the device, with GLOBAL column indices.  Concatenating the slabs
// of all ranks in rank order gives the CSR image of the whole matrix (== CCS of A^T): the
// bit-exact round trip of the partition.
extern "C" int b200_mat_local_csr(const b200_mat *A, int *rp, int *ci, double *va)
{
	B200_REQUIRE_INIT();
	B200_CHECK(A && rp, "b200_mat_local_csr: bad arguments");
	cudaStream_t st = g_b200.stream;
	B200_CUDA(cudaMemcpyAsync(rp, A->rp, sizeof(int) * ((size_t)A->nrows + 1), cudaMemcpyDeviceToHost, st));
	if (A->nnz) {
		B200_CUDA(cudaMemcpyAsync(ci, A->ci, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaMemcpyAsync(va, A->va, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost, st));
	}
	B200_CUDA(cudaStreamSynchronize(st));
	if (A->nhalo || A->row0)
		for (int e = 0; e < A->nnz; ++e)
			ci[e] = (A->halo_contiguous || ci[e] < A->nrows) ? ci[e] + A->row0 : A->halo_cols[ci[e] - A->nrows];
	return 0;
}

extern "C" int b200_mat_to_ccs(const b200_mat *A, int *j_col, int *i_row, double *data)
{
	B200_REQUIRE_INIT();
	B200_CHECK(A && j_col, "b200_mat_to_ccs: bad arguments");
	B200_CHECK(A->t_rp, "b200_mat_to_ccs: the matrix is row-partitioned over %d ranks; gather the slabs with "
	           "b200_mat_local_csr", g_b200.nranks);
	cudaStream_t st = g_b200.stream;
	B200_CUDA(cudaMemcpyAsync(j_col, A->t_rp, sizeof(int) * ((size_t)A->ncols + 1), cudaMemcpyDeviceToHost, st));
	if (A->nnz) {
		B200_CUDA(cudaMemcpyAsync(i_row, A->t_ci, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaMemcpyAsync(data, A->t_va, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost, st));
	}
	B200_CUDA(cudaStreamSynchronize(st));
	return 0;
}

// ---- Y = alpha X + beta Y, slot MatAxpby (reference src/ops.h:52; used by GCG for the in-place shift
// A + sigma B, src/ops_eig_sol_gcg.c:594-602).  The pattern of X must be a SUBSET of Y's (the reference's
// SLEPc back end passes SUBSET_NONZERO_PATTERN, app/app_slepc.c; e.g. A = stiffness, B = lumped mass).
// Everything is validated before anything is modified.

// one thread per row: merge-walk the ascending column lists of X's and Y's row; pos[e] = index of X's entry
// e in Y's arrays, or -1 (counted in *missing)
__global__ void mat_locate_kernel(int nrows, const int *__restrict__ xrp, const int *__restrict__ xci,
                                  const int *__restrict__ yrp, const int *__restrict__ yci, int *__restrict__ pos,
                                  int *missing)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nrows) return;
	int ey = yrp[r];
	const int ey1 = yrp[r + 1];
	int miss = 0;
	for (int ex = xrp[r]; ex < xrp[r + 1]; ++ex) {
		const int c = xci[ex];
		while (ey < ey1 && yci[ey] < c) ++ey;
		if (ey < ey1 && yci[ey] == c) pos[ex] = ey;
		else { pos[ex] = -1; ++miss; }
	}
	if (miss) atomicAdd(missing, miss);
}

__global__ void mat_scale_kernel(long long cnt, double beta, double *__restrict__ y)
{
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x)
		y[i] = beta * y[i];
}

// same operation order as a BLAS dscal followed by daxpy
__global__ void mat_scatter_add_kernel(int nnz, double alpha, const double *__restrict__ x, const int *__restrict__ pos,
                                       double *__restrict__ y)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += gridDim.x * blockDim.x)
		y[pos[i]] = fma(alpha, x[i], y[pos[i]]);
}

// the diagonal image of Y: entry (r, c) of X lands in slot start_g + (c - r - off_g) of row r
__global__ void mat_dia_add_kernel(int nrows, const int *__restrict__ xrp, const int *__restrict__ xci,
                                   const double *__restrict__ xva, double alpha, int ng, const int *__restrict__ off,
                                   const int *__restrict__ grp, int ndp, double *__restrict__ dia, int *missing)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nrows) return;
	for (int e = xrp[r]; e < xrp[r + 1]; ++e) {
		const int d = xci[e] - r;
		int slot = -1;
		for (int g = 0; g < ng; ++g)
			if (d >= off[g] && d < off[g] + grp[2 * g + 1]) slot = grp[2 * g] + (d - off[g]);
		if (slot < 0) { if (missing) atomicAdd(missing, 1); continue; }
		if (dia) dia[(size_t)r * ndp + slot] = fma(alpha, xva[e], dia[(size_t)r * ndp + slot]);
	}
}

extern "C" int b200_mat_axpby(double alpha, const b200_mat *X, double beta, b200_mat *Y)
{
	B200_REQUIRE_INIT();
	B200_CHECK(X && Y, "b200_mat_axpby: NULL matrix");
	B200_CHECK(X->nrows == Y->nrows && X->ncols == Y->ncols && X->row0 == Y->row0 &&
	           X->nrows_global == Y->nrows_global, "b200_mat_axpby: shapes differ");
	if (b200_multi()) {
		// the local column numbering must agree: contiguous halos number columns by global index; halo
		// lists number them by slot, so the lists must be the same
		bool same = X->halo_contiguous == Y->halo_contiguous && X->halo_below <= Y->halo_below;
		if (same && !X->halo_contiguous) {
			same = X->nhalo == Y->nhalo;
			for (int i = 0; same && i < X->nhalo; ++i) same = X->halo_cols[i] == Y->halo_cols[i];
		}
		B200_CHECK(same, "b200_mat_axpby: across ranks the two matrices need the same halo numbering");
	}
	cudaStream_t st = g_b200.stream;
	const int threads = 256;
	const int row_blocks = b200_ceil_div(Y->nrows > 0 ? Y->nrows : 1, threads);
	// ---- validate: every entry of X must exist in Y (CSR image, transpose image, diagonal image)
	int *pos = (int *)b200_scratch(2, sizeof(int) * ((size_t)2 * (X->nnz > 0 ? X->nnz : 1) + 4));
	if (!pos) return 1;
	int *pos_t = pos + (X->nnz > 0 ? X->nnz : 1), *missing = pos_t + (X->nnz > 0 ? X->nnz : 1);
	B200_CUDA(cudaMemsetAsync(missing, 0, sizeof(int), st));
	const bool own_t = !Y->t_shared && Y->t_va;
	if (X->nnz > 0) {
		mat_locate_kernel<<<row_blocks, threads, 0, st>>>(X->nrows, X->rp, X->ci, Y->rp, Y->ci, pos, missing);
		B200_KERNEL_CHECK();
		if (own_t) {
			B200_CHECK(X->t_rp && X->t_ci && X->t_va, "b200_mat_axpby: internal (no transpose image of X)");
			mat_locate_kernel<<<b200_ceil_div(X->ncols, threads), threads, 0, st>>>(X->ncols, X->t_rp, X->t_ci, Y->t_rp, Y->t_ci,
			                                                                     pos_t, missing);
			B200_KERNEL_CHECK();
		}
		if (Y->dia_nd) {
			mat_dia_add_kernel<<<row_blocks, threads, 0, st>>>(X->nrows, X->rp, X->ci, X->va, 0.0, Y->dia_ng, Y->dia_off,
			                                                  Y->dia_grp, Y->dia_ndp, nullptr, missing);
			B200_KERNEL_CHECK();
		}
	}
	int missing_h = 0;
	B200_CUDA(cudaMemcpyAsync(&missing_h, missing, sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	B200_CHECK(missing_h == 0, "b200_mat_axpby: %d entries of X have no counterpart in Y (the pattern of X must be a "
	           "subset of Y's, as SUBSET_NONZERO_PATTERN in the reference's SLEPc back end); Y is unchanged", missing_h);
	// ---- update
	const int sm8 = g_b200.num_sms * 8;
	auto blocks_for = [&](long long cnt) { long long b = (cnt + threads - 1) / threads; return (int)(b < sm8 ? (b > 0 ? b : 1) : sm8); };
	const long long dia_cnt = Y->dia_nd ? (long long)Y->nrows * Y->dia_ndp : 0;
	if (beta != 1.0) {
		if (Y->nnz > 0) { mat_scale_kernel<<<blocks_for(Y->nnz), threads, 0, st>>>(Y->nnz, beta, Y->va); B200_KERNEL_CHECK(); }
		if (own_t && Y->nnz > 0) { mat_scale_kernel<<<blocks_for(Y->nnz), threads, 0, st>>>(Y->nnz, beta, Y->t_va); B200_KERNEL_CHECK(); }
		if (dia_cnt > 0) { mat_scale_kernel<<<blocks_for(dia_cnt), threads, 0, st>>>(dia_cnt, beta, Y->dia_val); B200_KERNEL_CHECK(); }
	}
	if (X->nnz > 0 && alpha != 0.0) {
		mat_scatter_add_kernel<<<blocks_for(X->nnz), threads, 0, st>>>(X->nnz, alpha, X->va, pos, Y->va);
		B200_KERNEL_CHECK();
		if (own_t) {
			mat_scatter_add_kernel<<<blocks_for(X->nnz), threads, 0, st>>>(X->nnz, alpha, X->t_va, pos_t, Y->t_va);
			B200_KERNEL_CHECK();
		}
		if (dia_cnt > 0) {
			mat_dia_add_kernel<<<row_blocks, threads, 0, st>>>(X->nrows, X->rp, X->ci, X->va, alpha, Y->dia_ng, Y->dia_off,
			                                                  Y->dia_grp, Y->dia_ndp, Y->dia_val, nullptr);
			B200_KERNEL_CHECK();
		}
	}
	// the values changed: the constant-stencil coefficients kept on the host (lattice SpMM without value loads)
	// must follow them, and a constant stencil may have stopped being one (or have become one)
	return b200k_lat_detect(Y);
}
