/* b200_orth.c -- host driver of the device block orthogonalisation (C-ABI b200_mv_orth).
 *
 * Replaces ModifiedGramSchmidt + OrthSelf (reference src/ops_orth.c:203-393, :45-118).
 * Same block structure as the reference (blocks of block_size columns, each block
 * B-orthonormalised against everything before it and against itself, dependent columns
 * dropped and refilled from the tail, *end_x shrunk), but each block is processed as
 *
 *     twice:  X1 -= X0 (X0^T B X1)          one Gram (s1 x k) + one update, tensor-core tiles
 *             G = X1^T B X1 ; X1 <- X1 T     panel: k x k Gram, Cholesky recurrence with the
 *                                            reference's drop rule on device, one update
 *
 * i.e. block classical Gram-Schmidt with re-orthogonalisation and a Gram/Cholesky panel in
 * place of the column-by-column OrthSelf.  In exact arithmetic the result is the reference's
 * (QR factorisation is unique); in floating point it differs by normalising BEFORE the second
 * projection.  The reference stops re-projecting on an ABSOLUTE test, max|coef| < 50 eps
 * (src/ops_orth.c:262-267), which lets tiny columns (norm ~1e-12, routine for P and W near
 * convergence) keep an O(1e-2) relative component along X0 once they are normalised; the
 * test-suite reproduces that loss of orthogonality with the reference's own ops_orth.c on
 * identical input (DESIGN.md "Orthogonalisation").  The panel costs two reads
 * of the n x k block instead of OrthSelf's k reads.
 */
#include <stdlib.h>
#include <string.h>
#include "b200_dev.h"

#define ORTH_MAX_BLOCK 128
#define LINCOMB_INPLACE_MAX 64      /* LC_BN of b200_dense.cu: output columns of one LinearComb CTA */

static int orth_panel(long long n, double *x1, int ldx, int kb, const b200_mat *B, double *ws, int ldws,
                      double zero_tol, double *g_dev, double *t_dev, int *nlive_dev, int *n_live,
                      const double *scale_in, double *scale_out, int dist)
{
	const double *y = x1; int ldy = ldx;
	if (B) {
		if (b200k_spmm(B, 0, x1, ldx, ws, ldws, kb, NULL)) return 1;
		y = ws; ldy = ldws;
	}
	if (b200k_gram('S', n, kb, kb, 1.0, x1, ldx, y, ldy, g_dev, kb, 1, dist)) return 1;
	if (b200k_chol_drop(kb, g_dev, zero_tol, t_dev, nlive_dev, scale_in, scale_out)) return 1;
	/* X1 <- X1 T (T: element (i,j) at t[i*kb+j]).  A block of at most LINCOMB_INPLACE_MAX columns is one
	 * column tile of the LinearComb kernels: a CTA owns whole rows of X1, reads them completely through its
	 * tile ring and stores them afterwards, so the update runs in place; wider blocks go through the
	 * workspace like the reference's (src/ops_orth.c:108-114) */
	if (kb <= LINCOMB_INPLACE_MAX) {
		if (b200k_lincomb(n, kb, kb, x1, ldx, t_dev, kb, 1, NULL, 0, x1, ldx)) return 1;
	} else {
		if (b200k_lincomb(n, kb, kb, x1, ldx, t_dev, kb, 1, NULL, 0, ws, ldws)) return 1;
		if (b200k_axpby(n, kb, 1.0, ws, ldws, 0.0, x1, ldx)) return 1;
	}
	if (b200k_d2h(n_live, nlive_dev, sizeof(int))) return 1;
	return 0;
}

int b200_mv_orth(b200_mv *x, int start_x, int *end_x, const b200_mat *B,
                 const b200_orth_params *prm, b200_mv *ws)
{
	if (!x || !end_x || !prm || !ws) return b200_fail("b200_mv_orth: bad arguments");
	if (*end_x <= start_x) return 0;
	if (start_x < 0 || *end_x > x->ncols) return b200_fail("b200_mv_orth: range [%d,%d) outside %d columns", start_x, *end_x, x->ncols);
	if (ws->nrows != x->nrows) return b200_fail("b200_mv_orth: workspace row count differs");
	if (B && (B->nrows != x->nrows || B->ncols != x->nrows)) return b200_fail("b200_mv_orth: B is not n x n");
	if (B && b200k_spmm_check_halo(B, x)) return 1;
	const long long n = x->nrows;
	int end = *end_x;
	int init_start = start_x;
	int block = prm->block_size;
	if (block <= 0) {                                       /* reference src/ops_orth.c:275-278 */
		block = (end - init_start) / 2 > 2 ? (end - init_start) / 2 : 2;
	}
	if (block > ORTH_MAX_BLOCK) block = ORTH_MAX_BLOCK;
	if (block > ws->ncols) block = ws->ncols;
	if (block < 1) return b200_fail("b200_mv_orth: workspace has no columns");
	if (block > end - init_start) block = end - init_start;

	/* device scratch: coefficient block (<= ncols x block), G, T, ones, n_live */
	const size_t ncoef = (size_t)x->ncols * ORTH_MAX_BLOCK;
	const size_t npan = (size_t)ORTH_MAX_BLOCK * ORTH_MAX_BLOCK;
	double *base = (double *)b200_scratch(1, sizeof(double) * (ncoef + 2 * npan + 2 * ORTH_MAX_BLOCK + 8) + 64);
	if (!base) return 1;
	double *c_dev = base, *g_dev = base + ncoef, *t_dev = g_dev + npan, *one_dev = t_dev + npan;
	double *sc0_dev = one_dev + 4, *sc1_dev = sc0_dev + ORTH_MAX_BLOCK;
	int *nlive_dev = (int *)(sc1_dev + ORTH_MAX_BLOCK);
	const double one = 1.0;
	if (b200k_h2d(one_dev, &one, sizeof(double))) return 1;

	while (block > 0) {
		const int s1 = init_start;
		int e1 = s1 + block;
		for (int round = 0; round < 2 && e1 > s1; ++round) {
			const int kb = e1 - s1;
			double *x1 = x->d + s1;
			if (s1 > 0) {
				const double *y = x1; int ldy = x->ld;
				if (B) {
					if (b200k_spmm(B, 0, x1, x->ld, ws->d, ws->ld, kb, NULL)) return 1;
					y = ws->d; ldy = ws->ld;
				}
				/* C = -(X0^T B X1), row-major s1 x kb; X1 += X0 C */
				if (b200k_gram('N', n, s1, kb, -1.0, x->d, x->ld, y, ldy, c_dev, kb, 1, x->dist)) return 1;
				if (b200k_lincomb(n, s1, kb, x->d, x->ld, c_dev, kb, 1, one_dev, 0, x1, x->ld)) return 1;
			}
			int n_live = kb;
			if (orth_panel(n, x1, x->ld, kb, B, ws->d, ws->ld, prm->orth_zero_tol, g_dev, t_dev, nlive_dev, &n_live,
			               round == 0 ? NULL : sc0_dev, round == 0 ? sc0_dev : sc1_dev, x->dist))
				return 1;
			e1 = s1 + n_live;
		}
		const int init_end = e1;
		/* refill the dropped slots from the tail, reference src/ops_orth.c:293-307 */
		int length = block - (init_end - s1);
		if (length > end - init_end - length) length = end - init_end - length;
		if (length > 0) {
			if (b200k_axpby(n, length, 1.0, x->d + (end - length), x->ld, 0.0, x->d + init_end, x->ld)) return 1;
		}
		end -= block - (init_end - init_start);
		init_start = init_end;
		if (block > end - init_start) block = end - init_start;
	}
	*end_x = end;
	return 0;
}
