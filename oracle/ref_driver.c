/* TEST INFRASTRUCTURE ONLY -- C entry points over the UNMODIFIED reference.
 *
 * This file is linked with the reference's own sources (compiled where they
 * lie under /root/reference by oracle/Makefile) into oracle/_ref/libgcge_ref.so.
 * It contains no numerical code of its own: every function builds the
 * reference's data types (CCSMAT reference app/app_ccs.h:20-24, LAPACKVEC
 * reference app/app_lapack.h:17-20) around caller-owned arrays and calls
 * through the reference's own `struct OPS_` table as filled by OPS_CCS_Set
 * (reference app/app_ccs.c:213-249).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the library.
 *
 * ref_gcg_solve() follows the call sequence of the reference's driver
 * TestEigenSolverGCG (reference test/test_eig_sol_gcg.c:28-169): same
 * workspace creation order, same random fills before srand(0), same setup
 * and parameter calls, so the rand() stream that seeds X is the reference's.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <stdarg.h>
#include <omp.h>

#include "ops.h"
#include "app_ccs.h"
#include "app_lapack.h"
#include "ops_orth.h"
#include "ops_lin_sol.h"
#include "ops_eig_sol_gcg.h"

int gcge_ref_omp_threads = 1;

static void quiet_printf(const char *fmt, ...) { (void)fmt; }

int ref_set_threads(int n)
{
	if (n < 1) n = 1;
	gcge_ref_omp_threads = n;
	return gcge_ref_omp_threads;
}
int ref_get_max_threads(void) { return omp_get_num_procs(); }

void ref_srand(unsigned seed) { srand(seed); }

static OPS *make_ops(int quiet)
{
	OPS *ops = NULL;
	OPS_Create(&ops);
	OPS_CCS_Set(ops);
	OPS_Setup(ops);
	if (quiet) {
		ops->Printf = quiet_printf;
		ops->lapack_ops->Printf = quiet_printf;
	}
	return ops;
}

static void wrap_ccs(CCSMAT *m, int n, const int *j_col, const int *i_row, const double *data)
{
	m->nrows = n; m->ncols = n;
	m->j_col = (int *)j_col; m->i_row = (int *)i_row; m->data = (double *)data;
}
static void wrap_mv(LAPACKVEC *v, double *data, int nrows, int ncols)
{
	v->data = data; v->nrows = nrows; v->ncols = ncols; v->ldd = nrows;
}

/* ---- slot-level entry points (kernel parity) ---------------------------- */

void ref_mat_dot_multivec(int n, const int *j_col, const int *i_row, const double *data,
		double *x, int ncols_x, double *y, int ncols_y, int *start, int *end)
{
	OPS *ops = make_ops(1);
	CCSMAT A; LAPACKVEC X, Y;
	wrap_mv(&X, x, n, ncols_x); wrap_mv(&Y, y, n, ncols_y);
	if (j_col != NULL) {
		wrap_ccs(&A, n, j_col, i_row, data);
		ops->MatDotMultiVec((void *)&A, (void **)&X, (void **)&Y, start, end, ops);
	} else {
		ops->MatDotMultiVec(NULL, (void **)&X, (void **)&Y, start, end, ops);
	}
	OPS_Destroy(&ops);
}

void ref_multivec_axpby(int n, double alpha, double *x, int ncols_x, double beta,
		double *y, int ncols_y, int *start, int *end)
{
	OPS *ops = make_ops(1);
	LAPACKVEC X, Y;
	wrap_mv(&Y, y, n, ncols_y);
	if (x != NULL) {
		wrap_mv(&X, x, n, ncols_x);
		ops->MultiVecAxpby(alpha, (void **)&X, beta, (void **)&Y, start, end, ops);
	} else {
		ops->MultiVecAxpby(alpha, NULL, beta, (void **)&Y, start, end, ops);
	}
	OPS_Destroy(&ops);
}

void ref_multivec_linear_comb(int n, double *x, int ncols_x, double *y, int ncols_y,
		int *start, int *end, double *coef, int ldc, double *beta, int incb)
{
	OPS *ops = make_ops(1);
	LAPACKVEC X, Y;
	wrap_mv(&Y, y, n, ncols_y);
	if (x != NULL) wrap_mv(&X, x, n, ncols_x);
	ops->MultiVecLinearComb(x != NULL ? (void **)&X : NULL, (void **)&Y, 0,
			start, end, coef, ldc, beta, incb, ops);
	OPS_Destroy(&ops);
}

void ref_multivec_inner_prod(int n, char nsdIP, double *x, int ncols_x, double *y, int ncols_y,
		int *start, int *end, double *inner_prod, int ldIP)
{
	OPS *ops = make_ops(1);
	LAPACKVEC X, Y;
	wrap_mv(&X, x, n, ncols_x); wrap_mv(&Y, y, n, ncols_y);
	ops->MultiVecInnerProd(nsdIP, (void **)&X, (void **)&Y, 0, start, end, inner_prod, ldIP, ops);
	OPS_Destroy(&ops);
}

void ref_multivec_qtap(int n, char ntsA, char ntsdQAP,
		double *q, int ncols_q,
		const int *j_col, const int *i_row, const double *data,
		double *p, int ncols_p, int *start, int *end,
		double *qAp, int ldQAP, double *ws, int ncols_ws)
{
	OPS *ops = make_ops(1);
	CCSMAT A; LAPACKVEC Q, P, W;
	wrap_mv(&Q, q, n, ncols_q); wrap_mv(&P, p, n, ncols_p); wrap_mv(&W, ws, n, ncols_ws);
	if (j_col != NULL) wrap_ccs(&A, n, j_col, i_row, data);
	ops->MultiVecQtAP(ntsA, ntsdQAP, (void **)&Q, j_col != NULL ? (void *)&A : NULL,
			(void **)&P, 0, start, end, qAp, ldQAP, (void **)&W, ops);
	OPS_Destroy(&ops);
}

void ref_multivec_set_random(int n, double *x, int ncols_x, int start, int end)
{
	OPS *ops = make_ops(1);
	LAPACKVEC X;
	wrap_mv(&X, x, n, ncols_x);
	ops->MultiVecSetRandomValue((void **)&X, start, end, ops);
	OPS_Destroy(&ops);
}

/* method: 0 = ModifiedGramSchmidt, 1 = BinaryGramSchmidt (reference src/ops_orth.h:36-41).
 * ws must hold n*ncols_ws doubles, dbl_ws as the reference sizes it. */
void ref_multivec_orth(int n, int method, int block_size, int max_reorth, double orth_zero_tol,
		double *x, int ncols_x, int start_x, int *end_x,
		const int *j_col, const int *i_row, const double *data,
		double *ws, int ncols_ws, double *dbl_ws)
{
	OPS *ops = make_ops(1);
	CCSMAT B; LAPACKVEC X, W;
	wrap_mv(&X, x, n, ncols_x); wrap_mv(&W, ws, n, ncols_ws);
	if (j_col != NULL) wrap_ccs(&B, n, j_col, i_row, data);
	if (method == 1)
		MultiVecOrthSetup_BinaryGramSchmidt(block_size, max_reorth, orth_zero_tol,
				(void **)&W, dbl_ws, ops);
	else
		MultiVecOrthSetup_ModifiedGramSchmidt(block_size, max_reorth, orth_zero_tol,
				(void **)&W, dbl_ws, ops);
	ops->MultiVecOrth((void **)&X, start_x, end_x, j_col != NULL ? (void *)&B : NULL, ops);
	OPS_Destroy(&ops);
}

/* BlockPCG (reference src/ops_lin_sol.c:140-437).  r,p,w: three n*ncols_ws work blocks. */
void ref_block_pcg(int n, const int *j_col, const int *i_row, const double *data,
		double *b, int ncols_b, double *x, int ncols_x, int *start, int *end,
		int max_iter, double rate, double tol, const char *tol_type,
		double *r, double *p, double *w, int ncols_ws,
		int *niter, double *residual)
{
	OPS *ops = make_ops(1);
	CCSMAT A; LAPACKVEC Bv, Xv, R, P, W;
	void **mv_ws[3];
	int k = end[0] - start[0];
	double *dbl_ws = calloc(6 * (size_t)k + 8, sizeof(double));
	int *int_ws = calloc(2 * (size_t)k + 8, sizeof(int));
	wrap_ccs(&A, n, j_col, i_row, data);
	wrap_mv(&Bv, b, n, ncols_b); wrap_mv(&Xv, x, n, ncols_x);
	wrap_mv(&R, r, n, ncols_ws); wrap_mv(&P, p, n, ncols_ws); wrap_mv(&W, w, n, ncols_ws);
	mv_ws[0] = (void **)&R; mv_ws[1] = (void **)&P; mv_ws[2] = (void **)&W;
	MultiLinearSolverSetup_BlockPCG(max_iter, rate, tol, tol_type, mv_ws, dbl_ws, int_ws,
			NULL, NULL, ops);
	ops->MultiLinearSolver((void *)&A, (void **)&Bv, (void **)&Xv, start, end, ops);
	if (niter) *niter = ((BlockPCGSolver *)ops->multi_linear_solver_workspace)->niter;
	if (residual) *residual = ((BlockPCGSolver *)ops->multi_linear_solver_workspace)->residual;
	free(dbl_ws); free(int_ws);
	OPS_Destroy(&ops);
}

/* ---- whole solve -------------------------------------------------------- */

/* Follows reference test/test_eig_sol_gcg.c:28-169.  B_j_col == NULL => standard problem.
 * eval: nevMax doubles; evec: n*nevMax doubles column-major (may be NULL).
 * argc/argv are handed to EigenSolverSetParametersFromCommandLine_GCG
 * (reference src/ops_eig_sol_gcg.c:1737) exactly as the reference driver does. */
int ref_gcg_solve(int n,
		const int *A_j_col, const int *A_i_row, const double *A_data,
		const int *B_j_col, const int *B_i_row, const double *B_data,
		int nevConv, int nevMax, int block_size, int nevInit,
		double tol_abs, double tol_rel, int max_iter_gcg,
		int argc, char **argv, int quiet,
		double *eval_out, double *evec_out,
		int *numIter_out, int *nevConv_out, double *seconds_out,
		int nevGiven, const double *evec_given)
{
	OPS *ops = make_ops(quiet);
	CCSMAT ccsA, ccsB; void *A, *B = NULL;
	wrap_ccs(&ccsA, n, A_j_col, A_i_row, A_data); A = (void *)&ccsA;
	if (B_j_col != NULL) { wrap_ccs(&ccsB, n, B_j_col, B_i_row, B_data); B = (void *)&ccsB; }

	int multiMax = 1; double gapMin = 1e-5;
	if (nevMax <= 0) nevMax = 2 * nevConv;
	if (block_size <= 0) block_size = nevConv < 30 ? (nevMax - nevConv) : nevConv / 5;
	if (nevInit <= 0) nevInit = nevMax;
	nevInit = nevInit < nevMax ? nevInit : nevMax;
	double tol_gcg[2]; tol_gcg[0] = tol_abs; tol_gcg[1] = tol_rel;

	double *eval = calloc(nevMax, sizeof(double));
	void **evec;
	ops->MultiVecCreateByMat(&evec, nevMax, A, ops);
	ops->MultiVecSetRandomValue(evec, 0, nevMax, ops);
	/* warm start (reference src/ops_eig_sol_gcg.c:107-109): the caller's approximate eigenvectors
	 * in the first nevGiven columns of evec */
	if (nevGiven > 0 && evec_given != NULL)
		memcpy(((LAPACKVEC *)evec)->data, evec_given, (size_t)n * nevGiven * sizeof(double));
	else nevGiven = 0;
	void **gcg_mv_ws[4]; double *dbl_ws; int *int_ws;
	ops->MultiVecCreateByMat(&gcg_mv_ws[0], nevMax + 2 * block_size, A, ops);
	ops->MultiVecSetRandomValue(gcg_mv_ws[0], 0, nevMax + 2 * block_size, ops);
	for (int i = 1; i < 4; ++i) {
		ops->MultiVecCreateByMat(&gcg_mv_ws[i], block_size, A, ops);
		ops->MultiVecSetRandomValue(gcg_mv_ws[i], 0, block_size, ops);
	}
	int sizeV = nevInit + 2 * block_size;
	size_t length_dbl_ws = 2 * (size_t)sizeV * sizeV + 10 * (size_t)sizeV
		+ (nevMax + 2 * block_size) + (size_t)nevMax * block_size;
	size_t length_int_ws = 6 * (size_t)sizeV + 2 * (block_size + 3);
	dbl_ws = calloc(length_dbl_ws, sizeof(double));
	int_ws = calloc(length_int_ws, sizeof(int));

	srand(0);
	double t0 = omp_get_wtime();
	EigenSolverSetup_GCG(multiMax, gapMin, nevInit, nevMax, block_size,
			tol_gcg, max_iter_gcg, 0, gcg_mv_ws, dbl_ws, int_ws, ops);
	EigenSolverSetParameters_GCG(
			50,
			"mgs", 80, 2, 2 * DBL_EPSILON,
			"mgs", -1, 2, 2 * DBL_EPSILON,
			"mgs", 80, 2, 2 * DBL_EPSILON,
			30, 1e-2, 1e-14, "abs", 0,
			-1, gapMin, 2 * DBL_EPSILON, ops);
	{
		/* the solver object is a function-static struct of the reference (src/ops_eig_sol_gcg.c:1569):
		 * what EigenSolverSetParameters_GCG does not cover keeps the value a previous call's options
		 * left behind, so put the reference's own defaults (:1585-1596) back before parsing argv */
		GCGSolver *gs = (GCGSolver *)ops->eigen_solver_workspace;
		gs->compW_cg_order = 1; gs->compW_cg_shift = 0.0; gs->compW_cg_auto_shift = 0;
	}
	EigenSolverSetParametersFromCommandLine_GCG(argc, argv, ops);
	ops->EigenSolver(A, B, eval, evec, nevGiven, &nevConv, ops);
	double t1 = omp_get_wtime();

	if (numIter_out) *numIter_out = ((GCGSolver *)ops->eigen_solver_workspace)->numIter;
	if (nevConv_out) *nevConv_out = nevConv;
	if (seconds_out) *seconds_out = t1 - t0;
	memcpy(eval_out, eval, nevMax * sizeof(double));
	if (evec_out != NULL)
		memcpy(evec_out, ((LAPACKVEC *)evec)->data, (size_t)n * nevMax * sizeof(double));

	ops->MultiVecDestroy(&gcg_mv_ws[0], nevMax + 2 * block_size, ops);
	for (int i = 1; i < 4; ++i) ops->MultiVecDestroy(&gcg_mv_ws[i], block_size, ops);
	ops->MultiVecDestroy(&evec, nevMax, ops);
	free(dbl_ws); free(int_ws); free(eval);
	OPS_Destroy(&ops);
	return 0;
}

/* ---- BlockAMG (reference src/ops_lin_sol.c:466-715) over the reference's dense LAPACK back end ----------
 * The CCS back end cannot run it: its MatTransDotMultiVec assumes a square symmetric matrix (reference
 * app/app_ccs.c:140-150) and it has no MultiGridCreate.  The hierarchy comes from the caller: A_dense[l] is
 * n_l x n_l, P_dense[l] is n_l x n_{l+1} (prolongation, fine rows x coarse columns), all column-major.
 * Follows the sequence of the reference's TestMultiGrid (test/test_multi_grid.c:95-125). */
int ref_block_amg_dense(int num_levels, const int *n_l, double **A_dense, double **P_dense, int k,
		double *b, double *x, int *max_iter, double *rate, double *tol)
{
	OPS *ops = NULL;
	OPS_Create(&ops);
	OPS_LAPACK_Set(ops);
	OPS_Setup(ops);
	ops->Printf = quiet_printf;
	LAPACKMAT *A = calloc(num_levels, sizeof(LAPACKMAT)), *P = calloc(num_levels, sizeof(LAPACKMAT));
	void **A_array = calloc(num_levels, sizeof(void *)), **P_array = calloc(num_levels, sizeof(void *));
	for (int l = 0; l < num_levels; ++l) {
		A[l].nrows = n_l[l]; A[l].ncols = n_l[l]; A[l].ldd = n_l[l]; A[l].data = A_dense[l];
		A_array[l] = (void *)&A[l];
		if (l + 1 < num_levels) {
			P[l].nrows = n_l[l]; P[l].ncols = n_l[l + 1]; P[l].ldd = n_l[l]; P[l].data = P_dense[l];
			P_array[l] = (void *)&P[l];
		}
	}
	void ***mv_ws[5];
	for (int i = 0; i < 5; ++i) {
		mv_ws[i] = malloc(num_levels * sizeof(void **));
		for (int l = 0; l < num_levels; ++l) ops->MultiVecCreateByMat(&mv_ws[i][l], k, A_array[l], ops);
	}
	double *dbl_ws = calloc(4096 + 16 * (size_t)k, sizeof(double));
	int *int_ws = calloc(1024 + 4 * (size_t)k, sizeof(int));
	LAPACKVEC Bv, Xv;
	wrap_mv(&Bv, b, n_l[0], k); wrap_mv(&Xv, x, n_l[0], k);
	int start[2] = {0, 0}, end[2] = {k, k};
	MultiLinearSolverSetup_BlockAMG(max_iter, rate, tol, "abs", A_array, P_array, num_levels,
			mv_ws, dbl_ws, int_ws, NULL, ops);
	ops->MultiLinearSolver(A_array[0], (void **)&Bv, (void **)&Xv, start, end, ops);
	for (int i = 0; i < 5; ++i) {
		for (int l = 0; l < num_levels; ++l) ops->MultiVecDestroy(&mv_ws[i][l], k, ops);
		free(mv_ws[i]);
	}
	free(dbl_ws); free(int_ws); free(A); free(P); free(A_array); free(P_array);
	OPS_Destroy(&ops);
	return 0;
}
