/* TEST INFRASTRUCTURE ONLY -- drives the reference's OWN L3 code over OPS_B200_Set.
 *
 * Built (where the reference headers exist) into oracle/_ref/libdrive_b200.so and loaded by
 * tests/ after libgcge_ref.so (the unmodified reference) and libgcge_b200_ops.so (the B200
 * OPS adaptor).  drive_gcg_b200() repeats the call sequence of the reference driver
 * TestEigenSolverGCG (reference test/test_eig_sol_gcg.c:28-169) with one change: the ops
 * table comes from OPS_B200_Set and the matrices are B200MAT.  tier == 0 runs the
 * reference's GCG / ops_orth.c / ops_lin_sol.c unchanged over the device slots (the drop-in
 * claim); tier == 1 additionally installs the device GCG through EigenSolverSetup_GCG_B200.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <stdarg.h>

#include "ops.h"
#include "app_ccs.h"
#include "ops_eig_sol_gcg.h"
#include "app_b200.h"

static void quiet_printf(const char *fmt, ...) { (void)fmt; }

int drive_gcg_b200(int tier, int n,
		const int *A_j_col, const int *A_i_row, const double *A_data,
		const int *B_j_col, const int *B_i_row, const double *B_data,
		int nevConv, int nevMax, int block_size, int nevInit,
		double tol_abs, double tol_rel, int max_iter_gcg,
		int argc, char **argv, int quiet,
		double *eval_out, double *evec_out,
		int *numIter_out, int *nevConv_out, double *seconds_out,
		int nevGiven, const double *evec_given)
{
	OPS *ops = NULL;
	OPS_Create(&ops);
	OPS_B200_Set(ops);
	OPS_Setup(ops);
	if (quiet) { ops->Printf = quiet_printf; ops->lapack_ops->Printf = quiet_printf; }

	CCSMAT ccsA, ccsB; B200MAT mA, mB; void *A, *B = NULL;
	ccsA.nrows = n; ccsA.ncols = n; ccsA.j_col = (int *)A_j_col; ccsA.i_row = (int *)A_i_row; ccsA.data = (double *)A_data;
	B200_MatCreateFromCCS(&mA, &ccsA); A = (void *)&mA;
	if (B_j_col) {
		ccsB.nrows = n; ccsB.ncols = n; ccsB.j_col = (int *)B_j_col; ccsB.i_row = (int *)B_i_row; ccsB.data = (double *)B_data;
		B200_MatCreateFromCCS(&mB, &ccsB); B = (void *)&mB;
	}

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
	/* warm start (reference src/ops_eig_sol_gcg.c:107-109) */
	if (nevGiven > 0 && evec_given != NULL) {
		if (b200_mv_upload((b200_mv *)evec, 0, nevGiven, evec_given, n)) {
			fprintf(stderr, "drive_gcg_b200: %s\n", b200_last_error());
			abort();
		}
	} else nevGiven = 0;
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
	double t0 = ops->GetWtime();
	if (tier == 1)
		EigenSolverSetup_GCG_B200(multiMax, gapMin, nevInit, nevMax, block_size,
				tol_gcg, max_iter_gcg, 0, gcg_mv_ws, dbl_ws, int_ws, ops);
	else
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
	B200_SetOptionsFromCommandLine(argc, argv, ops);
	ops->EigenSolver(A, B, eval, evec, nevGiven, &nevConv, ops);
	double t1 = ops->GetWtime();

	if (numIter_out) *numIter_out = ((GCGSolver *)ops->eigen_solver_workspace)->numIter;
	if (nevConv_out) *nevConv_out = nevConv;
	if (seconds_out) *seconds_out = t1 - t0;
	memcpy(eval_out, eval, nevMax * sizeof(double));
	if (evec_out != NULL) {
		if (b200_mv_download((b200_mv *)evec, 0, nevMax, evec_out, n)) {
			fprintf(stderr, "drive_gcg_b200: %s\n", b200_last_error());
			abort();
		}
	}
	ops->MultiVecDestroy(&gcg_mv_ws[0], nevMax + 2 * block_size, ops);
	for (int i = 1; i < 4; ++i) ops->MultiVecDestroy(&gcg_mv_ws[i], block_size, ops);
	ops->MultiVecDestroy(&evec, nevMax, ops);
	free(dbl_ws); free(int_ws); free(eval);
	B200_MatDestroy(&mA);
	if (B) B200_MatDestroy(&mB);
	OPS_Destroy(&ops);
	return 0;
}

/* ---- the reference's BlockAMG (src/ops_lin_sol.c:466-715) UNCHANGED over OPS_B200_Set ------
 * Same call sequence as ref_block_amg_dense (oracle/ref_driver.c) and the reference's TestMultiGrid
 * (test/test_multi_grid.c:95-125), with the hierarchy as device matrices: A_l square, P_l rectangular
 * (n_l x n_{l+1}); restriction is the true transposed multiply of the device back end. */
#include "ops_lin_sol.h"
int drive_block_amg_b200(int num_levels, const int *n_l,
		int **A_j_col, int **A_i_row, double **A_data,
		int **P_j_col, int **P_i_row, double **P_data, int k,
		const double *b, double *x, int *max_iter, double *rate, double *tol)
{
	OPS *ops = NULL;
	OPS_Create(&ops);
	OPS_B200_Set(ops);
	OPS_Setup(ops);
	ops->Printf = quiet_printf; ops->lapack_ops->Printf = quiet_printf;
	B200MAT *A = calloc(num_levels, sizeof(B200MAT)), *P = calloc(num_levels, sizeof(B200MAT));
	void **A_array = calloc(num_levels, sizeof(void *)), **P_array = calloc(num_levels, sizeof(void *));
	for (int l = 0; l < num_levels; ++l) {
		CCSMAT c;
		c.nrows = n_l[l]; c.ncols = n_l[l]; c.j_col = A_j_col[l]; c.i_row = A_i_row[l]; c.data = A_data[l];
		B200_MatCreateFromCCS(&A[l], &c);
		A_array[l] = (void *)&A[l];
		if (l + 1 < num_levels) {
			c.nrows = n_l[l]; c.ncols = n_l[l + 1]; c.j_col = P_j_col[l]; c.i_row = P_i_row[l]; c.data = P_data[l];
			B200_MatCreateFromCCS(&P[l], &c);
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
	void **Bv, **Xv;
	ops->MultiVecCreateByMat(&Bv, k, A_array[0], ops);
	ops->MultiVecCreateByMat(&Xv, k, A_array[0], ops);
	if (b200_mv_upload((b200_mv *)Bv, 0, k, b, n_l[0]) || b200_mv_upload((b200_mv *)Xv, 0, k, x, n_l[0])) {
		fprintf(stderr, "drive_block_amg_b200: %s\n", b200_last_error());
		abort();
	}
	int start[2] = {0, 0}, end[2] = {k, k};
	MultiLinearSolverSetup_BlockAMG(max_iter, rate, tol, "abs", A_array, P_array, num_levels,
			mv_ws, dbl_ws, int_ws, NULL, ops);
	ops->MultiLinearSolver(A_array[0], Bv, Xv, start, end, ops);
	if (b200_mv_download((b200_mv *)Xv, 0, k, x, n_l[0])) {
		fprintf(stderr, "drive_block_amg_b200: %s\n", b200_last_error());
		abort();
	}
	ops->MultiVecDestroy(&Bv, k, ops); ops->MultiVecDestroy(&Xv, k, ops);
	for (int i = 0; i < 5; ++i) {
		for (int l = 0; l < num_levels; ++l) ops->MultiVecDestroy(&mv_ws[i][l], k, ops);
		free(mv_ws[i]);
	}
	for (int l = 0; l < num_levels; ++l) { B200_MatDestroy(&A[l]); if (l + 1 < num_levels) B200_MatDestroy(&P[l]); }
	free(dbl_ws); free(int_ws); free(A); free(P); free(A_array); free(P_array);
	OPS_Destroy(&ops);
	return 0;
}
