/* app_b200.h -- the B200 OPS back end, beside app_ccs.h / app_lapack.h.
 *
 * Drop this file and app_b200.c into the reference's app/ directory (or compile them
 * with -I<reference>/src -I<reference>/app) and link libgcge_b200.so.  The reference's
 * GCG (src/ops_eig_sol_gcg.c), orthogonalisation (src/ops_orth.c) and BlockPCG
 * (src/ops_lin_sol.c) then run unchanged on device-resident data.  See INTEGRATION.md.
 */
#ifndef _APP_B200_H_
#define _APP_B200_H_

#include "ops.h"          /* reference src/ops.h: struct OPS_ */
#include "app_ccs.h"      /* reference app/app_ccs.h: CCSMAT */
#include "gcge_b200.h"

/* Matrix handle handed to every slot as `void *mat`.  The reference has no
 * matrix-creation slot -- each app's driver builds its own type directly
 * (reference test/test_app_ccs.c:99-102) -- so the B200 app exports constructors. */
typedef struct B200MAT_ {
	b200_mat *dev;        /* device CSR image */
	int nrows, ncols;
} B200MAT;

void B200_MatCreateFromCCS(B200MAT *mat, const CCSMAT *ccs);
void B200_MatDestroy(B200MAT *mat);

/* Fill the OPS table: the distributed-style minimum slot set (reference
 * app/app_slepc.c:610-634) plus MultiVecInnerProd / MultiVecQtAP; creates a genuine
 * host lapack_ops exactly like OPS_CCS_Set (reference app/app_ccs.c:215-217). */
void OPS_B200_Set(struct OPS_ *ops);

/* -b200_<name> <int> on the command line -> the library's run-time switches (b200_option_set), read through the
 * table's own GetOptionFromCommandLine like every other option of the reference (src/ops_multi_vec.c:58-95,
 * src/ops_eig_sol_gcg.c:1737); call next to EigenSolverSetParametersFromCommandLine_GCG.  Returns the number set. */
int B200_SetOptionsFromCommandLine(int argc, char *argv[], struct OPS_ *ops);

/* Tier B: install the fused device providers in the three L3 slots.  Signatures mirror
 * the reference setups (reference src/ops_lin_sol.h:41-45, src/ops_orth.h:36-41,
 * src/ops_eig_sol_gcg.h:54-60); workspace arguments the device code does not need are
 * accepted and ignored so call sites need not change. */
void MultiLinearSolverSetup_BlockPCG_B200(int max_iter, double rate, double tol,
		const char *tol_type, void **mv_ws[3], double *dbl_ws, int *int_ws,
		void *pc, void *unused_matdot, struct OPS_ *ops);
void MultiVecOrthSetup_ModifiedGramSchmidt_B200(int block_size, int max_reorth,
		double orth_zero_tol, void **mv_ws, double *dbl_ws, struct OPS_ *ops);
void EigenSolverSetup_GCG_B200(int multiMax, double gapMin, int nevInit, int nevMax,
		int block_size, double tol[2], int numIterMax,
		int user_defined_multi_linear_solver,
		void **mv_ws[4], double *dbl_ws, int *int_ws, struct OPS_ *ops);

#endif
