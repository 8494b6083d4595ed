/* app_b200.c -- OPS back end on B200 (host C; all device work behind include/gcge_b200.h).
 *
 * Sits beside the reference's app_ccs.c / app_lapack.c and fills the same
 * `struct OPS_` (reference src/ops.h:43-152).  Multi-vector handles are opaque
 * single pointers (b200_mv*) cast to void**, exactly as the CCS/LAPACK apps cast a
 * LAPACKVEC* (reference app/app_ccs.c:198-203).  Every slot is void; on a device
 * error the message is printed and the process aborts, which is the reference's own
 * failure behaviour (assert(), reference app/app_ccs.c:53-55).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <assert.h>

#include "app_b200.h"
#include "app_lapack.h"
#include "ops_orth.h"
#include "ops_lin_sol.h"
#include "ops_eig_sol_gcg.h"

#define B200_DO(call)                                                          \
	do {                                                                       \
		if ((call) != 0) {                                                     \
			fprintf(stderr, "app_b200: %s\n  in %s\n", b200_last_error(), #call); \
			abort();                                                           \
		}                                                                      \
	} while (0)

#define MV(p) ((b200_mv *)(p))
#define MAT(p) ((p) ? ((B200MAT *)(p))->dev : NULL)

/* ---- matrix ---------------------------------------------------------------- */
void B200_MatCreateFromCCS(B200MAT *mat, const CCSMAT *ccs)
{
	mat->nrows = ccs->nrows; mat->ncols = ccs->ncols;
	B200_DO(b200_mat_create_from_ccs(ccs->nrows, ccs->ncols, ccs->j_col, ccs->i_row, ccs->data, &mat->dev));
}
void B200_MatDestroy(B200MAT *mat)
{
	B200_DO(b200_mat_destroy(mat->dev));
	mat->dev = NULL;
}

static void B200_MatView(void *mat, struct OPS_ *ops)
{
	B200MAT *m = (B200MAT *)mat;
	int nnz = 0, j, e;
	B200_DO(b200_mat_shape(m->dev, NULL, NULL, &nnz));
	int *j_col = malloc(sizeof(int) * (m->ncols + 1)), *i_row = malloc(sizeof(int) * (nnz + 1));
	double *data = malloc(sizeof(double) * (nnz + 1));
	B200_DO(b200_mat_to_ccs(m->dev, j_col, i_row, data));
	for (j = 0; j < m->ncols; ++j)
		for (e = j_col[j]; e < j_col[j + 1]; ++e)
			ops->Printf("(%d,%d) %6.4e\n", i_row[e], j, data[e]);
	free(j_col); free(i_row); free(data);
}
static void B200_MatAxpby(double alpha, void *matX, double beta, void *matY, struct OPS_ *ops)
{
	B200_DO(b200_mat_axpby(alpha, MAT(matX), beta, MAT(matY)));
}

/* ---- multi-vector life cycle ------------------------------------------------- */
static void B200_MultiVecCreateByMat(void ***mv, int num_vec, void *src_mat, struct OPS_ *ops)
{
	b200_mv *x = NULL;
	B200_DO(b200_mv_create(((B200MAT *)src_mat)->ncols, num_vec, &x));
	*mv = (void **)x;
}
static void B200_MultiVecCreateByMultiVec(void ***mv, int num_vec, void **src_mv, struct OPS_ *ops)
{
	b200_mv *x = NULL; int nrows = 0;
	B200_DO(b200_mv_shape(MV(src_mv), &nrows, NULL));
	B200_DO(b200_mv_create(nrows, num_vec, &x));
	*mv = (void **)x;
}
static void B200_MultiVecCreateByVec(void ***mv, int num_vec, void *src_vec, struct OPS_ *ops)
{
	B200_MultiVecCreateByMultiVec(mv, num_vec, (void **)src_vec, ops);
}
static void B200_MultiVecDestroy(void ***mv, int num_vec, struct OPS_ *ops)
{
	B200_DO(b200_mv_destroy(MV(*mv)));
	*mv = NULL;
}
static void B200_GetVecFromMultiVec(void **mv, int col, void **vec, struct OPS_ *ops)
{
	b200_mv *v = NULL;
	B200_DO(b200_mv_view(MV(mv), col, col + 1, &v));
	*vec = (void *)v;
}
static void B200_RestoreVecForMultiVec(void **mv, int col, void **vec, struct OPS_ *ops)
{
	B200_DO(b200_mv_destroy(MV(*vec)));
	*vec = NULL;
}
static void B200_MultiVecView(void **x, int start, int end, struct OPS_ *ops)
{
	int nrows = 0, row, col;
	if (end <= start) return;
	B200_DO(b200_mv_shape(MV(x), &nrows, NULL));
	double *h = malloc(sizeof(double) * (size_t)nrows * (end - start) + 8);
	B200_DO(b200_mv_download(MV(x), start, end, h, nrows));
	for (row = 0; row < nrows; ++row) {
		for (col = 0; col < end - start; ++col) ops->Printf("%6.4e\t", h[(size_t)col * nrows + row]);
		ops->Printf("\n");
	}
	free(h);
}

/* ---- multi-vector slots ------------------------------------------------------ */
static void B200_MultiVecInnerProd(char nsdIP, void **x, void **y, int is_vec, int *start, int *end,
		double *inner_prod, int ldIP, struct OPS_ *ops)
{
	/* one process drives the whole device: local == global */
	B200_DO(b200_mv_inner_prod(nsdIP, MV(x), MV(y), start, end, inner_prod, ldIP));
}
static void B200_MultiVecSetRandomValue(void **x, int start, int end, struct OPS_ *ops)
{
	B200_DO(b200_mv_set_random(MV(x), start, end));
}
static void B200_MultiVecAxpby(double alpha, void **x, double beta, void **y, int *start, int *end,
		struct OPS_ *ops)
{
	B200_DO(b200_mv_axpby(alpha, MV(x), beta, MV(y), start, end));
}
static void B200_MultiVecLinearComb(void **x, void **y, int is_vec, int *start, int *end,
		double *coef, int ldc, double *beta, int incb, struct OPS_ *ops)
{
	B200_DO(b200_mv_linear_comb(MV(x), MV(y), start, end, coef, ldc, beta, incb));
}
static void B200_MatDotMultiVec(void *mat, void **x, void **y, int *start, int *end, struct OPS_ *ops)
{
	B200_DO(b200_mat_dot_multivec(MAT(mat), 0, MV(x), MV(y), start, end));
}
static void B200_MatTransDotMultiVec(void *mat, void **x, void **y, int *start, int *end, struct OPS_ *ops)
{
	B200_DO(b200_mat_dot_multivec(MAT(mat), 1, MV(x), MV(y), start, end));
}
static void B200_MultiVecQtAP(char ntsA, char ntsdQAP, void **mvQ, void *matA, void **mvP, int is_vec,
		int *start, int *end, double *qAp, int ldQAP, void **mv_ws, struct OPS_ *ops)
{
	B200_DO(b200_mv_qtap(ntsA, ntsdQAP, MV(mvQ), MAT(matA), MV(mvP), start, end, qAp, ldQAP, MV(mv_ws)));
}

/* ---- single vectors: one-column multi-vectors --------------------------------- */
static void B200_VecCreateByMat(void **vec, void *src_mat, struct OPS_ *ops)
{
	B200_MultiVecCreateByMat((void ***)vec, 1, src_mat, ops);
}
static void B200_VecCreateByVec(void **vec, void *src_vec, struct OPS_ *ops)
{
	B200_MultiVecCreateByMultiVec((void ***)vec, 1, (void **)src_vec, ops);
}
static void B200_VecDestroy(void **vec, struct OPS_ *ops)
{
	B200_MultiVecDestroy((void ***)vec, 1, ops);
}
static void B200_VecView(void *x, struct OPS_ *ops) { B200_MultiVecView((void **)x, 0, 1, ops); }
static void B200_VecInnerProd(void *x, void *y, double *ip, struct OPS_ *ops)
{
	int s[2] = {0, 0}, e[2] = {1, 1};
	B200_MultiVecInnerProd('N', (void **)x, (void **)y, 1, s, e, ip, 1, ops);
}
static void B200_VecSetRandomValue(void *x, struct OPS_ *ops) { B200_MultiVecSetRandomValue((void **)x, 0, 1, ops); }
static void B200_VecAxpby(double alpha, void *x, double beta, void *y, struct OPS_ *ops)
{
	int s[2] = {0, 0}, e[2] = {1, 1};
	B200_MultiVecAxpby(alpha, (void **)x, beta, (void **)y, s, e, ops);
}
static void B200_MatDotVec(void *mat, void *x, void *y, struct OPS_ *ops)
{
	int s[2] = {0, 0}, e[2] = {1, 1};
	B200_MatDotMultiVec(mat, (void **)x, (void **)y, s, e, ops);
}
static void B200_MatTransDotVec(void *mat, void *x, void *y, struct OPS_ *ops)
{
	int s[2] = {0, 0}, e[2] = {1, 1};
	B200_MatTransDotMultiVec(mat, (void **)x, (void **)y, s, e, ops);
}

static double B200_GetWtime(void) { return b200_wtime(); }

void OPS_B200_Set(struct OPS_ *ops)
{
	assert(ops->lapack_ops == NULL);
	B200_DO(b200_init(-1));
	/* ComputeP and ComputeRayleighRitz work on HOST coefficient arrays through lapack_ops
	 * (reference src/ops_eig_sol_gcg.c:380-413, :943): keep it a genuine host LAPACK ops. */
	OPS_Create(&(ops->lapack_ops));
	OPS_LAPACK_Set(ops->lapack_ops);
	ops->Printf                   = DefaultPrintf;
	ops->GetOptionFromCommandLine = DefaultGetOptionFromCommandLine;
	ops->GetWtime                 = B200_GetWtime;
	ops->MatView                  = B200_MatView;
	ops->MatAxpby                 = B200_MatAxpby;
	/* vec */
	ops->VecCreateByMat           = B200_VecCreateByMat;
	ops->VecCreateByVec           = B200_VecCreateByVec;
	ops->VecDestroy               = B200_VecDestroy;
	ops->VecView                  = B200_VecView;
	ops->VecInnerProd             = B200_VecInnerProd;
	ops->VecLocalInnerProd        = B200_VecInnerProd;
	ops->VecSetRandomValue        = B200_VecSetRandomValue;
	ops->VecAxpby                 = B200_VecAxpby;
	ops->MatDotVec                = B200_MatDotVec;
	ops->MatTransDotVec           = B200_MatTransDotVec;
	/* multi-vec */
	ops->MultiVecCreateByMat      = B200_MultiVecCreateByMat;
	ops->MultiVecCreateByVec      = B200_MultiVecCreateByVec;
	ops->MultiVecCreateByMultiVec = B200_MultiVecCreateByMultiVec;
	ops->MultiVecDestroy          = B200_MultiVecDestroy;
	ops->GetVecFromMultiVec       = B200_GetVecFromMultiVec;
	ops->RestoreVecForMultiVec    = B200_RestoreVecForMultiVec;
	ops->MultiVecView             = B200_MultiVecView;
	ops->MultiVecLocalInnerProd   = B200_MultiVecInnerProd;
	ops->MultiVecInnerProd        = B200_MultiVecInnerProd;
	ops->MultiVecSetRandomValue   = B200_MultiVecSetRandomValue;
	ops->MultiVecAxpby            = B200_MultiVecAxpby;
	ops->MultiVecLinearComb       = B200_MultiVecLinearComb;
	ops->MatDotMultiVec           = B200_MatDotMultiVec;
	ops->MatTransDotMultiVec      = B200_MatTransDotMultiVec;
	ops->MultiVecQtAP             = B200_MultiVecQtAP;
}

int B200_SetOptionsFromCommandLine(int argc, char *argv[], struct OPS_ *ops)
{
	int nset = 0;
	for (int i = 0; i < b200_option_count(); ++i) {
		char opt[80];
		int value = 0;
		snprintf(opt, sizeof(opt), "-b200_%s", b200_option_name(i));
		if (ops->GetOptionFromCommandLine(opt, 'i', &value, argc, argv, ops)) {
			B200_DO(b200_option_set(b200_option_name(i), value));
			++nset;
		}
	}
	return nset;
}

/* ==== tier B: fused device providers behind the three L3 slots ===================== */

static int tol_type_code(const char *t) { return (t && 0 == strcmp(t, "rel")) ? 1 : 0; }

/* -- BlockPCG ---------------------------------------------------------------------- */
static BlockPCGSolver bpcg_b200_static;

static void BlockPCG_B200(void *mat, void **mv_b, void **mv_x, int *start_bx, int *end_bx, struct OPS_ *ops)
{
	BlockPCGSolver *s = (BlockPCGSolver *)ops->multi_linear_solver_workspace;
	b200_bpcg_params prm;
	prm.max_iter = s->max_iter; prm.rate = s->rate; prm.tol = s->tol;
	prm.tol_type = tol_type_code(s->tol_type); prm.shift = 0.0;
	B200_DO(b200_block_pcg(MAT(mat), NULL, MV(mv_b), MV(mv_x), start_bx, end_bx, &prm,
				MV(s->mv_ws[0]), MV(s->mv_ws[1]), MV(s->mv_ws[2]), &s->niter, &s->residual));
}
void MultiLinearSolverSetup_BlockPCG_B200(int max_iter, double rate, double tol,
		const char *tol_type, void **mv_ws[3], double *dbl_ws, int *int_ws,
		void *pc, void *unused_matdot, struct OPS_ *ops)
{
	bpcg_b200_static.max_iter = max_iter; bpcg_b200_static.rate = rate; bpcg_b200_static.tol = tol;
	strncpy(bpcg_b200_static.tol_type, tol_type, 7); bpcg_b200_static.tol_type[7] = 0;
	bpcg_b200_static.mv_ws[0] = mv_ws[0]; bpcg_b200_static.mv_ws[1] = mv_ws[1];
	bpcg_b200_static.mv_ws[2] = mv_ws[2];
	bpcg_b200_static.dbl_ws = dbl_ws; bpcg_b200_static.int_ws = int_ws;
	bpcg_b200_static.pc = pc; bpcg_b200_static.MatDotMultiVec = NULL;
	bpcg_b200_static.niter = 0; bpcg_b200_static.residual = -1.0;
	ops->multi_linear_solver_workspace = (void *)&bpcg_b200_static;
	ops->MultiLinearSolver = BlockPCG_B200;
}

/* -- orthogonalisation ---------------------------------------------------------------- */
static ModifiedGramSchmidtOrth mgs_b200_static;

static void ModifiedGramSchmidt_B200(void **x, int start_x, int *end_x, void *B, struct OPS_ *ops)
{
	ModifiedGramSchmidtOrth *s = (ModifiedGramSchmidtOrth *)ops->orth_workspace;
	b200_orth_params prm;
	prm.block_size = s->block_size; prm.max_reorth = s->max_reorth;
	prm.orth_zero_tol = s->orth_zero_tol; prm.reorth_tol = s->reorth_tol;
	B200_DO(b200_mv_orth(MV(x), start_x, end_x, MAT(B), &prm, MV(s->mv_ws)));
}
void MultiVecOrthSetup_ModifiedGramSchmidt_B200(int block_size, int max_reorth,
		double orth_zero_tol, void **mv_ws, double *dbl_ws, struct OPS_ *ops)
{
	mgs_b200_static.block_size = block_size; mgs_b200_static.max_reorth = max_reorth;
	mgs_b200_static.orth_zero_tol = orth_zero_tol;
	mgs_b200_static.reorth_tol = 50 * DBL_EPSILON;        /* reference src/ops_orth.c:401-404 */
	mgs_b200_static.mv_ws = mv_ws; mgs_b200_static.dbl_ws = dbl_ws;
	ops->orth_workspace = (void *)&mgs_b200_static;
	ops->MultiVecOrth = ModifiedGramSchmidt_B200;
}

/* -- GCG ------------------------------------------------------------------------------
 * The workspace object is the reference's own GCGSolver (src/ops_eig_sol_gcg.h:26-52), so
 * EigenSolverSetParameters_GCG / ...FromCommandLine_GCG (src/ops_eig_sol_gcg.c:1678,1737)
 * keep working on it unchanged. */
static GCGSolver gcg_b200_static;

static void GCG_B200(void *A, void *B, double *eval, void **evec, int nevGiven, int *nevConv, struct OPS_ *ops)
{
	GCGSolver *s = (GCGSolver *)ops->eigen_solver_workspace;
	b200_gcg_params p; b200_gcg_stats st;
	b200_gcg_default_params(*nevConv, &p);
	p.nevMax = s->nevMax; p.multiMax = s->multiMax; p.nevInit = s->nevInit;
	p.block_size = s->block_size; p.numIterMax = s->numIterMax; p.gapMin = s->gapMin;
	p.tol[0] = s->tol[0]; p.tol[1] = s->tol[1];
	p.check_conv_max_num = s->check_conv_max_num;
	p.initX_orth_block_size = s->initX_orth_block_size; p.initX_orth_max_reorth = s->initX_orth_max_reorth;
	p.initX_orth_zero_tol = s->initX_orth_zero_tol;
	p.compP_orth_block_size = s->compP_orth_block_size; p.compP_orth_max_reorth = s->compP_orth_max_reorth;
	p.compP_orth_zero_tol = s->compP_orth_zero_tol;
	p.compW_orth_block_size = s->compW_orth_block_size; p.compW_orth_max_reorth = s->compW_orth_max_reorth;
	p.compW_orth_zero_tol = s->compW_orth_zero_tol;
	p.compW_cg_max_iter = s->compW_cg_max_iter; p.compW_cg_rate = s->compW_cg_rate;
	p.compW_cg_tol = s->compW_cg_tol; p.compW_cg_tol_type = tol_type_code(s->compW_cg_tol_type);
	p.compW_cg_auto_shift = s->compW_cg_auto_shift; p.compW_cg_shift = s->compW_cg_shift;
	p.compRR_tol = s->compRR_tol;
	p.compW_cg_order = s->compW_cg_order;
	p.initX_orth_method = (0 == strcmp("bgs", s->initX_orth_method)) ? 1 : 0;     /* reference :1757-1785 */
	p.compP_orth_method = (0 == strcmp("bgs", s->compP_orth_method)) ? 1 : 0;
	p.compW_orth_method = (0 == strcmp("bgs", s->compW_orth_method)) ? 1 : 0;
	p.verbose = 1;
	b200_mv *ws[4];
	for (int i = 0; i < 4; ++i) ws[i] = MV(s->mv_ws[i]);
	B200_DO(b200_gcg_solve(MAT(A), MAT(B), eval, MV(evec), nevGiven, nevConv, &p, ws, &st));
	s->numIter = st.numIter; s->nevConv = st.nevConv;
}
void EigenSolverSetup_GCG_B200(int multiMax, double gapMin, int nevInit, int nevMax,
		int block_size, double tol[2], int numIterMax,
		int user_defined_multi_linear_solver,
		void **mv_ws[4], double *dbl_ws, int *int_ws, struct OPS_ *ops)
{
	/* take the reference's defaults for every internal parameter, then swap the driver */
	EigenSolverSetup_GCG(multiMax, gapMin, nevInit, nevMax, block_size, tol, numIterMax,
			user_defined_multi_linear_solver, mv_ws, dbl_ws, int_ws, ops);
	gcg_b200_static = *(GCGSolver *)ops->eigen_solver_workspace;
	ops->eigen_solver_workspace = (void *)&gcg_b200_static;
	ops->EigenSolver = GCG_B200;
}
