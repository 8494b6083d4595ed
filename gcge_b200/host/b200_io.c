/* b200_io.c -- on-disk matrices: MatrixMarket coordinate files and PETSc binary
 * matrices into the CCS
 * arrays every other entry point takes (reference app/app_ccs.h:20-24: data / i_row / j_col, 0-based;
 * the reference reads its real-world matrices through PETSc/SLEPc drivers, test/test_app_slepc.c:416-445,
 * which are outside this path -- this is the dependency-free equivalent for the CCS app).
 * Host only; no device call. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include "b200_dev.h"

typedef struct { int r, c; double v; long long seq; } triple;

static int cmp_triple(const void *a, const void *b)
{
	const triple *x = (const triple *)a, *y = (const triple *)b;
	if (x->c != y->c) return x->c < y->c ? -1 : 1;
	if (x->r != y->r) return x->r < y->r ? -1 : 1;
	return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0);      /* duplicates keep their file order */
}

static void lower(char *s) { for (; *s; ++s) *s = (char)tolower((unsigned char)*s); }

int b200_ccs_read_matrix_market(const char *path, int *nrows, int *ncols, int **j_col, int **i_row, double **data)
{
	if (!path || !nrows || !ncols || !j_col || !i_row || !data) return b200_fail("b200_ccs_read_matrix_market: bad arguments");
	FILE *f = fopen(path, "r");
	if (!f) return b200_fail("b200_ccs_read_matrix_market: cannot open %s", path);
	char line[1024], banner[64], obj[64], fmt[64], field[64], sym[64];
	if (!fgets(line, sizeof line, f) || sscanf(line, "%63s %63s %63s %63s %63s", banner, obj, fmt, field, sym) != 5) {
		fclose(f); return b200_fail("%s: not a MatrixMarket file (no banner)", path);
	}
	lower(banner); lower(obj); lower(fmt); lower(field); lower(sym);
	if (strcmp(banner, "%%matrixmarket") || strcmp(obj, "matrix") || strcmp(fmt, "coordinate")) {
		fclose(f); return b200_fail("%s: only 'matrix coordinate' MatrixMarket files are supported", path);
	}
	const int pattern = !strcmp(field, "pattern");
	if (!pattern && strcmp(field, "real") && strcmp(field, "integer") && strcmp(field, "double")) {
		fclose(f); return b200_fail("%s: field '%s' is not supported (real, integer, pattern)", path, field);
	}
	int mirror = 0; double mirror_sign = 1.0;
	if (!strcmp(sym, "symmetric")) mirror = 1;
	else if (!strcmp(sym, "skew-symmetric")) { mirror = 1; mirror_sign = -1.0; }
	else if (strcmp(sym, "general")) { fclose(f); return b200_fail("%s: symmetry '%s' is not supported", path, sym); }
	do {
		if (!fgets(line, sizeof line, f)) { fclose(f); return b200_fail("%s: no size line", path); }
	} while (line[0] == '%' || line[0] == '\n' || line[0] == '\r');
	long long m = 0, n = 0, nz = 0;
	if (sscanf(line, "%lld %lld %lld", &m, &n, &nz) != 3 || m < 0 || n < 0 || nz < 0 || m > 0x7fffffff || n > 0x7fffffff) {
		fclose(f); return b200_fail("%s: bad size line", path);
	}
	const long long cap = mirror ? 2 * nz : nz;
	if (cap > 0x7fffffff) { fclose(f); return b200_fail("%s: %lld entries exceed the int index range of CCSMAT", path, cap); }
	triple *t = (triple *)malloc(sizeof(triple) * (size_t)(cap > 0 ? cap : 1));
	if (!t) { fclose(f); return b200_fail("%s: out of host memory", path); }
	long long cnt = 0;
	for (long long e = 0; e < nz; ++e) {
		long long r, c; double v = 1.0;
		int got = pattern ? fscanf(f, "%lld %lld", &r, &c) : fscanf(f, "%lld %lld %lf", &r, &c, &v);
		if (got != (pattern ? 2 : 3) || r < 1 || r > m || c < 1 || c > n) {
			free(t); fclose(f); return b200_fail("%s: bad entry %lld", path, e + 1);
		}
		t[cnt].r = (int)(r - 1); t[cnt].c = (int)(c - 1); t[cnt].v = v; t[cnt].seq = cnt; ++cnt;
		if (mirror && r != c) { t[cnt].r = (int)(c - 1); t[cnt].c = (int)(r - 1); t[cnt].v = mirror_sign * v; t[cnt].seq = cnt; ++cnt; }
	}
	fclose(f);
	qsort(t, (size_t)cnt, sizeof(triple), cmp_triple);
	int *jc = (int *)calloc((size_t)n + 1, sizeof(int));
	int *ir = (int *)malloc(sizeof(int) * (size_t)(cnt > 0 ? cnt : 1));
	double *da = (double *)malloc(sizeof(double) * (size_t)(cnt > 0 ? cnt : 1));
	if (!jc || !ir || !da) { free(t); free(jc); free(ir); free(da); return b200_fail("%s: out of host memory", path); }
	for (long long e = 0; e < cnt; ++e) { ++jc[t[e].c + 1]; ir[e] = t[e].r; da[e] = t[e].v; }
	for (long long j = 0; j < n; ++j) jc[j + 1] += jc[j];
	free(t);
	*nrows = (int)m; *ncols = (int)n; *j_col = jc; *i_row = ir; *data = da;
	return 0;
}

/* PETSc binary matrix (MatView to a binary viewer; what the reference's SLEPc driver loads with
 * MatLoad, test/test_app_slepc.c:416-445): big-endian int32 header {1211216, rows, cols, nnz}, int32
 * entries per row [rows], int32 column indices [nnz] (row by row, 0-based), float64 values [nnz].
 * Converted to CCS by a counting sort over the columns (rows ascending inside a column, duplicates in
 * file order). */
static unsigned int be32(const unsigned char *p) { return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3]; }
static double be64(const unsigned char *p)
{
	unsigned long long u = 0;
	for (int i = 0; i < 8; ++i) u = (u << 8) | p[i];
	double d; memcpy(&d, &u, sizeof d);
	return d;
}

int b200_ccs_read_petsc_binary(const char *path, int *nrows, int *ncols, int **j_col, int **i_row, double **data)
{
	if (!path || !nrows || !ncols || !j_col || !i_row || !data) return b200_fail("b200_ccs_read_petsc_binary: bad arguments");
	FILE *f = fopen(path, "rb");
	if (!f) return b200_fail("b200_ccs_read_petsc_binary: cannot open %s", path);
	unsigned char h[16];
	if (fread(h, 1, 16, f) != 16 || be32(h) != 1211216u) { fclose(f); return b200_fail("%s: not a PETSc binary matrix (class id)", path); }
	const long long m = (int)be32(h + 4), n = (int)be32(h + 8), nz = (int)be32(h + 12);
	if (m < 0 || n < 0 || nz < 0) { fclose(f); return b200_fail("%s: bad header (%lld x %lld, %lld entries; dense PETSc files are not supported)", path, m, n, nz); }
	unsigned char *rl = (unsigned char *)malloc((size_t)(m > 0 ? m : 1) * 4);
	unsigned char *ci = (unsigned char *)malloc((size_t)(nz > 0 ? nz : 1) * 4);
	unsigned char *va = (unsigned char *)malloc((size_t)(nz > 0 ? nz : 1) * 8);
	int *jc = (int *)calloc((size_t)n + 1, sizeof(int));
	int *ir = (int *)malloc(sizeof(int) * (size_t)(nz > 0 ? nz : 1));
	double *da = (double *)malloc(sizeof(double) * (size_t)(nz > 0 ? nz : 1));
	int rc = 0;
	if (!rl || !ci || !va || !jc || !ir || !da) rc = b200_fail("%s: out of host memory", path);
	else if (fread(rl, 4, (size_t)m, f) != (size_t)m || fread(ci, 4, (size_t)nz, f) != (size_t)nz ||
	         fread(va, 8, (size_t)nz, f) != (size_t)nz) rc = b200_fail("%s: truncated file", path);
	fclose(f);
	if (!rc) {
		long long tot = 0;
		for (long long r = 0; r < m; ++r) tot += (int)be32(rl + 4 * r);
		if (tot != nz) rc = b200_fail("%s: row lengths sum to %lld, header says %lld", path, tot, nz);
	}
	if (!rc) {
		for (long long e = 0; e < nz; ++e) {
			const long long c = (int)be32(ci + 4 * e);
			if (c < 0 || c >= n) { rc = b200_fail("%s: column index %lld out of range at entry %lld", path, c, e); break; }
			++jc[c + 1];
		}
	}
	if (!rc) {
		for (long long j = 0; j < n; ++j) jc[j + 1] += jc[j];
		int *next = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
		if (!next) rc = b200_fail("%s: out of host memory", path);
		else {
			memcpy(next, jc, sizeof(int) * (size_t)n);
			long long e = 0;
			for (long long r = 0; r < m; ++r) {          /* rows ascending => ascending inside every column */
				const int len = (int)be32(rl + 4 * r);
				for (int k = 0; k < len; ++k, ++e) {
					const int c = (int)be32(ci + 4 * e);
					const int pos = next[c]++;
					ir[pos] = (int)r; da[pos] = be64(va + 8 * e);
				}
			}
			free(next);
		}
	}
	free(rl); free(ci); free(va);
	if (rc) { free(jc); free(ir); free(da); return rc; }
	*nrows = (int)m; *ncols = (int)n; *j_col = jc; *i_row = ir; *data = da;
	return 0;
}

void b200_ccs_free(int *j_col, int *i_row, double *data) { free(j_col); free(i_row); free(data); }
