// K2/K3: the two dense contractions of the GCG hot path, on FP64 tensor-core tiles.
//
//   gram    C(p x q)  = X(n x p)^T Y(n x q)            replaces DenseMatQtAP 'N'/'S'/'D',
//                                                       reference app/app_lapack.c:24-183
//   lincomb Y(n x q)  = X(n x p) C(p x q) + Y diag(b)   replaces MultiVecLinearComb,
//                                                       reference app/app_lapack.c:463-534
//
// sm_100a has no tcgen05 kind for f64; the FP64 tensor path is the warp-level
// mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4), which is what both kernels issue.
// Fragment layout (PTX ISA, m8n8k4 .f64): lane = 4*g + t
//   A (8x4, row):  a  = A[g][t]        B (4x8, col):  b = B[t][g]
//   C/D (8x8):     c0 = C[g][2t], c1 = C[g][2t+1]
//
// Multi-vector blocks are row-major (n x ld), so for gram both operands are read exactly
// as stored ([row][col] == [k][m] and [k][n]); for lincomb X is the row-major A operand.
// Shared-memory row strides are == 4 (mod 16) doubles, which makes every fragment load a
// two-wavefront (conflict-free) 64-bit access.
#include "b200_internal.h"

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
	             : "+d"(c0), "+d"(c1)
	             : "d"(a), "d"(b));
}

// =========================================================================== gram
constexpr int GR_BP = 64;     // tile of X columns (rows of C)
constexpr int GR_BQ = 64;     // tile of Y columns (cols of C)
constexpr int GR_BK = 32;     // multi-vector rows per stage
constexpr int GR_S  = 68;     // smem row stride (doubles), == 4 mod 16

// Each CTA: rows [chunk*rows_per_chunk, ...) x one (p-tile, q-tile); 8 warps = 4 p-slices
// (16 columns of X each) x 2 halves of the staged rows.  Partials go to
// part[chunk][q][p] (column-major p x q per chunk); a second kernel sums the chunks in a
// fixed order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(256)
gram_partial_kernel(long long n, int p, int q, const double *x, int ldx, const double *y, int ldy,
                    long long rows_per_chunk, double *part)
{
	__shared__ double Xs[GR_BK][GR_S];
	__shared__ double Ys[GR_BK][GR_S];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int g = lane >> 2, t = lane & 3;
	const int pw = warp & 3, kh = warp >> 2;
	const int p0 = blockIdx.y * GR_BP, q0 = blockIdx.z * GR_BQ;
	const int pt = min(GR_BP, p - p0), qt = min(GR_BQ, q - q0);
	const int nq8 = (qt + 7) >> 3;                       // active 8-column groups of Y
	const long long r_begin = (long long)blockIdx.x * rows_per_chunk;
	long long r_end = r_begin + rows_per_chunk; if (r_end > n) r_end = n;

	double acc[2][8][2];
#pragma unroll
	for (int i = 0; i < 2; ++i)
#pragma unroll
		for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

	for (long long r0 = r_begin; r0 < r_end; r0 += GR_BK) {
		// stage GR_BK rows of both operands (zero fill outside the block)
		for (int i = tid; i < GR_BK * GR_BP; i += 256) {
			const int rr = i / GR_BP, cc = i - rr * GR_BP;
			const long long r = r0 + rr;
			Xs[rr][cc] = (r < r_end && cc < pt) ? x[(size_t)r * ldx + p0 + cc] : 0.0;
		}
		for (int i = tid; i < GR_BK * GR_BQ; i += 256) {
			const int rr = i / GR_BQ, cc = i - rr * GR_BQ;
			const long long r = r0 + rr;
			Ys[rr][cc] = (r < r_end && cc < qt) ? y[(size_t)r * ldy + q0 + cc] : 0.0;
		}
		__syncthreads();
#pragma unroll
		for (int ks = 0; ks < GR_BK / 2; ks += 4) {
			const int kr = kh * (GR_BK / 2) + ks + t;
			const double a0 = Xs[kr][pw * 16 + g];
			const double a1 = Xs[kr][pw * 16 + 8 + g];
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				if (j < nq8) {
					const double b = Ys[kr][j * 8 + g];
					dmma_8x8x4(acc[0][j][0], acc[0][j][1], a0, b);
					dmma_8x8x4(acc[1][j][0], acc[1][j][1], a1, b);
				}
			}
		}
		__syncthreads();
	}
	// combine the two row halves through shared memory (reuse Xs as a 64 x 64 tile, stride 68)
	// the 64 x 64 C tile of the upper row half goes through the two 32 x 68 stage buffers
	// (Xs: C rows 0..31, Ys: C rows 32..63)
	if (kh == 1) {
#pragma unroll
		for (int i = 0; i < 2; ++i)
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const int cr = pw * 16 + i * 8 + g;        // row of C tile (0..63)
				double (*buf)[GR_S] = (cr < 32) ? Xs : Ys;
				buf[cr & 31][j * 8 + 2 * t]     = acc[i][j][0];
				buf[cr & 31][j * 8 + 2 * t + 1] = acc[i][j][1];
			}
	}
	__syncthreads();
	if (kh == 0) {
		double *out = part + (size_t)blockIdx.x * p * q;
#pragma unroll
		for (int i = 0; i < 2; ++i)
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const int cr = pw * 16 + i * 8 + g;
				double (*buf)[GR_S] = (cr < 32) ? Xs : Ys;
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					const int cc = j * 8 + 2 * t + h;
					if (cr < pt && cc < qt)
						out[(size_t)(q0 + cc) * p + (p0 + cr)] = acc[i][j][h] + buf[cr & 31][cc];
				}
			}
	}
}

// C[i + j*ldc] = sum_chunks part[chunk][i + j*p]; mode 'S' mirrors the lower triangle.
__global__ void gram_reduce_kernel(int p, int q, int chunks, const double *__restrict__ part, double alpha,
                                   double *__restrict__ c, int c_rs, int c_cs, int symmetric)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= p * q) return;
	const int i = idx % p, j = idx / p;
	if (symmetric && i < j) return;
	double s = 0.0;
	for (int ch = 0; ch < chunks; ++ch) s += part[(size_t)ch * p * q + idx];
	s *= alpha;
	c[(size_t)i * c_rs + (size_t)j * c_cs] = s;
	if (symmetric && i > j) c[(size_t)j * c_rs + (size_t)i * c_cs] = s;
}

// 'D': column dot products.  blockDim = (CX, 256/CX); thread (tx,ty) owns column tx (+CX...)
// for rows ty, ty+RY, ...; per-CTA partial per column, then the same fixed-order reduce.
__global__ void __launch_bounds__(256)
dots_partial_kernel(long long n, int k, const double *x, int ldx, const double *y, int ldy,
                    long long rows_per_chunk, double *part)
{
	extern __shared__ double sm[];           // [blockDim.y][k]
	const int cx = blockDim.x, ry = blockDim.y;
	const long long r_begin = (long long)blockIdx.x * rows_per_chunk;
	long long r_end = r_begin + rows_per_chunk; if (r_end > n) r_end = n;
	for (int c = threadIdx.x; c < k; c += cx) {
		double s = 0.0;
		for (long long r = r_begin + threadIdx.y; r < r_end; r += ry)
			s = fma(x[(size_t)r * ldx + c], y[(size_t)r * ldy + c], s);
		sm[threadIdx.y * k + c] = s;
	}
	__syncthreads();
	const int tid = threadIdx.y * cx + threadIdx.x;
	for (int c = tid; c < k; c += cx * ry) {
		double s = 0.0;
		for (int j = 0; j < ry; ++j) s += sm[j * k + c];
		part[(size_t)blockIdx.x * k + c] = s;
	}
}

__global__ void dots_reduce_kernel(int k, int chunks, const double *__restrict__ part, double alpha,
                                   double *__restrict__ c, int stride)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= k) return;
	double s = 0.0;
	for (int ch = 0; ch < chunks; ++ch) s += part[(size_t)ch * k + idx];
	c[(size_t)idx * stride] = alpha * s;
}

__global__ void fill2d_kernel(int p, int q, double v, double *c, int c_rs, int c_cs)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= p * q) return;
	c[(size_t)(idx % p) * c_rs + (size_t)(idx / p) * c_cs] = v;
}

// 'D': c_dev[i*c_rs] = alpha * x_i . y_i  (c_cs ignored)
int b200k_gram(char mode, long long n, int p, int q, double alpha, const double *x, int ldx,
               const double *y, int ldy, double *c_dev, int c_rs, int c_cs)
{
	if (p <= 0 || q <= 0) return 0;
	cudaStream_t st = g_b200.stream;
	if (mode == 'D') {
		const int k = p;
		if (n <= 0) {
			fill2d_kernel<<<b200_ceil_div(k, 128), 128, 0, st>>>(k, 1, 0.0, c_dev, c_rs, 0);
			B200_KERNEL_CHECK();
			return 0;
		}
		int cx = 1; while (cx < k && cx < 32) cx <<= 1;
		const int ry = 256 / cx;
		long long chunks = g_b200.num_sms * 4;
		long long rows_per_chunk = (n + chunks - 1) / chunks;
		if (rows_per_chunk < ry) rows_per_chunk = ry;
		chunks = (n + rows_per_chunk - 1) / rows_per_chunk;
		double *part = (double *)b200_scratch(0, sizeof(double) * (size_t)chunks * k);
		if (!part) return 1;
		dots_partial_kernel<<<(unsigned)chunks, dim3(cx, ry), sizeof(double) * (size_t)ry * k, st>>>(
			n, k, x, ldx, y, ldy, rows_per_chunk, part);
		B200_KERNEL_CHECK();
		dots_reduce_kernel<<<b200_ceil_div(k, 128), 128, 0, st>>>(k, (int)chunks, part, alpha, c_dev, c_rs);
		B200_KERNEL_CHECK();
		return 0;
	}
	if (n <= 0) {
		fill2d_kernel<<<b200_ceil_div((long long)p * q, 256), 256, 0, st>>>(p, q, 0.0, c_dev, c_rs, c_cs);
		B200_KERNEL_CHECK();
		return 0;
	}
	const int ptiles = b200_ceil_div(p, GR_BP), qtiles = b200_ceil_div(q, GR_BQ);
	long long chunks = (long long)(g_b200.num_sms * 4) / ((long long)ptiles * qtiles);
	if (chunks < 1) chunks = 1;
	long long rows_per_chunk = (n + chunks - 1) / chunks;
	rows_per_chunk = ((rows_per_chunk + GR_BK - 1) / GR_BK) * GR_BK;
	if (rows_per_chunk < 4 * GR_BK) rows_per_chunk = 4 * GR_BK;
	chunks = (n + rows_per_chunk - 1) / rows_per_chunk;
	double *part = (double *)b200_scratch(0, sizeof(double) * (size_t)chunks * p * q);
	if (!part) return 1;
	dim3 grid((unsigned)chunks, ptiles, qtiles);
	gram_partial_kernel<<<grid, 256, 0, st>>>(n, p, q, x, ldx, y, ldy, rows_per_chunk, part);
	B200_KERNEL_CHECK();
	gram_reduce_kernel<<<b200_ceil_div((long long)p * q, 256), 256, 0, st>>>(p, q, (int)chunks, part, alpha, c_dev,
	                                                                        c_rs, c_cs, mode == 'S' ? 1 : 0);
	B200_KERNEL_CHECK();
	return 0;
}

// ======================================================================== lincomb
constexpr int LC_BM = 128;    // multi-vector rows per CTA
constexpr int LC_BN = 64;     // output columns per CTA
constexpr int LC_BK = 16;     // slice of the contraction (columns of X / rows of C)
constexpr int LC_SX = 20;     // Xs row stride, == 4 mod 16
constexpr int LC_SC = 68;     // Cs row stride, == 4 mod 16

// C is device column-major (ldc).  8 warps, warp w owns rows [16w,16w+16) x 64 columns.
template <bool HAS_BETA>
__global__ void __launch_bounds__(256)
lincomb_kernel(long long n, int p, int q, const double *x, int ldx, const double *__restrict__ c, int c_rs,
               int c_cs, const double *__restrict__ beta, int incb, double *y, int ldy)
{
	__shared__ double Xs[LC_BM][LC_SX];
	__shared__ double Cs[LC_BK][LC_SC];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int g = lane >> 2, t = lane & 3;
	const long long r0 = (long long)blockIdx.x * LC_BM;
	const int n0 = blockIdx.y * LC_BN;
	const int nt = min(LC_BN, q - n0);
	const int nq8 = (nt + 7) >> 3;

	double acc[2][8][2];
#pragma unroll
	for (int i = 0; i < 2; ++i)
#pragma unroll
		for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

	for (int k0 = 0; k0 < p; k0 += LC_BK) {
		for (int i = tid; i < LC_BM * LC_BK; i += 256) {
			const int rr = i / LC_BK, kk = i - rr * LC_BK;
			const long long r = r0 + rr;
			Xs[rr][kk] = (r < n && k0 + kk < p) ? x[(size_t)r * ldx + k0 + kk] : 0.0;
		}
		if (c_rs == 1) {             // column-major C: walk down columns
			for (int i = tid; i < LC_BK * LC_BN; i += 256) {
				const int cc = i / LC_BK, kk = i - cc * LC_BK;
				Cs[kk][cc] = (k0 + kk < p && cc < nt) ? c[(size_t)(n0 + cc) * c_cs + k0 + kk] : 0.0;
			}
		} else {                     // row-major (or general strides): walk along rows
			for (int i = tid; i < LC_BK * LC_BN; i += 256) {
				const int kk = i / LC_BN, cc = i - kk * LC_BN;
				Cs[kk][cc] = (k0 + kk < p && cc < nt) ? c[(size_t)(k0 + kk) * c_rs + (size_t)(n0 + cc) * c_cs] : 0.0;
			}
		}
		__syncthreads();
#pragma unroll
		for (int ks = 0; ks < LC_BK; ks += 4) {
			const double a0 = Xs[warp * 16 + g][ks + t];
			const double a1 = Xs[warp * 16 + 8 + g][ks + t];
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				if (j < nq8) {
					const double b = Cs[ks + t][j * 8 + g];
					dmma_8x8x4(acc[0][j][0], acc[0][j][1], a0, b);
					dmma_8x8x4(acc[1][j][0], acc[1][j][1], a1, b);
				}
			}
		}
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 2; ++i) {
		const long long r = r0 + warp * 16 + i * 8 + g;
		if (r >= n) continue;
#pragma unroll
		for (int j = 0; j < 8; ++j)
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const int cc = j * 8 + 2 * t + h;
				if (cc < nt) {
					double v = acc[i][j][h];
					double *yp = y + (size_t)r * ldy + n0 + cc;
					if (HAS_BETA) {
						const double b = beta[(size_t)incb * (n0 + cc)];
						if (b != 0.0) v += b * (*yp);      // beta == 0 overwrites (no NaN carry-over)
					}
					*yp = v;
				}
			}
	}
}

// scaling only: y[:,c] *= beta[incb*c]  (beta == nullptr => y = 0)
__global__ void colscale_kernel(long long n, int q, int rows_per_cta, const double *__restrict__ beta, int incb,
                                double *y, int ldy)
{
	const long long r0 = (long long)blockIdx.x * rows_per_cta;
	long long nr = n - r0; if (nr > rows_per_cta) nr = rows_per_cta;
	const int total = (int)nr * q;
	for (int i = threadIdx.x; i < total; i += blockDim.x) {
		const int r = i / q, cidx = i - r * q;
		double *yp = y + (size_t)(r0 + r) * ldy + cidx;
		const double b = beta ? beta[(size_t)incb * cidx] : 0.0;
		*yp = (b == 0.0) ? 0.0 : b * (*yp);
	}
}

int b200k_lincomb(long long n, int p, int q, const double *x, int ldx, const double *c_dev, int c_rs, int c_cs,
                  const double *beta_dev, int incb, double *y, int ldy)
{
	if (n <= 0 || q <= 0) return 0;
	cudaStream_t st = g_b200.stream;
	if (x == nullptr || c_dev == nullptr || p <= 0) {
		// reference app/app_lapack.c:476-505: without x/coef only the dscal part runs, and
		// with beta == NULL nothing runs at all
		if (beta_dev == nullptr) return 0;
		int rows = 4096 / q; if (rows < 1) rows = 1;
		colscale_kernel<<<(unsigned)((n + rows - 1) / rows), 256, 0, st>>>(n, q, rows, beta_dev, incb, y, ldy);
		B200_KERNEL_CHECK();
		return 0;
	}
	dim3 grid((unsigned)((n + LC_BM - 1) / LC_BM), b200_ceil_div(q, LC_BN));
	if (beta_dev) lincomb_kernel<true><<<grid, 256, 0, st>>>(n, p, q, x, ldx, c_dev, c_rs, c_cs, beta_dev, incb, y, ldy);
	else          lincomb_kernel<false><<<grid, 256, 0, st>>>(n, p, q, x, ldx, c_dev, c_rs, c_cs, nullptr, 0, y, ldy);
	B200_KERNEL_CHECK();
	return 0;
}

// y[:,j] *= s[j]  (or /= s[j])
__global__ void colscale_vec_kernel(long long n, int q, int rows_per_cta, const double *__restrict__ sv, int invert,
                                    double *y, int ldy)
{
	const long long r0 = (long long)blockIdx.x * rows_per_cta;
	long long nr = n - r0; if (nr > rows_per_cta) nr = rows_per_cta;
	const int total = (int)nr * q;
	for (int i = threadIdx.x; i < total; i += blockDim.x) {
		const int r = i / q, cidx = i - r * q;
		double *yp = y + (size_t)(r0 + r) * ldy + cidx;
		const double f = invert ? 1.0 / sv[cidx] : sv[cidx];
		*yp = f * (*yp);
	}
}

int b200k_colscale(long long n, int q, const double *s_dev, int invert, double *y, int ldy)
{
	if (n <= 0 || q <= 0) return 0;
	int rows = 4096 / q; if (rows < 1) rows = 1;
	colscale_vec_kernel<<<(unsigned)((n + rows - 1) / rows), 256, 0, g_b200.stream>>>(n, q, rows, s_dev, invert, y, ldy);
	B200_KERNEL_CHECK();
	return 0;
}
