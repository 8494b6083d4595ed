// Shared pieces of the fused "stream + per-column dot" kernels (BlockPCG, residual norms):
// a CTA owns a contiguous chunk of rows; thread (tx,ty) owns columns tx, tx+CX, ... and rows
// ty, ty+RY, ...  Per-column partial sums are written per CTA, and the LAST CTA to finish
// (ticket counter) adds the partials in a fixed order, so results are deterministic and no
// second kernel launch is needed for the tiny per-column scalar update that follows.
#pragma once
#include "b200_internal.h"

constexpr int RED_THREADS = 256;

struct RedGeom {
	int cx, ry;               // blockDim = (cx, ry), cx*ry == RED_THREADS
	long long rows_per_chunk;
	int chunks;
};

static inline RedGeom red_geometry(long long n, int k)
{
	RedGeom g;
	g.cx = 1; while (g.cx < k && g.cx < 32) g.cx <<= 1;
	g.ry = RED_THREADS / g.cx;
	long long chunks = (long long)g_b200.num_sms * 4;
	long long rpc = (n + chunks - 1) / chunks;
	if (rpc < g.ry) rpc = g.ry;
	g.rows_per_chunk = rpc;
	g.chunks = (int)((n + rpc - 1) / rpc);
	if (g.chunks < 1) g.chunks = 1;
	return g;
}

// Sum NACC per-thread accumulators per owned column over the CTA and store them at
// part[(blockIdx.x*NACC + a)*k + c].  sm: RY*k doubles.  Returns true in the last CTA to
// arrive (all threads), after which every CTA's partials are visible.
template <int CPT, int NACC>
__device__ __forceinline__ bool red_block_and_elect(double (&acc)[NACC][CPT], int k, double *sm,
                                                    double *part, unsigned *ticket)
{
	__shared__ bool is_last;
	const int cx = blockDim.x, ry = blockDim.y;
	const int tid = threadIdx.y * cx + threadIdx.x;
#pragma unroll
	for (int a = 0; a < NACC; ++a) {
		__syncthreads();
#pragma unroll
		for (int i = 0; i < CPT; ++i) {
			const int c = threadIdx.x + i * cx;
			if (c < k) sm[threadIdx.y * k + c] = acc[a][i];
		}
		__syncthreads();
		for (int c = tid; c < k; c += cx * ry) {
			double s = 0.0;
			for (int j = 0; j < ry; ++j) s += sm[j * k + c];
			part[((size_t)blockIdx.x * NACC + a) * k + c] = s;
		}
	}
	__threadfence();
	__syncthreads();
	if (tid == 0) {
		const unsigned t = atomicAdd(ticket, 1u);
		is_last = (t == gridDim.x - 1);
		if (is_last) *ticket = 0;            // re-arm for the next launch
	}
	__syncthreads();
	if (is_last) __threadfence();
	return is_last;
}

// In the last CTA: total for column c, accumulator a, over all chunks; one warp per column,
// lane l adds chunks l, l+32, ... then a fixed shuffle tree.  Call with all 256 threads;
// the result is valid in every lane of the warp that owns column c (c % 8 == warp).
template <int NACC>
__device__ __forceinline__ double red_total(const double *part, int chunks, int k, int a, int c)
{
	const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
	double s = 0.0;
	for (int ch = lane; ch < chunks; ch += 32) s += part[((size_t)ch * NACC + a) * k + c];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	return s;
}
