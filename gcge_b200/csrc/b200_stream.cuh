// Geometry shared by the streaming kernels over row-major n x k blocks (axpby, the fused
// BlockPCG steps, residual norms, column dots).  These kernels are pure HBM streams, so the
// only things that matter are (a) every warp request is a full-width, 128-bit, contiguous
// access and (b) enough independent loads are in flight per SM to cover the ~1 us HBM
// latency at 6.5 TB/s (~45 KB per SM).
//
//   thread t of a 256-thread CTA owns the column group cg = t % kp (VEC = 2 adjacent columns
//   when everything is 16-byte aligned, else 1) and the rows rl, rl + rp, rl + 2 rp ... of the
//   CTA's contiguous row chunk, rp = 256 / kp rows per pass.  A warp therefore touches whole
//   consecutive row segments, and a thread always works on the same columns, so per-column
//   scalars (alpha, beta, masks) live in registers and per-column dot products accumulate in
//   registers.  Each thread keeps ST_UNROLL rows of every input stream in flight.
//
// The earlier geometry (blockDim = (32, 8), 8-byte accesses, column passes of 32) left 24 of 32
// lanes idle on the second pass at k = 40 and ran the fused BlockPCG update at 3.4 TB/s with
// 48 % occupancy stalled on the long scoreboard (profiles/ncu_r1a).
#pragma once
#include "b200_internal.h"

constexpr int ST_THREADS = 256;
constexpr int ST_UNROLL = 4;

struct StreamGeom {
	int vec;                 // columns per thread (1 or 2)
	int kp;                  // column groups = ceil(k / vec)
	int rp;                  // rows per pass = 256 / kp
	long long rows_per_chunk;
	int chunks;              // grid size
};

static inline bool stream_aligned16(const void *p, int ld) { return ((uintptr_t)p % 16 == 0) && (ld % 2 == 0); }

static inline StreamGeom stream_geometry(long long n, int k, bool aligned16, int ctas_per_sm = 6)
{
	StreamGeom g;
	g.vec = (aligned16 && k % 2 == 0) ? 2 : 1;
	g.kp = (k + g.vec - 1) / g.vec;
	g.rp = ST_THREADS / g.kp; if (g.rp < 1) g.rp = 1;
	long long chunks = (long long)g_b200.num_sms * ctas_per_sm;
	long long rpc = (n + chunks - 1) / chunks;
	const long long quantum = (long long)g.rp * ST_UNROLL;
	rpc = ((rpc + quantum - 1) / quantum) * quantum;
	if (rpc < quantum) rpc = quantum;
	g.rows_per_chunk = rpc;
	g.chunks = (int)((n + rpc - 1) / rpc);
	if (g.chunks < 1) g.chunks = 1;
	return g;
}

template <int VEC> struct StV { double v[VEC]; };

template <int VEC>
__device__ __forceinline__ StV<VEC> st_ld(const double *p)
{
	StV<VEC> r;
	if (VEC == 2) { const double2 t = *reinterpret_cast<const double2 *>(p); r.v[0] = t.x; r.v[VEC - 1] = t.y; }
	else r.v[0] = *p;
	return r;
}

template <int VEC>
__device__ __forceinline__ void st_st(double *p, const StV<VEC> &r)
{
	if (VEC == 2) *reinterpret_cast<double2 *>(p) = make_double2(r.v[0], r.v[VEC - 1]);
	else *p = r.v[0];
}

// Per-thread column ownership.  Threads with rl >= rp (the 256 % kp leftovers) are idle.
struct StreamThread {
	int cg, rl, c;
	bool active;
};

template <int VEC>
__device__ __forceinline__ StreamThread stream_thread(const StreamGeom &g)
{
	StreamThread t;
	t.cg = threadIdx.x % g.kp; t.rl = threadIdx.x / g.kp; t.c = t.cg * VEC;
	t.active = t.rl < g.rp;
	return t;
}

// CTA-wide sum of NACC per-thread accumulators per owned column; partials go to
// part[(blockIdx.x*NACC + a)*k + c].  Returns true (to all threads) in the last CTA to arrive,
// after which every CTA's partials are visible.  Summation order is fixed => deterministic.
template <int VEC, int NACC>
__device__ __forceinline__ bool stream_reduce_and_elect(const double (&acc)[NACC][VEC], int k, const StreamGeom &g,
                                                        const StreamThread &t, double *part, unsigned *ticket)
{
	__shared__ double sm[2 * ST_THREADS];         // rp * k <= 256 * VEC
	__shared__ bool is_last;
#pragma unroll
	for (int a = 0; a < NACC; ++a) {
		__syncthreads();
		if (t.active) {
#pragma unroll
			for (int i = 0; i < VEC; ++i) sm[t.rl * k + t.c + i] = acc[a][i];
		}
		__syncthreads();
		for (int c = threadIdx.x; c < k; c += ST_THREADS) {
			double s = 0.0;
			for (int j = 0; j < g.rp; ++j) s += sm[j * k + c];
			part[((size_t)blockIdx.x * NACC + a) * k + c] = s;
		}
	}
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned tk = atomicAdd(ticket, 1u);
		is_last = (tk == gridDim.x - 1);
		if (is_last) *ticket = 0;                 // re-arm for the next launch
	}
	__syncthreads();
	if (is_last) __threadfence();
	return is_last;
}

// In the last CTA: total of accumulator a for column c over all chunks.  One warp per column:
// lane l adds chunks l, l+32, ... then a fixed shuffle tree.  Valid in every lane.
template <int NACC>
__device__ __forceinline__ double stream_total(const double *part, int chunks, int k, int a, int c)
{
	const int lane = threadIdx.x & 31;
	double s = 0.0;
	for (int ch = lane; ch < chunks; ch += 32) s += part[((size_t)ch * NACC + a) * k + c];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	return s;
}

// Sum vals[0..cnt) over the ranks, in place, by the calling CTA (every thread must call; cnt <=
// B200_AR_MAX_COUNT; vals in global memory).  See b200_comm.cu "allreduce inside a kernel".
__device__ __forceinline__ void stream_allreduce_cta(const B200ArCtx &ar, double *vals, int cnt)
{
	__shared__ unsigned ar_epoch;
	__syncthreads();
	if (threadIdx.x == 0) ar_epoch = ++(*ar.seq);
	__syncthreads();
	const unsigned e = ar_epoch, par = e & 1u;
	const size_t slot = ((size_t)par * ar.nranks + ar.rank) * B200_AR_MAX_COUNT;
	for (int i = threadIdx.x; i < cnt * ar.nranks; i += blockDim.x) {
		const int q = i / cnt, j = i - q * cnt;
		ar.inbox[q][slot + j] = vals[j];                    // remote store (local for q == rank)
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x < ar.nranks) {
		volatile unsigned *f = ar.flags[threadIdx.x] + par * ar.nranks + ar.rank;
		*f = e;
	}
	if (threadIdx.x < ar.nranks) {
		volatile unsigned *mine = ar.flags[ar.rank] + par * ar.nranks + threadIdx.x;
		const long long t0 = clock64();
		while ((int)(*mine - e) < 0) {
			if (clock64() - t0 > (1ll << 36)) { *ar.status = 1; break; }      // ~35 s: give up instead of hanging
		}
	}
	__threadfence_system();
	__syncthreads();
	for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
		double s = 0.0;
		for (int q = 0; q < ar.nranks; ++q)
			s += ((volatile double *)ar.inbox[ar.rank])[((size_t)par * ar.nranks + q) * B200_AR_MAX_COUNT + j];
		vals[j] = s;
	}
	__syncthreads();
}

#define ST_DISPATCH_VEC(g, CALL)                       \
	do {                                               \
		if ((g).vec == 2) { constexpr int VEC = 2; CALL; } \
		else { constexpr int VEC = 1; CALL; }          \
	} while (0)
