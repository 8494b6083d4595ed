// K9: MultiVecSetRandomValue on device, bit-exact with the process's glibc rand() stream.
//
// The reference fills x[row, col] = rand()/(RAND_MAX+1.0) in column-major order from the
// process-wide glibc generator (reference app/app_lapack.c:322-333) that its drivers seed with
// srand(0) (reference test/test_eig_sol_gcg.c:87).  The initial block decides the iteration
// count, so the stream must be reproduced exactly -- but n x nevMax calls of rand() on one host
// core (3.2e9 at n = 8 M, nev = 200) would cost more than the whole solve on the GPU.
//
// glibc's rand() is random_r() TYPE_3: an additive lagged-Fibonacci generator over Z/2^32,
//     o_t = o_{t-31} + o_{t-3}  (mod 2^32),   rand() = o_t >> 1,
// kept in a 31-word table with a front and a rear pointer (glibc stdlib/random_r.c).  It is
// LINEAR, so the state can be advanced by any distance with a 31 x 31 matrix power
// (mod 2^32):  S_t = (o_{t-31} .. o_{t-1}),  S_{t+1} = M S_t.
//
//   host   : reads the live generator state through setstate() (initstate/setstate hand the
//            caller the state array, with the rear index encoded in word 0), builds
//            P_j = M^(L 2^j), launches the kernel, and writes the state advanced by the number
//            of values consumed back into glibc -- later rand() calls of the process continue
//            the same stream as if rand() had been called that often;
//   device : chunk c (L consecutive stream positions) gets its start state (M^L)^c S_0 from the
//            binary expansion of c -- the high bits once per CTA, the low bits per thread --
//            then runs the recurrence in registers and stores straight into the row-major
//            multi-vector.
#include "b200_internal.h"
#include <vector>

namespace {

constexpr int DEG = 31;                 // TYPE_3 degree
constexpr int SEP = 3;                  // TYPE_3 separation
constexpr int ROUNDS = 64;              // rounds of 31 values per thread
constexpr long long CHUNK = (long long)DEG * ROUNDS;
constexpr int RTHREADS = 128;           // chunks per CTA (power of two)
constexpr int LOWBITS = 7;              // log2(RTHREADS)
constexpr int NPOW = 48;                // P_0 .. P_47: chunk indices below 2^48

struct Mat31 { uint32_t a[DEG][DEG]; };

void mat_mul(const Mat31 &x, const Mat31 &y, Mat31 &z)
{
	static Mat31 t;
	for (int i = 0; i < DEG; ++i)
		for (int j = 0; j < DEG; ++j) {
			uint32_t s = 0;
			for (int k = 0; k < DEG; ++k) s += x.a[i][k] * y.a[k][j];
			t.a[i][j] = s;
		}
	z = t;
}

void mat_vec(const Mat31 &m, uint32_t *v)
{
	uint32_t t[DEG];
	for (int i = 0; i < DEG; ++i) {
		uint32_t s = 0;
		for (int k = 0; k < DEG; ++k) s += m.a[i][k] * v[k];
		t[i] = s;
	}
	memcpy(v, t, sizeof(t));
}

void step_matrix(Mat31 &m)
{
	memset(&m, 0, sizeof(m));
	for (int i = 0; i + 1 < DEG; ++i) m.a[i][i + 1] = 1;      // shift
	m.a[DEG - 1][0] = 1; m.a[DEG - 1][DEG - SEP] = 1;         // o_t = o_{t-31} + o_{t-3}
}

// S <- M^steps S
void advance(uint32_t *S, unsigned long long steps)
{
	Mat31 q; step_matrix(q);
	while (steps) {
		if (steps & 1ull) mat_vec(q, S);
		steps >>= 1;
		if (steps) mat_mul(q, q, q);
	}
}

// ---- the live glibc state -------------------------------------------------------------
// initstate() installs a scratch state and returns the array the generator was using, with
// word 0 = 5*rear + type (glibc stdlib/random_r.c, __setstate_r); setstate() puts it back.
struct LiveState {
	int32_t *words = nullptr;          // word 0: type info, words 1..31: table
	int rear = 0;
	bool ok = false;
};
char g_scratch_state[128];

LiveState grab_state()
{
	LiveState ls;
	char *old = initstate(1u, g_scratch_state, sizeof(g_scratch_state));
	if (!old) return ls;
	ls.words = reinterpret_cast<int32_t *>(old);
	const int type = ls.words[0] % 5;
	ls.rear = ls.words[0] / 5;
	ls.ok = (type == 3) && ls.rear >= 0 && ls.rear < DEG;
	return ls;
}

void history_from(const LiveState &ls, uint32_t *S)
{
	// the front slot (rear + 3) holds the oldest value o_{t-31}
	for (int j = 0; j < DEG; ++j) S[j] = (uint32_t)ls.words[1 + (ls.rear + SEP + j) % DEG];
}

void put_state(LiveState &ls, const uint32_t *S, unsigned long long consumed)
{
	const int new_rear = (int)((ls.rear + consumed) % DEG);
	for (int j = 0; j < DEG; ++j) ls.words[1 + (new_rear + SEP + j) % DEG] = (int32_t)S[j];
	ls.words[0] = 5 * new_rear + 3;
	ls.rear = new_rear;
}

void release_state(LiveState &ls) { setstate(reinterpret_cast<char *>(ls.words)); }

// ---- device ------------------------------------------------------------------------------
__device__ __forceinline__ void dev_mat_vec(const uint32_t *__restrict__ m, uint32_t (&s)[DEG])
{
	uint32_t t[DEG];
#pragma unroll
	for (int i = 0; i < DEG; ++i) {
		uint32_t acc = 0;
#pragma unroll
		for (int k = 0; k < DEG; ++k) acc += __ldg(m + i * DEG + k) * s[k];
		t[i] = acc;
	}
#pragma unroll
	for (int i = 0; i < DEG; ++i) s[i] = t[i];
}

// pw: NPOW matrices, P_j = M^(CHUNK 2^j), row-major 31 x 31.  x: row-major destination block
// (first element of the first column to fill); stream position t -> row t % n, column t / n.
// Several ranks: n is the GLOBAL row count and x holds the rows [row_lo, row_hi) only; every
// rank walks the same stream and keeps its own rows (chunks that miss the slab exit early).
__global__ void __launch_bounds__(RTHREADS)
rand_fill_kernel(const uint32_t *__restrict__ s0, const uint32_t *__restrict__ pw, unsigned long long total,
                 long long n, long long row_lo, long long row_hi, double *__restrict__ x, int ld)
{
	__shared__ uint32_t base[2][DEG];
	const unsigned long long c0 = (unsigned long long)blockIdx.x * RTHREADS;
	if (threadIdx.x < DEG) base[0][threadIdx.x] = s0[threadIdx.x];
	__syncthreads();
	// high bits of the chunk index, once per CTA: 31 threads own one output word each
	int cur = 0;
	for (int j = LOWBITS; j < NPOW; ++j) {
		if (!((c0 >> j) & 1ull)) continue;                 // uniform over the CTA
		if (threadIdx.x < DEG) {
			const uint32_t *m = pw + (size_t)j * DEG * DEG + threadIdx.x * DEG;
			uint32_t acc = 0;
			for (int k = 0; k < DEG; ++k) acc += __ldg(m + k) * base[cur][k];
			base[cur ^ 1][threadIdx.x] = acc;
		}
		__syncthreads();
		cur ^= 1;
	}
	const unsigned long long c = c0 + threadIdx.x;
	const unsigned long long t0 = c * (unsigned long long)CHUNK;
	if (t0 >= total) return;
	{
		const unsigned long long len = (total - t0 < (unsigned long long)CHUNK) ? total - t0 : (unsigned long long)CHUNK;
		const unsigned long long r_first = t0 % (unsigned long long)n;
		if (r_first + len <= (unsigned long long)n &&
		    (r_first >= (unsigned long long)row_hi || r_first + len <= (unsigned long long)row_lo))
			return;                                    // no row of this chunk is ours
	}
	uint32_t s[DEG];
#pragma unroll
	for (int i = 0; i < DEG; ++i) s[i] = base[cur][i];
	for (int j = 0; j < LOWBITS; ++j)
		if ((threadIdx.x >> j) & 1) dev_mat_vec(pw + (size_t)j * DEG * DEG, s);
	long long col = (long long)(t0 / (unsigned long long)n);
	long long row = (long long)(t0 - (unsigned long long)col * (unsigned long long)n);
	unsigned long long left = total - t0;
	for (int r = 0; r < ROUNDS && left > 0; ++r) {
#pragma unroll
		for (int i = 0; i < DEG; ++i) {
			s[i] += s[(i + DEG - SEP) % DEG];
			if (left > 0) {
				if (row >= row_lo && row < row_hi)
					x[(size_t)(row - row_lo) * ld + col] = (double)(s[i] >> 1) * (1.0 / 2147483648.0);
				--left;
				++row;
				if (row == n) { row = 0; ++col; }
			}
		}
	}
}

uint32_t *g_pw_dev = nullptr;          // NPOW x 31 x 31

int ensure_powers()
{
	if (g_pw_dev) return 0;
	std::vector<Mat31> pw(NPOW);
	Mat31 q; step_matrix(q);
	// M^CHUNK by square-and-multiply
	Mat31 acc; memset(&acc, 0, sizeof(acc));
	for (int i = 0; i < DEG; ++i) acc.a[i][i] = 1;
	for (unsigned long long e = (unsigned long long)CHUNK; e; e >>= 1) {
		if (e & 1ull) mat_mul(acc, q, acc);
		if (e > 1) mat_mul(q, q, q);
	}
	pw[0] = acc;
	for (int j = 1; j < NPOW; ++j) mat_mul(pw[j - 1], pw[j - 1], pw[j]);
	B200_CUDA(cudaMalloc(&g_pw_dev, sizeof(Mat31) * NPOW));
	B200_CUDA(cudaMemcpy(g_pw_dev, pw.data(), sizeof(Mat31) * NPOW, cudaMemcpyHostToDevice));
	return 0;
}

}  // namespace

// Host-only self check (no device): the jump-ahead arithmetic against glibc itself.  Predicts
// the generator state `steps` calls ahead, then really calls rand() that often and compares the
// next 64 outputs.  Leaves the stream advanced by steps + 64.  Returns 0 when they agree.
extern "C" int b200_rand_selfcheck(unsigned long long steps)
{
	LiveState ls = grab_state();
	if (!ls.ok) { if (ls.words) release_state(ls); return b200_fail("rand: the process generator is not glibc TYPE_3"); }
	uint32_t S[DEG];
	history_from(ls, S);
	release_state(ls);
	advance(S, steps);
	for (unsigned long long i = 0; i < steps; ++i) (void)rand();
	for (int i = 0; i < 64; ++i) {
		// one more step of the recurrence on the predicted state
		const uint32_t o = S[0] + S[DEG - SEP];
		memmove(S, S + 1, sizeof(uint32_t) * (DEG - 1));
		S[DEG - 1] = o;
		const int want = (int)(o >> 1), got = rand();
		if (want != got) return b200_fail("rand: jump-ahead mismatch at +%d: predicted %d, glibc %d", i, want, got);
	}
	// and the write-back path: advance the live state by 1000 without calling rand()
	ls = grab_state();
	if (!ls.ok) return b200_fail("rand: state grab failed");
	history_from(ls, S);
	uint32_t T[DEG]; memcpy(T, S, sizeof(T));
	advance(T, 1000);
	put_state(ls, T, 1000);
	release_state(ls);
	for (int i = 0; i < 64; ++i) {
		const uint32_t o = T[0] + T[DEG - SEP];
		memmove(T, T + 1, sizeof(uint32_t) * (DEG - 1));
		T[DEG - 1] = o;
		const int want = (int)(o >> 1), got = rand();
		if (want != got) return b200_fail("rand: write-back mismatch at +%d: predicted %d, glibc %d", i, want, got);
	}
	return 0;
}

extern "C" int b200_mv_set_random(b200_mv *x, int start, int end)
{
	B200_REQUIRE_INIT();
	B200_CHECK(x && start >= 0 && end <= x->ncols && start <= end, "b200_mv_set_random: bad arguments");
	const long long n = x->nrows_global;
	if (n == 0 || end == start) return 0;
	const unsigned long long total = (unsigned long long)n * (unsigned long long)(end - start);
	if (ensure_powers()) return 1;
	LiveState ls = grab_state();
	if (!ls.ok) {
		if (ls.words) release_state(ls);
		return b200_fail("b200_mv_set_random: the process generator is not glibc rand() TYPE_3 "
		                 "(initstate()/setstate() was used with another table size)");
	}
	uint32_t S[DEG];
	history_from(ls, S);
	uint32_t *s_dev = (uint32_t *)b200_scratch(3, sizeof(uint32_t) * DEG);
	if (!s_dev) { release_state(ls); return 1; }
	cudaError_t e = cudaMemcpyAsync(s_dev, S, sizeof(S), cudaMemcpyHostToDevice, g_b200.stream);
	if (e == cudaSuccess) {
		const unsigned long long chunks = (total + CHUNK - 1) / CHUNK;
		const unsigned long long ctas = (chunks + RTHREADS - 1) / RTHREADS;
		rand_fill_kernel<<<(unsigned)ctas, RTHREADS, 0, g_b200.stream>>>(s_dev, g_pw_dev, total, n, x->row0, x->row0 + x->nrows,
		                                                                 x->d + start, x->ld);
		B200_LAUNCHED();
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(g_b200.stream);   // S is a stack buffer
	// the process's generator moves on exactly as if rand() had been called `total` times
	advance(S, total);
	put_state(ls, S, total);
	release_state(ls);
	if (e != cudaSuccess) return b200_fail("b200_mv_set_random: %s", cudaGetErrorString(e));
	return 0;
}
