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
//   device : the stream runs down the columns, the multi-vector is stored by rows.  A warp owns 32
//            ADJACENT COLUMNS and a run of L rows: lane j jumps to its own stream position
//            t = column_j * n + first row exactly -- M^t S_0 from the binary expansion of t with the
//            powers M^(2^j) -- and then all lanes run the recurrence in registers in lockstep, one row
//            per step, so every store of the warp is one contiguous 256-byte row segment.  (The first
//            version gave a thread L consecutive stream positions = one column: 8-byte stores a row
//            pitch apart, 141 ms for the 8 M x 400 block of the headline solve, 0.7 TB/s of sectors.)
//            Several ranks fill only their own rows.
#include "b200_internal.h"
#include <vector>

namespace {

constexpr int DEG = 31;                 // TYPE_3 degree
constexpr int SEP = 3;                  // TYPE_3 separation
constexpr int ROUNDS = 512;             // rounds of 31 rows per warp tile
constexpr long long TILE_ROWS = (long long)DEG * ROUNDS;
constexpr int RWARPS = 4;               // warp tiles per CTA
constexpr int NPOW = 48;                // Q_0 .. Q_47 = M^(2^j): stream positions below 2^48

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
// s <- bit ? m s : s  for every lane (m: 31 x 31 row-major, the same for all lanes: broadcast loads)
__device__ __forceinline__ void dev_mat_vec_if(const uint32_t *__restrict__ m, uint32_t (&s)[DEG], bool bit)
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
	for (int i = 0; i < DEG; ++i) s[i] = bit ? t[i] : s[i];
}

// pw: NPOW matrices Q_j = M^(2^j), row-major 31 x 31.  x: first element of the first column to fill of the row-major
// block holding the rows [row_lo, row_hi) of the n global rows; stream position t -> row t % n, column t / n.
// Warp tile w: columns 32 (w % ncg) .. + 31, rows row_lo + TILE_ROWS (w / ncg) .. + TILE_ROWS - 1.
__global__ void __launch_bounds__(32 * RWARPS)
rand_fill_kernel(const uint32_t *__restrict__ s0, const uint32_t *__restrict__ pw, int ncols, int ncg, long long ntiles,
                 long long n, long long row_lo, long long row_hi, double *__restrict__ x, int ld)
{
	const int lane = threadIdx.x & 31;
	const long long w = (long long)blockIdx.x * RWARPS + (threadIdx.x >> 5);
	if (w >= ntiles) return;                                   // (whole warps)
	const int col = 32 * (int)(w % ncg) + lane;
	long long row = row_lo + TILE_ROWS * (w / ncg);
	const long long row_end = (row + TILE_ROWS < row_hi) ? row + TILE_ROWS : row_hi;
	const bool active = col < ncols;
	const unsigned long long t = (unsigned long long)col * (unsigned long long)n + (unsigned long long)row;
	uint32_t s[DEG];
#pragma unroll
	for (int i = 0; i < DEG; ++i) s[i] = __ldg(s0 + i);
	for (int j = 0; j < NPOW; ++j) {
		const bool bit = active && ((t >> j) & 1ull);
		if (!__any_sync(0xffffffffu, bit)) continue;
		dev_mat_vec_if(pw + (size_t)j * DEG * DEG, s, bit);
	}
	double *dst = x + (size_t)(row - row_lo) * ld + col;
	for (int r = 0; r < ROUNDS && row < row_end; ++r) {
#pragma unroll
		for (int i = 0; i < DEG; ++i) {
			s[i] += s[(i + DEG - SEP) % DEG];
			if (row < row_end) {                               // (warp-uniform)
				if (active) *dst = (double)(s[i] >> 1) * (1.0 / 2147483648.0);
				dst += ld; ++row;
			}
		}
	}
}

uint32_t *g_pw_dev = nullptr;          // NPOW x 31 x 31

int ensure_powers()
{
	if (g_pw_dev) return 0;
	std::vector<Mat31> pw(NPOW);
	step_matrix(pw[0]);
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
		const int ncols = end - start, ncg = (ncols + 31) / 32;
		const long long ntiles = (long long)ncg * ((x->nrows + TILE_ROWS - 1) / TILE_ROWS);
		if (ntiles > 0)
			rand_fill_kernel<<<(unsigned)((ntiles + RWARPS - 1) / RWARPS), 32 * RWARPS, 0, g_b200.stream>>>(
				s_dev, g_pw_dev, ncols, ncg, ntiles, n, x->row0, x->row0 + x->nrows, x->d + start, x->ld);
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
