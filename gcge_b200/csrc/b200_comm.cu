// Multi-GPU plumbing: one process per GPU, 1-D contiguous row blocks (SURVEY.md §8e -- the
// scheme of the reference's MPI back ends: every rank owns a row slab of A, B and of every
// multi-vector; reference app/app_slepc.c, app/app_phg.c:292-357, src/ops_multi_vec.c:214).
//
//   rows of rank g : [ floor(g n / G), floor((g+1) n / G) )
//   SpMM           : halo rows of x come from the slab neighbours (ncclSend/ncclRecv over NVLink)
//   Gram / dots    : ncclAllReduce(sum, f64) of the small block on the library stream
//   projected RR   : replicated (identical inputs after the allreduce => identical outputs)
//
// NCCL is loaded with dlopen() so that the library itself has no link-time dependency on it
// (single-GPU users and the CPU-side ABI tests never touch it).  The host process bootstraps
// the communicator: rank 0 calls b200_comm_unique_id(), ships the 128 bytes to the other ranks
// by whatever it has (torch.distributed in bench.py / tests), every rank calls b200_comm_init().
#include "b200_internal.h"
#include <dlfcn.h>

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
enum { NCCL_SUM = 0, NCCL_FLOAT64 = 8, NCCL_INT8 = 0 };

struct NcclApi {
	void *handle = nullptr;
	int (*GetUniqueId)(nccl_uid *) = nullptr;
	int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
	int (*CommDestroy)(nccl_comm) = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
	int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
	int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
} g_nccl;

nccl_comm g_comm = nullptr;

int load_nccl()
{
	if (g_nccl.handle) return 0;
	const char *cands[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2",
	                       "/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib/libnccl.so.2", "libnccl.so"};
	for (const char *c : cands) {
		if (!c || !*c) continue;
		g_nccl.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
		if (g_nccl.handle) break;
	}
	if (!g_nccl.handle) return b200_fail("NCCL: cannot dlopen libnccl.so.2 (%s); set B200_NCCL_LIB", dlerror());
#define LOAD(field, sym)                                                        \
	do {                                                                        \
		*(void **)(&g_nccl.field) = dlsym(g_nccl.handle, sym);                  \
		if (!g_nccl.field) return b200_fail("NCCL: symbol %s missing", sym);    \
	} while (0)
	LOAD(GetUniqueId, "ncclGetUniqueId"); LOAD(CommInitRank, "ncclCommInitRank"); LOAD(CommDestroy, "ncclCommDestroy");
	LOAD(AllReduce, "ncclAllReduce"); LOAD(Send, "ncclSend"); LOAD(Recv, "ncclRecv");
	LOAD(GroupStart, "ncclGroupStart"); LOAD(GroupEnd, "ncclGroupEnd"); LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
	return 0;
}

#define B200_NCCL(call)                                                                         \
	do {                                                                                        \
		int r_ = (call);                                                                        \
		if (r_ != 0) return b200_fail("%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
	} while (0)

}  // namespace

extern "C" int b200_comm_unique_id(char *id128)
{
	B200_CHECK(id128, "b200_comm_unique_id: NULL buffer");
	if (load_nccl()) return 1;
	nccl_uid u;
	B200_NCCL(g_nccl.GetUniqueId(&u));
	memcpy(id128, u.internal, 128);
	return 0;
}

extern "C" int b200_comm_init(int rank, int nranks, const char *id128)
{
	B200_REQUIRE_INIT();
	B200_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "b200_comm_init: rank %d of %d", rank, nranks);
	if (nranks == 1) { g_b200.rank = 0; g_b200.nranks = 1; return 0; }
	B200_CHECK(id128, "b200_comm_init: NULL unique id");
	if (load_nccl()) return 1;
	B200_CHECK(g_comm == nullptr, "b200_comm_init: communicator already initialised");
	nccl_uid u; memcpy(u.internal, id128, 128);
	B200_NCCL(g_nccl.CommInitRank(&g_comm, nranks, u, rank));
	g_b200.rank = rank; g_b200.nranks = nranks;
	return 0;
}

extern "C" int b200_comm_finalize(void)
{
	if (g_comm) {
		cudaStreamSynchronize(g_b200.stream);
		g_nccl.CommDestroy(g_comm);
		g_comm = nullptr;
	}
	g_b200.rank = 0; g_b200.nranks = 1;
	return 0;
}

extern "C" int b200_comm_rank(void) { return g_b200.rank; }
extern "C" int b200_comm_size(void) { return g_b200.nranks > 0 ? g_b200.nranks : 1; }

// Test hook: pretend to be rank `rank` of `nranks` WITHOUT a communicator, so the host-side
// partition logic can be exercised on a machine with no GPU (tests/, gloo, world size 2).
extern "C" int b200_comm_set_layout(int rank, int nranks)
{
	if (nranks < 1 || rank < 0 || rank >= nranks) return b200_fail("b200_comm_set_layout: rank %d of %d", rank, nranks);
	if (g_comm) return b200_fail("b200_comm_set_layout: a communicator is active");
	g_b200.rank = rank; g_b200.nranks = nranks;
	return 0;
}

extern "C" void b200_partition_range(long long n, int rank, int nranks, long long *lo, long long *hi)
{
	if (nranks < 1) nranks = 1;
	*lo = (long long)(((__int128)n * rank) / nranks);
	*hi = (long long)(((__int128)n * (rank + 1)) / nranks);
}

// ---- collectives on the library stream (no-ops on one rank) ---------------------------------
int b200k_allreduce_sum(double *buf_dev, size_t count)
{
	if (g_b200.nranks <= 1 || count == 0) return 0;
	B200_CHECK(g_comm, "allreduce: %d ranks but no communicator (b200_comm_init was not called)", g_b200.nranks);
	B200_NCCL(g_nccl.AllReduce(buf_dev, buf_dev, count, NCCL_FLOAT64, NCCL_SUM, g_comm, g_b200.stream));
	B200_LAUNCHED();
	return 0;
}

// Exchange with nnbr neighbours: send cnt doubles from send_dev + send_off[i] to nbr[i], receive
// into recv_dev + recv_off[i].  Offsets and counts in doubles.
int b200k_neighbor_exchange(int nnbr, const int *nbr, const double *send_dev, const size_t *send_off,
                            const size_t *send_cnt, double *recv_dev, const size_t *recv_off, const size_t *recv_cnt)
{
	if (g_b200.nranks <= 1 || nnbr == 0) return 0;
	B200_CHECK(g_comm, "halo exchange: %d ranks but no communicator (b200_comm_init was not called)", g_b200.nranks);
	B200_NCCL(g_nccl.GroupStart());
	for (int i = 0; i < nnbr; ++i) {
		if (send_cnt[i]) B200_NCCL(g_nccl.Send(send_dev + send_off[i], send_cnt[i], NCCL_FLOAT64, nbr[i], g_comm, g_b200.stream));
		if (recv_cnt[i]) B200_NCCL(g_nccl.Recv(recv_dev + recv_off[i], recv_cnt[i], NCCL_FLOAT64, nbr[i], g_comm, g_b200.stream));
	}
	B200_NCCL(g_nccl.GroupEnd());
	B200_LAUNCHED();
	return 0;
}
