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
#include <vector>

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
enum { NCCL_SUM = 0, NCCL_FLOAT64 = 8, NCCL_INT8 = 0 };

struct NcclApi {
	void *handle = nullptr;
	int (*GetUniqueId)(nccl_uid *) = nullptr;
	int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
	int (*CommDestroy)(nccl_comm) = nullptr;
	int (*AllGather)(const void *, void *, size_t, int, nccl_comm, cudaStream_t) = nullptr;
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
#define LOAD2(field, sym) LOAD(field, sym)
	LOAD(GetUniqueId, "ncclGetUniqueId"); LOAD(CommInitRank, "ncclCommInitRank"); LOAD(CommDestroy, "ncclCommDestroy");
	LOAD(AllReduce, "ncclAllReduce"); LOAD(Send, "ncclSend"); LOAD(Recv, "ncclRecv");
	LOAD(GroupStart, "ncclGroupStart"); LOAD(GroupEnd, "ncclGroupEnd"); LOAD(GetErrorString, "ncclGetErrorString");
	LOAD2(AllGather, "ncclAllGather");
#undef LOAD2
#undef LOAD
	return 0;
}

#define B200_NCCL(call)                                                                         \
	do {                                                                                        \
		int r_ = (call);                                                                        \
		if (r_ != 0) return b200_fail("%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
	} while (0)

}  // namespace

namespace { void p2p_release(); int ar_setup(); void ar_release(); }

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
	if (!b200_opt(B200_OPT_NO_P2P) && ar_setup()) return 1;
	// side stream for the SpMM halo exchange (copy engines over NVLink, see p2p_* below), so that it
	// overlaps the interior rows of the multiply
	g_b200.comm_stream = nullptr;
	if (!b200_opt(B200_OPT_NO_OVERLAP)) {
		B200_CUDA(cudaStreamCreateWithFlags(&g_b200.comm_stream, cudaStreamNonBlocking));
		B200_CUDA(cudaEventCreateWithFlags(&g_b200.ev_x_ready, cudaEventDisableTiming));
		B200_CUDA(cudaEventCreateWithFlags(&g_b200.ev_halo_done, cudaEventDisableTiming));
	}
	return 0;
}

extern "C" int b200_comm_finalize(void)
{
	if (g_comm) {
		cudaStreamSynchronize(g_b200.stream);
		if (g_b200.comm_stream) {
			cudaStreamSynchronize(g_b200.comm_stream);
			p2p_release();
			cudaStreamDestroy(g_b200.comm_stream);
			cudaEventDestroy(g_b200.ev_x_ready); cudaEventDestroy(g_b200.ev_halo_done);
			g_b200.comm_stream = nullptr;
		}
		ar_release();
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

// ================================================================ halo exchange by copy engines
// The slab neighbours' halo rows over NVLink WITHOUT kernels: a peer-to-peer 2-D copy
// (cudaMemcpy2DAsync into the neighbour's IPC-mapped mailbox) and stream memory operations
// (cuStreamWriteValue32 / cuStreamWaitValue32 on IPC-mapped flags) for the hand-shake.  Nothing of
// it needs an SM, so it runs underneath the persistent SpMM CTAs working on the interior rows --
// an NCCL send/recv kernel on a side stream would simply queue behind them.
//
//   sender s -> receiver r, message e (the e-th message on the link s -> r; both ends count it):
//     s: wait  s.free[r]    >= e-1      r has unpacked my previous message
//        copy  x rows -> r.mailbox[slot of s]        (2-D: k*8 bytes wide, pitch ldx*8 -> k*8)
//        write r.arrived[s] =  e
//     r: wait  r.arrived[s] >= e
//        copy  r.mailbox[slot of s] -> halo rows of x
//        write s.free[r]    =  e
//
// Applies to banded matrices split into slabs whose halos come from the two adjacent ranks only
// (stencils, FEM on lattices): slot 0 = the rank below, slot 1 = the rank above.  Everything else
// uses the NCCL exchange above on the library stream.
#include <cuda.h>
namespace {
typedef CUresult (*stream_value32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
struct P2P {
	bool ok = false;
	size_t region_bytes = 0;            // per slot
	double *mailbox = nullptr;          // 2 regions
	unsigned *flags = nullptr;          // [0..1] arrived from below / above, [2..3] free of below / above
	double *peer_mailbox[2] = {nullptr, nullptr};     // below, above
	unsigned *peer_flags[2] = {nullptr, nullptr};
	// messages sent to / received from the rank below [0] and above [1]: both ends of a link count the
	// same stream of messages, whatever mix of matrices (with different halos) the exchanges belong to
	unsigned sent[2] = {0, 0}, recvd[2] = {0, 0};
	stream_value32_fn wait32 = nullptr, write32 = nullptr;
} g_p2p;

void p2p_release()
{
	for (int i = 0; i < 2; ++i) {
		if (g_p2p.peer_mailbox[i]) cudaIpcCloseMemHandle(g_p2p.peer_mailbox[i]);
		if (g_p2p.peer_flags[i]) cudaIpcCloseMemHandle(g_p2p.peer_flags[i]);
		g_p2p.peer_mailbox[i] = nullptr; g_p2p.peer_flags[i] = nullptr;
	}
	if (g_p2p.mailbox) cudaFree(g_p2p.mailbox);
	if (g_p2p.flags) cudaFree(g_p2p.flags);
	g_p2p.mailbox = nullptr; g_p2p.flags = nullptr; g_p2p.region_bytes = 0; g_p2p.ok = false;
	g_p2p.sent[0] = g_p2p.sent[1] = g_p2p.recvd[0] = g_p2p.recvd[1] = 0;
	cudaGetLastError();
}

bool p2p_resolve_driver()
{
	if (g_p2p.wait32 && g_p2p.write32) return true;
	void *a = nullptr, *b = nullptr;
	cudaDriverEntryPointQueryResult q;
	if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &a, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return false; }
	if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &b, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return false; }
	g_p2p.wait32 = (stream_value32_fn)a; g_p2p.write32 = (stream_value32_fn)b;
	return true;
}
}  // namespace

// Collective (called by every rank while it creates a row-partitioned matrix): make sure the mailbox
// holds `rows` halo rows of 64 columns per neighbour; (re)allocates and re-exchanges the IPC handles
// when some rank needs more.  rows < 0: this rank cannot take the matrix through the mailboxes (its
// halos are not the two adjacent slabs, no diagonal image) -- then no rank does (*all_ranks_ok = 0)
// and an existing registration stays as it is.  Failure is not an error: the NCCL exchange stays.
int b200k_p2p_register(int rows, int *all_ranks_ok)
{
	*all_ranks_ok = 0;
	if (g_b200.nranks <= 1 || !g_comm || !g_b200.comm_stream || b200_opt(B200_OPT_NO_P2P)) return 0;
	cudaStream_t st = g_b200.stream;
	// global maximum of the rows needed
	double *dmax = (double *)b200_scratch(8, 256);
	if (!dmax) return 1;
	// max through a sum of one-hot slots would need nranks entries; use nranks doubles and gather
	std::vector<double> all((size_t)g_b200.nranks, 0.0);
	{
		double mine = (double)rows;
		double *dall = (double *)b200_scratch(8, sizeof(double) * ((size_t)g_b200.nranks + 8));
		if (!dall) return 1;
		B200_CUDA(cudaMemsetAsync(dall, 0, sizeof(double) * (size_t)g_b200.nranks, st));
		B200_CUDA(cudaMemcpyAsync(dall + g_b200.rank, &mine, sizeof(double), cudaMemcpyHostToDevice, st));
		B200_NCCL(g_nccl.AllReduce(dall, dall, (size_t)g_b200.nranks, NCCL_FLOAT64, NCCL_SUM, g_comm, st));
		B200_CUDA(cudaMemcpyAsync(all.data(), dall, sizeof(double) * (size_t)g_b200.nranks, cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
	}
	size_t need_rows = 0;
	bool shapes_ok = true;
	for (double v : all) {
		if (v < 0.0) shapes_ok = false;            // some rank cannot take this matrix through the mailboxes
		else if ((size_t)v > need_rows) need_rows = (size_t)v;
	}
	if (!shapes_ok) return 0;
	const size_t need = need_rows * 64 * sizeof(double);
	if (need == 0) return 0;
	if (g_p2p.ok && need <= g_p2p.region_bytes) { *all_ranks_ok = 1; return 0; }
	if (!p2p_resolve_driver()) return 0;
	// (re)build: every rank gets here together
	B200_CUDA(cudaStreamSynchronize(g_b200.comm_stream));
	p2p_release();
	if (cudaMalloc(&g_p2p.mailbox, 2 * need) != cudaSuccess || cudaMalloc(&g_p2p.flags, 256) != cudaSuccess) {
		cudaGetLastError(); p2p_release();
	}
	int good = (g_p2p.mailbox && g_p2p.flags) ? 1 : 0;
	struct Handles { cudaIpcMemHandle_t mb, fl; int good; int pad[3]; };
	Handles mine; memset(&mine, 0, sizeof(mine));
	if (good) {
		B200_CUDA(cudaMemsetAsync(g_p2p.flags, 0, 256, st));
		if (cudaIpcGetMemHandle(&mine.mb, g_p2p.mailbox) != cudaSuccess || cudaIpcGetMemHandle(&mine.fl, g_p2p.flags) != cudaSuccess) {
			cudaGetLastError(); good = 0;
		}
	}
	mine.good = good;
	const size_t hb = sizeof(Handles);
	char *dh = (char *)b200_scratch(8, hb * ((size_t)g_b200.nranks + 1));
	if (!dh) return 1;
	std::vector<Handles> hs((size_t)g_b200.nranks);
	B200_CUDA(cudaMemcpyAsync(dh + hb * g_b200.nranks, &mine, hb, cudaMemcpyHostToDevice, st));
	B200_NCCL(g_nccl.AllGather(dh + hb * g_b200.nranks, dh, hb, NCCL_INT8, g_comm, st));
	B200_CUDA(cudaMemcpyAsync(hs.data(), dh, hb * g_b200.nranks, cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	bool all_good = true;
	for (const Handles &h : hs) all_good = all_good && h.good;
	if (all_good) {
		const int nb[2] = {g_b200.rank - 1, g_b200.rank + 1};
		for (int i = 0; i < 2 && all_good; ++i) {
			if (nb[i] < 0 || nb[i] >= g_b200.nranks) continue;
			void *pm = nullptr, *pf = nullptr;
			if (cudaIpcOpenMemHandle(&pm, hs[nb[i]].mb, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
			    cudaIpcOpenMemHandle(&pf, hs[nb[i]].fl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
				cudaGetLastError(); all_good = false; break;
			}
			g_p2p.peer_mailbox[i] = (double *)pm; g_p2p.peer_flags[i] = (unsigned *)pf;
		}
	}
	// a rank that failed to map a peer must stop everybody from using the path
	{
		double flag = all_good ? 0.0 : 1.0;
		double *df = (double *)b200_scratch(8, 256);
		B200_CUDA(cudaMemcpyAsync(df, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
		B200_NCCL(g_nccl.AllReduce(df, df, 1, NCCL_FLOAT64, NCCL_SUM, g_comm, st));
		B200_CUDA(cudaMemcpyAsync(&flag, df, sizeof(double), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
		all_good = (flag == 0.0);
	}
	if (!all_good) { p2p_release(); return 0; }
	g_p2p.region_bytes = need; g_p2p.ok = true;
	g_p2p.sent[0] = g_p2p.sent[1] = g_p2p.recvd[0] = g_p2p.recvd[1] = 0;
	*all_ranks_ok = 1;
	return 0;
}

// 1 when M's halo exchange can go through the mailboxes (same answer on every rank for a given matrix:
// the plan is symmetric -- what I receive from a neighbour it sends to me)
int b200k_p2p_usable(const b200_mat *M, int k)
{
	if (!g_p2p.ok || !M->p2p_ok || !M->halo_contiguous || k > 64 || M->nnbr < 1 || M->nnbr > 2) return 0;
	for (int i = 0; i < M->nnbr; ++i) {
		if (M->nbr[i] != g_b200.rank - 1 && M->nbr[i] != g_b200.rank + 1) return 0;
		const int nrecv = M->recv_off[i + 1] - M->recv_off[i], nsend = M->send_off[i + 1] - M->send_off[i];
		if ((size_t)nrecv * 64 * sizeof(double) > g_p2p.region_bytes) return 0;
		// the rows sent must be one contiguous range
		for (int j = M->send_off[i] + 1; j < M->send_off[i + 1]; ++j)
			if (M->send_rows[j] != M->send_rows[j - 1] + 1) return 0;
		(void)nsend;
	}
	return 1;
}

#define B200_CU(call)                                                                            \
	do {                                                                                         \
		CUresult r_ = (call);                                                                    \
		if (r_ != CUDA_SUCCESS) return b200_fail("%s:%d %s: CUresult %d", __FILE__, __LINE__, #call, (int)r_); \
	} while (0)

// Enqueue the exchange of the k-column block x (window layout: halo rows in front of and behind the
// local rows) on the comm stream, after everything the library stream has enqueued so far; the
// caller makes the library stream wait for g_b200.ev_halo_done before it touches the halo rows.
int b200k_p2p_halo_exchange(const b200_mat *M, double *x, int ldx, int k)
{
	cudaStream_t cs = g_b200.comm_stream;
	B200_CUDA(cudaEventRecord(g_b200.ev_x_ready, g_b200.stream));
	B200_CUDA(cudaStreamWaitEvent(cs, g_b200.ev_x_ready, 0));
	const size_t width = (size_t)k * sizeof(double);
	// sends
	for (int i = 0; i < M->nnbr; ++i) {
		const int side = (M->nbr[i] < g_b200.rank) ? 0 : 1;           // the neighbour is below / above me
		const int nsend = M->send_off[i + 1] - M->send_off[i];
		if (nsend <= 0) continue;
		// at the neighbour I am the rank above (slot 1) if it is below me, and the other way round
		const int my_slot_there = 1 - side;
		const unsigned e = ++g_p2p.sent[side];                          // my e-th message on this link
		if (e > 1) B200_CU(g_p2p.wait32((CUstream)cs, (CUdeviceptr)(g_p2p.flags + 2 + side), e - 1, CU_STREAM_WAIT_VALUE_GEQ));
		double *dst = (double *)((char *)g_p2p.peer_mailbox[side] + (size_t)my_slot_there * g_p2p.region_bytes);
		const double *src = x + (size_t)M->send_rows[M->send_off[i]] * ldx;
		B200_CUDA(cudaMemcpy2DAsync(dst, width, src, (size_t)ldx * sizeof(double), width, (size_t)nsend, cudaMemcpyDeviceToDevice, cs));
		B200_CU(g_p2p.write32((CUstream)cs, (CUdeviceptr)(g_p2p.peer_flags[side] + my_slot_there), e, CU_STREAM_WRITE_VALUE_DEFAULT));
	}
	// receives
	for (int i = 0; i < M->nnbr; ++i) {
		const int side = (M->nbr[i] < g_b200.rank) ? 0 : 1;
		const int nrecv = M->recv_off[i + 1] - M->recv_off[i];
		if (nrecv <= 0) continue;
		const unsigned e = ++g_p2p.recvd[side];                         // the neighbour's e-th message to me
		B200_CU(g_p2p.wait32((CUstream)cs, (CUdeviceptr)(g_p2p.flags + side), e, CU_STREAM_WAIT_VALUE_GEQ));
		const double *src = (const double *)((const char *)g_p2p.mailbox + (size_t)side * g_p2p.region_bytes);
		double *dst = (side == 0) ? x - (size_t)nrecv * ldx : x + (size_t)M->nrows * ldx;
		B200_CUDA(cudaMemcpy2DAsync(dst, (size_t)ldx * sizeof(double), src, width, width, (size_t)nrecv, cudaMemcpyDeviceToDevice, cs));
		B200_CU(g_p2p.write32((CUstream)cs, (CUdeviceptr)(g_p2p.peer_flags[side] + 2 + (1 - side)), e, CU_STREAM_WRITE_VALUE_DEFAULT));
	}
	B200_CUDA(cudaEventRecord(g_b200.ev_halo_done, cs));
	B200_LAUNCHED();
	return 0;
}

// ================================================================ allreduce inside a kernel
// The CG step needs two tiny allreduces per iteration (diag(p^T A p) and diag(r^T r), k doubles each).
// Through NCCL each is a kernel launch of its own (~35 us at 8 GPUs) followed by a one-CTA kernel for the
// scalar update.  Instead, the last CTA of the streaming kernel that produced the partial sums writes
// its k values into an inbox on EVERY rank (remote stores over NVLink into IPC-mapped memory), raises
// a flag there, waits for the flags of all the others in its own memory, adds the nranks
// contributions in rank order -- the same bits on every rank -- and goes on with the scalar update.
// Two inbox parities: a rank can run at most one allreduce ahead of the slowest (it needs everybody's
// flag for the one in between).  A wait that lasts longer than ~35 s records a time-out in a status word
// (checked by the host at the end of the solve) instead of hanging the GPU.
namespace {
struct ArState {
	bool ok = false;
	double *inbox = nullptr; unsigned *flags = nullptr;      // local, exported
	void *peer_inbox[B200_AR_MAX_RANKS] = {}, *peer_flags[B200_AR_MAX_RANKS] = {};
} g_ar;

void ar_release()
{
	for (int q = 0; q < B200_AR_MAX_RANKS; ++q) {
		if (g_ar.peer_inbox[q]) cudaIpcCloseMemHandle(g_ar.peer_inbox[q]);
		if (g_ar.peer_flags[q]) cudaIpcCloseMemHandle(g_ar.peer_flags[q]);
		g_ar.peer_inbox[q] = g_ar.peer_flags[q] = nullptr;
	}
	if (g_ar.inbox) cudaFree(g_ar.inbox);
	if (g_ar.flags) cudaFree(g_ar.flags);
	g_ar.inbox = nullptr; g_ar.flags = nullptr; g_ar.ok = false;
	cudaGetLastError();
}

// collective, from b200_comm_init.  Failure to map is not an error: NCCL stays in place.
int ar_setup()
{
	const int nr = g_b200.nranks;
	if (nr < 2 || nr > B200_AR_MAX_RANKS) return 0;
	cudaStream_t st = g_b200.stream;
	const size_t inbox_bytes = sizeof(double) * 2 * nr * B200_AR_MAX_COUNT, flag_bytes = 1024;
	int good = 1;
	if (cudaMalloc(&g_ar.inbox, inbox_bytes) != cudaSuccess || cudaMalloc(&g_ar.flags, flag_bytes) != cudaSuccess) { cudaGetLastError(); good = 0; }
	struct Handles { cudaIpcMemHandle_t ib, fl; int good; int pad[3]; };
	Handles mine; memset(&mine, 0, sizeof(mine));
	if (good) {
		B200_CUDA(cudaMemsetAsync(g_ar.inbox, 0, inbox_bytes, st));
		B200_CUDA(cudaMemsetAsync(g_ar.flags, 0, flag_bytes, st));
		if (cudaIpcGetMemHandle(&mine.ib, g_ar.inbox) != cudaSuccess || cudaIpcGetMemHandle(&mine.fl, g_ar.flags) != cudaSuccess) { cudaGetLastError(); good = 0; }
	}
	mine.good = good;
	const size_t hb = sizeof(Handles);
	char *dh = (char *)b200_scratch(8, hb * ((size_t)nr + 1));
	if (!dh) return 1;
	std::vector<Handles> hs((size_t)nr);
	B200_CUDA(cudaMemcpyAsync(dh + hb * nr, &mine, hb, cudaMemcpyHostToDevice, st));
	B200_NCCL(g_nccl.AllGather(dh + hb * nr, dh, hb, NCCL_INT8, g_comm, st));
	B200_CUDA(cudaMemcpyAsync(hs.data(), dh, hb * nr, cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	bool all_good = true;
	for (const Handles &h : hs) all_good = all_good && h.good;
	for (int q = 0; q < nr && all_good; ++q) {
		if (q == g_b200.rank) continue;
		if (cudaIpcOpenMemHandle(&g_ar.peer_inbox[q], hs[q].ib, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
		    cudaIpcOpenMemHandle(&g_ar.peer_flags[q], hs[q].fl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
			cudaGetLastError(); all_good = false;
		}
	}
	{   // everybody or nobody
		double flag = all_good ? 0.0 : 1.0;
		double *df = (double *)b200_scratch(8, 256);
		B200_CUDA(cudaMemcpyAsync(df, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
		B200_NCCL(g_nccl.AllReduce(df, df, 1, NCCL_FLOAT64, NCCL_SUM, g_comm, st));
		B200_CUDA(cudaMemcpyAsync(&flag, df, sizeof(double), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
		all_good = (flag == 0.0);
	}
	if (!all_good) { ar_release(); return 0; }
	g_ar.ok = true;
	return 0;
}
}  // namespace

B200ArCtx b200k_ar_ctx()
{
	B200ArCtx c; memset(&c, 0, sizeof(c));
	if (!g_ar.ok || b200_opt(B200_OPT_NO_KERNEL_ALLREDUCE)) return c;
	c.nranks = g_b200.nranks; c.rank = g_b200.rank;
	for (int q = 0; q < g_b200.nranks; ++q) {
		c.inbox[q] = (q == g_b200.rank) ? g_ar.inbox : (double *)g_ar.peer_inbox[q];
		c.flags[q] = (q == g_b200.rank) ? g_ar.flags : (unsigned *)g_ar.peer_flags[q];
	}
	c.seq = g_ar.flags + 2 * B200_AR_MAX_RANKS;            // behind the [2][nranks] flags
	c.status = (int *)(g_ar.flags + 2 * B200_AR_MAX_RANKS + 1);
	return c;
}

extern "C" int b200k_ar_check(void)
{
	if (!g_ar.ok) return 0;
	int st = 0;
	B200_CUDA(cudaMemcpyAsync(&st, g_ar.flags + 2 * B200_AR_MAX_RANKS + 1, sizeof(int), cudaMemcpyDeviceToHost, g_b200.stream));
	B200_CUDA(cudaStreamSynchronize(g_b200.stream));
	B200_CHECK(st == 0, "in-kernel allreduce: a rank waited more than 35 s for its peers (the ranks no longer run the same sequence of kernels)");
	return 0;
}
