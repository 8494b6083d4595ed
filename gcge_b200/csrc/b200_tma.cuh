// TMA / mbarrier helpers shared by the kernels that stage tiles with the tensor memory accelerator
// (b200_spmm.cu, b200_dense.cu): inline PTX for sm_100a, and the host-side tensor-map encoder
// (fetched through cudaGetDriverEntryPoint so the library does not link against libcuda).
#pragma once
#include "b200_internal.h"
#include <cuda.h>      // CUtensorMap types only

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE;\n"
		"bra WAIT_LOOP;\n"
		"DONE:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
	             ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3,
                                            unsigned long long *bar)
{
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
	             ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): data read once per kernel should leave L2 first
__device__ __forceinline__ void tma_load_4d_hint(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3,
                                                 unsigned long long *bar, unsigned long long policy)
{
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
	             ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity)
{
	unsigned ok;
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		"selp.u32 %0, 1, 0, p;\n"
		"}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_spin(unsigned long long *bar, unsigned parity)
{
	while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static tmap_encode_fn tmap_encoder()
{
	static tmap_encode_fn fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
		    q == cudaDriverEntryPointSuccess)
			fn = (tmap_encode_fn)p;
		else
			cudaGetLastError();
	}
	return fn;
}

