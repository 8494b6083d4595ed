// Measurement utility: the FP64 tensor (DMMA) issue-rate ceiling of this GPU, the roofline
// denominator for the Gram / LinearComb kernels (MEASURED_PEAKS.json has no FP64 number).
// Register-resident mma.sync.m8n8k4.f64 chains, no memory traffic, all SMs, 8 warps/SMSP-quad.
#include "b200_internal.h"

__global__ void __launch_bounds__(256)
dmma_peak_kernel(int iters, double *sink)
{
	double c[16][2];
#pragma unroll
	for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
	double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 16; ++i)
			asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
			             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
	}
	double s = 0.0;
#pragma unroll
	for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
	if (s == 12345.678) sink[0] = s;
}

extern "C" int b200_measure_dmma_peak(double *tflops)
{
	B200_REQUIRE_INIT();
	double *sink = (double *)b200_scratch(3, 64);
	if (!sink) return 1;
	const int iters = 20000, blocks = g_b200.num_sms * 4;
	dmma_peak_kernel<<<blocks, 256, 0, g_b200.stream>>>(100, sink);
	B200_KERNEL_CHECK();
	double ms = 0, best = 1e30;
	for (int r = 0; r < 3; ++r) {
		if (b200_timer_start()) return 1;
		dmma_peak_kernel<<<blocks, 256, 0, g_b200.stream>>>(iters, sink);
		B200_KERNEL_CHECK();
		if (b200_timer_stop(&ms)) return 1;
		if (ms < best) best = ms;
	}
	const double flops = 2.0 * 256.0 * 16.0 * iters * 8.0 * blocks;   // 8x8x4 MACs per warp-level mma
	*tflops = flops / (best * 1e-3) / 1e12;
	return 0;
}
