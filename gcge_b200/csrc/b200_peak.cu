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

// ---- what bounds the DMMA pipe: operand patterns, the plain DFMA pipe, and both together ----
// mode 0: 16 accumulators, one shared (a, b) register pair         (the probe above)
// mode 1: 16 accumulators, 2 A x 8 B fragments, order (a0,bj),(a1,bj)   (gram / lincomb inner loop)
// mode 2: 16 accumulators, 2 A x 8 B fragments, order (a0,b0..7),(a1,b0..7)
// mode 3: 32 independent DFMA chains per thread                     (FP64 FMA pipe)
// mode 4: modes 1 and 3 interleaved 1 DMMA : 2 DFMA                 (do the pipes add up?)
template <int MODE>
__global__ void __launch_bounds__(256)
fp64_peak_kernel(int iters, double *sink)
{
	double c[16][2];
	double f[32];
#pragma unroll
	for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
	for (int i = 0; i < 32; ++i) f[i] = 1e-3 * i;
	double a[2], b[8];
	a[0] = 1.0 + threadIdx.x * 1e-9; a[1] = 1.0 - threadIdx.x * 2e-9;
#pragma unroll
	for (int j = 0; j < 8; ++j) b[j] = 1.0 + (threadIdx.x + j) * 3e-9;
	const double fa = 1.0 + 1e-12, fb = 1e-9;
	for (int it = 0; it < iters; ++it) {
		if (MODE == 0) {
#pragma unroll
			for (int i = 0; i < 16; ++i)
				asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
				             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
		} else if (MODE == 1 || MODE == 4) {
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
				             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a[0]), "d"(b[j]));
				if (MODE == 4) { f[4 * j] = fma(f[4 * j], fa, fb); f[4 * j + 1] = fma(f[4 * j + 1], fa, fb); }
				asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
				             : "+d"(c[8 + j][0]), "+d"(c[8 + j][1]) : "d"(a[1]), "d"(b[j]));
				if (MODE == 4) { f[4 * j + 2] = fma(f[4 * j + 2], fa, fb); f[4 * j + 3] = fma(f[4 * j + 3], fa, fb); }
			}
		} else if (MODE == 2) {
#pragma unroll
			for (int i = 0; i < 2; ++i)
#pragma unroll
				for (int j = 0; j < 8; ++j)
					asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
					             : "+d"(c[8 * i + j][0]), "+d"(c[8 * i + j][1]) : "d"(a[i]), "d"(b[j]));
		} else {
#pragma unroll
			for (int i = 0; i < 32; ++i) f[i] = fma(f[i], fa, fb);
		}
	}
	double s = 0.0;
#pragma unroll
	for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
#pragma unroll
	for (int i = 0; i < 32; ++i) s += f[i];
	if (s == 12345.678) sink[0] = s;
}

template <int MODE>
static int fp64_peak_run(double *sink, double *tflops)
{
	const int iters = 20000, blocks = g_b200.num_sms * 4;
	fp64_peak_kernel<MODE><<<blocks, 256, 0, g_b200.stream>>>(100, sink);
	B200_KERNEL_CHECK();
	double ms = 0, best = 1e30;
	for (int r = 0; r < 3; ++r) {
		if (b200_timer_start()) return 1;
		fp64_peak_kernel<MODE><<<blocks, 256, 0, g_b200.stream>>>(iters, sink);
		B200_KERNEL_CHECK();
		if (b200_timer_stop(&ms)) return 1;
		if (ms < best) best = ms;
	}
	const double mma = (MODE == 3) ? 0.0 : 2.0 * 256.0 * 16.0 * iters * 8.0 * blocks;       // 256 FMA per warp-level mma
	const double dfma = (MODE == 3 || MODE == 4) ? 2.0 * 32.0 * iters * 256.0 * blocks : 0.0;
	*tflops = (mma + dfma) / (best * 1e-3) / 1e12;
	return 0;
}

// out[0..4]: TFLOP/s of the five modes above
extern "C" int b200_measure_fp64_peaks(double *out)
{
	B200_REQUIRE_INIT();
	double *sink = (double *)b200_scratch(3, 64);
	if (!sink) return 1;
	if (fp64_peak_run<0>(sink, out + 0)) return 1;
	if (fp64_peak_run<1>(sink, out + 1)) return 1;
	if (fp64_peak_run<2>(sink, out + 2)) return 1;
	if (fp64_peak_run<3>(sink, out + 3)) return 1;
	if (fp64_peak_run<4>(sink, out + 4)) return 1;
	return 0;
}
