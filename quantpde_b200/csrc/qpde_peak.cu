// qpde_peak.cu — a measured FP64 peak for the rooflines of the control / jump kernels
// (bench.py: "bound": "fp64").  MEASURED_PEAKS.json has the HBM copy rate and a bf16 GEMM rate,
// nothing for the FP64 pipe, so the library times its own chain of independent DFMAs.
#include "qpde_kernels.cuh"

namespace qpde {

namespace {

// 8 independent fused multiply-adds per thread and trip: enough instruction-level parallelism to
// keep the FP64 pipe of an SM sub-partition fed from 16 resident warps
__global__ void __launch_bounds__(256) k_fp64_peak(double *sink, int trips, double seed) {
	double a0 = seed + threadIdx.x, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3.;
	double a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
	const double m = 0.999999, c = 1e-9;
	for(int k = 0; k < trips; ++k) {
		a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
		a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
	}
	const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
	if(s == 12345.6789) sink[0] = s;      // never true: keeps the chain alive
}

} // namespace

// TFLOP/s (2 flops per DFMA) of the best of `reps` timed launches
int measure_fp64_peak(int sm_count, cudaStream_t stream, double *tflops, long long *launches) {
	double *sink = nullptr;
	if(cudaMalloc(&sink, sizeof(double)) != cudaSuccess) return QPDE_ERR_CUDA;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int trips = 1 << 16, blocks = sm_count * 8;
	double best = 0.;
	for(int rep = 0; rep < 4; ++rep) {
		cudaEventRecord(e0, stream);
		k_fp64_peak<<<blocks, 256, 0, stream>>>(sink, trips, 1.0 + rep);
		cudaEventRecord(e1, stream);
		if(cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return QPDE_ERR_CUDA; }
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		*launches += 1;
		const double flops = 2. * 8. * (double)trips * 256. * (double)blocks;
		if(rep > 0 && ms > 0.f) best = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(sink);
	*tflops = best;
	return QPDE_OK;
}

} // namespace qpde
