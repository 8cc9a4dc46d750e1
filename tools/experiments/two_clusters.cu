// Experiment: do two 8-CTA clusters (512 threads, ~128 registers, one CTA per SM) on two streams run at the same
// time, alone and next to a grid of small CTAs that keeps every SM busy?  nvcc -arch=sm_100a two_clusters.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(512, 1) k_cluster(long long cycles, double *sink) {
	extern __shared__ double sm[];
	cg::cluster_group cl = cg::this_cluster();
	const long long t0 = clock64();
	double a = threadIdx.x;
	while(clock64() - t0 < cycles) { a = a * 1.0000001 + 1e-9; cl.sync(); }
	sm[threadIdx.x] = a;
	if(a == 42.) sink[0] = sm[0];
}
__global__ void __launch_bounds__(128) k_small(long long cycles, double *sink) {
	const long long t0 = clock64();
	double a = threadIdx.x;
	while(clock64() - t0 < cycles) a = a * 1.0000001 + 1e-9;
	if(a == 42.) sink[0] = a;
}
static void launch_cluster(cudaStream_t s, long long cycles, double *sink) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(8); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 100 * 1024; cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster, cycles, sink);
	if(e != cudaSuccess) printf("launch: %s\n", cudaGetErrorString(e));
}
int main() {
	double *sink; cudaMalloc(&sink, 8);
	cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
	int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
	cudaStream_t a, b, c;
	cudaStreamCreateWithPriority(&a, cudaStreamNonBlocking, hi); cudaStreamCreateWithPriority(&b, cudaStreamNonBlocking, hi);
	cudaStreamCreateWithPriority(&c, cudaStreamNonBlocking, lo);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const long long ms = 1965000;      // cycles per millisecond at 1965 MHz
	for(int withSmall = 0; withSmall < 2; ++withSmall) {
		for(int two = 0; two < 2; ++two) {
			cudaDeviceSynchronize();
			cudaEventRecord(e0, a);
			if(withSmall) for(int k = 0; k < 40; ++k) k_small<<<148 * 16, 128, 0, c>>>(ms / 20, sink);   // 40 x ~0.1 ms waves of small CTAs
			for(int k = 0; k < 4; ++k) { launch_cluster(a, ms, sink); if(two) launch_cluster(b, ms, sink); }
			cudaStreamSynchronize(b);
			cudaEventRecord(e1, a);
			cudaDeviceSynchronize();
			float t; cudaEventElapsedTime(&t, e0, e1);
			printf("small grid %d, streams %d: 4 x 1 ms clusters per stream took %.2f ms (line streams only)\n", withSmall, two + 1, t);
		}
	}
	return 0;
}
